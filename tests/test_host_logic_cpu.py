"""CPU: host-side logic of the drop-in mirrors - the CommonClustering label post-steps against the oracle restatement
of speakerlab/process/cluster.py:204-239, and construction of every mirror through the reference's own plug-in API
(speakerlab/utils/builder.py:9-12,52-91 resolves `obj:` dotted names from YAML)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import b200spk
from oracle import cluster_oracle, gen_golden

REF = "/root/reference"


def _post_steps_case(seed, n, k, d, minor):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((k, d)).astype(np.float32)
    centers[1] = centers[0] + 0.15 * rng.standard_normal(d).astype(np.float32)      # a pair that merge_by_cos joins
    truth = rng.integers(0, k, n)
    X = (centers[truth] + 0.4 * rng.standard_normal((n, d))).astype(np.float32)
    labels = truth.astype(np.int64) * 3 + 1                       # non-contiguous ids
    for i in range(minor):                                        # tiny clusters that must be dissolved
        labels[rng.integers(0, n, rng.integers(1, 4))] = 100 + i
    return X, labels


@pytest.mark.parametrize("seed,n,k,d,minor", [(0, 300, 5, 32, 3), (1, 80, 3, 16, 0), (2, 500, 8, 64, 6), (3, 12, 4, 8, 2)])
def test_filter_and_merge_match_oracle(seed, n, k, d, minor):
    X, labels = _post_steps_case(seed, n, k, d, minor)
    cc = b200spk.CommonClustering("spectral", mer_cos=0.8, min_cluster_size=4)
    got = cc.filter_minor_cluster(labels.copy(), X, 4)
    ref = cluster_oracle.filter_minor_cluster(labels.copy(), X, 4)
    assert np.array_equal(got, ref)
    for thr in (0.95, 0.8, 0.3):
        got_m = cc.merge_by_cos(got.copy(), X, thr)
        ref_m = cluster_oracle.merge_by_cos(ref.copy(), X, thr)
        assert np.array_equal(got_m, ref_m)
    if n >= 80:
        assert len(np.unique(cc.merge_by_cos(got.copy(), X, 0.8))) < len(np.unique(got))      # the planted pair merges


def test_filter_all_minor_gives_single_label():
    X = np.random.default_rng(5).standard_normal((6, 8)).astype(np.float32)
    cc = b200spk.CommonClustering("AHC", min_cluster_size=4)
    assert cc.filter_minor_cluster(np.array([0, 0, 1, 1, 2, 2]), X, 4).tolist() == [0] * 6
    assert np.array_equal(cc.filter_minor_cluster(np.zeros(6, dtype=np.int64), X, 4), np.zeros(6))


def test_trivial_inputs_and_argument_checks():
    cc = b200spk.CommonClustering("spectral", mer_cos=0.8)
    assert cc(np.zeros((0, 4), dtype=np.float32)).shape == (0,)
    assert cc(np.ones((1, 4), dtype=np.float32)).tolist() == [0]
    with pytest.raises(AssertionError):
        cc(np.zeros(5, dtype=np.float32))
    with pytest.raises(ValueError):
        b200spk.CommonClustering("umap_hdbscan")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_mirrors_build_through_the_reference_builder(tmp_path):
    """The one-line swap INTEGRATION.md describes: the recipe YAML with `obj:` pointing at b200spk classes goes
    through speakerlab.utils.config/build unchanged and yields objects with the attributes the callers read."""
    sys.path.insert(0, REF)
    try:
        from speakerlab.utils.builder import build
        from speakerlab.utils.config import build_config
    finally:
        sys.path.remove(REF)
    src = open(os.path.join(REF, "egs/3dspeaker/speaker-diarization/conf/diar.yaml")).read()
    src = src.replace("speakerlab.process.processor.FBank", "b200spk.FBank")
    src = src.replace("speakerlab.models.campplus.DTDNN.CAMPPlus", "b200spk.CAMPPlus")
    src = src.replace("speakerlab.process.cluster.CommonClustering", "b200spk.CommonClustering")
    assert src.count("obj: b200spk.") >= 3
    path = tmp_path / "diar_b200.yaml"
    path.write_text("sample_rate: 16000\n" + src)
    config = build_config(str(path))
    fe = build("feature_extractor", config)
    model = build("embedding_model", config)
    cluster = build("cluster", config)
    assert isinstance(fe, b200spk.FBank) and (fe.n_mels, fe.sample_rate, fe.mean_nor) == (80, 16000, True)
    assert isinstance(model, b200spk.CAMPPlus) and model.embedding_size == 192 and isinstance(model, torch.nn.Module)
    assert isinstance(cluster, b200spk.CommonClustering) and isinstance(cluster.cluster, b200spk.SpectralCluster)
    assert (cluster.cluster.max_num_spks, cluster.cluster.pval, cluster.mer_cos, cluster.min_cluster_size) == (15, 0.012, 0.8, 4)
    # the SV recipes name the networks the same way (egs/*/sv-*/conf/*.yaml: embedding_model obj + args)
    for obj, args in (("b200spk.ERes2NetV2", {"feat_dim": 80, "embedding_size": 192, "baseWidth": 26, "scale": 2, "expansion": 2}),
                      ("b200spk.ECAPA_TDNN", {"input_size": 80, "lin_neurons": 192})):
        cfg = build_config.__globals__["Config"]({"embedding_model": {"obj": obj, "args": dict(args)}})
        m = build("embedding_model", cfg)
        assert isinstance(m, torch.nn.Module) and not m.training


# --------------------------------------------------------------------------- turns / RTTM, ark/scp, trial metrics
def test_compress_segments_matches_the_sequential_reference_rule():
    from oracle import metrics_oracle, synth
    rng = np.random.default_rng(0)
    for trial in range(20):
        regions, t = [], 0.0
        for _ in range(rng.integers(1, 6)):                      # VAD regions separated by silences
            t += float(rng.uniform(0.2, 3.0))
            dur = float(rng.uniform(0.4, 30.0))
            regions.append([round(t, 3), round(t + dur, 3)])
            t += dur
        chunks = [c for st, ed in regions for c in synth.chunk(st, ed)]
        labels = np.repeat(rng.integers(0, 3, len(chunks) // 3 + 1), 3)[:len(chunks)]       # turns of >= 1 segment
        ref = metrics_oracle.compressed_seg([[st, ed, int(c)] for (st, ed), c in zip(chunks, labels)])
        got = b200spk.compress_segments(chunks, labels)
        assert len(got) == len(ref)
        assert np.allclose(np.array(got, dtype=np.float64), np.array(ref, dtype=np.float64), atol=1e-12)
    assert b200spk.compress_segments([], []) == []


def test_rttm_format(tmp_path):
    turns = b200spk.write_rttm(str(tmp_path / "a.rttm"), [[0.0, 1.5], [0.75, 2.25], [1.5, 3.0]], [0, 0, 1], "rec1")
    assert turns == [[0.0, 1.875, 0], [1.875, 3.0, 1]]
    assert open(tmp_path / "a.rttm").read() == ("SPEAKER rec1 0 0.000 1.875 <NA> <NA> 1 <NA> <NA>\n"
                                                "SPEAKER rec1 0 1.875 1.125 <NA> <NA> 2 <NA> <NA>\n")


def test_ark_scp_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    emb = rng.standard_normal((5, 192)).astype(np.float32)
    keys = ["utt%02d" % i for i in range(5)]
    ark, scp = str(tmp_path / "xvector_00.ark"), str(tmp_path / "xvector_00.scp")
    with b200spk.ArkWriter(ark, scp) as w:
        w.write_batch(keys[:4], emb[:4])
        w(keys[4], emb[4])                                   # a plain vector entry
    back = dict(b200spk.read_ark(ark))
    assert list(back) == keys
    assert all(back[k].shape == (1, 192) and np.array_equal(back[k][0], emb[i]) for i, k in enumerate(keys[:4]))
    assert back[keys[4]].shape == (192,) and np.array_equal(back[keys[4]], emb[4])
    by_scp = b200spk.read_scp(scp)
    assert all(np.array_equal(by_scp[k], back[k]) for k in keys)
    # byte layout of the first entry: key, space, \0B, 'FM ', \4 rows \4 cols
    raw = open(ark, "rb").read(6 + 2 + 3 + 10)
    assert raw == b"utt00 \0BFM \4\x01\x00\x00\x00\4\xc0\x00\x00\x00"


def test_det_metrics_oracle_on_a_separable_toy():
    from oracle import metrics_oracle
    scores = np.array([0.1, 0.2, 0.3, 0.4, 0.6, 0.7, 0.8, 0.9])
    labels = np.array([0, 0, 0, 1, 0, 1, 1, 1])
    fnr, fpr = metrics_oracle.pmiss_pfa(scores, labels)
    e, thr = metrics_oracle.eer(fnr, fpr, scores)
    assert abs(e - 0.25) < 1e-12 and thr in (0.4, 0.6)
    assert abs(metrics_oracle.c_norm(fnr, fpr, 0.5) - 0.25) < 1e-12


def test_bulk_chunk_table_is_circle_pad_then_slice():
    """IterWavList.load_wav / chunk_wav (infer_sv_batch.py:388-412): truncate to 90 s, circle-pad the recording to a
    whole number of 10 s chunks, slice.  The table must address exactly those samples."""
    fs, cs, cap = 16000, 160000, 90 * 16000
    lengths = [cs * 3 + 5, 100, fs * 95, cs]
    rng = np.random.default_rng(2)
    wavs = [rng.standard_normal(n).astype(np.float32) for n in lengths]
    buf = np.concatenate(wavs)
    starts, periods, phases, pos = b200spk.chunk_table(lengths)
    assert pos.tolist() == [0, 4, 5, 14, 15]
    row = 0
    for w in wavs:
        w = w[:cap]
        n = int(np.ceil(len(w) / cs))
        padded = np.tile(w, int(np.ceil(n * cs / len(w))))[:n * cs]
        for q in range(n):
            idx = starts[row] + (phases[row] + np.arange(cs)) % periods[row]
            assert np.array_equal(buf[idx], padded[q * cs:(q + 1) * cs])
            row += 1
    assert row == len(starts)


def test_bench_stdout_carries_only_the_result_line():
    """bench.py's contract is ONE JSON line on stdout.  Native libraries print there too (NCCL's version banner at
    N > 1), so bench.main() points file descriptor 1 at stderr and keeps a private handle for the result."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'banner from a native library\\n'); print('python-level noise'); bench.emit({'value': 1})" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    assert r.stdout == json.dumps({"value": 1}) + "\n"
    assert "banner from a native library" in r.stderr and "python-level noise" in r.stderr
