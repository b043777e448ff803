"""CPU: the oracle restatements against the golden vectors minted from the imported reference."""
import json
import os

import numpy as np
import pytest

from oracle import campplus_oracle, cluster_oracle, fbank_oracle, gen_golden, synth


@pytest.fixture(scope="module")
def fb_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "fbank.npz"))


@pytest.mark.parametrize("case", list(gen_golden.fbank_cases().keys()))
def test_fbank_oracle_matches_reference(fb_gold, case):
    wavs = gen_golden.fbank_cases()[case]
    ref32, ref64 = fb_gold[case + ".f32"], fb_gold[case + ".f64"]
    got32 = fbank_oracle.fbank_batch(wavs, dtype=np.float32)
    got64 = fbank_oracle.fbank_batch(wavs, dtype=np.float64)
    assert got32.shape == ref32.shape
    # fp64 restatement vs the reference run on float64 input: tight
    assert np.abs(got64 - ref64).max() < 1e-6
    # two-sided fp32 criterion: within 1e-4 of the fp64 truth everywhere; within
    # 1e-4 of ref-fp32 on >= 99.9 % of elements, the rest explained by ref-fp32's own error
    assert np.abs(got32 - ref64).max() <= 1e-4
    d = np.abs(got32 - ref32)
    assert (d > 1e-4).mean() <= 1e-3
    assert np.all(d <= 1e-4 + np.abs(ref32 - ref64))


def test_fbank_raw_no_cmn(fb_gold):
    wavs = gen_golden.fbank_cases()["noise_1p5s"]
    got = fbank_oracle.fbank_batch(wavs, mean_nor=False, dtype=np.float64)
    ref = fb_gold["noise_1p5s.raw_f32"]
    assert np.abs(got - ref).max() < 2e-3
    assert np.mean(np.abs(got - ref) > 1e-4) < 1e-3


def test_mel_bank_structure():
    mel = fbank_oracle.mel_banks(80)
    assert mel.shape == (80, 256)
    nnz = (mel > 0).sum(1)
    assert int((mel > 0).sum()) == 501          # SURVEY 8a a6 [probe]
    assert nnz.min() >= 1 and nnz.max() <= 16


def test_num_frames():
    assert fbank_oracle.num_frames(24000) == 148
    assert fbank_oracle.num_frames(48000) == 298
    assert fbank_oracle.num_frames(160000) == 998
    assert fbank_oracle.num_frames(399) == 0


@pytest.fixture(scope="module")
def cam_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "campplus.npz"))


@pytest.fixture(scope="module")
def layouts(golden_dir):
    with open(os.path.join(golden_dir, "state_dict_layouts.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", gen_golden.campplus_cases(), ids=lambda c: c[0])
def test_campplus_oracle_matches_reference(cam_gold, layouts, case):
    name, emb, batch, n_samples, wseed, bnrand = case
    shapes = layouts["campplus_e%d" % emb]
    sd = synth.fill_state_dict(shapes, wseed, randomize_bn=bnrand)
    feats = cam_gold[name + ".feats"]
    taps = {}
    got = campplus_oracle.forward(sd, feats, taps).numpy()
    ref = cam_gold[name + ".emb"]
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert rel < 1e-4, rel   # fp32 re-association noise through 52 dense layers
    for k in ("head", "xvector.tdnn", "xvector.block1", "xvector.transit1", "xvector.block2",
              "xvector.transit2", "xvector.block3", "xvector.transit3", "xvector.stats"):
        v = taps[k]
        fp = np.array([v.mean().item(), v.abs().mean().item(), v.double().norm().item()])
        np.testing.assert_allclose(fp, cam_gold[name + ".tap." + k], rtol=2e-4, atol=1e-5)


def test_campplus_input_regenerates(cam_gold):
    # the goldens' feats come from seeded waveforms: the seed path must reproduce them
    name, emb, batch, n_samples, wseed, bnrand = gen_golden.campplus_cases()[0]
    wavs = gen_golden.campplus_input(batch, n_samples, seed=wseed + 1000)
    feats = fbank_oracle.fbank_batch(wavs)
    assert np.abs(feats - cam_gold[name + ".feats"]).max() < 2e-3


def test_seg_pooling_partial_window():
    import torch
    x = torch.arange(149, dtype=torch.float32).view(1, 1, 149)
    s = campplus_oracle.seg_pooling(x)
    assert s.shape == x.shape
    assert s[0, 0, 0].item() == 49.5 and s[0, 0, 148].item() == 124.0   # SURVEY a14 [probe]


@pytest.fixture(scope="module")
def cl_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cluster.npz"))


@pytest.mark.parametrize("case", gen_golden.cluster_cases(), ids=lambda c: c[0])
def test_cluster_oracle_matches_reference(cl_gold, case):
    name, n, d, k, seed, kw = case
    X, truth = gen_golden.cluster_input(n, d, k, seed)
    X0 = X.copy()
    np.random.seed(0)
    labels, st = cluster_oracle.spectral_cluster(X, return_stages=True, **kw)
    assert np.array_equal(X, X0)                       # must not mutate X
    assert st["k"] == int(cl_gold[name + ".k"])
    np.testing.assert_allclose(st["lambdas"], cl_gold[name + ".lambdas"], atol=1e-3)
    np.testing.assert_allclose(np.diag(st["laplacian"]), cl_gold[name + ".lap_diag"], rtol=1e-5)
    ref = cl_gold[name + ".labels"]
    assert np.array_equal(cluster_oracle.match_labels(ref, labels), ref)


def test_prune_count():
    assert cluster_oracle.prune_count(4799, 0.012) == 4741     # keeps 58 per row
    assert cluster_oracle.prune_count(10, 0.02) == 4


def test_chunking_one_hour():
    ch = synth.chunk(0.0, 3600.0)
    assert len(ch) == 4799 and ch[-1] == [3598.5, 3600.0]       # SURVEY 3.2 [probe]
    assert synth.chunk(0.0, 1.0) == [[0.0, 1.0]]
    assert synth.chunk(2.0, 2.0) == []


def test_circle_pad():
    x = np.arange(5)
    assert synth.circle_pad(x, 12).tolist() == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1]
    assert synth.circle_pad(x, 3) is x


@pytest.mark.parametrize("case", gen_golden.eres2netv2_cases(), ids=lambda c: c[0])
def test_eres2netv2_oracle_matches_reference(golden_dir, layouts, case):
    from oracle import eres2netv2_oracle
    gold = np.load(os.path.join(golden_dir, "eres2netv2.npz"))
    name, kw, batch, n_samples, wseed = case
    key = "eres2netv2_w%ds%de%d" % (kw["baseWidth"], kw["scale"], kw["expansion"])
    sd = synth.fill_state_dict(layouts[key], wseed, randomize_bn=True, gain=gen_golden.ERES_GAIN)
    got = eres2netv2_oracle.forward(sd, gold[name + ".feats"], scale=kw["scale"]).numpy()
    ref = gold[name + ".emb"]
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-5


@pytest.mark.parametrize("case", gen_golden.ecapa_cases(), ids=lambda c: c[0])
def test_ecapa_oracle_matches_reference(golden_dir, layouts, case):
    from oracle import ecapa_oracle
    gold = np.load(os.path.join(golden_dir, "ecapa.npz"))
    name, kw, batch, n_samples, wseed = case
    sd = synth.fill_state_dict(layouts["ecapa_c%d" % kw["channels"][0]], wseed, randomize_bn=True, gain=gen_golden.ECAPA_GAIN)
    got = ecapa_oracle.forward(sd, gold[name + ".feats"]).numpy()
    ref = gold[name + ".emb"]
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-5
    # the fp32 reference itself is within 1e-6 of the same graph in fp64: these test networks are well conditioned
    assert np.linalg.norm(ref - gold[name + ".emb_f64"]) / np.linalg.norm(ref) < 1e-5


@pytest.mark.parametrize("case", gen_golden.common_cases(), ids=lambda c: c[0])
def test_common_clustering_oracle_matches_reference(golden_dir, case):
    from oracle import cluster_oracle
    gold = np.load(os.path.join(golden_dir, "common_clustering.npz"))
    name, n, d, k, seed, noise, outliers, kw = case
    X, _ = gen_golden.common_input(n, d, k, seed, noise, outliers)
    np.random.seed(0)
    got = cluster_oracle.common_clustering(X.copy(), **kw)
    ref = gold[name + ".labels"]
    assert np.array_equal(cluster_oracle.match_labels(ref, got), ref)
    if kw["cluster_type"] == "AHC":
        raw = cluster_oracle.ahc(X, kw.get("fix_cos_thr", 0.4))
        assert np.array_equal(cluster_oracle.match_labels(gold[name + ".raw_ahc"], raw), gold[name + ".raw_ahc"])


@pytest.mark.parametrize("case", gen_golden.eres2net_cases(), ids=lambda c: c[0])
def test_eres2net_v1_oracle_matches_reference(golden_dir, layouts, case):
    from oracle import eres2net_oracle
    gold = np.load(os.path.join(golden_dir, "eres2net.npz"))
    name, variant, kw, batch, n_samples, wseed = case
    sd = synth.fill_state_dict(layouts["eres2net_" + variant], wseed, randomize_bn=True, gain=gen_golden.ERES_GAIN)
    got = eres2net_oracle.forward(sd, gold[name + ".feats"], scale=kw.get("scale", 2)).numpy()
    ref = gold[name + ".emb"]
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-5
