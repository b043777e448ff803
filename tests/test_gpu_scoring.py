"""GPU: bulk extraction plumbing (chunk gather, per-recording mean) and trial scoring metrics (SURVEY 8f rows 3-4,
BASELINE config 5) against the oracle restatements."""
import numpy as np
import pytest
import torch

import b200spk
from oracle import metrics_oracle, synth

pytestmark = pytest.mark.gpu


def test_segment_mean_matches_numpy():
    rng = np.random.default_rng(3)
    E = rng.standard_normal((23, 192)).astype(np.float32)
    pos = np.array([0, 9, 10, 10, 23], dtype=np.int32)              # includes an empty recording
    got = b200spk.segment_mean(torch.from_numpy(E).cuda(), torch.from_numpy(pos).cuda()).cpu().numpy()
    for w in range(4):
        ref = E[pos[w]:pos[w + 1]].mean(0) if pos[w + 1] > pos[w] else np.zeros(192, np.float32)
        assert np.allclose(got[w], ref, atol=1e-6)


def test_bulk_extractor_matches_materialised_chunks():
    """Recordings of odd lengths back to back in one int16 buffer: the window-mode front end + per-recording mean
    against fbank on explicitly circle-padded, sliced chunks (infer_sv_batch.py:388-412) and a numpy mean."""
    fs, cs = 16000, 160000
    lengths = [cs + 777, 4000, 2 * cs]
    rng = np.random.default_rng(4)
    wavs = [np.clip(np.round(3000 * rng.standard_normal(n)), -32768, 32767).astype(np.int16) for n in lengths]
    buf = torch.from_numpy(np.concatenate(wavs)).cuda()
    torch.manual_seed(3)
    model = b200spk.CAMPPlus(embedding_size=192, precision="fp32").cuda().eval()
    fb = b200spk.FBank(80, 16000, mean_nor=True)
    bx = b200spk.BulkExtractor(fb, model, batchsize=2)
    got = bx(buf, lengths).cpu().numpy()
    chunks, owner = [], []
    for i, w in enumerate(wavs):
        n = int(np.ceil(len(w) / cs))
        padded = np.tile(w, int(np.ceil(n * cs / len(w))))[:n * cs]
        chunks += [padded[q * cs:(q + 1) * cs] for q in range(n)]
        owner += [i] * n
    with torch.no_grad():
        emb = model(fb.batch(torch.from_numpy(np.stack(chunks)).cuda())).cpu().numpy()
    owner = np.array(owner)
    ref = np.stack([emb[owner == i].mean(0) for i in range(len(wavs))])
    assert got.shape == (3, 192)
    assert np.allclose(got, ref, atol=1e-5 * np.abs(ref).max())


def test_eer_min_dcf_match_the_reference_formulas():
    rng = np.random.default_rng(5)
    n = 200_000
    labels = (rng.random(n) < 0.1).astype(np.int64)
    scores = (rng.standard_normal(n) * 0.2 + np.where(labels == 1, 0.55, 0.1)).astype(np.float32)
    fnr, fpr = metrics_oracle.pmiss_pfa(scores, labels)
    eer_ref, thr_ref = metrics_oracle.eer(fnr, fpr, scores)
    dcf_ref = metrics_oracle.c_norm(fnr, fpr, 0.01)
    eer, thr, dcf = b200spk.eer_min_dcf(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), p_target=0.01)
    assert abs(eer - eer_ref) < 1e-9 and abs(thr - thr_ref) < 1e-7 and abs(dcf - dcf_ref) < 1e-9


def test_cosine_pairs_then_metrics_end_to_end():
    # embeddings of 40 'speakers' x 25 utterances: same-speaker trials must score high, EER near zero
    rng = np.random.default_rng(6)
    centers = rng.standard_normal((40, 192))
    spk = np.repeat(np.arange(40), 25)
    E = (centers[spk] + 0.5 * rng.standard_normal((1000, 192))).astype(np.float32)
    a = rng.integers(0, 1000, 50_000).astype(np.int32)
    b = rng.integers(0, 1000, 50_000).astype(np.int32)
    sc = b200spk.cosine_pairs(torch.from_numpy(E).cuda(), torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    ref = np.einsum("ij,ij->i", E[a], E[b]) / (np.linalg.norm(E[a], axis=1) * np.linalg.norm(E[b], axis=1))
    assert np.abs(sc.cpu().numpy() - ref).max() < 1e-5
    eer, thr, dcf = b200spk.eer_min_dcf(sc, torch.from_numpy((spk[a] == spk[b]).astype(np.int64)).cuda())
    assert eer < 0.01 and 0.0 < thr < 1.0
