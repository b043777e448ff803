"""GPU: ERes2Net v1 (base / large / huge) forward through the C ABI vs the golden vectors minted from
the imported reference (speakerlab/models/eres2net/ERes2Net.py, ERes2Net_huge.py) and the oracle."""
import os

import numpy as np
import pytest
import torch

import b200spk
from oracle import eres2net_oracle, gen_golden, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "eres2net.npz"))


def _model(kw, wseed, precision="fp32", chunk=None):
    m = b200spk.ERes2Net(precision=precision, chunk=chunk, **kw)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=gen_golden.ERES_GAIN)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.cuda().eval(), sd


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _cos_min(a, b):
    return float(((a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))).min())


@pytest.mark.parametrize("case", gen_golden.eres2net_cases(), ids=lambda c: c[0])
def test_fp32_vs_golden(gold, case):
    name, variant, kw, batch, n_samples, wseed = case
    model, _ = _model(kw, wseed)
    feats = torch.from_numpy(gold[name + ".feats"]).cuda()
    with torch.no_grad():
        got = model(feats).cpu().numpy()
    ref = gold[name + ".emb"]
    assert got.shape == ref.shape
    assert _rel(got, ref) <= 1e-4, _rel(got, ref)
    assert _cos_min(got, ref) >= 0.9999


def test_fp32_fusion_chain_vs_oracle_taps(gold):
    """The three bottom-up fusions (ERes2Net.py:213-221) buffer by buffer against the oracle."""
    name, variant, kw, batch, n_samples, wseed = gen_golden.eres2net_cases()[0]
    model, sd = _model(kw, wseed)
    feats = gold[name + ".feats"]
    taps = {}
    eres2net_oracle.forward(sd, feats, taps=taps)
    with torch.no_grad():
        model(torch.from_numpy(feats).cuda())
    eng, T = model._engine, feats.shape[1]
    for buf in ("fuse12", "fuse123", "fuse1234"):
        t = taps[buf]                                   # [B, C, H, W]
        b_, c_, h_, w_ = t.shape
        x = eng.model.read_buffer(T, buf, b_).cpu().view(b_, h_, w_, c_).permute(0, 3, 1, 2)
        assert _rel(x.numpy(), t.numpy()) <= 1e-4, buf


@pytest.mark.parametrize("variant", ["base", "huge"])
def test_bf16_random_init_vs_cpu_oracle(variant):
    """bf16 at the north-star cosine on the model's own initialisation (cf. tests/test_gpu_bf16_parity.py)."""
    torch.manual_seed(13)
    ctor = b200spk.ERes2Net if variant == "base" else b200spk.ERes2Net_huge
    f32 = ctor(precision="fp32")
    b16 = ctor(precision="bf16")
    b16.load_state_dict(f32.state_dict())
    sd = {k: v.detach().clone() for k, v in f32.state_dict().items()}
    b16 = b16.cuda().eval()
    wavs = gen_golden.campplus_input(3, 48000, seed=57)
    feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
    ref = eres2net_oracle.forward(sd, feats.cpu().numpy(), scale=f32.scale).numpy()
    with torch.no_grad():
        got = b16(feats).cpu().numpy()
    assert _cos_min(got, ref) >= 0.999, _cos_min(got, ref)


def test_sub_batches_are_bit_identical():
    name, variant, kw, batch, n_samples, wseed = gen_golden.eres2net_cases()[0]
    feats = b200spk.fbank_batch(torch.from_numpy(gen_golden.campplus_input(5, n_samples, seed=8)).cuda())
    a, _ = _model(kw, wseed, chunk=2)
    b, _ = _model(kw, wseed, chunk=8)
    with torch.no_grad():
        assert torch.equal(a(feats), b(feats))
