"""GPU: ERes2NetV2 forward through the C ABI vs the oracle and the golden vectors."""
import json
import os

import numpy as np
import pytest
import torch

import b200spk
from oracle import eres2netv2_oracle, gen_golden, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "eres2netv2.npz"))


def _model(kw, wseed, precision="fp32", chunk=None):
    m = b200spk.ERes2NetV2(precision=precision, chunk=chunk, **kw)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=gen_golden.ERES_GAIN)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.cuda().eval(), sd


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _cos_min(a, b):
    return float(((a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))).min())


@pytest.mark.parametrize("case", gen_golden.eres2netv2_cases(), ids=lambda c: c[0])
def test_fp32_vs_golden(gold, case):
    name, kw, batch, n_samples, wseed = case
    model, _ = _model(kw, wseed)
    feats = torch.from_numpy(gold[name + ".feats"]).cuda()
    keep = feats.clone()
    with torch.no_grad():
        got = model(feats).cpu().numpy()
    assert torch.equal(feats, keep)                 # the caller's tensor is left alone
    ref = gold[name + ".emb"]
    assert got.shape == ref.shape
    assert _rel(got, ref) <= 1e-4, _rel(got, ref)
    assert _cos_min(got, ref) >= 0.9999


def test_bf16_vs_oracle_stress_weights():
    """Seeded weights with randomised BN statistics (the stress sets of the fp32 goldens) in bf16 against the fp32 CPU
    oracle.  The default variant holds the north-star cosine even here; the 158-conv w24s4ep4 stack on this weight
    set is ill-conditioned (PyTorch's own CPU bf16 run of the reference graph: 0.9968), so there the bound is on the
    relative error, at the level a correct bf16 implementation produces (6e-2 measured; a wrong layer gives O(1)).
    The north-star tolerance itself is asserted on random-init weights at T=298 in tests/test_gpu_bf16_parity.py."""
    for (name, kw, batch, n_samples, wseed), cos_bar, rel_bar in zip(gen_golden.eres2netv2_cases()[:2], (0.999, 0.99), (2e-2, 9e-2)):
        wavs = gen_golden.campplus_input(6, n_samples, seed=77)
        feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
        model, sd = _model(kw, wseed, precision="bf16")
        ref = eres2netv2_oracle.forward(sd, feats.cpu().numpy(), scale=kw["scale"]).numpy()
        with torch.no_grad():
            got = model(feats).cpu().numpy()
        assert _cos_min(got, ref) >= cos_bar, (name, _cos_min(got, ref))
        assert _rel(got, ref) <= rel_bar, (name, _rel(got, ref))


def test_chunking_is_invisible():
    name, kw, batch, n_samples, wseed = gen_golden.eres2netv2_cases()[0]
    feats = b200spk.fbank_batch(torch.from_numpy(gen_golden.campplus_input(5, n_samples, seed=8)).cuda())
    a, _ = _model(kw, wseed, chunk=2)
    b, _ = _model(kw, wseed, chunk=8)
    with torch.no_grad():
        assert torch.equal(a(feats), b(feats))


def test_bf16_many_tiles_per_cta_matches_single_tile_launches():
    """Same as the CAM++ test of this name: large sub-batches (several tiles per persistent CTA, the TMA-epilogue
    residual GEMMs with their staging-buffer hand-offs) against small ones; segments are independent, so the
    embeddings must be bit-identical."""
    torch.manual_seed(6)
    feats = torch.randn(48, 148, 80, device="cuda")
    big = b200spk.ERes2NetV2(precision="bf16", chunk=(48, 48)).cuda().eval()
    small = b200spk.ERes2NetV2(precision="bf16", chunk=(2, 2)).cuda().eval()
    small.load_state_dict(big.state_dict())
    with torch.no_grad():
        a = big(feats)
        b = small(feats)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
