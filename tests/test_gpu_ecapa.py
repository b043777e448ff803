"""GPU: ECAPA-TDNN forward through the C ABI vs the oracle and the golden vectors."""
import os

import numpy as np
import pytest
import torch

import b200spk
from oracle import ecapa_oracle, gen_golden, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ecapa.npz"))


def _model(kw, wseed, precision="fp32", chunk=None):
    m = b200spk.ECAPA_TDNN(80, lin_neurons=192, precision=precision, chunk=chunk, **kw)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=gen_golden.ECAPA_GAIN)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.cuda().eval(), sd


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _cos_min(a, b):
    return float(((a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))).min())


@pytest.mark.parametrize("case", gen_golden.ecapa_cases(), ids=lambda c: c[0])
def test_fp32_vs_golden(gold, case):
    name, kw, batch, n_samples, wseed = case
    model, _ = _model(kw, wseed)
    feats = torch.from_numpy(gold[name + ".feats"]).cuda()
    with torch.no_grad():
        got = model(feats).cpu().numpy()
    ref = gold[name + ".emb"]
    assert got.shape == ref.shape
    assert _rel(got, ref) <= 1e-4, _rel(got, ref)
    assert _cos_min(got, ref) >= 0.9999


def test_fp32_per_buffer_vs_oracle(gold):
    """Block outputs (the slices of the MFA concat buffer) and the MFA output against the oracle's taps."""
    name, kw, batch, n_samples, wseed = gen_golden.ecapa_cases()[0]
    model, sd = _model(kw, wseed)
    feats = gold[name + ".feats"]
    taps = {}
    ecapa_oracle.forward({k: torch.from_numpy(v) for k, v in sd.items()}, feats, taps=taps)
    with torch.no_grad():
        model(torch.from_numpy(feats).cuda())
    eng, T = model._engine, feats.shape[1]
    C, Cm = kw["channels"][1], kw["channels"][-1]
    cat = eng.model.read_buffer(T, "cat", batch).cpu().numpy().reshape(batch, T, Cm)
    for i in (1, 2, 3):
        ref = taps["blocks.%d" % i].numpy().transpose(0, 2, 1)
        assert _rel(cat[:, :, (i - 1) * C:i * C], ref) <= 1e-4, i
    mfa = eng.model.read_buffer(T, "mfa", batch).cpu().numpy().reshape(batch, T, Cm)
    assert _rel(mfa, taps["mfa"].numpy().transpose(0, 2, 1)) <= 1e-4


def test_fp32_sub_batches_are_bit_identical():
    name, kw, batch, n_samples, wseed = gen_golden.ecapa_cases()[0]
    wavs = gen_golden.campplus_input(5, n_samples, seed=91)
    feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
    a, _ = _model(kw, wseed)
    b, _ = _model(kw, wseed, chunk=2)
    with torch.no_grad():
        ea, eb = a(feats), b(feats)
    assert torch.equal(ea, eb)


@pytest.mark.parametrize("case", gen_golden.ecapa_cases()[:2], ids=lambda c: c[0])
def test_bf16_vs_oracle(case):
    """Tensor-core mode against the fp32 CPU oracle on more segments than the goldens hold (M >= 128 rows so the
    tcgen05 kernels, not the CUDA-core fallback, run)."""
    name, kw, batch, n_samples, wseed = case
    wavs = gen_golden.campplus_input(6, n_samples, seed=79)
    feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
    model, sd = _model(kw, wseed, precision="bf16")
    with torch.no_grad():
        got = model(feats).cpu().numpy()
    ref = ecapa_oracle.forward({k: torch.from_numpy(v) for k, v in sd.items()}, feats.cpu().numpy()).numpy()
    assert np.isfinite(got).all()
    assert _cos_min(got, ref) >= 0.999, _cos_min(got, ref)
    assert _rel(got, ref) <= 3e-2, _rel(got, ref)


def test_rejects_lengths_and_training():
    m = b200spk.ECAPA_TDNN(80).cuda()
    x = torch.zeros(1, 148, 80, device="cuda")
    with pytest.raises(AssertionError):
        m(x, lengths=torch.ones(1))
    m.train()
    with pytest.raises(AssertionError):
        m(x)
