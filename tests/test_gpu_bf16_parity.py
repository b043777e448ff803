"""GPU: the bf16 (tcgen05) mode at the north-star tolerance and at the shapes BASELINE.json quotes.

North star: "bf16 embeddings at cosine >= 0.999 versus the reference's own PyTorch CPU path on the same synthetic
inputs and random-init weights".  Random-init = the models' own initialisation (the reference ships no checkpoint
that can be fetched here), so that is what the literal-tolerance test uses, for every network family at its headline
shape; SURVEY 7-6 adds the stronger per-buffer check of bf16 against the fp32 path on the GPU.  The weight sets with
randomised BatchNorm statistics used elsewhere in tests/ are deliberately ill-conditioned stress cases (PyTorch's own
CPU bf16 run of the reference graph loses 1e-2 of cosine on them); their per-block error growth is bounded here so a
regression in any single layer still shows.
"""
import os

import numpy as np
import pytest
import torch

import b200spk
from oracle import campplus_oracle, ecapa_oracle, eres2netv2_oracle, gen_golden, synth

pytestmark = pytest.mark.gpu

COS_BF16 = 0.999          # north-star, literal
ECAPA_CH = [1024, 1024, 1024, 1024, 3072]


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _cos_min(a, b):
    return float(((a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))).min())


def _families():
    """(id, constructor(precision), oracle(sd, feats), seconds, segments)"""
    return [
        ("campplus_e192_t148", lambda p: b200spk.CAMPPlus(embedding_size=192, precision=p),
         lambda sd, f: campplus_oracle.forward(sd, f), 1.5, 8),
        ("campplus_e512_t148", lambda p: b200spk.CAMPPlus(embedding_size=512, precision=p),
         lambda sd, f: campplus_oracle.forward(sd, f), 1.5, 8),
        ("eres2netv2_w26s2e2_t298", lambda p: b200spk.ERes2NetV2(precision=p),
         lambda sd, f: eres2netv2_oracle.forward(sd, f, scale=2), 3.0, 2),
        ("eres2netv2_w24s4ep4_t298", lambda p: b200spk.ERes2NetV2(baseWidth=24, scale=4, expansion=4, precision=p),
         lambda sd, f: eres2netv2_oracle.forward(sd, f, scale=4), 3.0, 2),
        ("ecapa_c1024_t998", lambda p: b200spk.ECAPA_TDNN(80, lin_neurons=192, channels=ECAPA_CH, precision=p),
         lambda sd, f: ecapa_oracle.forward(sd, f), 10.0, 1),
    ]


@pytest.mark.parametrize("fam", _families(), ids=lambda f: f[0])
def test_bf16_random_init_vs_cpu_reference_path(fam):
    name, ctor, oracle, secs, n = fam
    torch.manual_seed(11)
    f32 = ctor("fp32")
    b16 = ctor("bf16")
    b16.load_state_dict(f32.state_dict())
    sd = {k: v.detach().clone() for k, v in f32.state_dict().items()}
    f32, b16 = f32.cuda().eval(), b16.cuda().eval()
    wavs = gen_golden.campplus_input(n, int(secs * 16000), seed=55)
    feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
    ref = oracle(sd, feats.cpu().numpy()).numpy()
    with torch.no_grad():
        e32, e16 = f32(feats).cpu().numpy(), b16(feats).cpu().numpy()
    assert _rel(e32, ref) <= 1e-4, _rel(e32, ref)                  # the fp32 path IS the reference path
    assert _cos_min(e32, ref) >= 0.9999
    assert _cos_min(e16, ref) >= COS_BF16, _cos_min(e16, ref)
    assert _rel(e16, ref) <= 1.5e-2, _rel(e16, ref)


def _buffers(model, T, B):
    names = model._engine.model.programs[T].names
    return {k: model._engine.model.read_buffer(T, k, B) for k, i in names.items() if i >= 2}


@pytest.mark.parametrize("fam", [_families()[0], _families()[3]], ids=lambda f: f[0])
def test_bf16_per_buffer_vs_fp32_on_gpu_random_init(fam):
    """every workspace buffer of the bf16 run within 2e-2 rel-L2 of the fp32 run (bf16 has 8 mantissa
    bits: 4e-3 per rounding; the bound leaves room for tens of layers, not for a wrong layer)."""
    name, ctor, oracle, secs, n = fam
    torch.manual_seed(12)
    f32 = ctor("fp32")
    b16 = ctor("bf16")
    b16.load_state_dict(f32.state_dict())
    f32, b16 = f32.cuda().eval(), b16.cuda().eval()
    wavs = gen_golden.campplus_input(4, int(secs * 16000), seed=56)
    feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
    with torch.no_grad():
        f32(feats)
        b16(feats)
    T, B = feats.shape[1], feats.shape[0]
    a, b = _buffers(f32, T, B), _buffers(b16, T, B)
    assert set(a) == set(b) and len(a) >= 8
    worst = {}
    for k in a:
        if k == "gate":             # scratch of the unfused CAM path: the fused bf16 kernel never writes it
            continue
        x, y = a[k], b[k]
        if k == "fcm_a":            # the bf16 program fuses the stem into the first block (SPK_OP_STEM_BLOCK): only the part of
            n = x.numel() // 4      # this buffer that layer2[0] re-uses for its output ([B, F/4, T, 32], packed) is written by both
            x, y = x.reshape(-1)[:n], y.reshape(-1)[:n]
        worst[k] = float(((x - y).norm() / x.norm()).item())
    assert max(worst.values()) <= 2e-2, worst


def test_bf16_stress_weights_error_growth_is_bounded():
    """CAM++ with randomised BatchNorm statistics (weight set 101, the ill-conditioned stress case): the error may
    grow through the 52 dense layers, but block by block it stays inside the envelope measured for a correct
    implementation (block1 1e-2, block2 3.4e-2, block3 5.5e-2, embedding 5.4e-2)."""
    def build(prec):
        m = b200spk.CAMPPlus(embedding_size=192, precision=prec)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = synth.fill_state_dict(shapes, 101, randomize_bn=True)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        return m.cuda().eval()
    f32, b16 = build("fp32"), build("bf16")
    feats = b200spk.fbank_batch(torch.from_numpy(gen_golden.campplus_input(16, 24000, seed=5)).cuda())
    with torch.no_grad():
        e32, e16 = f32(feats), b16(feats)
    a, b = _buffers(f32, 148, 16), _buffers(b16, 148, 16)
    def _rel_buf(k):
        x, y = a[k].reshape(-1), b[k].reshape(-1)
        if k == "fcm_a":            # fused stem in the bf16 program: compare the part layer2[0] re-uses for its output
            x, y = x[:x.numel() // 4], y[:y.numel() // 4]
        return float(((x - y).norm() / x.norm()).item())
    rel = {k: _rel_buf(k) for k in a if k != "gate"}
    bound = {"fcm_a": 1e-2, "fcm_out": 1.2e-2, "block1": 2.5e-2, "block2": 7e-2, "block3": 1.1e-1, "final": 1.2e-1, "stats": 1.1e-1}
    for k, v in bound.items():
        assert rel[k] <= v, (k, rel[k])
    assert float(((e32 - e16).norm() / e32.norm()).item()) <= 1.1e-1


@pytest.fixture(scope="module")
def headline(golden_dir):
    return np.load(os.path.join(golden_dir, "headline_shapes.npz"))


def _seeded(model, wseed, gain):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=gain)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return model.cuda().eval()


def test_headline_eres2netv2_w24s4ep4_t298_vs_reference_golden(headline):
    """BASELINE config 3's network at its own segment length, against the embedding minted from the imported
    reference (seeded weights with randomised BN, tests/golden/headline_shapes.npz)."""
    fam, name, kw, batch, n_samples, wseed = gen_golden.headline_cases()[0]
    feats = torch.from_numpy(headline[name + ".feats"]).cuda()
    ref = headline[name + ".emb"]
    got = _seeded(b200spk.ERes2NetV2(precision="fp32", **kw), wseed, gen_golden.ERES_GAIN)(feats).cpu().numpy()
    assert got.shape == ref.shape == (batch, 192)
    assert _rel(got, ref) <= 1e-4, _rel(got, ref)
    assert _cos_min(got, ref) >= 0.9999


def test_headline_ecapa_c1024_t998_vs_reference_golden(headline):
    """BASELINE config 5's network on a 10 s chunk (T = 998, the non-tiled kernel paths), fp32 against the reference
    golden and bf16 at the north-star cosine."""
    fam, name, kw, batch, n_samples, wseed = gen_golden.headline_cases()[1]
    feats = torch.from_numpy(headline[name + ".feats"]).cuda()
    ref = headline[name + ".emb"]
    got = _seeded(b200spk.ECAPA_TDNN(80, lin_neurons=192, precision="fp32", **kw), wseed, gen_golden.ECAPA_GAIN)(feats).cpu().numpy()
    assert got.shape == ref.shape == (batch, 192)
    assert _rel(got, ref) <= 1e-4, _rel(got, ref)
    assert _cos_min(got, ref) >= 0.9999
    b16 = _seeded(b200spk.ECAPA_TDNN(80, lin_neurons=192, precision="bf16", **kw), wseed, gen_golden.ECAPA_GAIN)(feats).cpu().numpy()
    assert _cos_min(b16, ref) >= COS_BF16, _cos_min(b16, ref)
