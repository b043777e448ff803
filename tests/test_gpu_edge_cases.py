"""GPU: batch sizes and segment lengths off the beaten path, for all three networks: a single segment (the tensor-core
kernels need >= 128 rows, so parts of the bf16 program drop to CUDA cores), odd and short frame counts (ragged last
tiles, odd T' after the stride-2 layers, one partial context window), more rows than one tile, and the empty batch."""
import numpy as np
import pytest
import torch

import b200spk
from oracle import campplus_oracle, ecapa_oracle, eres2netv2_oracle, synth

pytestmark = pytest.mark.gpu

NETS = {
    "campplus": (lambda p: b200spk.CAMPPlus(embedding_size=192, precision=p), lambda sd, f: campplus_oracle.forward(sd, f)),
    "ecapa": (lambda p: b200spk.ECAPA_TDNN(80, channels=[512, 512, 512, 512, 1536], precision=p), lambda sd, f: ecapa_oracle.forward(sd, f)),
    "eres2netv2": (lambda p: b200spk.ERes2NetV2(precision=p), lambda sd, f: eres2netv2_oracle.forward(sd, f, scale=2)),
}


def _pair(make):
    m32, m16 = make("fp32"), make("bf16")
    shapes = {k: tuple(v.shape) for k, v in m32.state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(shapes, 3, randomize_bn=True, gain=1.0).items()}
    m32.load_state_dict(sd)
    m16.load_state_dict(sd)
    return m32.cuda().eval(), m16.cuda().eval(), sd


@pytest.mark.parametrize("net", sorted(NETS))
def test_odd_shapes(net):
    make, oracle = NETS[net]
    m32, m16, sd = _pair(make)
    g = torch.Generator().manual_seed(11)
    for B, T in [(1, 148), (3, 149), (2, 57), (130, 100)]:
        feats = torch.randn(B, T, 80, generator=g).cuda()
        with torch.no_grad():
            e32 = m32(feats).cpu().numpy()
            e16 = m16(feats).cpu().numpy()
        assert e32.shape == (B, 192) and np.isfinite(e32).all() and np.isfinite(e16).all()
        if B <= 3:
            ref = oracle(sd, feats.cpu().numpy()).numpy()
            assert np.linalg.norm(e32 - ref) / np.linalg.norm(ref) <= 1e-4, (net, B, T)
        cos = ((e16 * e32).sum(1) / (np.linalg.norm(e16, axis=1) * np.linalg.norm(e32, axis=1))).min()
        assert cos >= 0.999, (net, B, T, float(cos))
    with torch.no_grad():
        assert tuple(m16(torch.zeros(0, 148, 80, device="cuda")).shape) == (0, 192)
        assert tuple(m32(torch.zeros(0, 148, 80, device="cuda")).shape) == (0, 192)
