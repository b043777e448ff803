"""CPU: the host-side compilers (state_dict -> fused-op list) run against a recording mock of the
device model, and every op is checked for internal consistency (weight sizes, buffer extents,
channel windows, alignment rules of the kernels)."""
import numpy as np
import pytest
import torch

import b200spk
from b200spk import _lib, campplus, ecapa_tdnn, eres2net, eres2netv2
from b200spk.program import conv_out


class MockModel:
    def __init__(self, precision):
        self.precision = precision
        self.act_dtype = _lib.DT_BF16 if precision == _lib.PREC_BF16 else _lib.DT_F32
        self.params = []
        self.programs = {}

    def param(self, t):
        self.params.append(tuple(t.shape) if hasattr(t, "shape") else None)
        self.sizes = getattr(self, "sizes", [])
        self.sizes.append(int(t.numel()))
        return len(self.params) - 1

    def set_program(self, T, prog):
        self.programs[T] = prog

    def close(self):
        pass


def _check_program(model, prog, T, F=80):
    bufs = prog.bufs
    written = {0}
    for i, op in enumerate(prog.ops):
        def extent(buf, ld, choff, C, H, W):
            assert 0 <= buf < len(bufs), (i, buf)
            assert choff + C <= ld, (i, "channel window exceeds the pixel pitch", choff, C, ld)
            assert H * W * ld <= bufs[buf].elems, (i, "buffer too small", buf, H, W, ld, bufs[buf].elems)
        if op.kind in (_lib.OP_CONV, _lib.OP_CAM_LOCAL):
            assert model.sizes[op.w] == op.Cout * op.KH * op.KW * op.Cin, (i, "weight size")
            assert op.Ho == conv_out(op.H, op.KH, op.sh, op.ph, op.dh) and op.Wo == conv_out(op.W, op.KW, op.sw, op.pw, op.dw), i
            extent(op.in_buf, op.in_ld, op.in_choff, op.Cin, op.H, op.W)
            extent(op.out_buf, op.out_ld, op.out_choff, op.Cout, op.Ho, op.Wo)
            assert op.Cin % 16 == 0 and op.in_ld % 8 == 0 and op.in_choff % 8 == 0, (i, "input alignment")
            assert op.out_ld % 8 == 0 and op.out_choff % 8 == 0, (i, "output alignment")
            if op.res_buf >= 0:
                extent(op.res_buf, op.res_ld, op.res_choff, op.Cout, op.Ho, op.Wo)
                assert op.res_buf in written, (i, "residual read before written")
            for pid, n in ((op.pro_scale, op.Cin), (op.pro_shift, op.Cin), (op.epi_scale, op.Cout), (op.epi_shift, op.Cout)):
                if pid >= 0:
                    assert model.sizes[pid] == n, (i, "affine vector length", model.sizes[pid], n)
            assert not (op.in_buf == op.out_buf and not (op.out_choff >= op.in_choff + op.Cin or op.out_choff + op.Cout <= op.in_choff)), \
                (i, "conv writes the channels it reads")
        elif op.kind == _lib.OP_STEM:
            assert model.sizes[op.w] == op.Cout * 9
            extent(op.out_buf, op.out_ld, op.out_choff, op.Cout, op.H, op.W)
        elif op.kind == _lib.OP_STEM_BLOCK:
            # stem fused with the first block's conv1 and shortcut: two outputs at half the frequency resolution
            assert model.precision == _lib.PREC_BF16 and op.in_buf == 0 and op.Cout == 32 and op.Ho == op.H // 2 and op.Wo == op.W
            assert model.sizes[op.w] == 32 * 9 and model.sizes[op.aux[0]] == 32 * 9 * 32 and model.sizes[op.aux[3]] == 32 * 32
            for pid in (op.epi_scale, op.epi_shift, op.aux[1], op.aux[2], op.iaux[0], op.iaux[1]):
                assert model.sizes[pid] == 32
            extent(op.out_buf, op.out_ld, op.out_choff, 32, op.Ho, op.Wo)
            extent(op.res_buf, op.res_ld, op.res_choff, 32, op.Ho, op.Wo)
            assert bufs[op.out_buf].dtype == _lib.DT_BF16 and bufs[op.res_buf].dtype == _lib.DT_BF16
            written.add(op.res_buf)
        elif op.kind == _lib.OP_AFF_BLEND:
            extent(op.in_buf, op.in_ld, op.in_choff, op.Cin, op.H, op.W)
            if op.res_buf >= 0:             # res_buf < 0: plain (dtype-converting) copy
                extent(op.res_buf, op.res_ld, op.res_choff, op.Cin, op.H, op.W)
            extent(op.out_buf, op.out_ld, op.out_choff, op.Cin, op.H, op.W)
            if op.gate_buf >= 0:
                extent(op.gate_buf, op.iaux[0], op.iaux[1], op.Cin, op.H, op.W)
        elif op.kind == _lib.OP_CAM_GATE:
            extent(op.in_buf, op.in_ld, op.in_choff, op.Cin, 1, op.W)
        elif op.kind == _lib.OP_SE_SCALE:
            extent(op.in_buf, op.in_ld, op.in_choff, op.Cin, op.H, op.W)
            extent(op.out_buf, op.out_ld, op.out_choff, op.Cin, op.H, op.W)
            assert op.gate_buf in written and bufs[op.gate_buf].elems >= op.Cin and bufs[op.gate_buf].dtype == _lib.DT_F32
            if op.res_buf >= 0:
                extent(op.res_buf, op.res_ld, op.res_choff, op.Cin, op.H, op.W)
                assert op.res_buf in written
        elif op.kind == _lib.OP_ASP_POOL:
            extent(op.in_buf, op.in_ld, op.in_choff, op.Cin, op.H, op.W)
            extent(op.res_buf, op.res_ld, op.res_choff, op.Cin, op.H, op.W)
            assert op.res_buf in written and bufs[op.out_buf].elems >= 2 * op.Cin and bufs[op.out_buf].dtype == _lib.DT_F32
        elif op.kind == _lib.OP_STATS_POOL:
            extent(op.in_buf, op.in_ld, op.in_choff, op.Cin, op.H, op.W)
            assert bufs[op.out_buf].elems >= 2 * op.H * op.Cin
        assert op.in_buf in written, (i, "input read before written", op.in_buf)
        written.add(op.out_buf)
    assert 1 in written            # the embedding buffer is produced


@pytest.mark.parametrize("prec", [_lib.PREC_F32, _lib.PREC_BF16])
@pytest.mark.parametrize("T", [148, 298, 61])
def test_campplus_program(prec, T):
    mod = b200spk.CAMPPlus(embedding_size=192)
    eng = campplus._Engine(mod, MockModel(prec))
    eng.compile(T)
    prog = eng.model.programs[T]
    _check_program(eng.model, prog, T)
    fused = sum(1 for o in prog.ops if o.kind == _lib.OP_STEM_BLOCK)
    assert fused == (1 if prec == _lib.PREC_BF16 and T <= 254 else 0)        # stem + layer1[0].conv1 + shortcut in one op
    assert sum(1 for o in prog.ops if o.kind == _lib.OP_STEM) == 1 - fused
    assert sum(1 for o in prog.ops if o.kind == _lib.OP_CONV) == 11 - 2 * fused + 1 + 52 + 3 + 1     # FCM (8 + 2 shortcuts + conv2), tdnn, 52 bottlenecks, 3 transit, dense
    assert sum(1 for o in prog.ops if o.kind == _lib.OP_CAM_LOCAL) == 52


@pytest.mark.parametrize("prec", [_lib.PREC_F32, _lib.PREC_BF16])
@pytest.mark.parametrize("kw", [dict(), dict(baseWidth=24, scale=4, expansion=4)], ids=["w26s2e2", "w24s4e4"])
@pytest.mark.parametrize("T", [148, 298])
def test_eres2netv2_program(prec, kw, T):
    mod = b200spk.ERes2NetV2(**kw)
    eng = eres2netv2._Engine(mod, MockModel(prec))
    eng.compile(T)
    _check_program(eng.model, eng.model.programs[T], T)


@pytest.mark.parametrize("prec", [_lib.PREC_F32, _lib.PREC_BF16])
@pytest.mark.parametrize("variant", ["base", "large", "huge"])
@pytest.mark.parametrize("T", [148, 298])
def test_eres2net_v1_program(prec, variant, T):
    mod = {"base": b200spk.ERes2Net, "large": lambda: b200spk.ERes2Net(m_channels=64), "huge": b200spk.ERes2Net_huge}[variant]()
    eng = eres2net._Engine(mod, MockModel(prec))
    eng.compile(T)
    prog = eng.model.programs[T]
    _check_program(eng.model, prog, T)
    # three stride-2 3x3 downsamples without BN / activation and three top-level fusions (ERes2Net.py:179-186)
    ds = [o for o in prog.ops if o.kind == _lib.OP_CONV and o.KH == 3 and o.sh == 2 and o.epi_scale < 0]
    assert len(ds) == 3
    S = mod.scale
    n_blend = sum(1 for o in prog.ops if o.kind == _lib.OP_AFF_BLEND)
    assert n_blend == sum(mod.num_blocks) * (S - 1) + 3


@pytest.mark.parametrize("prec", [_lib.PREC_F32, _lib.PREC_BF16])
@pytest.mark.parametrize("T,C", [(148, 512), (998, 1024), (37, 512)])
def test_ecapa_program(prec, T, C):
    mod = b200spk.ECAPA_TDNN(80, channels=[C, C, C, C, 3 * C])
    model = MockModel(prec)
    eng = ecapa_tdnn._Engine(mod, model)
    eng.compile(T)
    prog = model.programs[T]
    _check_program(model, prog, T)
    kinds = [op.kind for op in prog.ops]
    # 3 SE-Res2Net blocks: one gate + one scaling each; 7 dilated convs per block with reflect padding
    assert kinds.count(_lib.OP_CAM_GATE) == 3 and kinds.count(_lib.OP_SE_SCALE) == 3 and kinds.count(_lib.OP_ASP_POOL) == 1
    dil = [op for op in prog.ops if op.kind == _lib.OP_CONV and op.KW == 3]
    assert len(dil) == 21 and all(op.iaux[1] == 1 and op.pw == op.dw for op in dil)
    assert sorted({op.dw for op in dil}) == [2, 3, 4]
    # every TDNNBlock is conv -> ReLU -> BN: activation ReLU with a post-activation affine
    tdnn = [op for op in prog.ops if op.kind == _lib.OP_CONV and op.aux[0] >= 0]
    assert len(tdnn) == 1 + 3 * (2 + 7) + 1 + 1 and all(op.act == _lib.ACT_RELU for op in tdnn)
    # the SE scaling writes the three block outputs side by side into the MFA input
    offs = sorted(op.out_choff for op in prog.ops if op.kind == _lib.OP_SE_SCALE)
    assert offs == [0, C, 2 * C]
    # bf16 mode casts the features once; fp32 mode reads them in place
    casts = [op for op in prog.ops if op.kind == _lib.OP_AFF_BLEND and op.res_buf < 0 and op.in_buf == 0]
    assert len(casts) == (1 if prec == _lib.PREC_BF16 else 0)

