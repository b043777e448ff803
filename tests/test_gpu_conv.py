"""GPU: op-level parity of the conv kernels (CUDA-core fp32 and tcgen05 bf16) through the C ABI
against torch.nn.functional on the CPU.  A one-op program is fed through the same
spk_model_* entry points the networks use."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F
import torch.nn.functional as F_

from b200spk import _lib
from b200spk.program import Model, Program, conv_out

pytestmark = pytest.mark.gpu


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def run_conv(x, w, *, stride=(1, 1), pad=(0, 0), dil=(1, 1), pro=None, epi=None, act=_lib.ACT_NONE,
             residual=False, gate=None, precision="bf16", out_choff=0, out_extra=0, chunk=None):
    """x [B,H,W,Cin] f32 (channels-last), w [Cout,KH,KW,Cin].  Returns (y [B,Ho,Wo,Cout], ref)."""
    B, H, W, Cin = x.shape
    Cout, KH, KW, _ = w.shape
    Ho, Wo = conv_out(H, KH, stride[0], pad[0], dil[0]), conv_out(W, KW, stride[1], pad[1], dil[1])
    bf = precision == "bf16"
    AD = _lib.DT_BF16 if bf else _lib.DT_F32
    model = Model(_lib.PREC_BF16 if bf else _lib.PREC_F32, "cuda:0")
    Ctot = Cin + (Cout if residual else 0)
    ld_out = out_choff + Cout + out_extra
    prog = Program(H * W * Ctot, Ho * Wo * ld_out)
    xin = prog.buf("x", H * W * Ctot, AD)
    # cast op: identity 1x1 conv on CUDA cores turns the f32 input into the activation dtype
    eye = model.param(torch.eye(Ctot).reshape(Ctot, 1, 1, Ctot))
    prog.op(_lib.OP_CONV, in_buf=0, in_ld=Ctot, out_buf=xin, out_ld=Ctot, H=H, W=W, Cin=Ctot, Ho=H, Wo=W, Cout=Ctot, w=eye)
    # the conv under test writes the activation dtype (as inside a network) unless it targets a
    # channel slice of a wider buffer; a second identity conv widens the result to the f32 output
    staged = out_choff == 0 and out_extra == 0 and Cout % 16 == 0
    ybuf = prog.buf("y", Ho * Wo * ld_out, AD) if staged else 1
    kw = dict(in_buf=xin, in_ld=Ctot, out_buf=ybuf, out_ld=ld_out, out_choff=out_choff, H=H, W=W, Cin=Cin, Ho=Ho, Wo=Wo,
              Cout=Cout, KH=KH, KW=KW, sh=stride[0], sw=stride[1], ph=pad[0], pw=pad[1], dh=dil[0], dw=dil[1],
              w=model.param(w), act=act)
    if pro is not None:
        kw.update(pro_scale=model.param(pro[0]), pro_shift=model.param(pro[1]), pro_relu=1)
    if epi is not None:
        kw.update(epi_scale=model.param(epi[0]), epi_shift=model.param(epi[1]))
    if residual:
        kw.update(res_buf=xin, res_ld=Ctot, res_choff=Cin)
    gbuf = None
    if gate is not None:
        w1, b1, w2, b2, seg = gate
        nwin = math.ceil(W / seg)
        gbuf = prog.buf("gate", nwin * Cout, _lib.DT_F32)
        prog.op(_lib.OP_CAM_GATE, in_buf=xin, in_ld=Ctot, out_buf=gbuf, W=W, Cin=Cin, Cout=Cout,
                aux=[model.param(w1), model.param(b1), model.param(w2), model.param(b2)], iaux=[w1.shape[0], seg])
        kw.update(gate_buf=gbuf, gate_win=seg)
    prog.op(_lib.OP_CONV, **kw)
    if staged:
        eye_o = model.param(torch.eye(Cout).reshape(Cout, 1, 1, Cout))
        prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=Cout, out_buf=1, out_ld=Cout, H=Ho, W=Wo, Cin=Cout, Ho=Ho, Wo=Wo,
                Cout=Cout, w=eye_o)
    T = 1
    model.set_program(T, prog)
    res = torch.randn(B, H, W, Cout) if residual else None
    xfull = torch.cat([x, res], dim=-1) if residual else x
    out = model.forward(T, xfull.reshape(B, -1).cuda().contiguous(), Ho * Wo * ld_out, chunk or B)
    torch.cuda.synchronize()
    y = out.cpu().view(B, Ho, Wo, ld_out)[..., out_choff:out_choff + Cout]
    # reference on the CPU with the operand rounding the kernel applies
    xr = _bf16_round(x) if bf else x
    wr = _bf16_round(w) if bf else w
    a = xr
    if pro is not None:
        if bf:      # the tensor-core path runs the prologue in packed bf16 math: scale/shift are bf16
            a = _bf16_round(torch.relu(a * _bf16_round(pro[0]) + _bf16_round(pro[1])))
        else:
            a = torch.relu(a * pro[0] + pro[1])
    ref = F.conv2d(a.permute(0, 3, 1, 2).double(), wr.permute(0, 3, 1, 2).double(), stride=stride, padding=pad,
                   dilation=dil).permute(0, 2, 3, 1)
    if epi is not None:
        ref = ref * epi[0].double() + epi[1].double()
    if residual:
        ref = ref + (_bf16_round(res) if bf else res).double()
    if act == _lib.ACT_RELU:
        ref = torch.relu(ref)
    elif act == _lib.ACT_CLAMP20:
        ref = ref.clamp(0, 20)
    if gate is not None:
        xs = xr.double()                                              # [B,1,W,C]
        tot = xs.mean(dim=2, keepdim=True)
        g = torch.zeros(B, Ho, Wo, Cout, dtype=torch.float64)
        for wi in range(math.ceil(W / seg)):
            a0, a1 = wi * seg, min(W, (wi + 1) * seg)
            ctx = (tot + xs[:, :, a0:a1].mean(dim=2, keepdim=True)).squeeze(2).squeeze(1)     # [B,C]
            h = torch.relu(ctx @ w1.double().t() + b1.double())
            g[:, :, a0:a1, :] = torch.sigmoid(h @ w2.double().t() + b2.double())[:, None, None, :]
        ref = ref * g
    model.close()
    y = y.double()
    y.bf16_stored = bool(staged and bf)
    if y.bf16_stored:
        ref = _bf16_round(ref.float()).double()       # the conv stores bf16
    return y, ref


def _check(y, ref, tol):
    # a bf16-stored result may land one bf16 ulp away from the rounded reference when the fp32
    # accumulation order moves a value across a rounding boundary: allow one bf16 ulp (2^-7 relative) per element
    err = (y - ref).abs()
    scale = ref.abs().max().item() + 1e-6
    ok = err <= tol * scale + (2.0 ** -7 * ref.abs() if y.bf16_stored else 0.0)
    assert bool(ok.all()), (err.max().item(), scale)


CASES = [
    # name, B,H,W,Cin, Cout,KH,KW, stride, pad, dil
    ("1x1_k128", 3, 1, 74, 128, 128, 1, 1, (1, 1), (0, 0), (1, 1)),
    ("1x1_k992_growing_concat", 5, 1, 74, 992, 128, 1, 1, (1, 1), (0, 0), (1, 1)),
    ("1x1_n256", 4, 1, 74, 512, 256, 1, 1, (1, 1), (0, 0), (1, 1)),
    ("1x1_n512_two_ntiles", 4, 1, 74, 1024, 512, 1, 1, (1, 1), (0, 0), (1, 1)),
    ("k3_dil2_cam_local", 5, 1, 74, 128, 32, 1, 3, (1, 1), (0, 2), (1, 2)),
    ("3x3_fcm", 2, 40, 37, 32, 32, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("3x3_fcm_stride2h", 2, 40, 37, 32, 32, 3, 3, (2, 1), (1, 1), (1, 1)),
    ("1x1_shortcut_stride2h", 2, 40, 37, 32, 32, 1, 1, (2, 1), (0, 0), (1, 1)),
    ("tdnn_kh10_kw5_s2", 3, 10, 148, 32, 128, 10, 5, (1, 2), (0, 2), (1, 1)),
    ("ragged_m_not_multiple_of_128", 1, 1, 131, 64, 64, 1, 1, (1, 1), (0, 0), (1, 1)),
    ("k_not_multiple_of_64", 2, 1, 150, 160, 48, 1, 1, (1, 1), (0, 0), (1, 1)),
    # conv_slab4: Res2Net branch convs of ERes2NetV2 (3 s segments: column parts; 128-byte pixels; N = 48)
    ("3x3_c64_w149_parts", 2, 13, 149, 64, 64, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("3x3_c48_w149_parts", 2, 9, 149, 48, 48, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("3x3_c32_w298_parts", 2, 11, 298, 32, 32, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("3x3_c64_w38", 3, 10, 38, 64, 64, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("3x3_c16_w75", 2, 7, 75, 16, 16, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("3x3_c32_w1000_parts", 1, 5, 1000, 32, 32, 3, 3, (1, 1), (1, 1), (1, 1)),
    # TMA im2col path of conv_gemm: wide 3x3 convs (partial last 64-channel chunk), strided 1x1 / 3x3, tiles crossing images
    ("im2col_3x3_c112", 3, 20, 75, 112, 112, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("im2col_3x3_c208", 2, 10, 38, 208, 208, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("im2col_1x1_stride2", 3, 40, 149, 128, 256, 1, 1, (2, 2), (0, 0), (1, 1)),
    ("im2col_3x3_stride2", 2, 20, 75, 64, 128, 3, 3, (2, 2), (1, 1), (1, 1)),
    ("im2col_small_images", 40, 3, 5, 64, 64, 3, 3, (1, 1), (1, 1), (1, 1)),
    ("im2col_k3_dil3_1d", 4, 1, 200, 96, 96, 1, 3, (1, 1), (0, 3), (1, 3)),
]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_conv_plain(case, precision):
    name, B, H, W, Cin, Cout, KH, KW, stride, pad, dil = case
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, KH, KW, Cin, generator=g) / math.sqrt(KH * KW * Cin)
    y, ref = run_conv(x, w, stride=stride, pad=pad, dil=dil, precision=precision)
    _check(y, ref, 2e-5)      # same rounded operands, fp32 accumulate


@pytest.mark.parametrize("K", [128, 992, 4096])
def test_fp32_mode_tensor_core_sum_error_does_not_grow_with_k(K):
    """conv_f32x3.cu: 3xTF32 split products with the chunk partial sums drained into fp32 registers.  The error against
    float64 is flat in K (1.4e-7 .. 2.5e-7 rel-L2 measured) and carries no shrink towards zero: sums left in the TMEM
    accumulator, which truncates, measured -5.9e-6 mean signed relative error at K = 992 and -2.5e-5 at K = 4096."""
    g = torch.Generator().manual_seed(K)
    x = torch.randn(4, 1, 74, K, generator=g)
    w = torch.randn(128, 1, 1, K, generator=g) / math.sqrt(K)
    y, ref = run_conv(x, w, precision="fp32")
    rel = ((y - ref).norm() / ref.norm()).item()
    signed = (((y - ref) * ref.sign()).sum() / ref.abs().sum()).item()
    assert rel <= 5e-7, rel
    assert abs(signed) <= 2e-7, signed


@pytest.mark.parametrize("Cin,Cout,B", [(352, 128, 40), (96, 32, 24), (1024, 128, 8), (64, 256, 12), (160, 48, 3)])
def test_fp32_mode_gemm_prologue_many_tiles(Cin, Cout, B):
    """The fp32-mode GEMM over many M tiles per CTA (ring and accumulator-ring wrap-around), fp32 BN-ReLU prologue on the
    landed tile, partial last K stage, two N tiles (Cout = 256) and a Cout that is not a tile width."""
    g = torch.Generator().manual_seed(Cin + Cout)
    W = 74
    x = torch.randn(B, 1, W, Cin, generator=g)
    w = torch.randn(Cout, 1, 1, Cin, generator=g) / math.sqrt(Cin)
    pro = (torch.rand(Cin, generator=g) + 0.5, 0.1 * torch.randn(Cin, generator=g))
    epi = (torch.rand(Cout, generator=g) + 0.5, 0.1 * torch.randn(Cout, generator=g))
    y, ref = run_conv(x, w, pro=pro, epi=epi, act=_lib.ACT_RELU, precision="fp32")
    _check(y, ref, 2e-6)


def test_fp32_mode_runs_on_the_tensor_cores():
    """SPK_TRACE_DISPATCH=1 names the kernel family of every conv op: in the fp32 precision mode all convs of CAM++
    except the per-segment dense layer go through conv_f32x3 (none through the CUDA-core implicit GEMM)."""
    import os
    import subprocess
    import sys
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    env = dict(os.environ, SPK_TRACE_DISPATCH="1", SPK_GRAPH_MAX_BATCH="0")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "run_forward.py"), "--segments", "4", "--precision", "fp32", "--iters", "1"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    fams = [ln.split()[2] for ln in r.stderr.splitlines() if ln.startswith("[spk dispatch]")]
    assert fams.count("f32x3") >= 100, (fams.count("f32x3"), sorted(set(fams)))
    assert fams.count("simt") <= 2, fams.count("simt")          # the dense layer (routed to the split-K linear kernel)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_prologue_epilogue_relu(precision):
    g = torch.Generator().manual_seed(1)
    B, W, Cin, Cout = 4, 74, 352, 128
    x = torch.randn(B, 1, W, Cin, generator=g)
    w = torch.randn(Cout, 1, 1, Cin, generator=g) / math.sqrt(Cin)
    pro = (torch.rand(Cin, generator=g) + 0.5, 0.1 * torch.randn(Cin, generator=g))
    epi = (torch.rand(Cout, generator=g) + 0.5, 0.1 * torch.randn(Cout, generator=g))
    y, ref = run_conv(x, w, pro=pro, epi=epi, act=_lib.ACT_RELU, precision=precision)
    _check(y, ref, 1e-4 if precision == "fp32" else 2e-3)       # bf16: single- vs double-rounded fma moves a few A elements by one ulp


@pytest.mark.parametrize("Cin,Cout,B", [(352, 128, 40), (160, 64, 24), (96, 32, 24), (1024, 128, 8), (64, 256, 12)])
def test_gemm_prologue_many_tiles(Cin, Cout, B):
    """The TMA GEMM with the BN-ReLU prologue over many M tiles per CTA ring wrap-around: the prologue goes through
    tensor memory for Cout <= 128 (two transform groups on alternate stages, partial last K stage) and stays in
    shared memory for Cout = 256."""
    g = torch.Generator().manual_seed(Cin + Cout)
    W = 74
    x = torch.randn(B, 1, W, Cin, generator=g)
    w = torch.randn(Cout, 1, 1, Cin, generator=g) / math.sqrt(Cin)
    pro = (torch.rand(Cin, generator=g) + 0.5, 0.1 * torch.randn(Cin, generator=g))
    epi = (torch.rand(Cout, generator=g) + 0.5, 0.1 * torch.randn(Cout, generator=g))
    y, ref = run_conv(x, w, pro=pro, epi=epi, act=_lib.ACT_RELU, precision="bf16")
    _check(y, ref, 2e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_padding_stays_zero_after_prologue(precision):
    # padded taps must contribute 0, not relu(shift): the prologue applies to real pixels only
    g = torch.Generator().manual_seed(2)
    x = torch.randn(5, 1, 40, 64, generator=g)        # M = 200 >= 128 so bf16 mode takes the tcgen05 path
    w = torch.randn(32, 1, 3, 64, generator=g) / math.sqrt(192)
    pro = (torch.ones(64), torch.full((64,), 0.7))
    y, ref = run_conv(x, w, pad=(0, 1), pro=pro, precision=precision)
    _check(y, ref, 1e-4 if precision == "fp32" else 2e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_residual_relu(precision):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 20, 37, 32, generator=g)
    w = torch.randn(32, 3, 3, 32, generator=g) / math.sqrt(288)
    epi = (torch.rand(32, generator=g) + 0.5, 0.1 * torch.randn(32, generator=g))
    torch.manual_seed(33)
    y, ref = run_conv(x, w, pad=(1, 1), epi=epi, residual=True, act=_lib.ACT_RELU, precision=precision)
    _check(y, ref, 1e-4 if precision == "fp32" else 1e-4)


def test_slab4_residual_c64_parts():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 9, 149, 64, generator=g)
    w = torch.randn(64, 3, 3, 64, generator=g) / math.sqrt(576)
    epi = (torch.rand(64, generator=g) + 0.5, 0.1 * torch.randn(64, generator=g))
    torch.manual_seed(34)
    y, ref = run_conv(x, w, pad=(1, 1), epi=epi, residual=True, act=_lib.ACT_CLAMP20, precision="bf16", chunk=2)
    _check(y, ref, 1e-4)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("Cin,Cout,K,dil,T", [(128, 128, 3, 3, 200), (128, 128, 3, 4, 998), (80, 256, 5, 1, 148), (128, 128, 3, 2, 131)])
def test_reflect_padded_conv1d_im2col_plus_edge_fix(Cin, Cout, K, dil, T, precision):
    """ECAPA TDNNBlock (ECAPA_TDNN.py:49-77): Conv1d with reflect padding -> ReLU -> BN.  The tensor-core paths (bf16
    conv_gemm, fp32-mode conv_f32x3) run the zero-padded conv with TMA im2col loads and recompute the positions next to
    the segment ends with mirrored taps."""
    g = torch.Generator().manual_seed(9)
    B, pad = 5, dil * (K - 1) // 2
    bf = precision == "bf16"
    AD = _lib.DT_BF16 if bf else _lib.DT_F32
    x = torch.randn(B, 1, T, Cin, generator=g)
    w = torch.randn(Cout, 1, K, Cin, generator=g) / math.sqrt(K * Cin)
    ps, pb = torch.rand(Cout, generator=g) + 0.5, 0.1 * torch.randn(Cout, generator=g)
    model = Model(_lib.PREC_BF16 if bf else _lib.PREC_F32, "cuda:0")
    prog = Program(T * Cin, T * Cout)
    xin = prog.buf("x", T * Cin, AD)
    ybuf = prog.buf("y", T * Cout, AD)
    prog.op(_lib.OP_CONV, in_buf=0, in_ld=Cin, out_buf=xin, out_ld=Cin, H=1, W=T, Cin=Cin, Ho=1, Wo=T, Cout=Cin,
            w=model.param(torch.eye(Cin).reshape(Cin, 1, 1, Cin)))
    prog.op(_lib.OP_CONV, in_buf=xin, in_ld=Cin, out_buf=ybuf, out_ld=Cout, H=1, W=T, Cin=Cin, Ho=1, Wo=T, Cout=Cout, KH=1, KW=K,
            pw=pad, dw=dil, w=model.param(w), act=_lib.ACT_RELU, aux=[model.param(ps), model.param(pb)], iaux=[_lib.ACT_NONE, 1, 0])
    prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=Cout, out_buf=1, out_ld=Cout, H=1, W=T, Cin=Cout, Ho=1, Wo=T, Cout=Cout,
            w=model.param(torch.eye(Cout).reshape(Cout, 1, 1, Cout)))
    model.set_program(1, prog)
    y = model.forward(1, x.reshape(B, -1).cuda().contiguous(), T * Cout, B).cpu().view(B, T, Cout).double()
    model.close()
    rnd = _bf16_round if bf else (lambda t: t)
    xr = rnd(x)[:, 0].permute(0, 2, 1).double()                      # [B, Cin, T]
    wr = rnd(w)[:, 0].permute(0, 2, 1).double()                      # [Cout, Cin, K]
    ref = F.conv1d(F.pad(xr, (pad, pad), mode="reflect"), wr, dilation=dil).permute(0, 2, 1)
    ref = torch.relu(ref) * ps.double() + pb.double()
    err = (y - ref if not bf else y - _bf16_round(ref.float()).double()).abs()
    if bf:
        ref = _bf16_round(ref.float()).double()
        assert bool((err <= 1e-4 * ref.abs().max() + 2.0 ** -7 * ref.abs()).all()), (err.max().item(), err.argmax().item())
    else:
        assert bool((err <= 2e-6 * ref.abs().max()).all()), (err.max().item(), err.argmax().item())
        edge = err[:, list(range(pad)) + list(range(T - pad, T))]
        assert edge.max().item() <= 2e-6 * ref.abs().max().item()


@pytest.mark.parametrize("T,F,B", [(148, 80, 5), (61, 80, 3), (254, 80, 2), (148, 24, 4)])
def test_stem_block_fused_op(T, F, B):
    """SPK_OP_STEM_BLOCK (conv_stem.cu): stem conv + BN + ReLU (never stored) -> 3x3 stride (2,1) conv + BN + ReLU and the
    1x1 stride (2,1) shortcut + BN, vs torch in float64 with the kernel's roundings (stem and outputs stored as bf16,
    bf16 conv weights).  The stem itself runs as a split-bf16 tensor-core GEMM: fp32-accurate, so a stem value may land on
    the other side of a bf16 rounding boundary - hence rel-L2 and a few-ulp bound instead of an exact match."""
    g = torch.Generator().manual_seed(11)
    C = 32
    feats = 3.0 * torch.randn(B, T, F, generator=g)
    w0 = torch.randn(C, 3, 3, generator=g) / 3.0
    w1 = torch.randn(C, 3, 3, C, generator=g) / math.sqrt(9 * C)
    ws = torch.randn(C, 1, 1, C, generator=g) / math.sqrt(C)
    bn = [(torch.rand(C, generator=g) + 0.5, 0.2 * torch.randn(C, generator=g)) for _ in range(3)]
    model = Model(_lib.PREC_BF16, "cuda:0")
    Ho = F // 2
    prog = Program(T * F, Ho * T * 2 * C)
    y1 = prog.buf("y1", Ho * T * C, _lib.DT_BF16)
    y2 = prog.buf("y2", Ho * T * C, _lib.DT_BF16)
    prog.op(_lib.OP_STEM_BLOCK, in_buf=0, out_buf=y1, out_ld=C, res_buf=y2, res_ld=C, H=F, W=T, Ho=Ho, Wo=T, Cin=1, Cout=C,
            w=model.param(w0.reshape(-1)), epi_scale=model.param(bn[0][0]), epi_shift=model.param(bn[0][1]), act=_lib.ACT_RELU,
            aux=[model.param(w1), model.param(bn[1][0]), model.param(bn[1][1]), model.param(ws)], iaux=[model.param(bn[2][0]), model.param(bn[2][1])])
    eye = model.param(torch.eye(C).reshape(C, 1, 1, C))
    for i, src in enumerate((y1, y2)):     # widen both outputs into the two halves of the f32 output pixel
        prog.op(_lib.OP_CONV, in_buf=src, in_ld=C, out_buf=1, out_ld=2 * C, out_choff=i * C, H=Ho, W=T, Cin=C, Ho=Ho, Wo=T, Cout=C, w=eye)
    model.set_program(1, prog)
    out = model.forward(1, feats.reshape(B, -1).cuda().contiguous(), Ho * T * 2 * C, B).cpu().view(B, Ho, T, 2 * C).double()
    model.close()
    img = feats.permute(0, 2, 1).unsqueeze(1).double()                                        # [B, 1, F, T]
    stem = F_.conv2d(img, w0.unsqueeze(1).double(), padding=1)
    stem = _bf16_round(torch.relu(stem * bn[0][0].view(1, C, 1, 1) + bn[0][1].view(1, C, 1, 1)).float()).double()
    r1 = F_.conv2d(stem, _bf16_round(w1).permute(0, 3, 1, 2).double(), stride=(2, 1), padding=1)
    r1 = _bf16_round(torch.relu(r1 * bn[1][0].view(1, C, 1, 1) + bn[1][1].view(1, C, 1, 1)).float()).double().permute(0, 2, 3, 1)
    r2 = F_.conv2d(stem, _bf16_round(ws).permute(0, 3, 1, 2).double(), stride=(2, 1))
    r2 = _bf16_round((r2 * bn[2][0].view(1, C, 1, 1) + bn[2][1].view(1, C, 1, 1)).float()).double().permute(0, 2, 3, 1)
    for name, got, ref in (("conv1", out[..., :C], r1), ("shortcut", out[..., C:], r2)):
        rel = float((got - ref).norm() / ref.norm())
        assert rel <= 3e-3, (name, rel)
        assert float((got - ref).abs().max()) <= 2.0 ** -6 * float(ref.abs().max()), (name, float((got - ref).abs().max()))


def test_slab4_channel_windows_are_clipped():
    """A 48-channel conv reads a window of an 80-channel buffer and writes a window of a 96-channel buffer (the
    Res2Net `cat` layout, ERes2NetV2.py:77-81): the 64-channel TMA boxes must not use the 16 foreign input channels
    nor touch the neighbouring output channels."""
    g = torch.Generator().manual_seed(6)
    B, H, W, LDI, LDO, C, IN_OFF, OUT_OFF = 2, 12, 149, 80, 96, 48, 16, 48
    x = torch.randn(B, H, W, LDI, generator=g)
    wb = torch.randn(LDO, 1, 1, LDI, generator=g) / math.sqrt(LDI)
    w = torch.randn(C, 3, 3, C, generator=g) / math.sqrt(9 * C)
    model = Model(_lib.PREC_BF16, "cuda:0")
    prog = Program(H * W * LDI, H * W * LDO)
    xin = prog.buf("x", H * W * LDI, _lib.DT_BF16)
    ybuf = prog.buf("y", H * W * LDO, _lib.DT_BF16)
    prog.op(_lib.OP_CONV, in_buf=0, in_ld=LDI, out_buf=xin, out_ld=LDI, H=H, W=W, Cin=LDI, Ho=H, Wo=W, Cout=LDI,
            w=model.param(torch.eye(LDI).reshape(LDI, 1, 1, LDI)))
    prog.op(_lib.OP_CONV, in_buf=xin, in_ld=LDI, out_buf=ybuf, out_ld=LDO, H=H, W=W, Cin=LDI, Ho=H, Wo=W, Cout=LDO, w=model.param(wb))
    prog.op(_lib.OP_CONV, in_buf=xin, in_ld=LDI, in_choff=IN_OFF, out_buf=ybuf, out_ld=LDO, out_choff=OUT_OFF, H=H, W=W, Cin=C,
            Ho=H, Wo=W, Cout=C, KH=3, KW=3, ph=1, pw=1, w=model.param(w), act=_lib.ACT_RELU)
    prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=LDO, out_buf=1, out_ld=LDO, H=H, W=W, Cin=LDO, Ho=H, Wo=W, Cout=LDO,
            w=model.param(torch.eye(LDO).reshape(LDO, 1, 1, LDO)))
    model.set_program(1, prog)
    out = model.forward(1, x.reshape(B, -1).cuda().contiguous(), H * W * LDO, B).cpu().view(B, H, W, LDO).double()
    model.close()
    xr = _bf16_round(x)
    bg = _bf16_round((xr.double() @ _bf16_round(wb).reshape(LDO, LDI).double().t()).float()).double()
    conv = F.conv2d(xr[..., IN_OFF:IN_OFF + C].permute(0, 3, 1, 2).double(), _bf16_round(w).permute(0, 3, 1, 2).double(),
                    padding=1).permute(0, 2, 3, 1)
    ref = bg.clone()
    ref[..., OUT_OFF:OUT_OFF + C] = _bf16_round(torch.relu(conv).float()).double()
    err = (out - ref).abs()
    assert bool((err <= 1e-4 * ref.abs().max() + 2.0 ** -7 * ref.abs()).all()), err.max().item()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_clamp20(precision):
    g = torch.Generator().manual_seed(4)
    x = 8 * torch.randn(2, 1, 140, 64, generator=g)
    w = torch.randn(64, 1, 1, 64, generator=g)
    y, ref = run_conv(x, w, act=_lib.ACT_CLAMP20, precision=precision)
    assert ref.max().item() == 20.0 and ref.min().item() == 0.0
    _check(y, ref, 1e-4)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("W", [74, 149, 230])
def test_conv_cam_gate_and_concat_write(precision, W):
    # dilated k=3 conv x CAM gate, written in place at a channel offset of a wider buffer
    g = torch.Generator().manual_seed(5)
    B, C, G = 3, 128, 32
    x = torch.randn(B, 1, W, C, generator=g)
    w = torch.randn(G, 1, 3, C, generator=g) / math.sqrt(3 * C)
    gate = (torch.randn(64, C, generator=g) / math.sqrt(C), 0.1 * torch.randn(64, generator=g),
            torch.randn(G, 64, generator=g) / 8, 0.1 * torch.randn(G, generator=g), 100)
    y, ref = run_conv(x, w, pad=(0, 2), dil=(1, 2), gate=gate, precision=precision, out_choff=160, out_extra=64)
    _check(y, ref, 1e-4 if precision == "fp32" else 1e-4)


def test_conv_many_tiles_persistent_loop():
    # more tiles than resident CTAs: exercises the stage ring wrap, both TMEM buffers, phase flips
    g = torch.Generator().manual_seed(6)
    x = torch.randn(700, 1, 74, 256, generator=g)
    w = torch.randn(128, 1, 1, 256, generator=g) / 16
    y, ref = run_conv(x, w, precision="bf16")
    _check(y, ref, 2e-5)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("W,B,dil", [(74, 7, 2), (74, 3, 1), (149, 5, 2), (230, 2, 2), (61, 9, 2)])
def test_cam_local_fused_op(precision, W, B, dil):
    """The whole CAMLayer as one op (fused tcgen05 kernel in bf16 mode, gate + gated conv in fp32):
    ragged last item (B not a multiple of the segments per CTA), 1-3 context windows, both dilations."""
    g = torch.Generator().manual_seed(W * 10 + B)
    C, G, hidden, seg = 128, 32, 64, 100
    bf = precision == "bf16"
    AD = _lib.DT_BF16 if bf else _lib.DT_F32
    x = torch.randn(B, 1, W, C, generator=g)
    w = torch.randn(G, 1, 3, C, generator=g) / math.sqrt(3 * C)
    w1, b1 = torch.randn(hidden, C, generator=g) / math.sqrt(C), 0.1 * torch.randn(hidden, generator=g)
    w2, b2 = torch.randn(G, hidden, generator=g) / 8, 0.1 * torch.randn(G, generator=g)
    model = Model(_lib.PREC_BF16 if bf else _lib.PREC_F32, "cuda:0")
    nwin = math.ceil(W / seg)
    prog = Program(W * C, W * G)
    xin = prog.buf("x", W * C, AD)
    ybuf = prog.buf("y", W * G, AD)
    gbuf = prog.buf("gate", nwin * G, _lib.DT_F32)
    prog.op(_lib.OP_CONV, in_buf=0, in_ld=C, out_buf=xin, out_ld=C, H=1, W=W, Cin=C, Ho=1, Wo=W, Cout=C,
            w=model.param(torch.eye(C).reshape(C, 1, 1, C)))
    prog.op(_lib.OP_CAM_LOCAL, in_buf=xin, in_ld=C, out_buf=ybuf, out_ld=G, H=1, W=W, Cin=C, Ho=1, Wo=W, Cout=G, KH=1, KW=3,
            pw=dil, dw=dil, w=model.param(w), gate_buf=gbuf, gate_win=seg,
            aux=[model.param(w1), model.param(b1), model.param(w2), model.param(b2)],
            iaux=[hidden, seg, model.param(w1.t().contiguous()), model.param(w2.t().contiguous())])
    prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=G, out_buf=1, out_ld=G, H=1, W=W, Cin=G, Ho=1, Wo=W, Cout=G,
            w=model.param(torch.eye(G).reshape(G, 1, 1, G)))
    model.set_program(1, prog)
    out = model.forward(1, x.reshape(B, -1).cuda().contiguous(), W * G, B)
    torch.cuda.synchronize()
    y = out.cpu().view(B, 1, W, G).double()
    model.close()
    xr = _bf16_round(x) if bf else x
    wr = _bf16_round(w) if bf else w
    ref = F.conv2d(xr.permute(0, 3, 1, 2).double(), wr.permute(0, 3, 1, 2).double(), padding=(0, dil), dilation=(1, dil)).permute(0, 2, 3, 1)
    xs = xr.double()
    tot = xs.mean(dim=2, keepdim=True)
    gate = torch.zeros(B, 1, W, G, dtype=torch.float64)
    for wi in range(nwin):
        a0, a1 = wi * seg, min(W, (wi + 1) * seg)
        ctx = (tot + xs[:, :, a0:a1].mean(dim=2, keepdim=True)).squeeze(2).squeeze(1)
        h = torch.relu(ctx @ w1.double().t() + b1.double())
        gate[:, :, a0:a1, :] = torch.sigmoid(h @ w2.double().t() + b2.double())[:, None, None, :]
    ref = ref * gate
    y.bf16_stored = bf
    if bf:
        ref = _bf16_round(ref.float()).double()
    _check(y, ref, 1e-4)
