"""Micro-benchmark of one FCM conv (conv_slab3 path) with a per-item role timeline (debug aid).
MODE=s1 (3x3 stride 1, default) | s2 (3x3 stride 2) | sc (1x1 stride 2); RES=1 adds the residual read."""
import math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-speaker_b200"))
import torch
from b200spk import _lib
from b200spk.program import Model, Program

def main():
    mode, with_res = os.environ.get("MODE", "s1"), os.environ.get("RES", "0") == "1"
    B, H, W, C, reps = int(os.environ.get("B", "1024")), 40, 148, 32, int(os.environ.get("REPS", "6"))
    KS, S = (1, 2) if mode == "sc" else (3, 2 if mode == "s2" else 1)
    if S == 2:
        H = 80
    Ho = (H + 2 * (KS // 2) - KS) // S + 1
    g = torch.Generator().manual_seed(1)
    w = torch.randn(C, KS, KS, C, generator=g) / math.sqrt(KS * KS * C)
    model = Model(_lib.PREC_BF16, "cuda:0")
    res = {}
    for tag, k in (("base", 1), ("conv", reps)):
        prog = Program(H * W * C, Ho * W * C)
        xin = prog.buf("x", H * W * C, _lib.DT_BF16)
        rbuf = prog.buf("r", Ho * W * C, _lib.DT_BF16)
        ybuf = prog.buf("y", Ho * W * C, _lib.DT_BF16)
        prog.op(_lib.OP_CONV, in_buf=0, in_ld=C, out_buf=xin, out_ld=C, H=H, W=W, Cin=C, Ho=H, Wo=W, Cout=C,
                w=model.param(torch.eye(C).reshape(C, 1, 1, C)))
        for _ in range(k):
            kw = dict(in_buf=xin, in_ld=C, out_buf=ybuf, out_ld=C, H=H, W=W, Cin=C, Ho=Ho, Wo=W, Cout=C, KH=KS, KW=KS, sh=S, sw=1,
                      ph=KS // 2, pw=KS // 2, w=model.param(w), act=_lib.ACT_RELU,
                      epi_scale=model.param(torch.ones(C)), epi_shift=model.param(torch.zeros(C)))
            if with_res:
                kw.update(res_buf=rbuf, res_ld=C, res_choff=0)
            prog.op(_lib.OP_CONV, **kw)
        prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=C, out_buf=1, out_ld=C, H=Ho, W=W, Cin=C, Ho=Ho, Wo=W, Cout=C,
                w=model.param(torch.eye(C).reshape(C, 1, 1, C)))
        T = 1 if tag == "base" else 2
        model.set_program(T, prog)
        xd = torch.randn(B, H * W * C, generator=g).cuda()
        for _ in range(2):
            model.forward(T, xd, Ho * W * C, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model.forward(T, xd, Ho * W * C, B)
        e1.record()
        torch.cuda.synchronize()
        res[tag] = e0.elapsed_time(e1) / 5
    per = (res["conv"] - res["base"]) / (reps - 1) * 1e3
    mb = B * W * C * 2 * (H // (2 if mode == "sc" else 1) + Ho * (2 if with_res else 1)) / 1e6
    print("%s res=%d B=%d: %.1f us per launch, %.0f MB -> %.0f GB/s" % (mode, with_res, B, per, mb, mb / per * 1e3))
    if os.environ.get("SPK_SLAB_DBG"):
        import ctypes, numpy as np
        ts = np.zeros(64 * 8, dtype=np.int64)
        _lib.lib().spk_debug_slab_timeline(ctypes.c_void_p(ts.ctypes.data))
        ts = ts.reshape(64, 8)
        t0 = ts[0, 0]
        print("item  P.issue  M.accfree  M.slabfull  M.issued  E0.start  E0.done  E1.start  E1.done")
        for i in range(12):
            print("%4d " % i + " ".join("%9d" % (ts[i, j] - t0) for j in range(8)))

main()
