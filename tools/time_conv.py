"""Micro-benchmark of one conv op (any kernel the dispatcher picks) with the conv_tc2 timeline (debug aid).
env: B H W C (Cin = Cout) KS (kernel size) ST (stride) REPS, SPK_TC2_DBG=1 prints per-stage timestamps of CTA 0."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-speaker_b200"))
import numpy as np
import torch
from b200spk import _lib
from b200spk.program import Model, Program, conv_out

def main():
    E = lambda k, d: int(os.environ.get(k, d))
    B, H, W, C, CO, KS, ST, reps = E("B", "163"), E("H", "20"), E("W", "75"), E("C", "112"), E("CO", "0"), E("KS", "3"), E("ST", "1"), E("REPS", "8")
    CO = CO or C
    pad = KS // 2
    Ho, Wo = conv_out(H, KS, ST, pad), conv_out(W, KS, ST, pad)
    g = torch.Generator().manual_seed(1)
    w = torch.randn(CO, KS, KS, C, generator=g) / math.sqrt(KS * KS * C)
    model = Model(_lib.PREC_BF16, "cuda:0")
    model.graph_max_batch = 0
    res = {}
    for tag, k in (("base", 1), ("conv", reps)):
        prog = Program(H * W * 16, Ho * Wo * CO)
        xin = prog.buf("x", H * W * C, _lib.DT_BF16)
        ybuf = prog.buf("y", Ho * Wo * CO, _lib.DT_BF16)
        wide = torch.zeros(C, 1, 1, 16)
        wide[torch.arange(C), 0, 0, torch.arange(C) % 16] = 1.0
        prog.op(_lib.OP_CONV, in_buf=0, in_ld=16, out_buf=xin, out_ld=C, H=H, W=W, Cin=16, Ho=H, Wo=W, Cout=C, w=model.param(wide))
        for _ in range(k):
            prog.op(_lib.OP_CONV, in_buf=xin, in_ld=C, out_buf=ybuf, out_ld=CO, H=H, W=W, Cin=C, Ho=Ho, Wo=Wo, Cout=CO, KH=KS, KW=KS,
                    sh=ST, sw=ST, ph=pad, pw=pad, w=model.param(w), act=_lib.ACT_RELU)
        prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=CO, out_buf=1, out_ld=CO, H=Ho, W=Wo, Cin=CO, Ho=Ho, Wo=Wo, Cout=CO,
                w=model.param(torch.eye(CO).reshape(CO, 1, 1, CO)))
        T = 1 if tag == "base" else 2
        model.set_program(T, prog)
        xd = torch.randn(B, H * W * 16, generator=g).cuda()
        if os.environ.get("SPK_TC2_DBG") and tag == "conv":
            _lib.lib().spk_debug_tc2_enable(1)
        for _ in range(2):
            model.forward(T, xd, Ho * Wo * CO, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model.forward(T, xd, Ho * Wo * CO, B)
        e1.record()
        torch.cuda.synchronize()
        res[tag] = e0.elapsed_time(e1) / 5
    per = (res["conv"] - res["base"]) / (reps - 1) * 1e3
    fl = 2.0 * B * Ho * Wo * KS * KS * C * CO
    print("B=%d %dx%d C=%d->%d k%d s%d: %.1f us per launch, %.0f TFLOP/s" % (B, H, W, C, CO, KS, ST, per, fl / per / 1e6))
    if os.environ.get("SPK_TC2_DBG"):
        ts = np.zeros(256 * 8, dtype=np.int64)
        _lib.lib().spk_debug_tc2_timeline(ctypes.c_void_p(ts.ctypes.data))
        ts = ts.reshape(256, 8)
        t0 = ts[0, 0]
        print("stage  P.start  P.free  P.done | M.wait  M.full  M.issued")
        for i in range(16, 56):
            print("%4d " % i + " ".join("%8d" % (ts[i, j] - t0) for j in range(6)))

main()
