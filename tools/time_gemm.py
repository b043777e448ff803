"""Micro-benchmark of the bottleneck 1x1 conv (conv_gemm path, BN-ReLU prologue) with a per-stage timeline (debug aid)."""
import math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-speaker_b200"))
import torch
from b200spk import _lib
from b200spk.program import Model, Program

def main():
    B, W, K, N, reps = int(os.environ.get("B", "2048")), int(os.environ.get("W", "74")), int(os.environ.get("K", "512")), int(os.environ.get("N", "128")), int(os.environ.get("REPS", "8"))
    PRO, RES = int(os.environ.get("PRO", "1")), int(os.environ.get("RES", "0"))
    LD = int(os.environ.get("LD", "0")) or K          # pixel pitch of the input buffer (>= K: a channel window of a wider buffer)
    g = torch.Generator().manual_seed(1)
    w = torch.randn(N, 1, 1, K, generator=g) / math.sqrt(K)
    ps, pb = torch.rand(K, generator=g) + 0.5, 0.1 * torch.randn(K, generator=g)
    model = Model(_lib.PREC_BF16, "cuda:0")
    res = {}
    for tag, k in (("base", 1), ("gemm", reps)):
        prog = Program(W * 32, W * N)
        xin = prog.buf("x", W * LD, _lib.DT_BF16)
        ybuf = prog.buf("y", W * N, _lib.DT_BF16)
        rbuf = prog.buf("r", W * N, _lib.DT_BF16)
        # widen the 32-channel input into a K-channel buffer (values do not matter for timing)
        wide = torch.zeros(K, 1, 1, 32)
        wide[torch.arange(K), 0, 0, torch.arange(K) % 32] = 1.0
        prog.op(_lib.OP_CONV, in_buf=0, in_ld=32, out_buf=xin, out_ld=LD, H=1, W=W, Cin=32, Ho=1, Wo=W, Cout=K, w=model.param(wide))
        for _ in range(k):
            prog.op(_lib.OP_CONV, in_buf=xin, in_ld=LD, out_buf=ybuf, out_ld=N, H=1, W=W, Cin=K, Ho=1, Wo=W, Cout=N,
                    w=model.param(w), act=_lib.ACT_RELU, **(dict(pro_scale=model.param(ps), pro_shift=model.param(pb), pro_relu=1) if PRO else {}),
                    **(dict(res_buf=rbuf, res_ld=N) if RES else {}))
        prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=N, out_buf=1, out_ld=N, H=1, W=W, Cin=N, Ho=1, Wo=W, Cout=N,
                w=model.param(torch.eye(N).reshape(N, 1, 1, N)))
        T = 1 if tag == "base" else 2
        model.set_program(T, prog)
        xd = torch.randn(B, W * 32, generator=g).cuda()
        for _ in range(2):
            model.forward(T, xd, W * N, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model.forward(T, xd, W * N, B)
        e1.record()
        torch.cuda.synchronize()
        res[tag] = e0.elapsed_time(e1) / 5
    per = (res["gemm"] - res["base"]) / (reps - 1) * 1e3
    print("K=%d N=%d pro=%d res=%d: %.1f us per launch, %.0f GB/s in+out, %.0f TFLOP/s" % (K, N, PRO, RES, per, B * W * (K + N + RES * N) * 2 / per / 1e3, 2.0 * B * W * K * N / per / 1e6))
    if os.environ.get("SPK_GEMM_DBG"):
        import ctypes, numpy as np
        ts = np.zeros(256 * 8, dtype=np.int64)
        _lib.lib().spk_debug_gemm_timeline(ctypes.c_void_p(ts.ctypes.data))
        ts = ts.reshape(256, 8)
        t0 = ts[0, 0]
        print("stage  P.issue  X.landed  X.done  M.wait  M.ready  M.issued   | per tile: E.wait E.accf")
        for i in range(8, 48):
            print("%4d " % i + " ".join("%8d" % (ts[i, j] - t0) for j in range(6)) + "   | " + " ".join("%8d" % (ts[i, j] - t0) for j in (6, 7)))

main()
