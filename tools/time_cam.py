"""Micro-benchmark of the fused CAM layer op (debug aid): K launches on one staged input."""
import math, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-speaker_b200"))
import torch
from b200spk import _lib
from b200spk.program import Model, Program

def main():
    B, W, C, G, hidden, seg, dil, K = 2048, 74, 128, 32, 64, 100, 2, int(os.environ.get("K", "16"))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, W * C, generator=g)
    w = torch.randn(G, 1, 3, C, generator=g) / math.sqrt(3 * C)
    w1, b1 = torch.randn(hidden, C, generator=g) / math.sqrt(C), 0.1 * torch.randn(hidden, generator=g)
    w2, b2 = torch.randn(G, hidden, generator=g) / 8, 0.1 * torch.randn(G, generator=g)
    model = Model(_lib.PREC_BF16, "cuda:0")
    nwin = math.ceil(W / seg)
    res = {}
    for tag, k in (("base", 0), ("cam", K)):
        prog = Program(W * C, W * G)
        xin = prog.buf("x", W * C, _lib.DT_BF16)
        ybuf = prog.buf("y", W * G, _lib.DT_BF16)
        gbuf = prog.buf("gate", nwin * G, _lib.DT_F32)
        prog.op(_lib.OP_CONV, in_buf=0, in_ld=C, out_buf=xin, out_ld=C, H=1, W=W, Cin=C, Ho=1, Wo=W, Cout=C,
                w=model.param(torch.eye(C).reshape(C, 1, 1, C)))
        for _ in range(max(k, 1)):
            prog.op(_lib.OP_CAM_LOCAL, in_buf=xin, in_ld=C, out_buf=ybuf, out_ld=G, H=1, W=W, Cin=C, Ho=1, Wo=W, Cout=G, KH=1, KW=3,
                    pw=dil, dw=dil, w=model.param(w), gate_buf=gbuf, gate_win=seg,
                    aux=[model.param(w1), model.param(b1), model.param(w2), model.param(b2)],
                    iaux=[hidden, seg, model.param(w1.t().contiguous()), model.param(w2.t().contiguous())])
        prog.op(_lib.OP_CONV, in_buf=ybuf, in_ld=G, out_buf=1, out_ld=G, H=1, W=W, Cin=G, Ho=1, Wo=W, Cout=G,
                w=model.param(torch.eye(G).reshape(G, 1, 1, G)))
        T = 1 if tag == "base" else 2
        model.set_program(T, prog)
        xd = x.cuda()
        for _ in range(3):
            model.forward(T, xd, W * G, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model.forward(T, xd, W * G, B)
        e1.record()
        torch.cuda.synchronize()
        res[tag] = e0.elapsed_time(e1) / 10
    if int(os.environ.get("SPK_CAM_DBG", "0")) & 64:
        import ctypes, numpy as np
        ts = np.zeros(16 * 12, dtype=np.int64)
        _lib.lib().spk_debug_cam_timeline(ctypes.c_void_p(ts.ctypes.data))
        ts = ts.reshape(16, 12)
        t0 = ts[0, 0]
        names = ["P.start", "P.issued", "G.tmemld", "M.start", "M.issued", "G.start", "G.sums", "G.mlp0", "G.done", "E.start", "E.done"]
        print("item " + " ".join("%9s" % n for n in names) + "   (cycles since first producer start)")
        for i in range(6):
            print("%4d " % i + " ".join("%9d" % (ts[i, j] - t0) for j in range(11)))
    print("SPK_CAM_DBG=%s  per cam_local launch: %.1f us" % (os.environ.get("SPK_CAM_DBG", "0"), (res["cam"] - res["base"]) / (K - 1) * 1e3))

main()
