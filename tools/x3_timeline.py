"""Role timeline of one fp32-mode conv (conv_f32x3.cu, CTA 0, stages 512..767): SPK_F32X3_DBG=<K of the conv> python tools/x3_timeline.py
env: B H W C CO KS (defaults: the first FCM conv of CAM++ at 256 segments)."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-speaker_b200"))
import numpy as np
import torch
from b200spk import _lib
from b200spk.program import Model, Program, conv_out

E = lambda k, d: int(os.environ.get(k, d))
B, H, W, C, CO, KS = E("B", "256"), E("H", "40"), E("W", "148"), E("C", "32"), E("CO", "32"), E("KS", "3")
pad = KS // 2
g = torch.Generator().manual_seed(1)
w = torch.randn(CO, KS, KS, C, generator=g) / math.sqrt(KS * KS * C)
model = Model(_lib.PREC_F32, "cuda:0")
model.graph_max_batch = 0
prog = Program(H * W * C, H * W * CO)
ybuf = prog.buf("y", H * W * CO, _lib.DT_F32)
for i in range(3):
    prog.op(_lib.OP_CONV, in_buf=0, in_ld=C, out_buf=ybuf if i < 2 else 1, out_ld=CO, H=H, W=W, Cin=C, Ho=H, Wo=W, Cout=CO, KH=KS, KW=KS,
            ph=pad, pw=pad, w=model.param(w), act=_lib.ACT_RELU)
model.set_program(1, prog)
xd = torch.randn(B, H * W * C, generator=g).cuda()
for _ in range(2):
    model.forward(1, xd, H * W * CO, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.forward(1, xd, H * W * CO, B)
e1.record()
torch.cuda.synchronize()
print("3 convs: %.1f us each" % (e0.elapsed_time(e1) / 3 * 1e3))
ts = np.zeros(256 * 8, dtype=np.int64)
_lib.lib().spk_debug_f32x3_timeline(ctypes.c_void_p(ts.ctypes.data))
ts = ts.reshape(256, 8)
t0 = ts[0, 0]
print("stage | P.start P.free | X.landed X.done | M.wait M.ready M.issued | D.done      (cycles from stage 512's P.start)")
for i in range(0, 40):
    print("%4d  " % (512 + i) + " ".join("%8d" % (ts[i, j] - t0) for j in range(8)))
d = np.diff(ts[8:200, 6])
print("MMA issue period: mean %.0f cycles/stage" % d.mean())
