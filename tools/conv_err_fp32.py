"""Error of one fp32-mode conv against float64, for the tensor-core 3xTF32 path (default) or the CUDA-core path
(SPK_NO_F32X3=1): max error over max |ref| and rel-L2, per test case of tests/test_gpu_conv.py."""
import math
import os
import sys
import zlib

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "3d-speaker_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_conv as tg  # noqa: E402

for case in tg.CASES + [("1x1_k4096", 4, 1, 74, 4096, 128, 1, 1, (1, 1), (0, 0), (1, 1))]:
    name, B, H, W, Cin, Cout, KH, KW, stride, pad, dil = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    x = torch.randn(B, H, W, Cin, generator=g)
    if os.environ.get("POSITIVE"):
        x = x.abs()
    w = torch.randn(Cout, KH, KW, Cin, generator=g) / math.sqrt(KH * KW * Cin)
    if os.environ.get("POSITIVE"):
        w = w.abs()
    y, ref = tg.run_conv(x, w, stride=stride, pad=pad, dil=dil, precision="fp32")
    err = (y - ref).abs()
    print("%-32s K=%5d  max/scale %.2e  rel-L2 %.2e  mean signed rel %.2e" % (name, KH * KW * Cin, err.max().item() / ref.abs().max().item(),
          ((y - ref).norm() / ref.norm()).item(), (((y - ref) * ref.sign()).sum() / ref.abs().sum()).item()))
