"""Per-item role timeline of the fused stem block kernel (CTA 0): debug aid."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np, torch
import b200spk, bench
from b200spk import _lib
model = b200spk.CAMPPlus(embedding_size=512, precision="bf16")
tsd, _ = bench.make_weights(model); model.load_state_dict(tsd); model = model.cuda().eval()
feats = torch.randn(2048, 148, 80, device="cuda")
with torch.no_grad():
    model(feats)
    _lib.lib().spk_debug_stem_enable(1)
    model(feats)
torch.cuda.synchronize()
ts = np.zeros(64 * 8, dtype=np.int64)
_lib.lib().spk_debug_stem_timeline(ctypes.c_void_p(ts.ctypes.data))
ts = ts.reshape(64, 8); t0 = ts[0, 0]
print("item  S.wait  S.free  S.done | M.start M.issued | E.start E.done | store.done")
for i in range(4, 40):
    print("%3d " % i + " ".join("%8d" % (ts[i, j] - t0) for j in range(8)))
