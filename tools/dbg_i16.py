import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np, torch
import b200spk
rng = np.random.default_rng(84)
pcm = rng.integers(-20000, 20000, size=(5, 24000), dtype=np.int16)
xf = torch.from_numpy(pcm.astype(np.float32) / 32768.0).cuda()
xi = torch.from_numpy(pcm).cuda()
for mn in (False, True):
    a = b200spk.fbank_batch(xi, 80, mn); b = b200spk.fbank_batch(xf, 80, mn); b2 = b200spk.fbank_batch(xf, 80, mn); a2 = b200spk.fbank_batch(xi, 80, mn)
    d = (a - b).abs()
    print("mean_nor", mn, "i16 vs f32 max diff", d.max().item(), "n diff", (d > 0).sum().item(), "f32 rerun equal", torch.equal(b, b2), "i16 rerun equal", torch.equal(a, a2))
    idx = torch.nonzero(d > 0)[:10]
    print(idx.tolist())
    for i in idx[:5].tolist():
        print(i, a[tuple(i)].item(), b[tuple(i)].item())
