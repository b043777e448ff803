"""fbank alone on 1024 x 3 s utterances (set A: 0.1 * randn), HBM-resident input.
Prints utterances/s, the HBM roofline fraction (287,360 algorithmic bytes per utterance) and the two-sided parity
numbers against the CPU oracle (fp64 and fp32) on a 64-utterance subsample (NSUB)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np
import torch
import b200spk
from oracle import fbank_oracle

PEAK_GBS = 6460.5
torch.manual_seed(1)
wav = (0.1 * torch.randn(1024, 48000)).cuda()
fb = b200spk.FBank(80, 16000, mean_nor=True)
for _ in range(5):
    out = fb.batch(wav)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
big = torch.empty(40 * 1024 * 1024, dtype=torch.float32, device="cuda")     # 160 MB: flushes L2 between iterations
ms = []
for _ in range(20):
    big.zero_()
    e0.record(); out = fb.batch(wav); e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = float(np.median(ms))
bytes_per_utt = 4 * 48000 + 4 * 298 * 80
gbs = 1024 * bytes_per_utt / ms / 1e6
NSUB = int(os.environ.get('NSUB', '64'))
sub = wav[:NSUB].cpu().numpy()
r64 = fbank_oracle.fbank_batch(sub, dtype=np.float64)
r32 = fbank_oracle.fbank_batch(sub, dtype=np.float32)
got = out[:NSUB].cpu().numpy()
d64, d32 = np.abs(got - r64), np.abs(got - r32)
print(json.dumps({"config": "fbank alone, 1024 x 3 s, set A", "ms": round(ms, 4), "utterances_per_s": round(1024 / ms * 1e3, 1),
                  "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK_GBS, 4),
                  "max_abs_vs_fp64": float(d64.max()), "mean_abs_vs_fp64": float(d64.mean()),
                  "frac_gt_1e-4_vs_fp32": float((d32 > 1e-4).mean()), "ref_fp32_vs_fp64_max": float(np.abs(r32 - r64).max())}))
