"""One CAM++ extraction pass for profiling (ncu launch lists / full captures)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import torch
import b200spk
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=2048)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--iters", type=int, default=2)
args = ap.parse_args()
model = b200spk.CAMPPlus(embedding_size=512, precision=args.precision, chunk=args.chunk or None)
tsd, _ = bench.make_weights(model)
model.load_state_dict(tsd)
model = model.cuda().eval()
fb = b200spk.FBank(80, 16000, mean_nor=True)
wav = torch.from_numpy(bench.make_windows(args.segments, seed=1)).cuda()
with torch.no_grad():
    for _ in range(args.iters):
        emb = model(fb.batch(wav))
torch.cuda.synchronize()
print("ok", tuple(emb.shape), float(emb.abs().mean()), "launches", b200spk.lib().spk_launch_count())
