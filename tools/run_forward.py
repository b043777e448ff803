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
ap.add_argument("--model", default="campplus", choices=["campplus", "eres", "eres_w24", "ecapa"])
ap.add_argument("--seconds", type=float, default=1.5)
args = ap.parse_args()
if args.model == "campplus":
    model = b200spk.CAMPPlus(embedding_size=512, precision=args.precision, chunk=args.chunk or None)
    tsd, _ = bench.make_weights(model)
else:
    from oracle import synth
    if args.model == "eres":
        model = b200spk.ERes2NetV2(precision=args.precision, chunk=args.chunk or None)
    elif args.model == "eres_w24":
        model = b200spk.ERes2NetV2(baseWidth=24, scale=4, expansion=4, precision=args.precision, chunk=args.chunk or None)
    else:
        model = b200spk.ECAPA_TDNN(80, channels=[1024, 1024, 1024, 1024, 3072], precision=args.precision, chunk=args.chunk or None)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    tsd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(shapes, 7, randomize_bn=True, gain=1.0).items()}
model.load_state_dict(tsd)
model = model.cuda().eval()
fb = b200spk.FBank(80, 16000, mean_nor=True)
if args.model == "campplus" and args.seconds == 1.5:
    wav = torch.from_numpy(bench.make_windows(args.segments, seed=1)).cuda()
else:
    wav = 0.1 * torch.randn(args.segments, int(args.seconds * 16000), generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")
with torch.no_grad():
    for _ in range(args.iters):
        emb = model(fb.batch(wav))
torch.cuda.synchronize()
print("ok", tuple(emb.shape), float(emb.abs().mean()), "launches", b200spk.lib().spk_launch_count())
