"""Throughput of the other embedding networks of SURVEY section 8(d) (configs 3 and 5) on one GPU: ERes2NetV2 (both
variants, 3 s segments, T=298) and ECAPA-TDNN C=1024 (10 s chunks, T=998), bf16, random-init weights, synthetic
features resident in HBM.  Prints one JSON line per model with TFLOP/s against the minimal FLOP counts of the survey."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np
import torch
import b200spk
from oracle import synth

PEAK = 1423.2


def run(name, model, T, n_seg, gflop_per_seg, batch, reps=3):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = synth.fill_state_dict(shapes, 7, randomize_bn=True, gain=1.0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(3)
    feats = torch.randn(n_seg, T, 80, generator=g, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            for i in range(0, n_seg, batch):
                model(feats[i:i + batch])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for i in range(0, n_seg, batch):
                emb = model(feats[i:i + batch])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = n_seg * gflop_per_seg / ms
    print(json.dumps({"model": name, "T": T, "segments": n_seg, "batch": batch, "ms": round(ms, 2), "segments_per_s": round(n_seg / ms * 1e3, 1),
                      "gflop_per_segment": gflop_per_seg, "tflops": round(tf, 1), "frac_of_sustained_peak": round(tf / PEAK, 4),
                      "finite": bool(torch.isfinite(emb).all())}), flush=True)


CH = int(os.environ.get("CHUNK", "0")) or None      # engine sub-batch override for experiments
BS = float(os.environ.get("BATCH_SCALE", "1"))      # batch-size multiplier for experiments
which = sys.argv[1:] or ["eres", "eres_w24", "ecapa", "ecapa_1p5"]
if "eres" in which:
    run("ERes2NetV2 (26,2,2)", b200spk.ERes2NetV2(precision="bf16", chunk=CH), 298, int(1024 * max(BS, 1)), 24.934, int(256 * BS))
if "eres_w24" in which:
    run("ERes2NetV2 w24s4ep4", b200spk.ERes2NetV2(baseWidth=24, scale=4, expansion=4, precision="bf16", chunk=CH), 298, int(512 * max(BS, 1)), 73.850, int(128 * BS))
if "ecapa" in which:
    run("ECAPA-TDNN C=1024, 10 s", b200spk.ECAPA_TDNN(80, channels=[1024, 1024, 1024, 1024, 3072], precision="bf16"), 998, int(512 * max(BS, 1)), 35.848, int(128 * BS))
if "ecapa_1p5" in which:
    run("ECAPA-TDNN C=1024, 1.5 s", b200spk.ECAPA_TDNN(80, channels=[1024, 1024, 1024, 1024, 3072], precision="bf16"), 148, 4096, 5.552 * 35.848 / 37.416, 512)
