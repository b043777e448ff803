"""Forward-only throughput of CAM++ (512-d, bf16) for coarse/fine sub-batch pairs (features resident)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import torch
import b200spk
import bench

pairs = [tuple(int(x) for x in p.split("/")) for p in (sys.argv[1:] or ["2048/1024", "1024/1024", "768/768", "512/512", "384/384", "256/256"])]
feats = torch.randn(int(os.environ.get("NSEG", "8192")), 148, 80, device="cuda")
for coarse, fine in pairs:
    model = b200spk.CAMPPlus(embedding_size=512, precision="bf16", chunk=(coarse, fine))
    tsd, _ = bench.make_weights(model)
    model.load_state_dict(tsd)
    model = model.cuda().eval()
    with torch.no_grad():
        for _ in range(2):
            model(feats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            model(feats)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"coarse": coarse, "fine": fine, "ms_per_8192": round(ms, 2), "emb_per_s": round(feats.shape[0] / ms * 1e3)}), flush=True)
    del model
