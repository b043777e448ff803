"""CAM++ (512-d, bf16) through EmbeddingExtractor at the reference's own batch size (64 windows per fbank/forward
call, infer_diarization.py:629-635) and at larger batches: host buffers in, host embeddings out."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import torch
import b200spk
import bench

model = b200spk.CAMPPlus(embedding_size=512, precision="bf16")
tsd, _ = bench.make_weights(model)
model.load_state_dict(tsd)
model = model.cuda().eval()
fb = b200spk.FBank(80, 16000, mean_nor=True)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
host = torch.from_numpy(bench.make_windows(N, seed=3)).pin_memory()
for bs in (64, 256, 2048):
    ex = b200spk.EmbeddingExtractor(fb, model, batchsize=bs, reuse_output=True, head=min(bs, 512))
    for _ in range(2):
        ex(host)
    torch.cuda.synchronize()
    l0 = b200spk.lib().spk_launch_count()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ex(host)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"batch": bs, "windows": N, "ms": round(dt * 1e3, 2), "emb_per_s": round(N / dt, 1),
                      "launches_per_call": (b200spk.lib().spk_launch_count() - l0) // reps // ((N + bs - 1) // bs)}), flush=True)
