#!/bin/bash
# Quick regression + speed check on the GPU box: parity tests, then the three model benches and the CAM++ forward time.
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/bench_models.py eres eres_w24 ecapa 2>&1 | tail -3
python - <<'PY'
import os, sys, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "3d-speaker_b200")]
import b200spk, bench
model = b200spk.CAMPPlus(embedding_size=512, precision="bf16")
tsd, _ = bench.make_weights(model); model.load_state_dict(tsd); model = model.cuda().eval()
feats = torch.randn(8192, 148, 80, device="cuda")
with torch.no_grad():
    for _ in range(2): model(feats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): model(feats)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("campplus forward: %.2f ms per 8192 segments = %.0f seg/s, %.2f ms per 2048" % (ms, 8192 / ms * 1e3, ms / 4))
PY
