"""fp32 precision mode: time the network forward with the 3xTF32 tensor-core convs (default) against the CUDA-core
convs (SPK_NO_F32X3=1, set by the caller per process) and print the embedding so two runs can be compared.

usage: python tools/time_fp32.py [campplus|eres2netv2|ecapa] [batch]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-speaker_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import b200spk  # noqa: E402


def main():
    fam = sys.argv[1] if len(sys.argv) > 1 else "campplus"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    torch.manual_seed(3)
    if fam == "campplus":
        m, T = b200spk.CAMPPlus(embedding_size=192, precision="fp32"), 148
    elif fam == "eres2netv2":
        m, T = b200spk.ERes2NetV2(precision="fp32"), 298
    else:
        m, T = b200spk.ECAPA_TDNN(80, lin_neurons=192, channels=[1024, 1024, 1024, 1024, 3072], precision="fp32"), 998
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(5)
    feats = torch.randn(B, T, 80, device="cuda", generator=g)
    with torch.no_grad():
        for _ in range(2):
            e = m(feats)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            e = m(feats)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
    out = os.environ.get("OUT")
    if out:
        np.save(out, e.cpu().numpy())
    print(json.dumps({"family": fam, "batch": B, "no_f32x3": os.environ.get("SPK_NO_F32X3", "0"), "ms": round(dt * 1e3, 2),
                      "segments_per_s": round(B / dt, 1), "finite": bool(torch.isfinite(e).all()), "emb0": e[0, :4].tolist()}))


if __name__ == "__main__":
    main()
