"""Per-buffer rel-L2 of the bf16 (tcgen05) forward against the fp32 (CUDA-core) forward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np, torch
import b200spk
from oracle import gen_golden, synth

def build(prec, wseed, bnrand, emb=192):
    m = b200spk.CAMPPlus(embedding_size=emb, precision=prec)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth.fill_state_dict(shapes, wseed, randomize_bn=bnrand)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.cuda().eval()

for wseed, bnrand, inp in ((101, True, "fm"), (7, True, "noise"), (104, False, "fm")):
    if inp == "fm":
        wavs = gen_golden.campplus_input(16, 24000, seed=5)
    else:
        wavs = synth.white_noise(16, 24000, seed=123)
    feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
    a, b = build("fp32", wseed, bnrand), build("bf16", wseed, bnrand)
    with torch.no_grad():
        ea, eb = a(feats), b(feats)
    print("== weights", wseed, "bnrand", bnrand, "input", inp)
    T, B = feats.shape[1], feats.shape[0]
    for name in ("fcm_a", "fcm_d", "fcm_b", "block1", "block2", "block3", "bottleneck", "gate", "final", "stats"):
        x = a._engine.model.read_buffer(T, name, B)
        y = b._engine.model.read_buffer(T, name, B)
        print("  %-10s rel-L2 %.3e   max|fp32| %.3e" % (name, ((x - y).norm() / x.norm()).item(), x.abs().max().item()))
    rel = ((ea - eb).norm() / ea.norm()).item()
    cos = torch.nn.functional.cosine_similarity(ea, eb).min().item()
    print("  embedding  rel-L2 %.3e   min cos %.6f  |emb| %.3e" % (rel, cos, ea.norm(dim=1).mean().item()))
