"""per-buffer rel-L2 of the bf16 (tcgen05) forward against the fp32 forward ON THE GPU, every named
workspace buffer as the last forward left it, for CAM++ (weight set 101) and ERes2NetV2 w24s4ep4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np, torch
import b200spk
from oracle import gen_golden, synth


def load(m, wseed, gain=None):
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    kw = {} if gain is None else {"gain": gain}
    sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, **kw)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.cuda().eval()


def compare(name, a, b, feats):
    with torch.no_grad():
        ea, eb = a(feats), b(feats)
    T, B = feats.shape[1], feats.shape[0]
    print("==", name)
    names = a._engine.model.programs[T].names
    for buf in sorted(names, key=lambda k: names[k]):
        if names[buf] < 2:
            continue
        x = a._engine.model.read_buffer(T, buf, B)
        y = b._engine.model.read_buffer(T, buf, B)
        print("  %-12s rel-L2 %.3e   max|fp32| %.3e" % (buf, ((x - y).norm() / x.norm()).item(), x.abs().max().item()))
    cos = torch.nn.functional.cosine_similarity(ea, eb).min().item()
    print("  embedding    rel-L2 %.3e   min cos %.6f" % (((ea - eb).norm() / ea.norm()).item(), cos))


wavs = gen_golden.campplus_input(16, 24000, seed=5)
feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
compare("CAM++ weights 101", load(b200spk.CAMPPlus(embedding_size=192, precision="fp32"), 101),
        load(b200spk.CAMPPlus(embedding_size=192, precision="bf16"), 101), feats)
kw = dict(baseWidth=24, scale=4, expansion=4)
wavs = gen_golden.campplus_input(6, 24000, seed=77)
feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
compare("ERes2NetV2 w24s4ep4 weights 202", load(b200spk.ERes2NetV2(precision="fp32", **kw), 202, gen_golden.ERES_GAIN),
        load(b200spk.ERes2NetV2(precision="bf16", **kw), 202, gen_golden.ERES_GAIN), feats)

# ---- the models' own default initialisation (what "random-init" means in BASELINE.json): bf16 vs fp32 on the GPU
def default_init(ctor, seed, **kw):
    torch.manual_seed(seed)
    a = ctor(precision="fp32", **kw)
    b = ctor(precision="bf16", **kw)
    b.load_state_dict(a.state_dict())
    return a.cuda().eval(), b.cuda().eval()

g = torch.Generator().manual_seed(9)
for name, ctor, kw, secs, n in (("CAM++ e192 default init", b200spk.CAMPPlus, dict(embedding_size=192), 1.5, 16),
                                ("CAM++ e512 default init", b200spk.CAMPPlus, dict(embedding_size=512), 1.5, 16),
                                ("ERes2NetV2 (26,2,2) default init", b200spk.ERes2NetV2, dict(), 3.0, 8),
                                ("ERes2NetV2 w24s4ep4 default init", b200spk.ERes2NetV2, dict(baseWidth=24, scale=4, expansion=4), 3.0, 8),
                                ("ECAPA C=1024 default init", lambda precision, **k: b200spk.ECAPA_TDNN(80, channels=[1024, 1024, 1024, 1024, 3072], precision=precision), dict(), 10.0, 4)):
    a, b = default_init(ctor, 11, **kw)
    wavs = gen_golden.campplus_input(n, int(secs * 16000), seed=55)
    feats = b200spk.fbank_batch(torch.from_numpy(wavs).cuda())
    with torch.no_grad():
        ea, eb = a(feats), b(feats)
    cos = torch.nn.functional.cosine_similarity(ea, eb).min().item()
    print("%-36s bf16 vs fp32: rel-L2 %.3e  min cos %.6f  |emb| %.3e" % (name, ((ea - eb).norm() / ea.norm()).item(), cos, ea.norm(dim=1).mean().item()))
