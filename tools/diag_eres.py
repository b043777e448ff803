"""Per-stage rel-L2 of the ERes2NetV2 GPU forward against the CPU oracle taps."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np, torch
import b200spk
from b200spk import _lib
from oracle import eres2netv2_oracle, gen_golden, synth

g = np.load(os.path.join(ROOT, "tests/golden/eres2netv2.npz"))
for name, kw, batch, n_samples, wseed in gen_golden.eres2netv2_cases():
    for prec in ("fp32", "bf16"):
        m = b200spk.ERes2NetV2(precision=prec, **kw)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = synth.fill_state_dict(shapes, wseed, randomize_bn=True, gain=gen_golden.ERES_GAIN)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        m = m.cuda().eval()
        feats = torch.from_numpy(g[name + ".feats"]).cuda()
        taps = {}
        ref = eres2netv2_oracle.forward(sd, feats.cpu().numpy(), scale=kw["scale"], taps=taps).numpy()
        with torch.no_grad():
            got = m(feats).cpu().numpy()
        eng = m._engine
        T, B = feats.shape[1], feats.shape[0]
        prog = eng.model.programs[T]
        print("==", name, prec, "emb rel %.3e" % (np.linalg.norm(got - ref) / np.linalg.norm(ref)))
        def cmp(bufname, tap, C):
            x = eng.model.read_buffer(T, bufname, B).cpu()
            t = taps[tap]                                  # [B, C, H, W]
            b_, c_, h_, w_ = t.shape
            x = x[: b_ * h_ * w_ * C].view(b_, h_, w_, C).permute(0, 3, 1, 2)
            print("   %-8s rel %.3e  (|ref| max %.2f)" % (tap, ((x - t).norm() / t.norm()).item(), t.abs().max().item()))
        cmp("out3", "layer3", taps["layer3"].shape[1])
        cmp("fuse34", "fuse34", taps["fuse34"].shape[1])
        # layer4 output lives in whichever ping buffer the last conv3 wrote
        last = [o for o in prog.ops if o.kind == _lib.OP_CONV and o.Cout == taps["layer4"].shape[1] and o.res_buf >= 0][-1]
        nm = [k for k, v in prog.names.items() if v == last.out_buf][0]
        cmp(nm, "layer4", taps["layer4"].shape[1])
