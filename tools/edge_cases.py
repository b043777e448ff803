import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "3d-speaker_b200")]
import numpy as np, torch, b200spk
from oracle import synth, campplus_oracle, ecapa_oracle, eres2netv2_oracle
torch.manual_seed(0)
def mk(cls, prec, **kw):
    m = cls(precision=prec, **kw) if cls is not b200spk.ECAPA_TDNN else cls(80, precision=prec, **kw)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth.fill_state_dict(shapes, 3, randomize_bn=True, gain=1.0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.cuda().eval(), {k: torch.from_numpy(v) for k, v in sd.items()}
cases = [("campplus", b200spk.CAMPPlus, dict(embedding_size=192), lambda sd, f: campplus_oracle.forward(sd, f)),
         ("ecapa", b200spk.ECAPA_TDNN, dict(channels=[512, 512, 512, 512, 1536]), lambda sd, f: ecapa_oracle.forward(sd, f)),
         ("eres", b200spk.ERes2NetV2, dict(), lambda sd, f: eres2netv2_oracle.forward(sd, f, scale=2))]
for name, cls, kw, orc in cases:
    m32, sd = mk(cls, "fp32", **kw)
    m16, _ = mk(cls, "bf16", **kw)
    for B, T in [(1, 148), (3, 149), (2, 57), (5, 200), (130, 148)]:
        f = torch.randn(B, T, 80, device="cuda")
        with torch.no_grad():
            e32 = m32(f).cpu().numpy(); e16 = m16(f).cpu().numpy()
        msg = ""
        if B <= 3:
            ref = orc(sd, f.cpu().numpy()).numpy()
            msg = "fp32 rel vs oracle %.1e" % (np.linalg.norm(e32 - ref) / np.linalg.norm(ref))
        cos = ((e16 * e32).sum(1) / (np.linalg.norm(e16, axis=1) * np.linalg.norm(e32, axis=1))).min()
        print(name, "B=%d T=%d" % (B, T), msg, "bf16 cos vs fp32 %.5f" % cos, "finite", bool(np.isfinite(e16).all() and np.isfinite(e32).all()), flush=True)
    e = m16(torch.zeros(0, 148, 80, device="cuda"))
    print(name, "empty batch ->", tuple(e.shape))
