"""Small end-to-end case for compute-sanitizer (racecheck / synccheck, ONE tool per gpurun call):
fbank (fused + frame-range + repair paths), CAM++ bf16 with several tiles per persistent CTA, ERes2NetV2 bf16,
spectral back end with the cooperative Lanczos kernel.  Usage on the GPU box:
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "3d-speaker_b200")]
import numpy as np
import torch
import b200spk
from oracle import gen_golden, synth

which = sys.argv[1:] or ["fbank", "campplus", "eres", "ecapa", "cluster"]
if "fbank" in which:
    x = torch.from_numpy(synth.white_noise(700, 24000, seed=1)).cuda()
    a = b200spk.fbank_batch(x, 80, True)                 # fused CMN path
    b = b200spk.fbank_batch(x[:3], 80, True)             # frame ranges + cmn kernel
    assert torch.equal(a[:3], b)
    b200spk.lib().spk_fbank_set_repair(5e-2, 64)         # force the float64 repair path to run
    c = b200spk.fbank_batch(x[:8], 80, True)
    b200spk.lib().spk_fbank_set_repair(5e-4, 8)
    assert (c - a[:8]).abs().max().item() < 1e-4
    print("fbank ok")
if "campplus" in which:
    torch.manual_seed(6)
    feats = torch.randn(48, 148, 80, device="cuda")
    m = b200spk.CAMPPlus(embedding_size=192, precision="bf16", chunk=(48, 48)).cuda().eval()
    with torch.no_grad():
        e = m(feats)
    assert torch.isfinite(e).all()
    print("campplus ok")
if "eres" in which:
    torch.manual_seed(7)
    feats = torch.randn(6, 148, 80, device="cuda")
    m = b200spk.ERes2NetV2(precision="bf16").cuda().eval()
    with torch.no_grad():
        e = m(feats)
    assert torch.isfinite(e).all()
    print("eres ok")
if "ecapa" in which:
    torch.manual_seed(8)
    feats = torch.randn(3, 298, 80, device="cuda")
    m = b200spk.ECAPA_TDNN(80, channels=[512, 512, 512, 512, 1536], precision="bf16").cuda().eval()
    with torch.no_grad():
        e = m(feats)
    assert torch.isfinite(e).all()
    print("ecapa ok")
if "cluster" in which:
    X, _ = gen_golden.cluster_input(300, 64, 3, 22)
    np.random.seed(0)
    lab = b200spk.SpectralCluster()(X)
    assert len(np.unique(lab)) == 3
    print("cluster ok")
