"""Numerical prototype (CPU, float32 arithmetic) of the fbank kernel's cancellation-free formulation.

The reference computes X = DFT(w * preemph(d)).  For low bins the pre-emphasised signal is ~30x smaller than
the raw one, so a float32 FFT's absolute error (relative to the raw spectrum level) becomes a large relative
error there.  Identity used by the kernel (w[0] = w[399] = 0 for the povey window):

    X[k] = (1 - 0.97 e^{-2 pi i k/512}) * B[k] + C[k],
    B = DFT(d[m] * w[m+1]),   C = DFT(d[m] * (w[m] - w[m+1]))

B carries no cancellation; C is ~1/128 of B's level, so packing z = d*w+ + i*128*d*dw into ONE complex
512-point FFT gives both to float32 relative accuracy.  Run: python tools/fbank_bc_proto.py [n_utts]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import scipy.fft
from oracle import fbank_oracle as fo

S = np.float32(128.0)


def tables():
    n = np.arange(400, dtype=np.float64)
    w = (0.5 - 0.5 * np.cos(2 * np.pi * n / 399)) ** 0.85
    wp = np.concatenate([w[1:], [0.0]])
    dw = w - wp
    k = np.arange(256)
    H = 1.0 - 0.97 * np.exp(-2j * np.pi * k / 512)
    return wp.astype(np.float32), (dw * 128.0).astype(np.float32), H.astype(np.complex64)


def fbank_bc(wav, mean_nor=True):
    wav = np.asarray(wav, np.float32)
    m = fo.num_frames(wav.shape[0])
    idx = np.arange(m)[:, None] * 160 + np.arange(400)[None, :]
    x = wav[idx]
    mu = x.sum(axis=1, dtype=np.float32, keepdims=True) * np.float32(1 / 400)
    d = x - mu
    wp, dws, H = tables()
    z = np.zeros((m, 512), np.complex64)
    z.real[:, :400] = d * wp
    z.imag[:, :400] = d * dws
    Z = scipy.fft.fft(z, axis=1)
    assert Z.dtype == np.complex64
    Zk = Z[:, :256]
    Zn = np.conj(np.concatenate([Z[:, :1], Z[:, :0:-1]], axis=1)[:, :256])      # conj(Z[512-k])
    B = (Zk + Zn) * np.float32(0.5)
    Cs = (Zk - Zn) * np.complex64(-0.5j)
    X = H[None, :] * B + Cs * np.float32(1 / 128)
    p = X.real * X.real + X.imag * X.imag
    mel = fo.mel_banks(80, np.float32)
    e = p.astype(np.float32) @ mel.T
    e = np.log(np.maximum(e, fo.EPS)).astype(np.float32)
    if mean_nor:
        e = e - e.mean(axis=0, keepdims=True, dtype=np.float32)
    return e


def fbank_std32(wav, mean_nor=True):
    """the straightforward float32 pipeline with a float32 FFT (what torchaudio fp32 does)"""
    fr = fo.frames(wav, np.float32)
    X = scipy.fft.rfft(fr, axis=1)[:, :256]
    p = X.real * X.real + X.imag * X.imag
    mel = fo.mel_banks(80, np.float32)
    e = np.log(np.maximum(p @ mel.T, fo.EPS)).astype(np.float32)
    if mean_nor:
        e = e - e.mean(axis=0, keepdims=True, dtype=np.float32)
    return e


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    rng = np.random.default_rng(1)
    worst_bc = worst_std = 0.0
    where = None
    for i in range(n):
        wav = (0.1 * rng.standard_normal(48000)).astype(np.float32)
        t = fo.fbank(wav, dtype=np.float64)
        a = np.abs(fbank_bc(wav) - t)
        b = np.abs(fbank_std32(wav) - t)
        if a.max() > worst_bc:
            worst_bc, where = a.max(), np.unravel_index(a.argmax(), a.shape)
        worst_std = max(worst_std, b.max())
    print("utts %d  max|bc - fp64| %.3e (frame, mel) %s   max|std32 - fp64| %.3e" % (n, worst_bc, where, worst_std))
