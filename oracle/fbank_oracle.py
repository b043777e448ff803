"""numpy restatement of the reference front end (Kaldi fbank + utterance CMN).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The arithmetic lives in a third-party
dependency of the reference: torchaudio.compliance.kaldi (reference pin
``torchaudio>=0.10.1``, container 2.11.0), reached from
speakerlab/process/processor.py:143-158.  Cited below as ``kaldi.py:LINE``
(= site-packages/torchaudio/compliance/kaldi.py).

Pinned by tests/golden/fbank_*.npz, minted from the imported reference by
oracle/gen_golden.py (the reference itself holds no golden vectors).
``dtype=np.float64`` gives the 'truth' side of the two-sided criterion in
SURVEY.md section 7-3.
"""
import numpy as np

SAMPLE_RATE = 16000
FRAME_LEN = 400      # 25 ms   kaldi.py:138
FRAME_SHIFT = 160    # 10 ms   kaldi.py:137
NFFT = 512           # kaldi.py:139 round_to_power_of_two
PREEMPH = 0.97
EPS = np.float32(1.1920928955078125e-07)  # kaldi.py:21-22


def num_frames(n_samples):
    """kaldi.py:67 (snip_edges=True)."""
    if n_samples < FRAME_LEN:
        return 0
    return 1 + (n_samples - FRAME_LEN) // FRAME_SHIFT


def povey_window(dtype=np.float32):
    """kaldi.py:98-100: hann_window(400, periodic=False) ** 0.85."""
    n = np.arange(FRAME_LEN, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / (FRAME_LEN - 1))
    return (hann ** 0.85).astype(dtype)


def mel_scale(f):
    return 1127.0 * np.log(1.0 + f / 700.0)


def mel_banks(num_bins=80, dtype=np.float32, low_freq=20.0, high_freq=0.0):
    """kaldi.py:436-511 (no VTLN).  Returns [num_bins, NFFT/2].  torchaudio ALWAYS builds
    the bank in fp32 and casts afterwards (kaldi.py:621-624), so the fp32 arithmetic is
    mirrored for every dtype - the fp64 'truth' run uses fp32-rounded weights too."""
    import torch   # fp32 log must round exactly like the reference's torch.log (numpy's differs by 1 ulp,
    #                which moves narrow-filter weights by ~1e-5 and the log-energy by 2e-5)
    num_fft_bins = NFFT // 2
    nyquist = 0.5 * SAMPLE_RATE
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = SAMPLE_RATE / NFFT
    mel_low = float(mel_scale(low_freq))     # python float (double), kaldi.py:462-463
    mel_high = float(mel_scale(high_freq))
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return bins.numpy().astype(dtype)


def frames(wav, dtype=np.float32):
    """kaldi.py:44-83 + 154-217: strided frames, DC removal, pre-emphasis, povey window,
    zero pad 400->512.  wav: [n] -> [m, 512]."""
    wav = np.asarray(wav, dtype=dtype)
    m = num_frames(wav.shape[0])
    idx = np.arange(m)[:, None] * FRAME_SHIFT + np.arange(FRAME_LEN)[None, :]
    x = wav[idx]                                                   # kaldi.py:82-83
    x = x - x.mean(axis=1, keepdims=True, dtype=dtype)             # kaldi.py:183-186
    prev = np.concatenate([x[:, :1], x[:, :-1]], axis=1)           # replicate pad kaldi.py:193-198
    x = x - dtype(PREEMPH) * prev
    x = x * povey_window(dtype)[None, :]                           # kaldi.py:200-204
    out = np.zeros((m, NFFT), dtype=dtype)                         # kaldi.py:207-211
    out[:, :FRAME_LEN] = x
    return out


def fbank(wav, num_mel_bins=80, mean_nor=True, dtype=np.float32):
    """FBank.__call__ (speakerlab/process/processor.py:143-158) for one utterance.

    wav [n] float in [-1,1] scale -> [m, num_mel_bins].  rfft -> |.|^2 (kaldi.py:616-618),
    mel matmul with a zero Nyquist column (kaldi.py:621-630), log(max(., eps))
    (kaldi.py:633), then CMN over frames (processor.py:156-157)."""
    fr = frames(wav, dtype)
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    spec = np.fft.rfft(fr.astype(np.float64), axis=1).astype(cdt)
    power = (np.abs(spec).astype(dtype)) ** 2
    mel = mel_banks(num_mel_bins, dtype)
    mel = np.concatenate([mel, np.zeros((num_mel_bins, 1), dtype=dtype)], axis=1)
    e = power.astype(dtype) @ mel.T
    e = np.log(np.maximum(e, dtype(EPS))).astype(dtype)
    if mean_nor:
        e = e - e.mean(axis=0, keepdims=True, dtype=dtype)
    return e.astype(dtype)


def fbank_batch(wavs, num_mel_bins=80, mean_nor=True, dtype=np.float32):
    """torch.vmap(FBank) call sites (speakerlab/bin/infer_diarization.py:634): [B,n] -> [B,m,80]."""
    return np.stack([fbank(w, num_mel_bins, mean_nor, dtype) for w in wavs])
