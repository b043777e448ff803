"""Seeded synthetic inputs and weights shared by tests, bench and golden minting.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything here is numpy-only and
driven by ``np.random.Generator(PCG64)`` streams, which are stable across numpy
versions and machines, so the GPU box regenerates bit-identical inputs/weights
from a seed instead of shipping them.

Restated reference logic (cited where used):
  * sub-segment chunking  speakerlab/bin/infer_diarization.py:606-619
  * circle_pad            speakerlab/utils/utils.py:232-238
  * window slicing        speakerlab/bin/infer_diarization.py:624-627
"""
import zlib

import numpy as np

FS = 16000


# --------------------------------------------------------------------------- audio
def white_noise(n_utts, n_samples, seed, scale=0.1):
    rng = np.random.default_rng([seed, 0xA0D10])
    return (scale * rng.standard_normal((n_utts, n_samples))).astype(np.float32)


# FM rates are multiples of 8/3 Hz: FM period and the AM period (rate r/2) both divide the 0.75 s
# sub-segment hop, so every window of a speaker sees the same modulation phase.  With free-running
# rates the within-speaker embeddings trace a phase loop, the 58-nearest-neighbour graph of a speaker
# becomes a ring with many near-zero Laplacian modes, and the reference's eigengap rule returns
# k = 14 instead of the number of speakers (measured with the CPU oracle).
_FM_RATES = tuple(8.0 / 3.0 * n for n in range(1, 9))


def _fm_voice(k, n_spk, t):
    """One spectro-temporal 'speaker' (SURVEY.md section 7-4): three FM partials at
    base*{1,1.9,3.1}, FM depth 0.35 at rate r_k, AM at r_k/2.  Stationary spectra
    are erased by utterance CMN, FM/AM patterns are not."""
    base = 250.0 * (3000.0 / 250.0) ** (k / max(n_spk - 1, 1))
    r = _FM_RATES[k % len(_FM_RATES)]
    x = np.zeros_like(t)
    for mult, amp in ((1.0, 1.0), (1.9, 0.6), (3.1, 0.35)):
        f0 = base * mult
        if f0 * 1.35 > 7600.0:
            continue
        # phase of f(t) = f0 (1 + 0.35 sin(2 pi r t))
        ph = 2 * np.pi * f0 * (t - 0.35 / (2 * np.pi * r) * np.cos(2 * np.pi * r * t))
        x += amp * np.sin(ph)
    x *= 0.6 + 0.4 * np.sin(2 * np.pi * (r / 2.0) * t)
    return x


def fm_meeting(seconds, n_spk, seed, fs=FS, noise_db=-25.0):
    """Synthetic meeting: random turns U(2,20) s, no immediate repeats, ~100 % speech, peak 0.2,
    white noise floor at ``noise_db`` re peak.  The noise floor matters: with phase-locked voices the
    windows of a speaker differ only through it, and at -40 dB they are so alike that the
    58-nearest-neighbour pruning is an arbitrary pick among near-ties (cosines equal to 1e-7);
    at -25 dB the CPU oracle returns k = 8 with 100 % pure-segment accuracy for 8 speakers / 1 hour
    and is stable under 1e-6 relative perturbations of the embeddings (measured).
    Returns (wav float32 [n], turns)."""
    rng = np.random.default_rng([seed, 0x3EE7])
    n = int(round(seconds * fs))
    t = np.arange(n, dtype=np.float64) / fs
    voices = [_fm_voice(k, n_spk, t) for k in range(n_spk)]
    wav = np.zeros(n, dtype=np.float64)
    turns = []
    pos, prev = 0.0, -1
    while pos < seconds:
        dur = float(rng.uniform(2.0, 20.0))
        spk = int(rng.integers(0, n_spk))
        if n_spk > 1:
            while spk == prev:
                spk = int(rng.integers(0, n_spk))
        end = min(pos + dur, seconds)
        a, b = int(round(pos * fs)), int(round(end * fs))
        wav[a:b] = voices[spk][a:b]
        turns.append((pos, end, spk))
        pos, prev = end, spk
    wav *= 0.2 / max(np.abs(wav).max(), 1e-9)
    wav += 0.2 * 10 ** (noise_db / 20) * rng.standard_normal(n)
    return wav.astype(np.float32), turns


def chunk(st, ed, dur=1.5, step=0.75):
    """speakerlab/bin/infer_diarization.py:606-619 (Diarization3Dspeaker.chunk)."""
    chunks = []
    if ed - st <= 0:
        return chunks
    s = st
    made = False
    while s + dur < ed + step:
        chunks.append([s, min(s + dur, ed)])
        s += step
        made = True
    if not made:
        chunks.append([st, ed])
    return chunks


def circle_pad(x, target_len):
    """speakerlab/utils/utils.py:232-238."""
    n = x.shape[0]
    if n >= target_len:
        return x
    reps = int(np.ceil(target_len / n))
    return np.concatenate([x] * reps)[:target_len]


def cut_windows(wav, chunks, fs=FS):
    """speakerlab/bin/infer_diarization.py:624-627: slice, circle-pad to the longest, stack."""
    segs = [wav[int(st * fs):int(ed * fs)] for st, ed in chunks]
    max_len = max(s.shape[0] for s in segs)
    return np.stack([circle_pad(s, max_len) for s in segs]).astype(np.float32)


def turn_labels(chunks, turns):
    """Ground-truth speaker of each chunk and whether it is 'pure' (inside one turn)."""
    starts = np.array([t[0] for t in turns])
    ends = np.array([t[1] for t in turns])
    spk = np.array([t[2] for t in turns])
    lab = np.zeros(len(chunks), dtype=np.int64)
    pure = np.zeros(len(chunks), dtype=bool)
    for i, (st, ed) in enumerate(chunks):
        mid = 0.5 * (st + ed)
        j = int(np.searchsorted(ends, mid, side="right"))
        j = min(j, len(turns) - 1)
        lab[i] = spk[j]
        pure[i] = (st >= starts[j] - 1e-9) and (ed <= ends[j] + 1e-9)
    return lab, pure


# --------------------------------------------------------------------------- weights
def _key_rng(seed, key):
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def fill_state_dict(shapes, seed, randomize_bn=True, gain=2.0 ** 0.5):
    """Deterministic weights for a {name: shape} table (a model's state_dict layout).

    conv/linear weights ~ N(0, gain^2/fan_in) (He for the default gain sqrt(2); deep residual nets
    need a smaller gain to stay out of the chaotic regime), biases ~ N(0, 0.05); with
    ``randomize_bn`` every BatchNorm gets running_mean~N(0,0.1), running_var~U(0.5,1.5),
    weight~U(0.5,1.5), bias~N(0,0.1) so BN folding is exercised (SURVEY.md section 8d
    config 1); otherwise fresh-BN values (mean 0, var 1, weight 1, bias 0).
    Returns {name: np.ndarray} (float32; num_batches_tracked int64)."""
    out = {}
    for key, shape in shapes.items():
        shape = tuple(int(s) for s in shape)
        rng = _key_rng(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = np.zeros(shape, dtype=np.int64)
        elif leaf == "running_mean":
            v = 0.1 * rng.standard_normal(shape) if randomize_bn else np.zeros(shape)
            out[key] = v.astype(np.float32)
        elif leaf == "running_var":
            v = rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)
            out[key] = v.astype(np.float32)
        elif leaf == "weight" and len(shape) == 1:  # BN affine weight
            v = rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)
            out[key] = v.astype(np.float32)
        elif leaf == "bias" and _is_bn_bias(key, shapes):
            v = 0.1 * rng.standard_normal(shape) if randomize_bn else np.zeros(shape)
            out[key] = v.astype(np.float32)
        elif leaf == "bias":
            out[key] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        else:  # conv / linear weight
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            out[key] = (gain / np.sqrt(fan_in) * rng.standard_normal(shape)).astype(np.float32)
    return out


def _is_bn_bias(key, shapes):
    return key.rsplit(".", 1)[0] + ".running_mean" in shapes
