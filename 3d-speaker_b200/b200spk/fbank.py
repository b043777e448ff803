"""Drop-in for ``speakerlab.process.processor.FBank`` (speakerlab/process/processor.py:133-158).

Same constructor, attributes and call convention; the arithmetic (torchaudio.compliance.kaldi
fbank + utterance CMN) runs in the sm_100a kernel behind ``spk_fbank_f32``.  Works under
``torch.vmap`` (speakerlab/bin/infer_diarization.py:634) through a vmap rule on the custom op,
which forwards the whole batch to ONE kernel launch.
"""
import ctypes as C
import threading

import torch

from . import _lib

_FRAME_LEN, _FRAME_SHIFT, _NFFT = 400, 160, 512
_tables_lock = threading.Lock()
_tables_for = None


def num_frames(n_samples):
    """kaldi.py:67 (snip_edges=True)."""
    return int(_lib.lib().spk_fbank_num_frames(int(n_samples)))


def _host_tables(n_mels):
    """Mel bank computed with the SAME torch fp32 arithmetic torchaudio uses (kaldi.py:436-511, always built
    on CPU in fp32 whatever the waveform dtype, kaldi.py:621-624), so the kernel's filter weights agree with
    the reference's to the last bit.  The povey window is left to the library, which evaluates it in float64
    (its cancellation-free formulation needs w[n] - w[n+1] to full precision; see csrc/fbank.cu)."""
    num_fft_bins = _NFFT // 2
    nyquist = 8000.0
    low_freq, high_freq = 20.0, nyquist
    fft_bin_width = 16000.0 / _NFFT
    import math
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (n_mels + 1)
    b = torch.arange(n_mels).unsqueeze(1)
    left, center, right = mel_low + b * delta, mel_low + (b + 1.0) * delta, mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bank = torch.max(torch.zeros(1), torch.min(up, down)).to(torch.float32).contiguous()
    return bank


def _ensure_tables(n_mels):
    global _tables_for
    with _tables_lock:
        if _tables_for == n_mels:
            return
        bank = _host_tables(n_mels)
        _lib.check(_lib.lib().spk_fbank_set_tables(None, C.c_void_p(bank.data_ptr()), int(n_mels)))
        _tables_for = n_mels


def fbank_batch(wavs, n_mels=80, mean_nor=True):
    """wavs [B, n] float32 in [-1, 1] scale or int16 PCM (scaled by 1/32768 in the kernel, what
    speakerlab/utils/fileio.py:115-117 does on the host) -> [B, m, n_mels] float32 on the same device.
    CUDA: one kernel on the current stream; CPU float32: the library's host-buffer entry point."""
    assert wavs.dim() == 2 and wavs.dtype in (torch.float32, torch.int16)
    _ensure_tables(n_mels)
    B, n = wavs.shape
    if wavs.stride(1) != 1:
        wavs = wavs.contiguous()
    m = num_frames(n)
    L = _lib.lib()
    if wavs.is_cuda:
        out = torch.empty((B, m, n_mels), dtype=torch.float32, device=wavs.device)
        fn = L.spk_fbank_i16 if wavs.dtype == torch.int16 else L.spk_fbank_f32
        with torch.cuda.device(wavs.device):
            _lib.check(fn(C.c_void_p(wavs.data_ptr()), B, n, wavs.stride(0) if B > 1 else n,
                          C.c_void_p(out.data_ptr()), n_mels, int(bool(mean_nor)), _lib.current_stream_ptr()))
        return out
    assert wavs.dtype == torch.float32, "the host-buffer entry point takes float32"
    out = torch.empty((B, m, n_mels), dtype=torch.float32)
    _lib.check(L.spk_fbank_host_f32(C.c_void_p(wavs.data_ptr()), B, n, wavs.stride(0) if B > 1 else n,
                                    C.c_void_p(out.data_ptr()), n_mels, int(bool(mean_nor))))
    return out


def fbank_windows(wav, starts, lens, n_samples, n_mels=80, mean_nor=True, phases=None):
    """Front end on pieces of recordings resident in ONE device buffer ``wav`` [n_total] (float32 or int16): sample
    i of row b is ``wav[starts[b] + (phases[b] + i) % lens[b]]`` -> [B, m, n_mels].  With ``phases=None`` row b is
    the window [starts[b], starts[b] + lens[b]) circle-padded to ``n_samples`` (the slice + circle_pad + stack of
    speakerlab/bin/infer_diarization.py:621-627); with starts/lens = a recording and phases = chunk offsets it is
    the circle-pad-then-slice chunking of speakerlab/bin/infer_sv_batch.py:388-412.  No [B, n_samples] tensor is
    ever built: the kernel's loads do the gather.  starts: int64 [B]; lens, phases: int32 [B]; all on the device."""
    assert wav.dim() == 1 and wav.is_cuda and wav.dtype in (torch.float32, torch.int16) and wav.is_contiguous()
    assert starts.dtype == torch.int64 and lens.dtype == torch.int32 and starts.is_cuda and lens.is_cuda
    assert phases is None or (phases.dtype == torch.int32 and phases.is_cuda and phases.shape == lens.shape)
    _ensure_tables(n_mels)
    B = starts.shape[0]
    m = num_frames(n_samples)
    out = torch.empty((B, m, n_mels), dtype=torch.float32, device=wav.device)
    with torch.cuda.device(wav.device):
        _lib.check(_lib.lib().spk_fbank_windows(C.c_void_p(wav.data_ptr()), int(wav.dtype == torch.int16), wav.shape[0],
                                                C.c_void_p(starts.data_ptr()), C.c_void_p(lens.data_ptr()),
                                                None if phases is None else C.c_void_p(phases.data_ptr()), B,
                                                int(n_samples), C.c_void_p(out.data_ptr()), n_mels,
                                                int(bool(mean_nor)), _lib.current_stream_ptr()))
    return out


@torch.library.custom_op("b200spk::fbank", mutates_args=())
def _fbank_op(wav: torch.Tensor, n_mels: int, mean_nor: bool) -> torch.Tensor:
    return fbank_batch(wav, n_mels, mean_nor)


@_fbank_op.register_fake
def _(wav, n_mels, mean_nor):
    n = wav.shape[1]
    m = 0 if n < _FRAME_LEN else 1 + (n - _FRAME_LEN) // _FRAME_SHIFT
    return wav.new_empty((wav.shape[0], m, n_mels))


def _fbank_vmap(info, in_dims, wav, n_mels, mean_nor):
    wav = wav.movedim(in_dims[0], 0)            # [V, B, n]
    V, B, n = wav.shape
    out = _fbank_op(wav.reshape(V * B, n), n_mels, mean_nor)
    return out.reshape(V, B, out.shape[1], out.shape[2]), 0


torch.library.register_vmap(_fbank_op, _fbank_vmap)


class FBank(object):
    """Same interface as the reference class: ``FBank(n_mels, sample_rate, mean_nor=False)``,
    ``__call__(wav, dither=0)`` with wav ``[T]`` or ``[C, T]`` (channel 0 is used) -> ``[m, n_mels]``."""

    def __init__(self, n_mels, sample_rate, mean_nor: bool = False):
        self.n_mels = n_mels
        self.sample_rate = sample_rate
        self.mean_nor = mean_nor

    def __call__(self, wav, dither=0):
        sr = 16000
        assert sr == self.sample_rate
        assert dither == 0, "the B200 front end is dither-free (the reference call sites pass dither=0)"
        if len(wav.shape) == 1:
            wav = wav.unsqueeze(0)
        if wav.shape[0] > 1:
            wav = wav[0, :].unsqueeze(0)
        assert len(wav.shape) == 2 and wav.shape[0] == 1
        assert wav.shape[1] >= _FRAME_LEN, \
            "choose a window size {} that is [2, {}]".format(_FRAME_LEN, wav.shape[1])   # kaldi.py:142
        if wav.dtype != torch.int16:
            wav = wav.to(torch.float32)
        feat = _fbank_op(wav, int(self.n_mels), bool(self.mean_nor))
        return feat.squeeze(0)

    def batch(self, wavs):
        """[B, n] -> [B, m, n_mels] in one launch (what torch.vmap(self) dispatches to)."""
        if wavs.dtype != torch.int16:
            wavs = wavs.to(torch.float32)
        return fbank_batch(wavs, int(self.n_mels), bool(self.mean_nor))

    def windows(self, wav, starts, lens, n_samples, phases=None):
        """Pieces of resident recordings (see ``fbank_windows``)."""
        return fbank_windows(wav, starts, lens, n_samples, int(self.n_mels), bool(self.mean_nor), phases)
