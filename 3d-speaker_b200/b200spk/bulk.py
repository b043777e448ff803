"""Bulk embedding extraction and trial scoring: the loop of ``speakerlab/bin/infer_sv_batch.py`` (every recording is
truncated to ``max_load_len`` seconds, circle-padded to a whole number of ``chunk_size``-second chunks, all chunks go
through fbank + network in batches, and a recording's embedding is the mean of its chunk embeddings,
infer_sv_batch.py:308-323, 388-412), followed by cosine scoring of trial pairs
(``speakerlab/bin/compute_score_metrics.py:113-114``).

B200 layout: the recordings of a shard live back to back in ONE device buffer (int16 PCM or float32); the fbank
kernel's window mode does the truncate / circle-pad / slice itself (no [chunks, 160000] tensor exists), the network
runs on sub-batches of chunks, ``spk_segment_mean`` folds chunk embeddings into recording embeddings.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .fbank import fbank_windows


def chunk_table(lengths, fs=16000, chunk_size=10.0, max_load_len=90.0):
    """Host-side index arithmetic of IterWavList.load_wav/chunk_wav: for recordings of ``lengths`` samples stored
    back to back, returns (starts int64 [n_chunks], periods int32, phases int32, pos int32 [n_wav + 1]) where
    chunk q of recording w reads ``buffer[start_w + (q * cs + i) % min(len_w, max_len)]``."""
    cs, cap = int(chunk_size * fs), int(max_load_len * fs)
    lengths = np.asarray(lengths, dtype=np.int64)
    origin = np.concatenate([[0], np.cumsum(lengths)[:-1]]) if len(lengths) else np.zeros(0, np.int64)
    kept = np.minimum(lengths, cap)
    n_chunks = (kept + cs - 1) // cs
    pos = np.concatenate([[0], np.cumsum(n_chunks)]).astype(np.int32)
    owner = np.repeat(np.arange(len(lengths)), n_chunks)
    q = np.arange(int(pos[-1])) - pos[owner]
    return origin[owner].astype(np.int64), kept[owner].astype(np.int32), (q * cs).astype(np.int32), pos


def segment_mean(emb, pos):
    """emb [N, D] f32 (device), pos int32 [n_wav + 1] (device) -> [n_wav, D]: per-recording mean of chunk embeddings."""
    assert emb.is_cuda and emb.dtype == torch.float32 and emb.is_contiguous() and pos.dtype == torch.int32
    n_wav = pos.shape[0] - 1
    out = torch.empty((n_wav, emb.shape[1]), dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.check(_lib.lib().spk_segment_mean(C.c_void_p(emb.data_ptr()), emb.shape[1], C.c_void_p(pos.data_ptr()), n_wav,
                                               C.c_void_p(out.data_ptr()), _lib.current_stream_ptr()))
    return out


class BulkExtractor:
    """feature_extractor: b200spk.FBank (n_mels, mean_nor); embedding_model: a b200spk network on ``device``."""

    def __init__(self, feature_extractor, embedding_model, device="cuda:0", batchsize=256, fs=16000, chunk_size=10.0,
                 max_load_len=90.0):
        self.fe, self.model = feature_extractor, embedding_model
        self.device = torch.device(device)
        self.batchsize, self.fs, self.chunk_size, self.max_load_len = batchsize, fs, chunk_size, max_load_len

    def chunk_embeddings(self, buffer, lengths):
        """buffer: device [sum(lengths)] int16/float32 (recordings back to back) -> (chunk embeddings [n_chunks, E]
        on the device, pos int32 [n_wav + 1] on the device)."""
        starts, periods, phases, pos = chunk_table(lengths, self.fs, self.chunk_size, self.max_load_len)
        dev = buffer.device
        starts_d, periods_d = torch.from_numpy(starts).to(dev), torch.from_numpy(periods).to(dev)
        phases_d, pos_d = torch.from_numpy(phases).to(dev), torch.from_numpy(pos).to(dev)
        n, cs = starts.shape[0], int(self.chunk_size * self.fs)
        outs = []
        with torch.no_grad():
            for lo in range(0, n, self.batchsize):
                hi = min(n, lo + self.batchsize)
                feats = fbank_windows(buffer, starts_d[lo:hi], periods_d[lo:hi], cs, int(self.fe.n_mels),
                                      bool(self.fe.mean_nor), phases_d[lo:hi])
                outs.append(self.model(feats))
        emb = torch.cat(outs, dim=0) if outs else torch.zeros((0, self.model.embedding_size), device=dev)
        return emb, pos_d

    def __call__(self, buffer, lengths):
        """-> per-recording embeddings [n_wav, E] on the device."""
        emb, pos = self.chunk_embeddings(buffer, lengths)
        return segment_mean(emb, pos)
