"""Sub-segment diarization pipeline on the GPU: the hot loop of
``Diarization3Dspeaker.__call__`` (speakerlab/bin/infer_diarization.py:256-315) without VAD and
RTTM writing: chunk -> windows -> fbank -> embeddings -> spectral clustering.

Multi-GPU (new behaviour): the N sub-segments of ONE recording are sharded
contiguously over the ranks of a ``torch.distributed`` group (no data-path collective during
extraction), then ONE ``all_gather`` of the ``[N/G, E]`` embedding shards feeds the clustering
stage, which runs on every rank (identical inputs -> identical labels) or only where needed.
"""
import math
import time

import numpy as np
import torch

FS = 16000


def chunk(st, ed, dur=1.5, step=0.75):
    """Sub-segment boundaries of one voiced region (infer_diarization.py:606-619)."""
    out = []
    if ed - st <= 0:
        return out
    s, made = st, False
    while s + dur < ed + step:
        out.append([s, min(s + dur, ed)])
        s += step
        made = True
    if not made:
        out.append([st, ed])
    return out


def circle_pad(x, target_len):
    """speakerlab/utils/utils.py:232-238 for 1-D tensors."""
    n = x.shape[0]
    if n >= target_len:
        return x
    reps = int(math.ceil(target_len / n))
    return x.repeat(reps)[:target_len]


def cut_windows(wav, chunks, fs=FS):
    """infer_diarization.py:624-627 on a resident waveform: slice every chunk, circle-pad to the
    longest and stack -> [N, L].  wav: 1-D tensor (any device); the gather runs on that device,
    so a waveform already in HBM never travels back to the host."""
    starts = torch.tensor([int(st * fs) for st, _ in chunks], dtype=torch.long)
    ends = torch.tensor([int(ed * fs) for _, ed in chunks], dtype=torch.long)
    lens = ends - starts
    L = int(lens.max())
    full = lens == L
    idx = starts.to(wav.device)[:, None] + torch.arange(L, device=wav.device)[None, :]
    idx = torch.minimum(idx, torch.tensor(wav.shape[0] - 1, device=wav.device))
    out = wav[idx]
    for i in torch.nonzero(~full).flatten().tolist():       # ragged tail windows: the reference's circle_pad
        out[i] = circle_pad(wav[int(starts[i]):int(ends[i])], L)
    return out


def window_table(chunks, fs=FS):
    """Sample ranges of the sub-segments, with the reference's truncating float -> int conversion
    (infer_diarization.py:624: ``wav[int(st * fs):int(ed * fs)]``) -> (starts int64, lens int32, longest)."""
    starts = np.array([int(st * fs) for st, _ in chunks], dtype=np.int64)
    ends = np.array([int(ed * fs) for _, ed in chunks], dtype=np.int64)
    lens = (ends - starts).astype(np.int32)
    return starts, lens, int(lens.max()) if len(chunks) else 0


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n items for this rank; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_embeddings(local, n_total, group=None):
    """all_gather the per-rank embedding shards (padded to equal rows) -> [n_total, E] on every rank."""
    import torch.distributed as dist
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    per = (n_total + world - 1) // world
    pad = torch.zeros((per, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        rows.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(rows, dim=0)


class Diarizer:
    """feature_extractor: b200spk.FBank; embedding_model: any b200spk network (on `device`);
    cluster: b200spk.SpectralCluster, AHCluster or CommonClustering (any callable taking the [N, E] embeddings)."""

    def __init__(self, feature_extractor, embedding_model, cluster, device="cuda:0", batchsize=8192,
                 seg_dur=1.5, seg_shift=0.75, group=None):
        self.fe, self.model, self.cluster = feature_extractor, embedding_model, cluster
        self.device = torch.device(device)
        self.batchsize, self.seg_dur, self.seg_shift = batchsize, seg_dur, seg_shift
        self.group = group
        self.fs = feature_extractor.sample_rate
        self.last = {}                 # wall-clock seconds per stage of the last call
        self.last_embeddings = None    # the gathered [N, E] embeddings of the last call (device)

    def subsegments(self, vad_segments):
        return [c for st, ed in vad_segments for c in chunk(st, ed, self.seg_dur, self.seg_shift)]

    def extract(self, wav_dev, chunks):
        """Embeddings of this rank's shard of the sub-segments -> [n_local, E] on the device."""
        import torch.distributed as dist
        rank, world = 0, 1
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        lo, hi = shard_range(len(chunks), rank, world)
        # the windows are never materialised: the fbank kernel gathers them (and circle-pads short tails) from the
        # resident recording; L is the longest window of the WHOLE recording, as in the reference's single stack
        starts, lens, L = window_table(chunks, self.fs)
        assert len(chunks) == 0 or int((starts + lens).max()) <= wav_dev.shape[0], "sub-segment outside the recording"
        starts_d = torch.from_numpy(starts[lo:hi]).to(wav_dev.device, non_blocking=True)
        lens_d = torch.from_numpy(lens[lo:hi]).to(wav_dev.device, non_blocking=True)
        outs = []
        with torch.no_grad():
            for st in range(0, hi - lo, self.batchsize):
                ed = min(hi - lo, st + self.batchsize)
                outs.append(self.model(self.fe.windows(wav_dev, starts_d[st:ed], lens_d[st:ed], L)))
        if not outs:
            return torch.zeros((0, self.model.embedding_size), device=self.device)
        return torch.cat(outs, dim=0)

    def __call__(self, wav, vad_segments=None, speaker_num=None):
        """wav: [T] or [1,T] float tensor/array in [-1, 1] scale, or int16 PCM (sent to the GPU as 2 bytes per
        sample and scaled by 1/32768 in the fbank kernel, fileio.py:115-117), at 16 kHz; vad_segments:
        [[st, ed], ...] in seconds (default: the whole recording).  Returns (chunks, labels)."""
        wav = torch.as_tensor(wav)
        if wav.dtype != torch.int16:
            wav = wav.to(torch.float32)
        if wav.dim() == 2:
            wav = wav[0]
        if vad_segments is None:
            vad_segments = [[0.0, wav.shape[0] / self.fs]]
        t0 = time.perf_counter()
        chunks = self.subsegments(vad_segments)
        wav_dev = wav.contiguous().to(self.device, non_blocking=True)
        local = self.extract(wav_dev, chunks)
        emb = gather_embeddings(local, len(chunks), self.group)
        torch.cuda.synchronize(self.device)          # stage boundary (the back end needs the embeddings anyway)
        t1 = time.perf_counter()
        kw = {} if speaker_num is None else {"speaker_num": speaker_num}
        labels = self.cluster(emb, **kw)
        t2 = time.perf_counter()
        self.last = {"h2d_extract_gather": t1 - t0, "cluster": t2 - t1}
        self.last.update({"cluster." + k: v for k, v in getattr(self.cluster, "last", {}).get("stages_s", {}).items()})
        self.last_embeddings = emb
        return chunks, labels
