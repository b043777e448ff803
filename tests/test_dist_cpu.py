"""CPU: the multi-GPU host logic (sub-segment sharding + embedding all-gather) on a world_size-2
gloo group, and the window slicing / chunking mirrors."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200spk import diarize
from oracle import synth


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 4799, 4800):
        for world in (1, 2, 3, 8):
            spans = [diarize.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_chunk_and_windows_mirror_reference():
    assert diarize.chunk(0.0, 3600.0) == synth.chunk(0.0, 3600.0)
    assert diarize.chunk(1.0, 1.9) == synth.chunk(1.0, 1.9)
    wav, _ = synth.fm_meeting(7.3, 2, seed=3)
    ch = synth.chunk(0.0, 7.3)                     # last window is ragged -> circle_pad
    ref = synth.cut_windows(wav, ch)
    got = diarize.cut_windows(torch.from_numpy(wav), ch).numpy()
    assert got.shape == ref.shape and np.array_equal(got, ref)


def _worker(rank, world, port, n, e, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * e, dtype=torch.float32).view(n, e)
        lo, hi = diarize.shard_range(n, rank, world)
        out = diarize.gather_embeddings(full[lo:hi].clone(), n)
        q.put((rank, bool(torch.equal(out, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11, 4799])
def test_gather_embeddings_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
