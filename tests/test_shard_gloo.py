"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py -- channel partition, max-over-ranks of the
per-rank time, and the optional gather of the decimated outputs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from liquiddsp.shard import channel_range, gather_audio


def test_channel_range_partitions_exactly():
    for total in (0, 1, 7, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [channel_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert channel_range(65536, 3, 8) == (24576, 32768)
    with pytest.raises(ValueError):
        channel_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = channel_range(total, rank, world)
    # stand-in for a rank's demodulated block: value encodes (channel, sample)
    local = (torch.arange(lo, hi, dtype=torch.float32)[:, None] * 1000 + torch.arange(5, dtype=torch.float32)[None, :])
    t = torch.tensor([10.0 + rank], dtype=torch.float64)          # per-rank step time
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    g = gather_audio(local, dst=0)
    if rank == 0:
        q.put((float(t.item()), g.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_max_time_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port, total = _free_port(), 7
    ps = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in ps:
        p.start()
    tmax, g = q.get()
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 11.0                                            # slowest rank decides
    ref = np.arange(total, dtype=np.float32)[:, None] * 1000 + np.arange(5, dtype=np.float32)[None, :]
    assert g.shape == (total, 5) and np.array_equal(g, ref)       # channel order preserved across ranks
