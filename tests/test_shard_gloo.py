"""CPU suite, part 3: the multi-GPU host logic on two gloo ranks (no GPU needed).

Batched transforms shard by contiguous batch ranges with no exchange step. Each rank asks the
library for its range (fftb200_shard_range), transforms only that slice (here with the CPU oracle standing in for the
per-device plan), and the max-over-ranks timing reduction and result gathering that bench.py performs are exercised.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, batch, outdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    import fftb200_loader
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F = fftb200_loader.load()
    first, count = F.shard_range(batch, world, rank)
    p = O.port()
    x = p.fill(43, first * n, count * n).reshape(count, n)     # rank-local slice of the global stream
    y = p.fft_batch(x, -1) if count else x
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)   # stand-in for the rank's device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([first, count]))
    np.save(os.path.join(outdir, "y%d.npy" % rank), y)
    if rank == 0:
        np.save(os.path.join(outdir, "meta.npy"), np.array([t.item()] + [int(v) for s in sizes for v in s]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [7, 64])
def test_batch_sharding_two_ranks(tmp_path, port, batch):
    import torch.multiprocessing as mp
    n, world = 256, 2
    mp.spawn(_worker, args=(world, _free_port(), n, batch, str(tmp_path)), nprocs=world, join=True)
    meta = np.load(tmp_path / "meta.npy")
    assert meta[0] == world                      # max over ranks
    ranges = meta[1:].reshape(world, 2).astype(int)
    assert ranges[0][0] == 0 and ranges[-1].sum() == batch
    assert all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))   # contiguous, disjoint
    y = np.concatenate([np.load(tmp_path / ("y%d.npy" % r)) for r in range(world)])
    x = port.fill(43, 0, n * batch).reshape(batch, n)
    assert np.array_equal(y, port.fft_batch(x, -1))


def test_shard_range_properties(F):
    for batch in (0, 1, 5, 8, 65536, 10 ** 9 + 7):
        for world in (1, 2, 3, 4, 8):
            got = [F.shard_range(batch, world, r) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == batch
            assert all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        F.shard_range(8, 2, 2)
    with pytest.raises(ValueError):
        F.shard_range(8, 0, 0)
