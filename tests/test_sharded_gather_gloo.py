"""world_size-2 gloo test of ShardedBacktest.gather: per-rank rows of unequal length (the first rank has one return /
turnover row fewer than rebalance dates, :1132) come back in date order on every rank.  No GPU: the engine is not
touched, the rows are synthetic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from incorporating_different_sources_b200.sharding import ShardedBacktest, partition


def _worker(rank, world, port, n_dates, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sb = ShardedBacktest.__new__(ShardedBacktest)          # gather() needs the partition only
    sb.rank, sb.world = rank, world
    sb.counts = [hi - lo for lo, hi in partition(n_dates, world)]
    sb.lo, sb.hi = partition(n_dates, world)[rank]
    N = 5
    w = torch.arange(sb.lo, sb.hi, dtype=torch.float64)[:, None] * 10 + torch.arange(N, dtype=torch.float64)[None, :]
    first = 1 if rank == 0 else 0
    r = torch.arange(sb.lo + first, sb.hi, dtype=torch.float64) + 0.5
    gw, gr = sb.gather([w, r], dist)
    if rank == 1:
        np.savez(out_path, w=gw.numpy(), r=gr.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_gather_orders_unequal_rows_by_date(tmp_path):
    n_dates = 11
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "g.npz")
    mp.spawn(_worker, args=(2, port, n_dates, out), nprocs=2, join=True)
    z = np.load(out)
    assert np.array_equal(z["w"], np.arange(n_dates)[:, None] * 10.0 + np.arange(5)[None, :])
    assert np.array_equal(z["r"], np.arange(1, n_dates) + 0.5)
