"""world_size-2 gloo test of the sharded path (host logic + gather), runnable without a GPU.

Each rank evaluates its contiguous range of rebalance dates on ITS SLICE of the market (own days plus
halo) — here with the CPU oracle standing in for the CUDA engine — and the weights are all-gathered.
The result must equal the unsharded evaluation bit for bit, which proves the halo bookkeeping
(row offsets, intraday look-back) and the gather order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from incorporating_different_sources_b200.sharding import run_sharded
from incorporating_different_sources_b200.synthetic import SyntheticMarket, generate_market
from oracle import bayes_oracle as bo

SPEC = dict(weighting_strategy="conjugate_hf_vix_vw", size=6, risk_aversion=5, rolling_window=60,
            rolling_window_frequency="daily", mcm_scaling=1)
HF_DAYS = 7


def _slice_market(mkt, d0, d1):
    bars = len(mkt.hf_ts) // mkt.n_days
    return SyntheticMarket(tickers=mkt.tickers, dates=mkt.dates[d0:d1], prices=mkt.prices[d0:d1],
                           caps=mkt.caps[d0:d1], hf_ts=mkt.hf_ts[d0 * bars:d1 * bars],
                           hf_prices=mkt.hf_prices[d0 * bars:d1 * bars], vix=mkt.vix[d0:d1], epu=mkt.epu[d0:d1],
                           rf=mkt.rf[d0:d1], sp500=mkt.sp500[d0:d1])


def _oracle_compute(mkt):
    def compute(shard):
        bars = len(mkt.hf_ts) // mkt.n_days
        d0 = min(shard.day_lo, shard.hf_lo // bars)
        sub = _slice_market(mkt, d0, shard.day_hi)            # only the shard's rows are visible
        cols = np.arange(mkt.n_assets)
        rows = [bo.conjugate_window(SPEC, sub, int(d) - d0, cols, hf_lookback_days=HF_DAYS)["weights"]
                for d in shard.d_indices]
        return torch.from_numpy(np.asarray(rows).reshape(len(rows), mkt.n_assets))
    return compute


def _worker(rank, world, port, d_idx, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mkt = generate_market(6, 100, seed=11)
    full = run_sharded(d_idx, SPEC["rolling_window"], _oracle_compute(mkt), hf_ts=mkt.hf_ts, dates=mkt.dates,
                       hf_lookback_days=HF_DAYS)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_dates", [11, 12])
def test_two_rank_sharded_backtest_equals_serial(tmp_path, n_dates):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    d_idx = list(range(100 - n_dates, 100))
    out = str(tmp_path / "w.npy")
    mp.spawn(_worker, args=(2, port, d_idx, out), nprocs=2, join=True)
    got = np.load(out)
    mkt = generate_market(6, 100, seed=11)
    cols = np.arange(6)
    ref = np.asarray([bo.conjugate_window(SPEC, mkt, d, cols, hf_lookback_days=HF_DAYS)["weights"] for d in d_idx])
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
