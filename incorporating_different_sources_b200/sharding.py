"""Multi-GPU sharding of a backtest: one process per GPU, contiguous date ranges with a halo.

Weights at date d depend only on data <= d (``portfolio_calculations.py:964-983``), never on
earlier weights, so rebalance windows are independent units.  Each rank owns a
contiguous range of rebalance dates, keeps only the market rows that range needs resident (its own
days plus the ``rolling_window - 1`` preceding days and the intraday look-back), and the per-window
weights are all-gathered with ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU
tests).  There is no data-path collective besides that gather.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def partition(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [lo, hi) ranges; the first ``n_items % world`` ranks get one more."""
    base, extra = divmod(n_items, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


@dataclass
class Shard:
    rank: int
    world: int
    lo: int                   # first owned position in the list of rebalance dates
    hi: int                   # one past the last owned position
    d_indices: np.ndarray     # owned trade-date rows (global daily row numbers)
    day_lo: int               # first daily row that must be resident (halo included)
    day_hi: int               # one past the last resident daily row
    hf_lo: int                # first / one-past-last resident intraday row
    hf_hi: int


def make_shard(d_indices: Sequence[int], rolling_window: int, rank: int, world: int,
               hf_ts: Optional[np.ndarray] = None, dates: Optional[np.ndarray] = None,
               hf_lookback_days: int = 1) -> Shard:
    """The rows rank ``rank`` needs for its share of ``d_indices`` (sorted global daily rows)."""
    d_indices = np.asarray(d_indices, dtype=np.int64)
    lo, hi = partition(len(d_indices), world)[rank]
    mine = d_indices[lo:hi]
    if len(mine) == 0:
        return Shard(rank, world, lo, hi, mine, 0, 0, 0, 0)
    day_lo = int(mine.min()) - (rolling_window - 1)
    day_hi = int(mine.max()) + 1
    if day_lo < 0:
        raise ValueError("not enough history before the first rebalance date of this shard")
    h_lo = h_hi = 0
    if hf_ts is not None:
        day = np.timedelta64(1, "D")
        start = dates[int(mine.min())] - hf_lookback_days * day + day
        h_lo = int(np.searchsorted(hf_ts, start, side="right"))
        h_hi = int(np.searchsorted(hf_ts, dates[int(mine.max())] + day, side="right"))
    return Shard(rank, world, lo, hi, mine, day_lo, day_hi, h_lo, h_hi)


def gather_rows(local, counts: Sequence[int], dist=None, group=None):
    """All-gather per-window rows ([W_local, ...] torch tensors) from every rank, in rank order.

    ``counts[r]`` is rank r's number of windows; ranks with fewer rows are padded to the maximum so
    a single ``all_gather_into_tensor`` moves everything (NCCL needs equal sizes).
    """
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size(group)
    wmax = max(counts)
    pad = torch.zeros((wmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world, wmax) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    return torch.cat([out[r, : counts[r]] for r in range(world)], dim=0)


def run_sharded(d_indices: Sequence[int], rolling_window: int, compute: Callable[[Shard], "object"],
                hf_ts=None, dates=None, hf_lookback_days: int = 1, dist=None):
    """Evaluate ``compute(shard) -> tensor [W_local, N]`` on every rank and gather the full [W, N]."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    shard = make_shard(d_indices, rolling_window, rank, world, hf_ts, dates, hf_lookback_days)
    local = compute(shard)
    counts = [hi - lo for lo, hi in partition(len(d_indices), world)]
    return gather_rows(local, counts, dist)


def engine_compute(engine, mkt, spec, hf_lookback_days=None, outputs=("weights",)):
    """``compute`` callback for :func:`run_sharded` on the CUDA engine: upload the shard's slice of a
    :class:`SyntheticMarket` (halo included) and evaluate its windows."""
    from .engine import upload_synthetic
    from .windows import hf_lookback, plan_daily_windows

    conj = spec["weighting_strategy"].startswith("conjugate")

    def compute(shard: Shard):
        bars = len(mkt.hf_ts) // mkt.n_days
        # resident slice: whole days, so that intraday rows stay aligned with daily rows
        d0, d1 = shard.day_lo, shard.day_hi
        if conj:
            d0 = min(d0, shard.hf_lo // bars)
        row_off, hf_off = upload_synthetic(engine, mkt, day_slice=slice(d0, d1))
        batch = plan_daily_windows(spec, mkt.dates, shard.d_indices, mkt.hf_ts if conj else None,
                                   hf_lookback_days=hf_lookback_days, need_hf=conj,
                                   row_offset=row_off, hf_row_offset=hf_off)
        fn = engine.conjugate if conj else engine.jeffreys
        res = fn(batch, outputs=tuple(outputs), device_out=True)
        return res[outputs[0]]
    return compute
