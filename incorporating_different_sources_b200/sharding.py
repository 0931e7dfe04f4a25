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


class ShardedBacktest:
    """One rank's share of ONE backtest split by contiguous date range (the north-star partition): conjugate and
    Jeffreys weights of its rebalance dates plus the loop body (daily returns, turnover, :1127-1219) of its range.

    Resident rows = the rank's own trading days, the ``rolling_window - 1`` daily rows before them (halo of the longer
    of the two windows) and the intraday bars of its dates' look-backs.  The loop body at the first own date needs
    the weights chosen at the previous rebalance date (:1134-1159, :1212), so every rank but the first also
    evaluates that ONE predecessor window (the one-window halo of SURVEY 8(e)) instead of receiving it.
    ``gather()`` all-gathers ``[2][W_local][N]`` weights and the ``[W_local]`` return / turnover rows — the only
    communication of the path."""

    def __init__(self, engine, mkt, conj_spec, jeff_spec, d_indices, rank: int, world: int,
                 hf_lookback_days: Optional[int] = None, pin=None):
        from .windows import ffill_rows, plan_daily_windows, trim_intraday
        self.engine, self.rank, self.world = engine, rank, world
        d_all = np.asarray(d_indices, dtype=np.int64)
        if len(d_all) < world or np.any(np.diff(d_all) != 1):
            raise ValueError("ShardedBacktest splits a DAILY-rebalance backtest: consecutive trading dates, >= 1 per rank")
        self.counts = [hi - lo for lo, hi in partition(len(d_all), world)]
        lo, hi = partition(len(d_all), world)[rank]
        self.lo, self.hi = lo, hi
        self.halo = 1 if lo > 0 else 0
        ext = d_all[lo - self.halo:hi]                       # predecessor window first (ranks > 0)
        self.n_ext = len(ext)
        n_max = max(int(conj_spec["rolling_window"]), int(jeff_spec["rolling_window"]))
        day_lo, day_hi = int(ext.min()) - (n_max - 1), int(ext.max()) + 1
        if day_lo < 0:
            raise ValueError("not enough history before the first rebalance date of this shard")
        dates = mkt.dates[day_lo:day_hi]
        self.cb = plan_daily_windows(conj_spec, dates, ext - day_lo, mkt.hf_ts, hf_lookback_days=hf_lookback_days)
        self.jb = plan_daily_windows(jeff_spec, dates, ext - day_lo, need_hf=False)
        h_lo, h_hi = trim_intraday(self.cb)                  # rows of mkt.hf_prices this rank's windows read
        rf_dates = getattr(mkt, "rf_dates", mkt.dates)
        pin = pin or (lambda a: np.ascontiguousarray(a))
        self.host = dict(prices=pin(mkt.prices[day_lo:day_hi]), caps=pin(mkt.caps[day_lo:day_hi]),
                         hf_prices=pin(mkt.hf_prices[h_lo:h_hi]),
                         mcm=pin(np.stack([mkt.vix[day_lo:day_hi], mkt.epu[day_lo:day_hi]])),
                         rf_row=pin(ffill_rows(dates, rf_dates, mkt.rf)))
        self.h2d_bytes = int(sum(v.nbytes for v in self.host.values()))
        self.reb_rows = (ext - day_lo).astype(np.int32)
        self.gamma_c = float(conj_spec["risk_aversion"])
        self.gamma_j = float(jeff_spec["risk_aversion"])
        self.cost = float(conj_spec["turnover_cost"])
        self.n_assets = mkt.n_assets
        self._fractions = None

    def upload(self, async_copy: bool = False):
        """Host slice -> HBM.  ``async_copy`` (page-locked host arrays, see ``pin``): the intraday block travels in
        wave-aligned segments on the copy stream, so that :meth:`compute` overlaps it (Jeffreys does not read intraday
        data; the conjugate stages run on the segments that have arrived)."""
        if async_copy:
            if self._fractions is None:
                self._fractions = self.engine.plan_upload_fractions(self.cb, self.host["hf_prices"].shape[0])
            self.engine.set_upload_fractions(self._fractions)
        self.engine.upload_market(**self.host, async_copy=async_copy)
        if async_copy:
            self.engine.set_upload_fractions(None)

    def compute(self, out_c, out_j, loop: bool = True):
        """Weights of the shard's windows into ``out_c`` / ``out_j`` (dicts with 'weights' [n_ext][N] and 'status'
        CUDA tensors) and, with ``loop``, the loop body of both strategies.  Returns the local rows to gather:
        (weights_c, weights_j, returns_c, returns_j, turnover_c, turnover_j) without the halo window."""
        eng = self.engine
        eng.jeffreys(self.jb, outputs=("weights", "status"), into=out_j)      # first: needs no intraday data
        eng.conjugate(self.cb, outputs=("weights", "status"), into=out_c)
        k = self.halo
        res = [out_c["weights"][k:], out_j["weights"][k:]]
        if loop:
            for w, g in ((out_c["weights"], self.gamma_c), (out_j["weights"], self.gamma_j)):
                r, to, _ = eng.backtest_loop(self.reb_rows, w, distance_scale=g, turnover_cost_bps=self.cost,
                                             device_out=True)
                res += [r, to]
            res = [res[0], res[1], res[2], res[4], res[3], res[5]]
        return res

    def gather(self, rows, dist, group=None):
        """All-gather per-date rows of every rank in date order with ONE collective: the rows of a rank (weights of both
        priors [W_local][N], returns and turnover [W_local] or, on the first rank, one row fewer -- the backtest's first
        date has no return, :1132) are packed side by side into one [W_max][columns] buffer, gathered, and unpacked."""
        import torch
        if dist is None:
            import torch.distributed as dist
        world = dist.get_world_size(group)
        wmax = max(self.counts)
        own = self.hi - self.lo
        widths = [1 if x.dim() == 1 else int(x.shape[1]) for x in rows]
        short = [x.dim() == 1 and x.shape[0] == own - (1 if self.rank == 0 else 0) for x in rows]   # return-like rows
        ref = rows[0]
        pack = torch.zeros((wmax, sum(widths)), dtype=ref.dtype, device=ref.device)
        c0 = 0
        for x, wd, sh in zip(rows, widths, short):
            first = 1 if (sh and self.rank == 0) else 0              # the first rank's rows start at its second date
            view = x if x.dim() == 2 else x[:, None]
            pack[first:first + x.shape[0], c0:c0 + wd] = view
            c0 += wd
        out = torch.empty((world, wmax, sum(widths)), dtype=ref.dtype, device=ref.device)
        dist.all_gather_into_tensor(out.view(-1), pack.view(-1), group=group)
        res, c0 = [], 0
        for x, wd, sh in zip(rows, widths, short):
            parts = [out[r, (1 if (sh and r == 0) else 0):self.counts[r], c0:c0 + wd] for r in range(world)]
            full = torch.cat(parts, dim=0)
            res.append(full if x.dim() == 2 else full[:, 0].contiguous())
            c0 += wd
        return res
