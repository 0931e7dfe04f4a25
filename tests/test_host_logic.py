"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares, host-side window
planning is bit-exact with pandas, sharding covers every window exactly once."""
import os
import re

import numpy as np
import pandas as pd
import pytest

from incorporating_different_sources_b200.sharding import make_shard, partition
from incorporating_different_sources_b200.synthetic import generate_market
from incorporating_different_sources_b200.windows import cap_descending_order, ffill_rows, plan_daily_windows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from incorporating_different_sources_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "bayes_portfolio.h")).read()
    declared = set(re.findall(r"\b(bp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/bayes_portfolio.h but not exported"
    assert declared == set(_lib.EXPORTED), "ctypes binding list out of sync with the header"
    assert lib.bp_version() >= 100


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from incorporating_different_sources_b200.engine import BayesEngine, BayesPortfolioError
    with pytest.raises(BayesPortfolioError):
        BayesEngine(0, use_torch_stream=False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "incorporating_different_sources_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{fn} mentions the oracle"


def test_window_planning_matches_pandas():
    mkt = generate_market(4, 330, seed=3)
    spec = dict(weighting_strategy="conjugate_hf_epu_ew", size=4, risk_aversion=2, rolling_window=252,
                rolling_window_frequency="daily", mcm_scaling=3)
    d_idx = [251, 300, 329]
    b = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts)
    prices = pd.DataFrame(mkt.prices, index=pd.DatetimeIndex(mkt.dates))
    intr = pd.DataFrame(mkt.hf_prices, index=pd.DatetimeIndex(mkt.hf_ts))
    for i, d in enumerate(d_idx):
        ts = pd.Timestamp(mkt.dates[d])
        win = prices.loc[:ts].iloc[-252:]
        assert b.day_row[i] == d and win.index[-1] == ts
        gaps = win.index.to_series().diff().dt.days.dropna()
        assert b.span_days[i] == int(gaps.sum())
        start = ts - pd.Timedelta(days=1)
        sel = intr[(intr.index > start + pd.Timedelta(days=1)) & (intr.index <= ts + pd.Timedelta(days=1))]
        assert (b.hf_lo[i], b.hf_hi[i]) == (intr.index.get_loc(sel.index[0]), intr.index.get_loc(sel.index[-1]) + 1)
        assert b.hf_hi[i] - b.hf_lo[i] == 78
    assert (b.mcm_index, b.prior_weights, b.mcm_scaling, b.risk_aversion) == (1, 1, 3.0, 2.0)
    b7 = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
    assert list(b7.hf_hi - b7.hf_lo) == [390, 390, 390]
    with pytest.raises(ValueError):
        plan_daily_windows(spec, mkt.dates, [100], mkt.hf_ts)


def test_gap_assertion_matches_reference_rule():
    mkt = generate_market(3, 300, seed=3)
    dates = mkt.dates.copy()
    dates[200:] += np.timedelta64(20, "D")          # a 3-week hole
    spec = dict(weighting_strategy="jeffreys", size=3, risk_aversion=1, rolling_window=252,
                rolling_window_frequency="daily", mcm_scaling=None)
    with pytest.raises(AssertionError):
        plan_daily_windows(spec, dates, [299], need_hf=False)


def test_ffill_and_cap_order_match_pandas():
    src_d = np.array(["2020-01-01", "2020-01-03", "2020-01-08"], dtype="datetime64[ns]")
    tgt = np.array(["2019-12-31", "2020-01-01", "2020-01-02", "2020-01-09"], dtype="datetime64[ns]")
    s = pd.Series([1.0, 2.0, 3.0], index=pd.DatetimeIndex(src_d))
    ref = s.reindex(pd.DatetimeIndex(tgt), method="ffill").to_numpy()
    got = ffill_rows(tgt, src_d, np.array([1.0, 2.0, 3.0]))
    assert np.array_equal(np.isnan(ref), np.isnan(got)) and np.array_equal(ref[1:], got[1:])
    caps = np.array([5.0, 9.0, 1.0, 7.0, 3.0])
    ser = pd.Series(caps, index=list("abcde"))
    assert [ser.index[i] for i in cap_descending_order(caps, 3)] == list(ser.nlargest(3).index)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_covers_every_window_once(world):
    d_idx = np.arange(1007, 1007 + 4150)
    parts = partition(len(d_idx), world)
    assert parts[0][0] == 0 and parts[-1][1] == len(d_idx)
    assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    mkt_dates = np.arange(0, 6000).astype("datetime64[D]").astype("datetime64[ns]")
    for r in range(world):
        sh = make_shard(d_idx, 1008, r, world)
        assert sh.day_lo == sh.d_indices[0] - 1007 and sh.day_hi == sh.d_indices[-1] + 1


def test_upload_segment_planner_and_intraday_trim():
    """Pure host logic of the end-to-end pipeline: wave-aligned segment cuts and the intraday row range a batch reads."""
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows, plan_wave_fractions, trim_intraday
    mkt = generate_market(4, 400, seed=1, bars_per_day=6)
    spec = dict(weighting_strategy="conjugate_hf_vix_vw", size=4, risk_aversion=5, rolling_window=100,
                rolling_window_frequency="daily", mcm_scaling=1)
    d_idx = list(range(150, 400))
    b = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
    R = mkt.hf_prices.shape[0]
    fr = plan_wave_fractions(b.hf_hi, R, wave=60)
    # 250 windows, waves of 60: 4 full waves + 10 windows -> 5 segments, each ending at the last bar of its last window
    assert len(fr) == 5 and fr[-1] == 1.0 and all(x < y for x, y in zip(fr, fr[1:]))
    assert [int(round(f * R)) for f in fr] == [int(b.hf_hi[k - 1]) for k in (60, 120, 180, 240, 250)]
    # more waves than segments: several waves per segment, never more than max_segments
    fr = plan_wave_fractions(b.hf_hi, R, wave=10, max_segments=8)
    assert len(fr) <= 8 and fr[-1] == 1.0
    # optional: the remainder cut into several tail segments (measured slower on the C2 workload, default 1)
    hi = np.arange(1, 1461) * 10
    assert [int(round(f * 14600)) // 10 for f in plan_wave_fractions(hi, 14600, wave=400)] == [400, 800, 1200, 1460]
    fr = plan_wave_fractions(hi, 14600, wave=400, tail_split=2)
    assert [int(round(f * 14600)) // 10 for f in fr] == [400, 800, 1200, 1330, 1460]
    assert plan_wave_fractions(b.hf_hi, R, wave=200) is None                       # less than two waves
    assert plan_wave_fractions(b.hf_hi[::-1], R, wave=60) is None                  # not sorted by date
    # sub-batches along the segment boundaries: a partition of the windows in date order; the windows of sub-batch k
    # end inside segment k (so that it only waits for segments <= k)
    from incorporating_different_sources_b200.windows import split_batch_by_fractions
    fr = plan_wave_fractions(b.hf_hi, R, wave=60)
    subs = split_batch_by_fractions(b, fr, R)
    assert [i1 - i0 for i0, i1, _ in subs] == [60, 60, 60, 60, 10] and subs[0][0] == 0 and subs[-1][1] == 250
    for (i0, i1, sb), f in zip(subs, fr):
        assert sb.n_windows == i1 - i0 and np.array_equal(sb.hf_hi, b.hf_hi[i0:i1]) and np.array_equal(sb.day_row, b.day_row[i0:i1])
        assert sb.hf_hi.max() <= int(np.ceil(f * R)) and sb.rolling_window == b.rolling_window
    lo0, hi0 = b.hf_lo.copy(), b.hf_hi.copy()
    lo, hi = trim_intraday(b)
    assert lo == lo0.min() > 0 and hi == hi0.max() == R
    assert np.array_equal(b.hf_lo + lo, lo0) and np.array_equal(b.hf_hi + lo, hi0) and b.hf_lo.min() == 0
    # the trimmed rows are exactly the bars after (first trade date - 7 days + 1 day)
    first = mkt.dates[d_idx[0]] - np.timedelta64(6, "D")
    assert lo == int(np.searchsorted(mkt.hf_ts, first, side="right"))
