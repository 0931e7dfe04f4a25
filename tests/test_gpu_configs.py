"""BASELINE.json configs 2-4 as parity cases (the bench line is config 1):
 * config 2 (HF-heavy): N=100, 252-day intraday window (19,655 five-minute returns per rebalance) injected
   through the HF look-back override, the reference's ``conjugate_prior_S_df=`` route;
 * config 3 (strategy sweep): N in {5,10,25,50,100} x {conjugate with constant MCM, conjugate+VIX,
   conjugate+EPU, Jeffreys};
 * config 4 shape (independent paths): two seeds evaluated back to back on one engine.
All against the CPU oracle, 1e-9 relative."""
import numpy as np
import pytest

from oracle import bayes_oracle as bo
from tests._golden import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def engine():
    from incorporating_different_sources_b200.engine import BayesEngine
    eng = BayesEngine(0)
    yield eng
    eng.close()


def _spec(strategy, n, **kw):
    s = dict(weighting_strategy=strategy, size=n, risk_aversion=5, turnover_cost=15, rebalancing_frequency="daily",
             rolling_window=252, rolling_window_frequency="daily", mcm_scaling=1, display_name=strategy)
    s.update(kw)
    return s


def test_config2_hf_heavy_window(engine):
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    n = 100
    mkt = generate_market(n, 262, seed=3)
    spec = _spec("conjugate_hf_vix_vw", n)
    d_idx = [259, 261]
    look = 366          # calendar days >= 252 trading days of 78 bars
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=look)
    m = batch.hf_hi - batch.hf_lo - 1
    assert m.min() >= 252 * 78 - 1
    got = engine.conjugate(batch, outputs=("weights", "w1", "scalars", "status", "S0"))
    assert not got["status"].any()
    cols = np.arange(n)
    for i, d in enumerate(d_idx):
        ref = bo.conjugate_window(spec, mkt, d, cols, hf_lookback_days=look)
        assert ref["hf_returns"] == m[i]
        assert relerr(got["S0"][i], ref["S0"]) <= TOL
        assert relerr(got["w1"][i], ref["w1"]) <= TOL
        assert relerr(got["weights"][i], ref["weights"]) <= TOL
        assert abs(got["scalars"][i][4] - ref["c"]) <= TOL * abs(ref["c"])


@pytest.mark.parametrize("n", [5, 10, 25, 50, 100])
def test_config3_strategy_sweep(engine, n):
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    d_idx = [262, 270, 279]
    cols = np.arange(n)
    for strategy, mcm_mode in (("conjugate_hf_vix_vw", "constant"), ("conjugate_hf_vix_vw", "varying"),
                               ("conjugate_hf_epu_vw", "varying"), ("jeffreys", "varying")):
        mkt = generate_market(n, 280, seed=500 + n, mcm_mode=mcm_mode)
        spec = _spec(strategy, n)
        upload_synthetic(engine, mkt)
        if strategy == "jeffreys":
            batch = plan_daily_windows(spec, mkt.dates, d_idx, need_hf=False)
            got = engine.jeffreys(batch, outputs=("weights", "status"))
            refs = [bo.jeffreys_window(spec, mkt, d, cols)["weights"] for d in d_idx]
        else:
            hf_days = 1 if n <= 50 else 7       # m - 1 + n - 1 >= N with margin (SURVEY 8(d) well-posed grid)
            batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=hf_days)
            got = engine.conjugate(batch, outputs=("weights", "scalars", "status"))
            refs = []
            for i, d in enumerate(d_idx):
                r = bo.conjugate_window(spec, mkt, d, cols, hf_lookback_days=hf_days)
                refs.append(r["weights"])
                if mcm_mode == "constant":
                    assert got["scalars"][i][0] == spec["rolling_window"] * spec["mcm_scaling"]   # f == 1 exactly
        assert not got["status"].any()
        for i in range(len(d_idx)):
            assert relerr(got["weights"][i], refs[i]) <= TOL, (strategy, mcm_mode, n, i)


def test_config4_independent_paths(engine):
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    n = 64
    spec = _spec("conjugate_hf_epu_ew", n, rolling_window=100)
    cols = np.arange(n)
    d_idx = list(range(110, 130))
    outs = []
    for path in (0, 1):
        mkt = generate_market(n, 130, seed=1000 * path + 2)
        upload_synthetic(engine, mkt)
        batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
        got = engine.conjugate(batch, outputs=("weights", "status"))
        assert not got["status"].any()
        ref = np.asarray([bo.conjugate_window(spec, mkt, d, cols, hf_lookback_days=7)["weights"] for d in d_idx])
        assert relerr(got["weights"], ref) <= TOL
        outs.append(got["weights"])
    assert not np.allclose(outs[0], outs[1])


@pytest.mark.parametrize("n_windows", [3, 40])
def test_wide_universe_beyond_512_columns(engine, n_windows):
    """N = 520 (> 512 columns: two column passes in the streaming prep, 5 x 5 Gram tiles, 17 Cholesky panels),
    conjugate with a 7-day intraday look-back and Jeffreys with n = 640, single windows and a batch of consecutive
    dates (block reuse, banded-GEMM prep, inner day blocks)."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    n = 520
    mkt = generate_market(n, 700, seed=41)
    upload_synthetic(engine, mkt)
    d_idx = list(range(700 - n_windows, 700))
    cols = np.arange(n)
    check = d_idx if n_windows <= 3 else [d_idx[0], d_idx[17], d_idx[-1]]
    cspec = _spec("conjugate_hf_vix_vw", n, rolling_window=252)
    cb = plan_daily_windows(cspec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
    got = engine.conjugate(cb, outputs=("weights", "scalars", "status"))
    assert not got["status"].any()
    for d in check:
        ref = bo.conjugate_window(cspec, mkt, d, cols, hf_lookback_days=7)
        assert relerr(got["weights"][d_idx.index(d)], ref["weights"]) <= TOL
    jspec = _spec("jeffreys", n, rolling_window=640)
    jb = plan_daily_windows(jspec, mkt.dates, d_idx, need_hf=False)
    gotj = engine.jeffreys(jb, outputs=("weights", "status"))
    assert not gotj["status"].any()
    for d in check:
        ref = bo.jeffreys_window(jspec, mkt, d, cols)
        assert relerr(gotj["weights"][d_idx.index(d)], ref["weights"]) <= TOL


def _upload(engine, mkt):
    """Upload a market whose intraday calendar need not be regular (upload_synthetic assumes 78 bars every day)."""
    from incorporating_different_sources_b200.windows import ffill_rows
    engine.upload_market(prices=mkt.prices, rf_row=ffill_rows(mkt.dates, mkt.dates, mkt.rf), caps=mkt.caps,
                         hf_prices=mkt.hf_prices, mcm=np.stack([mkt.vix, mkt.epu]))


def test_config2_batched_long_lookback_presummed_day_blocks(engine):
    """BASELINE config 3 on the BATCHED path: N=100, 64 consecutive rebalance dates, 366-calendar-day intraday
    look-back (>= 252 trading days, ~20k five-minute returns per window).  The day blocks are pre-summed (suffix /
    prefix scans), so a window adds <= 3 intraday tiles + 1 daily tile per output tile instead of ~260; every 8th
    window against the oracle, all windows against the one-tile-per-day path of the same library."""
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    n, W, look = 100, 64, 366
    mkt = generate_market(n, 270 + W, seed=33)
    spec = _spec("conjugate_hf_vix_vw", n)
    d_idx = list(range(270, 270 + W))
    cols = np.arange(n)
    _upload(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=look)
    assert (batch.hf_hi - batch.hf_lo - 1).min() >= 252 * 78 - 1
    outs = ("weights", "w1", "rhs", "scalars", "status")
    try:
        engine.set_hf_presum_min_days(0)
        engine.gram_work()
        plain = engine.conjugate(batch, outputs=outs)
        work_plain = engine.gram_work()
        engine.set_hf_presum_min_days(8)
        got = engine.conjugate(batch, outputs=outs)
        work = engine.gram_work()
    finally:
        engine.set_hf_presum_min_days(8)
    assert work_plain["add_blocks"] / W > 200          # one tile per trading day of the look-back
    assert work["add_blocks"] / W <= 4.0                # suffix' + whole chunk + prefix + the daily run tile
    assert work["k_rows"] < work_plain["k_rows"]        # the overnight rows are inside the scanned tiles
    assert not got["status"].any() and not plain["status"].any()
    for k in ("weights", "w1", "rhs"):
        assert relerr(got[k], plain[k]) <= 1e-11, k
    assert np.max(np.abs(got["scalars"] - plain["scalars"]) / np.maximum(np.abs(plain["scalars"]), 1e-300)) <= 1e-11
    for i in range(0, W, 8):
        ref = bo.conjugate_window(spec, mkt, d_idx[i], cols, hf_lookback_days=look)
        assert relerr(got["w1"][i], ref["w1"]) <= TOL
        assert relerr(got["weights"][i], ref["weights"]) <= TOL
        assert abs(got["scalars"][i][4] - ref["c"]) <= TOL * abs(ref["c"])
        assert abs(got["scalars"][i][0] - ref["n0"]) <= TOL * abs(ref["n0"])
    # the moment-only entry points take the same route (S0 through the Gram epilogue, rhs / c / v0 through the
    # mat-vec by-product) -- calculate_conjugate_prior_S / calculate_conjugate_c of the reference (:285-333, :382-430)
    sub = plan_daily_windows(spec, mkt.dates, d_idx[:40], mkt.hf_ts, hf_lookback_days=look)
    n0, S0 = engine.hf_cov(sub)
    mom = engine.moments(sub, outputs=("rhs", "scalars", "w0"))
    for i in (0, 17, 39):
        ref = bo.conjugate_window(spec, mkt, d_idx[i], cols, hf_lookback_days=look)
        assert relerr(S0[i], ref["S0"]) <= TOL
        assert abs(n0[i] - ref["n0"]) <= TOL * ref["n0"]
        assert abs(mom["scalars"][i][5] - ref["v0"]) <= TOL * abs(ref["v0"])
        assert abs(mom["scalars"][i][4] - ref["c"]) <= TOL * abs(ref["c"])
        assert relerr(mom["rhs"][i], ref["c"] * (ref["S0"] @ ref["w0"]) + ref["t"]) <= TOL


@pytest.mark.parametrize("look", [7, 31])
def test_irregular_intraday_calendar_day_blocks(engine, look):
    """Half days, missing bars and a day with a single surviving morning: the day blocks follow the hf_lo / hf_hi
    boundaries of the windows, not a fixed stride.  7 calendar days = 5 blocks per window (one tile per day + the
    gathered overnight rows); 31 days = 21..23 blocks (pre-summed runs of varying length)."""
    import dataclasses
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    n, W, nwin = 24, 48, 60
    base = generate_market(n, 40 + nwin + W, seed=77, bars_per_day=26)
    rng = np.random.default_rng(5)
    keep = np.ones(len(base.hf_ts), dtype=bool)
    B = 26
    for day in rng.choice(base.n_days, size=12, replace=False):
        keep[day * B + 13: (day + 1) * B] = False             # half days
    keep[rng.choice(len(keep), size=60, replace=False)] = False   # isolated missing bars
    keep[(base.n_days - 20) * B + 9: (base.n_days - 19) * B] = False   # a day cut down to 9 bars
    mkt = dataclasses.replace(base, hf_ts=base.hf_ts[keep], hf_prices=np.ascontiguousarray(base.hf_prices[keep]))
    spec = _spec("conjugate_hf_epu_vw", n, rolling_window=nwin)
    d_idx = list(range(mkt.n_days - W, mkt.n_days))
    cols = np.arange(n)
    _upload(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=look)
    rows = batch.hf_hi - batch.hf_lo
    assert len(set(rows.tolist())) > 3                         # the windows really differ in length
    engine.gram_work()
    got = engine.conjugate(batch, outputs=("weights", "scalars", "status"))
    work = engine.gram_work()
    assert not got["status"].any()
    assert work["add_blocks"] / W <= (4.0 if look == 31 else 6.0)    # block reuse engaged on the irregular calendar
    for i in range(W):
        ref = bo.conjugate_window(spec, mkt, d_idx[i], cols, hf_lookback_days=look)
        assert ref["hf_returns"] == rows[i] - 1
        assert relerr(got["weights"][i], ref["weights"]) <= TOL, i
        assert abs(got["scalars"][i][5] - ref["v0"]) <= TOL * abs(ref["v0"])


def test_headline_shapes_c2_conjugate_and_chained_jeffreys(engine):
    """The exact shapes of the bench line (BASELINE config 2) at a size the oracle can follow: N=500, 64 consecutive
    conjugate windows (n=252, 7-day look-back: inner day blocks, banded prep, block reuse) and 64 consecutive
    Jeffreys windows with n=1008 (every 8th factorised, the others chained); every 8th window against the oracle."""
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    n, W = 500, 64
    cols = np.arange(n)
    mkt = generate_market(n, 256 + W, seed=2)
    _upload(engine, mkt)
    cspec = _spec("conjugate_hf_vix_vw", n)
    d_idx = list(range(256, 256 + W))
    cb = plan_daily_windows(cspec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
    got = engine.conjugate(cb, outputs=("weights", "status"))
    assert not got["status"].any()
    for i in range(0, W, 8):
        ref = bo.conjugate_window(cspec, mkt, d_idx[i], cols, hf_lookback_days=7)
        assert relerr(got["weights"][i], ref["weights"]) <= TOL, i
    del mkt
    mktj = generate_market(n, 1010 + W, seed=2002, bars_per_day=2)
    _upload(engine, mktj)
    jspec = _spec("jeffreys", n, rolling_window=1008, mcm_scaling=None)
    dj = list(range(1010, 1010 + W))
    jb = plan_daily_windows(jspec, mktj.dates, dj, need_hf=False)
    engine.solve_work()
    gotj = engine.jeffreys(jb, outputs=("weights", "status"))
    assert engine.solve_work() == {"factored": 8.0, "chained": 56.0}
    assert not gotj["status"].any()
    for i in list(range(0, W, 8)) + [1, 7, 63]:
        ref = bo.jeffreys_window(jspec, mktj, dj[i], cols)
        assert relerr(gotj["weights"][i], ref["weights"]) <= TOL, i
