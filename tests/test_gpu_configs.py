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
