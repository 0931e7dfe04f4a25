"""Sibling estimators on the CUDA path, through the C-ABI (``bp_estimator_batched``).

* Jorion (:851-895) against the UNMODIFIED reference's outputs (tests/golden/est_*.npz, ``jorion_*`` keys) and,
  batched, against the oracle; tolerance 1e-9 relative as for the Bayesian path.
* Shrinkage (:703-758): the closed form of the reference's own CHECK block with sklearn's ``ledoit_wolf``
  (``lw_*`` keys; pypfopt / cvxpy are not installed, so the reference's rounded solver output is not available).
"""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import bayes_oracle as bo
from tests._golden import estimator_golden_names, load_golden, market_for, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9
ANNUAL = {"daily": 252, "weekly": 52}


@pytest.fixture(scope="module")
def engine():
    from incorporating_different_sources_b200.engine import BayesEngine
    eng = BayesEngine(0)
    yield eng
    eng.close()


def _plan(engine, mkt, spec, d_idx):
    from incorporating_different_sources_b200.windows import plan_daily_windows, plan_weekly_windows
    if spec["rolling_window_frequency"] == "daily":
        return plan_daily_windows(spec, mkt.dates, d_idx, need_hf=False)
    rows, batch = plan_weekly_windows(spec, mkt.dates, d_idx, mkt.dates, mkt.rf, None, need_hf=False)
    engine.set_resampled(rows)
    return batch


@pytest.mark.parametrize("name", estimator_golden_names())
def test_gpu_estimators_match_golden(engine, name):
    from incorporating_different_sources_b200._lib import SCAL_JORION, SCAL_LW
    from incorporating_different_sources_b200.engine import upload_synthetic
    z, meta = load_golden(name)
    mkt = market_for(meta)
    spec = meta["spec"]
    m = spec["rolling_window"] - 1
    f = ANNUAL[spec["rolling_window_frequency"]]
    for wi, w in enumerate(meta["windows"]):
        pre = f"w{wi}_"
        cols = z[pre + "cols"]
        upload_synthetic(engine, mkt, cols=cols)
        batch = _plan(engine, mkt, spec, [w["d_idx"]])
        got = engine.jorion(batch, outputs=("weights", "scalars", "status"))
        assert got["status"][0] == 0
        assert relerr(got["weights"][0], z[pre + "jorion_weights"]) <= TOL
        s = got["scalars"][0]
        for k, g in (("mu_g", "jorion_mu_g"), ("lambda_hat", "jorion_lambda"), ("v_hat", "jorion_v")):
            assert abs(s[SCAL_JORION[k]] - float(z[pre + g])) <= TOL * abs(float(z[pre + g])), k
        got = engine.shrinkage(batch, outputs=("weights", "scalars", "status", "S1"))
        assert got["status"][0] == 0
        assert abs(got["scalars"][0][SCAL_LW["shrinkage"]] - float(z[pre + "lw_shrinkage"])) <= TOL * float(z[pre + "lw_shrinkage"])
        cov = got["S1"][0] * (1.0 - got["scalars"][0][SCAL_LW["shrinkage"]]) / m * f     # the device factors m Sigma_LW / (1 - shrinkage)
        assert relerr(np.diag(cov), z[pre + "lw_cov_diag"]) <= TOL
        assert relerr(cov[0], z[pre + "lw_cov_row0"]) <= TOL
        assert relerr(got["weights"][0], z[pre + "lw_weights"]) <= TOL


@pytest.mark.parametrize("n_assets,rolling_window,n_windows", [(10, 252, 40), (50, 252, 33), (130, 252, 12), (200, 400, 35)])
def test_gpu_estimators_batched_match_oracle(engine, n_assets, rolling_window, n_windows):
    """Many overlapping windows in one call (reuse of block Gram tiles, banded prep) vs the oracle window by window."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    D = rolling_window + 50
    mkt = generate_market(n_assets, D, seed=6000 + n_assets, bars_per_day=2)
    spec = dict(weighting_strategy="jorion", size=n_assets, risk_aversion=4, turnover_cost=15,
                rebalancing_frequency="daily", rolling_window=rolling_window, rolling_window_frequency="daily",
                mcm_scaling=None, display_name="x")
    d_idx = list(range(D - n_windows, D))
    cols = np.arange(n_assets)
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, need_hf=False)
    gj = engine.jorion(batch, outputs=("weights", "status"))
    gs = engine.shrinkage(batch, outputs=("weights", "status"))
    assert not gj["status"].any() and not gs["status"].any()
    for i, d in enumerate(d_idx):
        assert relerr(gj["weights"][i], bo.jorion_window(spec, mkt, d, cols)["weights"]) <= TOL, f"jorion window {i}"
        assert relerr(gs["weights"][i], bo.shrinkage_window(spec, mkt, d, cols)["weights"]) <= TOL, f"shrinkage window {i}"


def test_jorion_after_jeffreys_does_not_disturb_it(engine):
    """The estimator batches reuse the Jeffreys machinery with another centring coefficient: a Jeffreys batch
    before and after must give bit-identical results."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(40, 320, seed=77, bars_per_day=2)
    spec = dict(weighting_strategy="jeffreys", size=40, risk_aversion=5, turnover_cost=15, rebalancing_frequency="daily",
                rolling_window=252, rolling_window_frequency="daily", mcm_scaling=None, display_name="x")
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, list(range(270, 320)), need_hf=False)
    a = engine.jeffreys(batch, outputs=("weights",))["weights"].copy()
    engine.jorion(batch, outputs=("weights",))
    engine.shrinkage(batch, outputs=("weights",))
    b = engine.jeffreys(batch, outputs=("weights",))["weights"]
    assert np.array_equal(a, b)


def test_jorion_rejects_short_windows(engine):
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(30, 80, seed=5, bars_per_day=2)
    spec = dict(weighting_strategy="jorion", size=30, risk_aversion=5, turnover_cost=15, rebalancing_frequency="daily",
                rolling_window=33, rolling_window_frequency="daily", mcm_scaling=None, display_name="x")   # T - N - 2 = 0
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, [79], need_hf=False)
    with pytest.raises(ValueError):
        engine.jorion(batch)


def test_facade_jorion_and_shrinkage_frames():
    """Reference signatures and containers: index named 'Stock' (:744, :893), column 'Weight', the caller's column
    order; shrinkage returns clean_weights() (cutoff 1e-4, 5 decimals, :743)."""
    from incorporating_different_sources_b200 import portfolio_calculations as pc
    z, meta = load_golden("est_n12of30_topk")
    mkt = market_for(meta)
    md = mkt.market_data()
    spec = meta["spec"]
    d = pd.Timestamp(mkt.dates[meta["windows"][0]["d_idx"]])
    names = [mkt.tickers[c] for c in z["w0_cols"]]
    prices_df = md["stock_prices_df"][names].loc[:d]
    rf_df = md["risk_free_rate_df"]
    before = prices_df.copy()
    w = pc.calculate_jorion_portfolio(spec, d, prices_df, rf_df)
    assert list(w.index) == names and w.index.name == "Stock" and list(w.columns) == ["Weight"]
    assert relerr(w["Weight"].to_numpy(), z["w0_jorion_weights"]) <= TOL
    assert prices_df.equals(before)                                   # inputs are never mutated
    full = pc.calculate_portfolio_weights(d, spec, md)                 # dispatcher route (:1036-1040)
    assert list(full.index) == names and relerr(full["Weight"].to_numpy(), z["w0_jorion_weights"]) <= TOL
    s = pc.calculate_shrinkage_portfolio(dict(spec, weighting_strategy="shrinkage"), d, prices_df, rf_df)
    assert list(s.index) == names and s.index.name == "Stock"
    assert np.array_equal(s["Weight"].to_numpy(), bo.clean_weights(s["Weight"].to_numpy()))
    assert np.max(np.abs(s["Weight"].to_numpy() - z["w0_lw_weights"])) <= 0.5e-5 + 1e-12
    raw = pc.calculate_shrinkage_portfolio(dict(spec, weighting_strategy="shrinkage"), d, prices_df, rf_df, clean=False)
    assert relerr(raw["Weight"].to_numpy(), z["w0_lw_weights"]) <= TOL
    with pytest.raises(ValueError):                                    # last row must be the trading date (:145)
        pc.calculate_jorion_portfolio(spec, d - pd.Timedelta(days=1), prices_df, rf_df)


@pytest.mark.parametrize("n_assets,rolling_window,n_windows", [(20, 80, 45), (75, 200, 26)])
def test_jorion_chain_matches_per_window_path(engine, n_assets, rolling_window, n_windows):
    """Jorion batches of consecutive dates factorise every 8th window and obtain C^-1 t and C^-1 1 of the others by the
    Woodbury identity (jeffreys_chain.cu): same weights and Bayes-Stein scalars as the per-window path (1e-10)."""
    from incorporating_different_sources_b200._lib import SCAL_JORION
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    D = rolling_window + n_windows + 2
    mkt = generate_market(n_assets, D, seed=6500 + n_assets, bars_per_day=2)
    spec = dict(weighting_strategy="jorion", size=n_assets, risk_aversion=2, turnover_cost=15,
                rebalancing_frequency="daily", rolling_window=rolling_window, rolling_window_frequency="daily",
                mcm_scaling=None, display_name="x")
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, list(range(D - n_windows, D)), need_hf=False)
    try:
        engine.set_jeffreys_chain(0)
        plain = engine.jorion(batch, outputs=("weights", "w1", "scalars", "status"))
        engine.set_jeffreys_chain(8)
        engine.solve_work()
        got = engine.jorion(batch, outputs=("weights", "w1", "scalars", "status"))
        work = engine.solve_work()
    finally:
        engine.set_jeffreys_chain(8)
    assert work["chained"] == n_windows - (-(-n_windows // 8))
    assert not got["status"].any() and not plain["status"].any()
    assert relerr(got["weights"], plain["weights"]) <= 1e-10 and relerr(got["w1"], plain["w1"]) <= 1e-10
    for k in ("mu_g", "lambda_hat", "v_hat"):
        i = SCAL_JORION[k]
        assert np.max(np.abs(got["scalars"][:, i] - plain["scalars"][:, i]) / np.abs(plain["scalars"][:, i])) <= 1e-9, k
