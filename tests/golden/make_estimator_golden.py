"""Golden vectors for the sibling estimators under tests/golden/est_*.npz.

Run in the build container only (needs /root/reference):

    python tests/golden/make_estimator_golden.py

* Jorion: the UNMODIFIED reference ``calculate_jorion_portfolio`` (:851-895) and, through the dispatcher,
  ``calculate_portfolio_weights`` with ``weighting_strategy="jorion"`` on frames sliced as at :954-983.
* Shrinkage: ``calculate_shrinkage_portfolio`` (:703-758) cannot run here (pypfopt / cvxpy are not installed).  The
  vectors are the closed form of its own CHECK block (:748-756) evaluated on the REFERENCE's excess returns
  (``adjust_stock_prices_window`` + ``calculate_excess_log_returns_from_prices``, unmodified) with the installed
  ``sklearn.covariance.ledoit_wolf`` — the routine pypfopt 1.5.5 ``CovarianceShrinkage.ledoit_wolf()`` calls — times the
  reference's annualisation factor (:116-124), ``np.linalg.inv`` and ``np.dot`` as at :750-753.  Parity of this
  estimator against the reference's rounded solver output is therefore UNPINNED; what is pinned is the closed form.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from incorporating_different_sources_b200.synthetic import generate_market  # noqa: E402
from oracle.ref_import import load_reference, set_universe  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def base_spec(**kw):
    s = dict(weighting_strategy="jorion", size=10, risk_aversion=5, turnover_cost=15,
             rebalancing_frequency="daily", rolling_window=252, rolling_window_frequency="daily",
             mcm_scaling=None, display_name="Jorion")
    s.update(kw)
    return s


CASES = [
    dict(name="est_n10_daily", market=dict(n_assets=10, n_days=300, seed=3001), spec=base_spec(), dates=[-1, -9]),
    dict(name="est_n50_daily", market=dict(n_assets=50, n_days=300, seed=3002), spec=base_spec(size=50, risk_aversion=3),
         dates=[-1, -21]),
    dict(name="est_n12of30_topk", market=dict(n_assets=30, n_days=290, seed=3003), spec=base_spec(size=12), dates=[-1]),
    dict(name="est_n25_weekly", market=dict(n_assets=25, n_days=420, seed=3004),
         spec=base_spec(size=25, rolling_window=60, rolling_window_frequency="weekly",
                        rebalancing_frequency="weekly"), dates=[-1, -4]),
    dict(name="est_n500_n1008", market=dict(n_assets=500, n_days=1012, seed=3005, bars_per_day=2),
         spec=base_spec(size=500, rolling_window=1008), dates=[-1, -3]),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run_case(pc, case):
    from sklearn.covariance import ledoit_wolf
    import sklearn
    mkt = generate_market(**case["market"])
    md = mkt.market_data()
    set_universe(mkt.tickers)
    spec = case["spec"]
    out = {}
    meta = dict(name=case["name"], market=case["market"], spec=spec, sha_prices=sha(mkt.prices),
                numpy=np.__version__, pandas=pd.__version__, sklearn=sklearn.__version__, windows=[])
    for wi, off in enumerate(case["dates"]):
        d_idx = mkt.n_days + off
        d = pd.Timestamp(mkt.dates[d_idx])
        kcaps = pc.get_k_largest_stocks_market_caps(
            md["stock_market_caps_df"], md["stock_prices_df"], md["stock_intraday_prices_df"], d,
            spec["size"], pc.get_window_trading_days(spec), spec["rebalancing_frequency"])
        names = list(kcaps.index)
        prices_df = md["stock_prices_df"][names].loc[:d]
        rf_df = md["risk_free_rate_df"]
        pre = f"w{wi}_"
        out[pre + "cols"] = np.array([mkt.tickers.index(nm) for nm in names], dtype=np.int64)
        # ---- Jorion: the reference itself
        w_df = pc.calculate_jorion_portfolio(spec, d, prices_df, rf_df)
        full = pc.calculate_portfolio_weights(d, spec, md)
        assert np.array_equal(full.values, w_df.values) and list(full.index) == names and list(w_df.index) == names
        assert w_df.index.name == "Stock" and list(w_df.columns) == ["Weight"]
        out[pre + "jorion_weights"] = w_df["Weight"].values.copy()
        # intermediate scalars, recomputed with the reference's own expressions on the reference's returns
        win = pc.adjust_stock_prices_window(spec, d, prices_df)
        X_df = pc.calculate_excess_log_returns_from_prices(spec, win, rf_df)
        N, T = len(prices_df.columns), len(X_df)
        mu_hat = X_df.mean().to_frame()
        V_bar = T / (T - N - 2) * X_df.cov()
        V_bar_inv = pd.DataFrame(np.linalg.inv(V_bar.to_numpy()), index=V_bar.index, columns=V_bar.columns)
        one = pd.DataFrame(np.ones(N), index=V_bar_inv.index)
        mu_g = (one.T.dot(V_bar_inv).dot(mu_hat) / one.T.dot(V_bar_inv).dot(one)).values[0, 0]
        diff = mu_hat.sub(mu_g * one.values, axis=0)
        q = diff.T.dot(V_bar_inv).dot(diff).values[0, 0]
        out[pre + "jorion_mu_g"] = np.array(mu_g)
        out[pre + "jorion_lambda"] = np.array((N + 2) / q)
        out[pre + "jorion_v"] = np.array((N + 2) / ((N + 2) + T * q))
        out[pre + "cond_V"] = np.array(np.linalg.cond(V_bar.to_numpy()))
        # ---- shrinkage: closed form of the reference's CHECK (:748-756) with sklearn's ledoit_wolf
        f = pc.get_window_annualization_factor(spec)
        mean = X_df.mean() * f                                       # mean_historical_return(returns_data, compounding=False)
        cov, shrink = ledoit_wolf(np.nan_to_num(X_df.dropna(how="all").values))
        cov = cov * f
        w_lw = 1 / spec["risk_aversion"] * np.dot(np.linalg.inv(cov), mean.values)
        out[pre + "lw_shrinkage"] = np.array(shrink)
        out[pre + "lw_weights"] = w_lw
        out[pre + "lw_cov_diag"] = np.diag(cov).copy()
        out[pre + "lw_cov_row0"] = cov[0].copy()
        out[pre + "cond_LW"] = np.array(np.linalg.cond(cov))
        meta["windows"].append(dict(d_idx=int(d_idx), date=str(d.date())))
        print(f"  {case['name']} {d.date()} jorion |w|max={np.abs(w_df.values).max():.4g} cond(V)={float(out[pre + 'cond_V']):.3g} "
              f"lw shrink={shrink:.4f} |w|max={np.abs(w_lw).max():.4g} cond={float(out[pre + 'cond_LW']):.3g}")
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **out)


def main():
    pc = load_reference(check=True)
    only = sys.argv[1:]
    for case in CASES:
        if only and case["name"] not in only:
            continue
        print(case["name"])
        run_case(pc, case)


if __name__ == "__main__":
    main()
