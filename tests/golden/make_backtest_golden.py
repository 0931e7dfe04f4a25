"""Golden vectors of the LOOP level: the unmodified reference's ``backtest_portfolio``
(portfolio_calculations.py:1221-1238) on small seeded markets.  Run in the build container only."""
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


def spec(**kw):
    s = dict(weighting_strategy="conjugate_hf_vix_vw", size=8, risk_aversion=5, turnover_cost=15,
             rebalancing_frequency="daily", rolling_window=60, rolling_window_frequency="daily",
             mcm_scaling=1, display_name="Conjugate HF-VIX VW")
    s.update(kw)
    return s


CASES = [
    dict(name="bt_conj_daily_n8", market=dict(n_assets=8, n_days=90, seed=3001), spec=spec(), start=-20, end=-1),
    dict(name="bt_jeffreys_monthly_n6", market=dict(n_assets=6, n_days=110, seed=3002),
         spec=spec(weighting_strategy="jeffreys", size=6, rolling_window=40, rebalancing_frequency="monthly",
                   mcm_scaling=None, display_name="Jeffreys"), start=-55, end=-3),
    dict(name="bt_conj_epu_ew_weekly_top6of10", market=dict(n_assets=10, n_days=100, seed=3003),
         spec=spec(weighting_strategy="conjugate_hf_epu_ew", size=6, rolling_window=50, rebalancing_frequency="weekly",
                   risk_aversion=3, turnover_cost=5, display_name="Conjugate HF-EPU EW"), start=-30, end=-1),
    dict(name="bt_conj_weeklywin_monthly_n7", market=dict(n_assets=7, n_days=330, seed=3005),
         spec=spec(size=7, rolling_window=40, rolling_window_frequency="weekly", rebalancing_frequency="monthly"),
         start=-70, end=-2),
    dict(name="bt_jorion_weekly_top6of9", market=dict(n_assets=9, n_days=120, seed=3006),
         spec=spec(weighting_strategy="jorion", size=6, rolling_window=45, rebalancing_frequency="weekly",
                   risk_aversion=4, mcm_scaling=None, display_name="Jorion"), start=-40, end=-2),
    dict(name="bt_vw_daily_top5of9", market=dict(n_assets=9, n_days=60, seed=3004),
         spec=spec(weighting_strategy="vw", size=5, risk_aversion=None, mcm_scaling=None, rolling_window=20,
                   display_name="VW"), start=-15, end=-1),
    # late listing: the stock with the largest cap has NaN prices until row 60 -> not eligible (NaN inside the 30-day
    # window) until row 89, then it enters the top 5: NaN prices in a NON-HELD column during the first dates, a universe
    # that changes inside the backtest
    dict(name="bt_jeffreys_daily_top5of8_late_listing", market=dict(n_assets=8, n_days=100, seed=3007),
         spec=spec(weighting_strategy="jeffreys", size=5, rolling_window=30, mcm_scaling=None, display_name="Jeffreys"),
         start=-25, end=-1, damage=dict(kind="late_listing_of_largest_cap", until_row=60)),
]


def apply_damage(md, damage):
    """The same edit on the reference's and on the repo's market_data (tests/test_gpu_backtest.py imports this)."""
    if not damage:
        return md
    if damage["kind"] == "late_listing_of_largest_cap":
        caps = md["stock_market_caps_df"]
        col = caps.columns[int(np.argmax(caps.iloc[-1].to_numpy()))]
        md = dict(md)
        for key in ("stock_prices_df", "stock_simple_returns_df", "stock_log_returns_df"):
            df = md[key].copy()
            df.iloc[:damage["until_row"], df.columns.get_loc(col)] = np.nan
            md[key] = df
        return md
    raise KeyError(damage["kind"])


def main():
    pc = load_reference(check=True)
    only = sys.argv[1:]
    for case in CASES:
        if only and case["name"] not in only:
            continue
        mkt = generate_market(**case["market"])
        set_universe(mkt.tickers)
        md = apply_damage(mkt.market_data(), case.get("damage"))
        d0, d1 = pd.Timestamp(mkt.dates[mkt.n_days + case["start"]]), pd.Timestamp(mkt.dates[mkt.n_days + case["end"]])
        res = pc.backtest_portfolio(case["spec"], d0, d1, md)
        r, t, m = (res["portfolio_simple_returns_series"], res["portfolio_turnover_series"],
                   res["portfolio_weights_metrics_df"])
        meta = dict(name=case["name"], market=case["market"], spec=case["spec"], start=str(d0.date()), end=str(d1.date()),
                    metrics_columns=list(m.columns), series_name=r.name, damage=case.get("damage"))
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), meta=np.array(json.dumps(meta)),
                            returns=r.to_numpy(), returns_idx=r.index.values.astype("int64"),
                            turnover=t.to_numpy(), turnover_idx=t.index.values.astype("int64"),
                            metrics=m.to_numpy(), metrics_idx=m.index.values.astype("int64"))
        print(case["name"], len(r), len(t), m.shape, float(r.sum()))


if __name__ == "__main__":
    main()
