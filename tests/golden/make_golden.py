"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case below the seeded synthetic market is generated, the reference's own
functions (``src/portfolio_calculations.py``) are called with ``CHECK=True`` on frames
sliced exactly as its dispatcher slices them (:954-983), and every posterior moment
named in SURVEY.md §8(a) is stored: ``t, T, n0, n1, S0, S1, w0, c, w1, v1, nu,
weights`` plus the output index order.  Matrices of the N=500 cases are stored as
(diagonal, three full rows, Frobenius norm) to keep the fixtures small.

The HF look-back D is the reference's table {daily:1, weekly:7, monthly:31}
(:299-304).  Cases with ``hf_lookback_days`` different from the window frequency's
own D obtain ``S0`` from the reference's ``calculate_conjugate_prior_S`` called with a
spec copy whose ``rolling_window_frequency`` selects that D, and inject it through
the reference's ``conjugate_prior_S_df=`` parameters — the injection point SURVEY F5
names.
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
D_TO_FREQ = {1: "daily", 7: "weekly", 31: "monthly"}


def base_spec(**kw):
    s = dict(weighting_strategy="conjugate_hf_vix_vw", size=10, risk_aversion=5, turnover_cost=15,
             rebalancing_frequency="daily", rolling_window=252, rolling_window_frequency="daily",
             mcm_scaling=1, display_name="Conjugate HF-VIX VW")
    s.update(kw)
    return s


CASES = [
    # name, market kwargs, spec, date offsets from the end, hf_lookback_days (None = reference table)
    dict(name="c1_n10_conj_vix_vw", market=dict(n_assets=10, n_days=300, seed=1001),
         spec=base_spec(), dates=[-1, -7, -20], hf_days=None),
    dict(name="n10_conj_epu_ew", market=dict(n_assets=10, n_days=300, seed=1002),
         spec=base_spec(weighting_strategy="conjugate_hf_epu_ew", risk_aversion=3, mcm_scaling=5,
                        display_name="Conjugate HF-EPU EW"), dates=[-1, -11], hf_days=None),
    dict(name="n12of30_conj_vix_vw_topk", market=dict(n_assets=30, n_days=290, seed=1003),
         spec=base_spec(size=12), dates=[-1, -15], hf_days=None),
    dict(name="n25_weekly_conj_vix_vw", market=dict(n_assets=25, n_days=420, seed=1004),
         spec=base_spec(size=25, rolling_window=60, rolling_window_frequency="weekly",
                        rebalancing_frequency="weekly"), dates=[-1, -3, -9], hf_days=None),
    dict(name="n50_conj_vix_vw_hf7", market=dict(n_assets=50, n_days=300, seed=1005),
         spec=base_spec(size=50), dates=[-1, -30], hf_days=7),
    dict(name="n100_conj_epu_vw_hf31", market=dict(n_assets=100, n_days=300, seed=1006),
         spec=base_spec(size=100, weighting_strategy="conjugate_hf_epu_vw", mcm_scaling=1,
                        display_name="Conjugate HF-EPU VW"), dates=[-1], hf_days=31),
    dict(name="n10_jeffreys", market=dict(n_assets=10, n_days=300, seed=1007),
         spec=base_spec(weighting_strategy="jeffreys", mcm_scaling=None, display_name="Jeffreys"),
         dates=[-1, -13], hf_days=None),
    dict(name="n50_jeffreys", market=dict(n_assets=50, n_days=300, seed=1008),
         spec=base_spec(size=50, weighting_strategy="jeffreys", mcm_scaling=None, display_name="Jeffreys"),
         dates=[-1], hf_days=None),
    dict(name="n25_weekly_jeffreys", market=dict(n_assets=25, n_days=420, seed=1009),
         spec=base_spec(size=25, weighting_strategy="jeffreys", mcm_scaling=None, rolling_window=60,
                        rolling_window_frequency="weekly", display_name="Jeffreys"), dates=[-2], hf_days=None),
    dict(name="n10_conj_const_mcm", market=dict(n_assets=10, n_days=280, seed=1010, mcm_mode="constant",
                                                rf_mode="constant"),
         spec=base_spec(), dates=[-1], hf_days=None),
    dict(name="c2_n500_conj_vix_vw_hf7", market=dict(n_assets=500, n_days=262, seed=2002),
         spec=base_spec(size=500), dates=[-1, -5], hf_days=7),
    dict(name="c2_n500_jeffreys_n1008", market=dict(n_assets=500, n_days=1012, seed=2003, bars_per_day=2),
         spec=base_spec(size=500, weighting_strategy="jeffreys", mcm_scaling=None, rolling_window=1008,
                        display_name="Jeffreys"), dates=[-1, -3], hf_days=None),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def compact(name, M, out):
    """Store a matrix fully (N<=100) or as diag / 3 rows / Frobenius norm."""
    M = np.asarray(M, dtype=np.float64)
    if M.shape[0] <= 100:
        out[name] = M
    else:
        out[name + "_diag"] = np.diag(M).copy()
        out[name + "_rows"] = M[[0, M.shape[0] // 2, M.shape[0] - 1]].copy()
        out[name + "_fro"] = np.array(np.linalg.norm(M))


def run_case(pc, case):
    mkt = generate_market(**case["market"])
    md = mkt.market_data()
    set_universe(mkt.tickers)
    spec = case["spec"]
    out = {}
    meta = dict(name=case["name"], market=case["market"], spec=spec, hf_days=case["hf_days"],
                sha_prices=sha(mkt.prices), sha_hf=sha(mkt.hf_prices), sha_caps=sha(mkt.caps),
                numpy=np.__version__, pandas=pd.__version__, windows=[])
    is_conj = spec["weighting_strategy"].startswith("conjugate")
    for wi, off in enumerate(case["dates"]):
        d_idx = mkt.n_days + off
        d = pd.Timestamp(mkt.dates[d_idx])
        # --- slicing exactly as calculate_portfolio_weights :954-983
        kcaps = pc.get_k_largest_stocks_market_caps(
            md["stock_market_caps_df"], md["stock_prices_df"], md["stock_intraday_prices_df"], d,
            spec["size"], pc.get_window_trading_days(spec), spec["rebalancing_frequency"])
        names = list(kcaps.index)
        caps_df = md["stock_market_caps_df"][names].loc[:d]
        prices_df = md["stock_prices_df"][names].loc[:d]
        d_incl = d.replace(hour=23, minute=59, second=59)
        intr_df = md["stock_intraday_prices_df"][names]
        intr_df = intr_df.loc[intr_df.index <= d_incl]
        rf_df = md["risk_free_rate_df"]
        pre = f"w{wi}_"
        out[pre + "cols"] = np.array([mkt.tickers.index(nm) for nm in names], dtype=np.int64)
        t_df = pc.calculate_canonical_statistics_t(spec, d, prices_df, rf_df)
        T_df = pc.calculate_canonical_statistics_T(spec, d, prices_df, rf_df)
        assert list(t_df.index) == names and list(T_df.index) == names
        out[pre + "t"] = t_df.values[:, 0].copy()
        compact(pre + "T", T_df.values, out)
        if is_conj:
            mcm_key = "vix_prices_df" if "vix" in spec["weighting_strategy"] else "epu_prices_df"
            mcm_df = md[mcm_key].loc[md[mcm_key].index <= d]
            n0 = pc.calculate_conjugate_prior_n(spec, d, mcm_df)
            n1 = pc.calculate_conjugate_posterior_n(spec, d, mcm_df)
            hf_days = case["hf_days"]
            if hf_days is None:
                S0_df = pc.calculate_conjugate_prior_S(spec, d, intr_df, mcm_df)
            else:
                spec_d = dict(spec, rolling_window_frequency=D_TO_FREQ[hf_days])
                S0_df = pc.calculate_conjugate_prior_S(spec_d, d, intr_df, mcm_df, conjugate_prior_n=n0)
            w0_df = pc.calculate_conjugate_prior_w(spec, d, prices_df, caps_df, mcm_df)
            c = pc.calculate_conjugate_c(spec, d, prices_df, caps_df, intr_df, mcm_df,
                                         conjugate_prior_S_df=S0_df)
            S1_df = pc.calculate_conjugate_posterior_S(spec, d, prices_df, intr_df, mcm_df, rf_df,
                                                       conjugate_prior_S_df=S0_df)
            w1_df = pc.calculate_conjugate_posterior_w(spec, d, prices_df, caps_df, intr_df, mcm_df, rf_df,
                                                       conjugate_c=c, conjugate_prior_S_df=S0_df,
                                                       conjugate_posterior_S_df=S1_df)
            v1 = pc.calculate_portfolio_variance(w1_df, S1_df)
            nu_df = pc.calculate_mean_conjugate_posterior_nu(
                spec, d, prices_df, caps_df, intr_df, mcm_df, rf_df, conjugate_c=c,
                conjugate_prior_S_df=S0_df, conjugate_posterior_S_df=S1_df)
            if hf_days is None:
                # the un-injected public entry point must agree with the staged computation
                w_df = pc.calculate_conjugate_hf_mcm_portfolio(spec, d, caps_df, prices_df, intr_df, mcm_df, rf_df)
                full = pc.calculate_portfolio_weights(d, spec, md)
                assert np.array_equal(full.values, w_df.values) and list(full.index) == names
            else:
                w_df = 1 / spec["risk_aversion"] * nu_df
            assert list(w0_df.index) == names and list(w1_df.index) == names and list(w_df.index) == names
            out[pre + "n0"] = np.array(n0)
            out[pre + "n1"] = np.array(n1)
            out[pre + "c"] = np.array(c)
            out[pre + "v1"] = np.array(v1)
            out[pre + "w0"] = w0_df["Weight"].values.copy()
            out[pre + "w1"] = w1_df["Weight"].values.copy()
            out[pre + "nu"] = nu_df["Weight"].values.copy()
            compact(pre + "S0", S0_df.loc[names, names].values, out)
            compact(pre + "S1", S1_df.loc[names, names].values, out)
            out[pre + "cond_S1"] = np.array(np.linalg.cond(S1_df.values))
        else:
            nu_df = pc.calculate_mean_jeffreys_posterior_nu(spec, d, prices_df, rf_df)
            w_df = pc.calculate_jeffreys_portfolio(spec, d, prices_df, rf_df)
            full = pc.calculate_portfolio_weights(d, spec, md)
            assert np.array_equal(full.values, w_df.values) and list(full.index) == names
            out[pre + "nu"] = nu_df["Weight"].values.copy()
            J = T_df.values - 1 / spec["rolling_window"] * np.outer(out[pre + "t"], out[pre + "t"])
            out[pre + "cond_J"] = np.array(np.linalg.cond(J))
        out[pre + "weights"] = w_df["Weight"].values.copy()
        meta["windows"].append(dict(d_idx=int(d_idx), date=str(d.date()), order=names[:5] + ["..."]))
        print(f"  {case['name']} {d.date()} |w|max={np.abs(out[pre + 'weights']).max():.4g} "
              + " ".join(f"{k[len(pre):]}={float(out[k]):.3g}" for k in out if k.startswith(pre + "cond")))
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
