"""TEST / BASELINE INFRASTRUCTURE — drives the UNMODIFIED reference (``oracle/ref_import.py``) one window at a
time on frames cut from a :class:`SyntheticMarket`, the way SURVEY §8(c) prescribes ("minimal pre-sliced frames":
bit-identical to full-history frames, probe there).  Used by ``bench.py --impl reference`` / ``cpu_baseline``
(the CPU arm: the reference's own NumPy/pandas path timed on the GPU box's host cores) and by the golden makers.
Never imported by the product (``incorporating_different_sources_b200/``).

Levels:
  L0     weight-function level, one process, NumPy's default BLAS threads, CHECK=False
  L0-mp  the same, date-sharded over P worker processes x 1 BLAS thread (windows are independent, :964-983)
  L1     loop level: ``backtest_portfolio`` as shipped (CHECK=True) on the full ``market_data`` dict
"""
from __future__ import annotations

import time

import numpy as np
import pandas as pd

from .ref_import import load_reference, set_universe

D_TO_FREQ = {1: "daily", 7: "weekly", 31: "monthly"}     # portfolio_calculations.py:299-304


def _frames(mkt, d_idx, n, cols, hf_rows=None):
    """prices[-n:], caps[-1:], intraday rows of the HF window, MCM[-n:], full risk-free frame."""
    names = [mkt.tickers[c] for c in cols]
    lo = d_idx - n + 1
    didx = pd.DatetimeIndex(mkt.dates[lo:d_idx + 1])
    prices = pd.DataFrame(mkt.prices[lo:d_idx + 1][:, cols], index=didx, columns=names)
    caps = pd.DataFrame(mkt.caps[d_idx:d_idx + 1][:, cols], index=didx[-1:], columns=names)
    intr = None
    if hf_rows is not None:
        a, b = hf_rows
        intr = pd.DataFrame(mkt.hf_prices[a:b][:, cols], index=pd.DatetimeIndex(mkt.hf_ts[a:b]), columns=names)
    rf = pd.DataFrame({"DTB3": mkt.rf}, index=pd.DatetimeIndex(mkt.dates))
    return names, prices, caps, intr, rf, didx


def conjugate_window(pc, spec, mkt, d_idx, cols, hf_lookback_days=None):
    """``calculate_conjugate_hf_mcm_portfolio`` (:819-836) of the unmodified reference for one date.  A look-back
    other than the window frequency's own goes through the reference's ``conjugate_prior_S_df=`` injection (F5):
    1 / 7 / 31 days via ``calculate_conjugate_prior_S`` with a spec copy selecting that look-back; any other length
    (BASELINE config 3's 252 trading days) with the reference's own expressions ``cov(h) * len(h) * n0``
    (:314-318, :333) evaluated by pandas on the longer row range."""
    n = int(spec["rolling_window"])
    d = pd.Timestamp(mkt.dates[d_idx])
    own = D_TO_FREQ_INV[spec["rolling_window_frequency"]]
    D = own if hf_lookback_days is None else int(hf_lookback_days)
    day = np.timedelta64(1, "D")
    a = int(np.searchsorted(mkt.hf_ts, mkt.dates[d_idx] - D * day + day, side="right"))
    b = int(np.searchsorted(mkt.hf_ts, mkt.dates[d_idx] + day, side="right"))
    names, prices, caps, intr, rf, didx = _frames(mkt, d_idx, n, cols, (a, b))
    key = mkt.vix if "vix" in spec["weighting_strategy"] else mkt.epu
    mcm = pd.DataFrame({"MCM": key[d_idx - n + 1:d_idx + 1]}, index=didx)
    if D == own:
        return pc.calculate_conjugate_hf_mcm_portfolio(spec, d, caps, prices, intr, mcm, rf)
    n0 = pc.calculate_conjugate_prior_n(spec, d, mcm)
    if D in D_TO_FREQ:
        S0 = pc.calculate_conjugate_prior_S(dict(spec, rolling_window_frequency=D_TO_FREQ[D]), d, intr, mcm,
                                            conjugate_prior_n=n0)
    else:
        h = np.log(intr / intr.shift(1)).dropna()                 # :314
        S0 = h.cov() * len(h) * n0                                # :317-318, :333
    # calculate_mean_conjugate_posterior_nu computes c WITHOUT forwarding an injected prior S (:517-523), so c is
    # computed with the injection first, as tests/golden/make_golden.py does
    c = pc.calculate_conjugate_c(spec, d, prices, caps, intr, mcm, conjugate_prior_n=n0, conjugate_prior_S_df=S0)
    nu = pc.calculate_mean_conjugate_posterior_nu(spec, d, prices, caps, intr, mcm, rf, conjugate_c=c,
                                                  conjugate_prior_n=n0, conjugate_prior_S_df=S0)
    return 1 / spec["risk_aversion"] * nu                         # :836


D_TO_FREQ_INV = {"daily": 1, "weekly": 7, "monthly": 31}


def jeffreys_window(pc, spec, mkt, d_idx, cols):
    """``calculate_jeffreys_portfolio`` (:838-849) of the unmodified reference for one date."""
    n = int(spec["rolling_window"])
    d = pd.Timestamp(mkt.dates[d_idx])
    names, prices, caps, intr, rf, didx = _frames(mkt, d_idx, n, cols)
    return pc.calculate_jeffreys_portfolio(spec, d, prices, rf)


def loop_level(mkt, spec, d_indices, check=True):
    """L1: the reference's ``backtest_portfolio`` (:1221-1238) as shipped over consecutive trading dates
    ``d_indices`` on the FULL market_data dict.  Returns (seconds, rebalances)."""
    pc = load_reference(check=check)
    set_universe(mkt.tickers)
    md = mkt.market_data()
    d0, d1 = pd.Timestamp(mkt.dates[d_indices[0]]), pd.Timestamp(mkt.dates[d_indices[-1]])
    t0 = time.perf_counter()
    res = pc.backtest_portfolio(spec, d0, d1, md)
    dt = time.perf_counter() - t0
    return dt, len(res["portfolio_weights_metrics_df"])
