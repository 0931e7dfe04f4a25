"""Seeded synthetic market generator.

The reference loads real Alpha Vantage / FMP / Yahoo data through
``src/data_handling.py:270-291`` (``get_market_data``); none of that is reachable
offline, so every parity test, golden fixture and benchmark is fed by this
generator instead.  It emits the same ten containers with the same index types:

* daily frames on a business-day ``DatetimeIndex`` (``stock_prices_df``,
  ``stock_market_caps_df``, VIX / EPU / S&P / DTB3 single-column frames);
* a 5-minute intraday frame with 78 bars per day, 09:35 ... 16:00
  (``stock_intraday_prices_df``), whose last bar of each day *is* the daily close.

Everything is float64, strictly positive, NaN-free and the market caps have no
ties, so the cap-descending asset order (``portfolio_calculations.py:654``) is
unambiguous.  The arrays are plain NumPy; :meth:`SyntheticMarket.market_data`
wraps them in pandas objects for the reference / the drop-in facade.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

BARS_PER_DAY = 78            # 09:35 ... 16:00 every 5 minutes
_NS_PER_DAY = 86_400_000_000_000
_NS_PER_MIN = 60_000_000_000
_FIRST_BAR_MIN = 9 * 60 + 35


def business_days(start: str, periods: int) -> np.ndarray:
    """``pd.bdate_range(start, periods=periods)`` without pandas: Mon-Fri, no holidays."""
    d0 = np.datetime64(start, "D")
    # numpy busday_offset rolls forward to the first business day, then steps.
    first = np.busday_offset(d0, 0, roll="forward")
    days = np.busday_offset(first, np.arange(periods))
    return days.astype("datetime64[ns]")


def make_tickers(n: int) -> List[str]:
    width = max(3, len(str(n - 1)))
    return [f"S{str(i).zfill(width)}" for i in range(n)]


@dataclass
class SyntheticMarket:
    tickers: List[str]
    dates: np.ndarray          # (D,) datetime64[ns], business days
    prices: np.ndarray         # (D, N) daily close
    caps: np.ndarray           # (D, N) market capitalisation
    hf_ts: np.ndarray          # (D*78,) datetime64[ns]
    hf_prices: np.ndarray      # (D*78, N) 5-minute prices
    vix: np.ndarray            # (D,)
    epu: np.ndarray            # (D,)
    rf: np.ndarray             # (D,) annualised risk-free rate, decimal (DTB3/100)
    sp500: np.ndarray          # (D,)
    seed: int = 0
    meta: Dict[str, object] = field(default_factory=dict)

    @property
    def n_assets(self) -> int:
        return self.prices.shape[1]

    @property
    def n_days(self) -> int:
        return self.prices.shape[0]

    def date_index_of(self, date) -> int:
        d = np.datetime64(date, "ns")
        i = int(np.searchsorted(self.dates, d))
        if i >= len(self.dates) or self.dates[i] != d:
            raise KeyError(f"{date} is not a trading date of this market")
        return i

    def market_data(self, columns: Optional[List[int]] = None):
        """The 10-key dict of ``data_handling.py:282-291`` as pandas objects."""
        import pandas as pd

        cols = list(range(self.n_assets)) if columns is None else list(columns)
        names = [self.tickers[c] for c in cols]
        didx = pd.DatetimeIndex(self.dates)
        prices = pd.DataFrame(self.prices[:, cols], index=didx, columns=names)
        caps = pd.DataFrame(self.caps[:, cols], index=didx, columns=names)
        intraday = pd.DataFrame(self.hf_prices[:, cols], index=pd.DatetimeIndex(self.hf_ts), columns=names)
        vix = pd.DataFrame({"VIX": self.vix}, index=didx)
        epu = pd.DataFrame({"EPU": self.epu}, index=didx)
        sp = pd.DataFrame({"S&P 500": self.sp500}, index=didx)
        rf = pd.DataFrame({"DTB3": self.rf}, index=didx)
        return {
            "stock_prices_df": prices,
            "stock_simple_returns_df": prices.pct_change(),
            "stock_log_returns_df": np.log(prices / prices.shift(1)),
            "stock_intraday_prices_df": intraday,
            "stock_market_caps_df": caps,
            "vix_prices_df": vix,
            "epu_prices_df": epu,
            "sp500_prices_df": sp,
            "sp500_simple_returns_df": sp.pct_change(),
            "risk_free_rate_df": rf,
        }


def generate_market(
    n_assets: int,
    n_days: int,
    seed: int = 0,
    start: str = "2006-01-02",
    n_factors: int = 5,
    bars_per_day: int = BARS_PER_DAY,
    rf_mode: str = "varying",
    mcm_mode: str = "varying",
) -> SyntheticMarket:
    """Seeded factor-model market.

    Daily log-return volatility is 1-3 %, drift ~4e-4/day, split evenly over the
    ``bars_per_day`` intraday bars (the first bar of a day carries the overnight
    move).  ``rf_mode`` / ``mcm_mode`` = ``"constant"`` produce a flat DTB3 / VIX /
    EPU series (a flat MCM series makes the MCM fraction exactly 1, i.e. the
    "plain conjugate" case of SURVEY §8(d)).
    """
    rng = np.random.default_rng(seed)
    N, D, B = n_assets, n_days, bars_per_day
    dates = business_days(start, D)

    # Only exactly-rounded elementwise arithmetic and sequential cumulative products are used below
    # (no BLAS matmul, no exp/sin/log ufuncs): the same seed must give bit-identical arrays on the
    # build container and on the GPU box, whose CPUs / BLAS thread counts differ.  The golden
    # fixtures store SHA-256 prefixes of these arrays and the tests refuse to run on a mismatch.
    beta = rng.normal(0.0, 1.0, size=(n_factors, N))
    beta[0] = np.abs(beta[0]) * 0.5 + 0.6          # market factor: all positive
    fvol = np.concatenate([[0.010], np.full(n_factors - 1, 0.004)])
    ivol = rng.uniform(0.008, 0.025, size=N)
    drift = rng.uniform(1e-4, 7e-4, size=N)

    rows = D * B
    sb = 1.0 / float(np.sqrt(float(B)))
    # intraday simple returns: idiosyncratic + factor part (+ drift); price = cumulative product
    r = rng.standard_normal(size=(rows, N))
    r *= ivol * sb
    f = rng.standard_normal(size=(rows, n_factors))
    for k in range(n_factors):
        r += (f[:, k:k + 1] * (fvol[k] * sb)) * beta[k]
    r += drift / B
    r += 1.0
    np.cumprod(r, axis=0, out=r)
    p0 = rng.uniform(20.0, 400.0, size=N)
    r *= p0
    hf_prices = r
    prices = np.ascontiguousarray(hf_prices[B - 1 :: B])      # close = last bar of the day

    shares = rng.uniform(5e7, 5e9, size=N)
    # distinct share counts -> no ties in caps
    shares = np.sort(shares)[rng.permutation(N)] * (1.0 + 1e-6 * np.arange(N))
    caps = prices * shares

    day_ns = dates.astype("int64")
    bar_off = (_FIRST_BAR_MIN + 5 * np.arange(B)) * _NS_PER_MIN
    hf_ts = (day_ns[:, None] + bar_off[None, :]).reshape(-1).astype("datetime64[ns]")

    tt = np.arange(D, dtype=np.float64)

    def tri(x):                                    # triangle wave in [-1, 1], arithmetic only
        fr = x - np.floor(x)
        return 2.0 * np.abs(2.0 * fr - 1.0) - 1.0

    if mcm_mode == "constant":
        vix = np.full(D, 20.0)
        epu = np.full(D, 100.0)
    else:
        vix = 20.0 * (1.0 + 0.35 * tri(tt / 517.0)) * (1.0 + 0.08 * np.clip(rng.standard_normal(D), -3, 3))
        epu = 100.0 * (1.0 + 0.45 * tri(tt / 731.0 + 0.3)) * (1.0 + 0.12 * np.clip(rng.standard_normal(D), -3, 3))
    if rf_mode == "constant":
        rf = np.full(D, 0.02)
    else:
        rf = 0.025 + 0.024 * tri(tt / 1900.0) + 0.0005 * rng.standard_normal(D)
        rf = np.clip(rf, 0.0, 0.05)
    sp500 = 1200.0 * np.cumprod(1.0 + rng.normal(3e-4, 0.011, size=D))

    return SyntheticMarket(
        tickers=make_tickers(N),
        dates=dates,
        prices=prices,
        caps=caps,
        hf_ts=hf_ts,
        hf_prices=hf_prices,
        vix=vix,
        epu=epu,
        rf=rf,
        sp500=sp500,
        seed=seed,
        meta={"n_factors": n_factors, "bars_per_day": B, "start": start,
              "rf_mode": rf_mode, "mcm_mode": mcm_mode},
    )
