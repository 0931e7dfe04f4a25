"""Output formats and evaluation statistics.

* CSV layer of ``main.py:47-81``: three files per portfolio spec (``<name>_simple_returns_<start>_<end>.csv``,
  ``…_turnover_…``, ``…_portfolio_weights_metrics_…``), written with the reference's own ``to_csv(header=True)`` calls
  and read back the way ``main.py:56-66`` does (``index_col=0, parse_dates=True``; ``squeeze=True`` was removed from
  pandas 2, ``.squeeze("columns")`` is its documented replacement).
* ``performance_metrics`` of ``portfolio_evaluation.py:464-701`` as ONE batched call over an ensemble of return series
  (BASELINE config 5: 64 paths x strategies): ``path_metrics`` runs ``path_metrics_kernel`` through the C-ABI
  (``bp_path_metrics``); the host keeps the label / calendar work (``compute_excess_returns`` :703-719,
  ``adjust_returns`` :46-72, ``get_insolvent_date`` :27-33) and the probabilistic Sharpe ratio (:78-120), which is a
  scalar formula over three of the kernel's outputs.

Plots (:122-462) and the LaTeX-style highlighting (:408-462) are presentation and out of scope.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Sequence

import numpy as np
import pandas as pd

METRIC_ROWS = ["Cum. Return", "CAGR", "Sharpe", "Sortino", "Max. DD", "Calmar", "Avg. Loss", "Avg. Return", "Avg. Win",
               "Best Day", "Worst Day", "Ann. Vol.", "Daily VaR", "skew", "kurtosis", "sharpe_per_period"]
PM = {name: i for i, name in enumerate(METRIC_ROWS)}


# ----------------------------------------------------------------------------- CSV layer (main.py:47-81)
def result_files(results_dir: str, portfolio_spec_name: str, str_start_date: str, str_end_date: str):
    """The three file names of ``main.py:48-50``."""
    mk = lambda kind: os.path.join(results_dir, f"{portfolio_spec_name}_{kind}_{str_start_date}_{str_end_date}.csv")
    return mk("simple_returns"), mk("turnover"), mk("portfolio_weights_metrics")


def save_results(results_dir, portfolio_spec_name, str_start_date, str_end_date, portfolio_performance: Dict):
    """``main.py:79-81``."""
    r, t, m = result_files(results_dir, portfolio_spec_name, str_start_date, str_end_date)
    portfolio_performance["portfolio_simple_returns_series"].to_csv(r, header=True)
    portfolio_performance["portfolio_turnover_series"].to_csv(t, header=True)
    portfolio_performance["portfolio_weights_metrics_df"].to_csv(m, header=True)
    return r, t, m


def load_results(results_dir, portfolio_spec_name, str_start_date, str_end_date) -> Optional[Dict]:
    """``main.py:52-66``: the cached containers, or None when one of the three files is missing."""
    files = result_files(results_dir, portfolio_spec_name, str_start_date, str_end_date)
    if not all(os.path.exists(f) for f in files):
        return None
    r = pd.read_csv(files[0], index_col=0, parse_dates=True).squeeze("columns")
    t = pd.read_csv(files[1], index_col=0, parse_dates=True).squeeze("columns")
    m = pd.read_csv(files[2], index_col=0, parse_dates=True)
    return {"portfolio_simple_returns_series": r, "portfolio_turnover_series": t, "portfolio_weights_metrics_df": m}


# ----------------------------------------------------------------------------- label / calendar work on the host
def compute_excess_returns(portfolio_simple_returns_series, risk_free_rate_df):
    """:703-719 — risk-free rate forward-filled (then back-filled) onto the return dates, de-annualised with 1/252."""
    rf = risk_free_rate_df["DTB3"].reindex(portfolio_simple_returns_series.index).ffill().bfill()
    out = portfolio_simple_returns_series - ((rf + 1) ** (1 / 252) - 1)
    out.name = portfolio_simple_returns_series.name
    return out


def get_insolvent_date(returns_series):
    """:27-33 — first date whose cumulative return is below -99 %."""
    cum = (1 + returns_series).cumprod() - 1
    return cum[cum < -0.99].first_valid_index()


def adjust_returns(series):
    """:46-72 — once the cumulative return would fall below -100 %, the return of that day is replaced
    (``0.000001 / previous cumulative return - 1``, sic) and every later return is 0.  O(T) instead of the
    reference's O(T^2) re-multiplication; same values."""
    v = series.to_numpy(dtype=np.float64).copy()
    cum = np.cumprod(1.0 + v) - 1.0
    hit = np.nonzero(cum < -1.0)[0]
    if len(hit):
        i = int(hit[0])
        v[i] = (0.000001 / cum[i - 1] - 1.0) if i > 0 else -1.0
        v[i + 1:] = 0.0
    return pd.Series(v, index=series.index, name=series.name)


def prob_sharpe_ratio(sharpe_1, skewness, kurt, n, benchmark_sharpe_1):
    """:78-120 from per-period Sharpe ratios, skewness and (non-excess) kurtosis; arrays broadcast."""
    var = (1.0 - skewness * sharpe_1 + ((kurt - 1.0) / 4.0) * sharpe_1 ** 2) / (n - 1.0)
    z = (sharpe_1 - benchmark_sharpe_1) / np.sqrt(var)
    return 0.5 * (1.0 + np.vectorize(math.erf)(z / math.sqrt(2.0)))


# ----------------------------------------------------------------------------- the batched statistics (CUDA)
def path_metrics(returns: np.ndarray, excess: np.ndarray, years: float, engine=None) -> np.ndarray:
    """[P][T] simple and excess returns -> [P][16] statistics in ``METRIC_ROWS`` order (``bp_path_metrics``)."""
    from . import portfolio_calculations as pc
    from .engine import _raise
    eng = engine or pc._engine()
    r = np.ascontiguousarray(returns, dtype=np.float64)
    x = np.ascontiguousarray(excess, dtype=np.float64)
    if r.ndim != 2 or r.shape != x.shape:
        raise ValueError("returns and excess must be [paths][observations] arrays of one shape")
    if np.isnan(r).any() or np.isnan(x).any():
        raise ValueError("NaN in a return series (QuantStats would fillna(0) silently)")
    out = np.empty((r.shape[0], len(METRIC_ROWS)))
    dp = C.POINTER(C.c_double)
    rc = eng._lib.bp_path_metrics(eng._h, r.shape[0], r.shape[1], r.ctypes.data_as(dp), x.ctypes.data_as(dp), float(years),
                                  out.ctypes.data_as(dp))
    if rc:
        _raise(rc)
    return out


def performance_metrics(portfolio_specs_simple_returns: Dict[str, pd.Series], risk_free_rate_df,
                        portfolio_specs_turnover: Optional[Dict[str, pd.Series]] = None, benchmark: str = "S&P 500",
                        engine=None) -> pd.DataFrame:
    """The metrics table of :464-701 for series on ONE common index (``check_indexes_and_convert_to_datetime``,
    :721-736): rows as in the reference ('Cum. Return' … 'Daily VaR', 'Prob. Sharpe', 'Avg. Turnover'), one column
    per series.  Insolvent series (:505-510) get None where the reference writes None; their 'Worst Day', 'Ann. Vol.'
    and 'Daily VaR' use the returns before the insolvency date as the reference does."""
    names = list(portfolio_specs_simple_returns)
    idx = portfolio_specs_simple_returns[names[0]].index
    for k in names:
        if not portfolio_specs_simple_returns[k].index.equals(idx):
            raise ValueError("all series must share one index (:721-736)")
    adj = {k: adjust_returns(portfolio_specs_simple_returns[k]) for k in names}
    exc = {k: adjust_returns(compute_excess_returns(portfolio_specs_simple_returns[k], risk_free_rate_df)) for k in names}
    years = (idx[-1] - idx[0]).days / 365
    M = path_metrics(np.stack([adj[k].to_numpy() for k in names]), np.stack([exc[k].to_numpy() for k in names]), years, engine)
    table = pd.DataFrame(index=["Cum. Return", "CAGR", "Sharpe", "Prob. Sharpe", "Sortino", "Calmar", "Max. DD", "Avg. Loss",
                                "Avg. Return", "Avg. Win", "Best Day", "Worst Day", "Ann. Vol.", "Daily VaR", "Avg. Turnover"],
                         columns=names, dtype=object)
    bench = M[names.index(benchmark), PM["sharpe_per_period"]] if benchmark in names else None
    for j, k in enumerate(names):
        insolvent = get_insolvent_date(adj[k])
        row = {name: M[j, PM[name]] for name in METRIC_ROWS}
        if insolvent is not None:
            for name in ("CAGR", "Sharpe", "Sortino", "Calmar"):
                row[name] = None
            before = adj[k][:insolvent - pd.Timedelta(days=1)]
            nz = adj[k][abs(adj[k]) > 1e-7]
            row["Avg. Return"] = nz.mean()
            row["Worst Day"] = before.min()
            row["Ann. Vol."] = before.std() * 252 ** 0.5
            row["Daily VaR"] = before.mean() - 1.6448536269514729 * before.std()
        for name in table.index:
            if name in row:
                table.at[name, k] = row[name]
        if bench is not None and insolvent is None:
            table.at["Prob. Sharpe", k] = float(prob_sharpe_ratio(M[j, PM["sharpe_per_period"]], M[j, PM["skew"]],
                                                                  M[j, PM["kurtosis"]], len(idx), bench))
        if portfolio_specs_turnover is not None and k in portfolio_specs_turnover:
            t = portfolio_specs_turnover[k]
            table.at["Avg. Turnover", k] = (t if insolvent is None else t[:insolvent]).mean()
    return table
