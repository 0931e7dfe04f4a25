"""SURVEY 8(f) rank 4 on the CPU: (1) the CSV layer writes what main.py:79-81 writes and reads it back as main.py:56-66
does; (2) oracle/eval_oracle.py (the restated QuantStats 0.0.62 formulas) against the reference's OWN CHECK expressions,
evaluated with pandas exactly as portfolio_evaluation.py writes them (:520-524, :537-541, :586-590, :600-604, :617-621,
:648-652, :85-108); (3) the host-side helpers against the UNMODIFIED reference functions (adjust_returns,
get_insolvent_date, compute_excess_returns, prob_sharpe_ratio_with_benchmark), imported with stub modules for the
plotting / QuantStats dependencies that are not installed (the stubbed qs.stats.sharpe is the formula the reference
itself asserts at :85-98)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pandas as pd
import pytest

from incorporating_different_sources_b200 import evaluation as ev
from oracle import eval_oracle as eo
from oracle.ref_import import REF_SRC_CANDIDATES


def _series(seed, T=700, name="S", scale=0.012):
    rng = np.random.default_rng(seed)
    idx = pd.bdate_range("2015-01-05", periods=T)
    return pd.Series(rng.normal(4e-4, scale, T), index=idx, name=name)


def _rf(idx):
    return pd.DataFrame({"DTB3": np.linspace(0.01, 0.03, len(idx) // 3)}, index=idx[::3][: len(idx) // 3])


def test_csv_round_trip_matches_main_py_schema(tmp_path):
    r = _series(1, 40, "Conjugate HF-VIX VW")
    t = pd.Series(np.abs(_series(2, 40).to_numpy()[1:]), index=r.index[1:], name=r.name)
    m = pd.DataFrame(np.random.default_rng(3).normal(size=(40, 5)), index=r.index,
                     columns=["max_long", "max_short", "avg_long", "avg_short", "average_distance_to_comparison_portfolio"])
    perf = {"portfolio_simple_returns_series": r, "portfolio_turnover_series": t, "portfolio_weights_metrics_df": m}
    assert ev.load_results(str(tmp_path), "spec", "2015-01-05", "2015-02-27") is None
    files = ev.save_results(str(tmp_path), "spec", "2015-01-05", "2015-02-27", perf)
    assert [os.path.basename(f) for f in files] == ["spec_simple_returns_2015-01-05_2015-02-27.csv",
                                                    "spec_turnover_2015-01-05_2015-02-27.csv",
                                                    "spec_portfolio_weights_metrics_2015-01-05_2015-02-27.csv"]
    with open(files[0]) as f:
        assert f.readline().strip() == ",Conjugate HF-VIX VW"          # to_csv(header=True): blank index label + series name
    back = ev.load_results(str(tmp_path), "spec", "2015-01-05", "2015-02-27")
    for k in perf:
        got, ref = back[k], perf[k]
        assert type(got) is type(ref) and got.index.equals(ref.index)
        assert np.allclose(got.to_numpy(), ref.to_numpy(), rtol=0, atol=1e-15)
    assert back["portfolio_simple_returns_series"].name == r.name
    assert list(back["portfolio_weights_metrics_df"].columns) == list(m.columns)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_oracle_matches_the_reference_check_expressions(seed):
    s = _series(seed)
    x = ev.compute_excess_returns(s, _rf(s.index))
    r, xv = s.to_numpy(), x.to_numpy()
    # :522  CAGR
    cagr_check = ((1 + s).prod()) ** (1 / ((s.index[-1] - s.index[0]).days / 365)) - 1
    assert abs(eo.cagr(r, s.index) - cagr_check) <= 1e-13
    # :539  Sharpe
    assert abs(eo.sharpe(xv) - (x.mean()) / x.std() * (252 ** 0.5)) <= 1e-12
    # :588, :602, :619  averages (the reference's CHECK uses a 1e-7 dead band; no return of these series is that small)
    assert abs(eo.avg_loss(r) - s[s < -1e-7].mean()) <= 1e-16
    assert abs(eo.avg_win(r) - s[s > 1e-7].mean()) <= 1e-16
    assert abs(eo.avg_return(r) - s.mean()) <= 1e-16
    # :650  volatility
    assert abs(eo.volatility(r) - s.std() * 252 ** 0.5) <= 1e-14
    # :85-108  per-period Sharpe, skewness, kurtosis.  The reference's CHECK divides the BIASED central moment by the
    # ddof=1 standard deviation (hence its 1e-3 tolerance): the exact relation to scipy's biased estimators is a
    # factor ((n-1)/n)^(k/2)
    assert abs(eo.sharpe(xv, 1) - np.mean(x) / np.std(x, ddof=1)) <= 1e-14
    row = eo.path_row(r, xv, s.index)
    n = len(x)
    assert abs(row[13] * ((n - 1) / n) ** 1.5 - ((x - np.mean(x)) ** 3).mean() / (np.std(x, ddof=1) ** 3)) <= 1e-12
    assert abs(row[14] * ((n - 1) / n) ** 2 - ((x - np.mean(x)) ** 4).mean() / (np.std(x, ddof=1) ** 4)) <= 1e-11
    # drawdown: running maximum of the wealth curve, first point included (published formula)
    wealth = np.cumprod(1 + r)
    assert abs(eo.max_drawdown(r) - min(wealth[i] / wealth[: i + 1].max() - 1 for i in range(len(r)))) <= 1e-15


def _load_reference_evaluation():
    src = next((p for p in REF_SRC_CANDIDATES if p and os.path.isfile(os.path.join(p, "portfolio_evaluation.py"))), None)
    if src is None:
        pytest.skip("reference sources not present on this machine")
    stubs = {}
    qs = types.ModuleType("quantstats")
    qs.stats = types.SimpleNamespace(sharpe=lambda s, periods=252: s.mean() / s.std() * np.sqrt(1 if periods is None else periods))
    stubs["quantstats"] = qs
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.dates", "matplotlib.ticker", "seaborn", "portfolio_specs"):
        stubs[name] = types.ModuleType(name)
    stubs["matplotlib.ticker"].FuncFormatter = stubs["matplotlib.ticker"].PercentFormatter = object
    stubs["matplotlib"].pyplot, stubs["matplotlib"].dates, stubs["matplotlib"].ticker = (
        stubs["matplotlib.pyplot"], stubs["matplotlib.dates"], stubs["matplotlib.ticker"])
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_ref_portfolio_evaluation", os.path.join(src, "portfolio_evaluation.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def test_host_helpers_match_the_unmodified_reference_functions():
    pe = _load_reference_evaluation()
    s = _series(21, 300)
    rf = _rf(s.index)
    # compute_excess_returns (:703-719), incl. the back-fill of the dates before the first risk-free observation
    late_rf = rf.iloc[5:]
    for frame in (rf, late_rf):
        got, ref = ev.compute_excess_returns(s, frame), pe.compute_excess_returns(s, frame)
        assert got.name == ref.name and got.index.equals(ref.index) and np.array_equal(got.to_numpy(), ref.to_numpy())
    # adjust_returns / get_insolvent_date (:27-72): a solvent series, a series that goes below -100 %, a first-day wipe-out
    crash = s.copy()
    crash.iloc[120] = -1.7
    first = s.copy()
    first.iloc[0] = -1.2
    for series in (s, crash, first):
        got, ref = ev.adjust_returns(series), pe.adjust_returns(series)
        assert np.allclose(got.to_numpy(), ref.to_numpy(), rtol=1e-13, atol=0) and got.index.equals(ref.index)
        assert ev.get_insolvent_date(got) == pe.get_insolvent_date(ref)
    # prob_sharpe_ratio_with_benchmark (:78-120) with CHECK = True inside the reference (its own kurtosis assertion mixes
    # biased moments with a ddof=1 deviation and only holds for long series: 4,149 returns, the backtest's length)
    s = _series(23, 4149)
    rf = _rf(s.index)
    x, xb = ev.compute_excess_returns(s, rf), ev.compute_excess_returns(_series(22, 4149, "S&P 500"), rf)
    ref = pe.prob_sharpe_ratio_with_benchmark(x, xb)
    assert abs(eo.prob_sharpe(x.to_numpy(), xb.to_numpy()) - ref) <= 1e-13
    row, rowb = eo.path_row(s.to_numpy(), x.to_numpy(), s.index), eo.path_row(xb.to_numpy(), xb.to_numpy(), s.index)
    assert abs(float(ev.prob_sharpe_ratio(row[15], row[13], row[14], len(x), rowb[15])) - ref) <= 1e-13
