"""path_metrics_kernel through the C-ABI (bp_path_metrics) against oracle/eval_oracle.py: an ensemble of 64 series of
4,149 returns (BASELINE config 5's shape), ragged cases (no losing day, a series with zeros), and the metrics table of
``performance_metrics`` with an insolvent series."""
import numpy as np
import pandas as pd
import pytest

from incorporating_different_sources_b200 import evaluation as ev
from oracle import eval_oracle as eo

pytestmark = pytest.mark.gpu


def _ensemble(P, T, seed):
    rng = np.random.default_rng(seed)
    idx = pd.bdate_range("2007-01-03", periods=T)
    r = rng.normal(4e-4, 0.011, size=(P, T)) * (1 + 2 * rng.random((P, 1)))
    rf = pd.DataFrame({"DTB3": 0.02 + 0.01 * np.sin(np.arange(T) / 200.0)}, index=idx)
    x = np.stack([ev.compute_excess_returns(pd.Series(row, index=idx), rf).to_numpy() for row in r])
    return idx, r, x, rf


@pytest.mark.parametrize("P,T", [(64, 4149), (3, 2), (5, 257)])
def test_path_metrics_match_oracle(P, T):
    idx, r, x, _ = _ensemble(P, T, 5)
    r[0, T // 2:] = np.abs(r[0, T // 2:])                     # a long winning streak
    if T > 10:
        r[1, ::7] = 0.0                                        # exact zeros (avg_return skips them)
    years = (idx[-1] - idx[0]).days / 365
    got = ev.path_metrics(r, x, years)
    assert got.shape == (P, 16)
    for j in range(P):
        ref = eo.path_row(r[j], x[j], idx)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got[j]), ok), j
        # (the skewness of a symmetric sample is 0 in exact arithmetic: absolute floor for rounding noise)
        assert np.all(np.abs(got[j][ok] - ref[ok]) <= 1e-10 * np.abs(ref[ok]) + 1e-13), (j, got[j], ref)


def test_no_losing_day_gives_nan_average_loss():
    idx, r, x, _ = _ensemble(2, 40, 9)
    r[0] = np.abs(r[0]) + 1e-4
    got = ev.path_metrics(r, x, 0.15)
    assert np.isnan(got[0, ev.PM["Avg. Loss"]]) and got[0, ev.PM["Max. DD"]] == 0.0
    assert np.isfinite(got[1, ev.PM["Avg. Loss"]])


def test_performance_metrics_table_with_insolvent_series():
    idx, r, x, rf = _ensemble(4, 600, 13)
    names = ["A", "B", "C", "S&P 500"]
    sr = {k: pd.Series(r[j], index=idx, name=k) for j, k in enumerate(names)}
    sr["B"].iloc[300] = -1.4                                   # below -100 %: adjusted, later returns 0 (:46-72)
    to = {k: pd.Series(np.abs(r[j][1:]) * 3, index=idx[1:], name=k) for j, k in enumerate(names)}
    tab = ev.performance_metrics(sr, rf, to)
    assert list(tab.columns) == names and "Prob. Sharpe" in tab.index
    for k in names:
        a = ev.adjust_returns(sr[k])
        xa = ev.adjust_returns(ev.compute_excess_returns(sr[k], rf))
        row = eo.path_row(a.to_numpy(), xa.to_numpy(), idx)
        assert abs(tab.at["Cum. Return", k] - row[0]) <= 1e-12 * max(1, abs(row[0]))
        assert abs(tab.at["Max. DD", k] - row[4]) <= 1e-12
        if k == "B":
            assert tab.at["Sharpe", k] is None and tab.at["CAGR", k] is None and tab.at["Calmar", k] is None
            assert abs(tab.at["Cum. Return", k] + 1.0) < 1e-3        # the "sic" adjustment of :62 leaves ~1e-5 of the capital
            ins = ev.get_insolvent_date(a)
            assert abs(tab.at["Avg. Turnover", k] - to[k][:ins].mean()) <= 1e-15
        else:
            assert abs(tab.at["Sharpe", k] - row[2]) <= 1e-10 * abs(row[2])
            assert abs(tab.at["Sortino", k] - row[3]) <= 1e-10 * abs(row[3])
            assert abs(tab.at["Daily VaR", k] - row[12]) <= 1e-12
            xb = ev.adjust_returns(ev.compute_excess_returns(sr["S&P 500"], rf)).to_numpy()
            assert abs(tab.at["Prob. Sharpe", k] - eo.prob_sharpe(xa.to_numpy(), xb)) <= 1e-10
            assert abs(tab.at["Avg. Turnover", k] - to[k].mean()) <= 1e-15
