"""TEST INFRASTRUCTURE — CPU restatement of the evaluation statistics of ``portfolio_evaluation.py:464-701``.

The reference delegates most of them to QuantStats 0.0.62 (``requirements.txt:10``; third-party, not vendored under
/root/reference, not installed here).  Its published algorithms are restated below, one function per call site, and
pinned on the reference's OWN ``CHECK`` expressions (:520-524 CAGR, :537-541 Sharpe, :586-590 average loss, :600-604
average return, :617-621 average win, :648-652 volatility, :85-108 per-period Sharpe / skewness / kurtosis), which
tests/test_evaluation.py evaluates with pandas exactly as the reference writes them.  Sortino, maximum drawdown and VaR
have no CHECK in the reference: for those three parity is pinned on the published QuantStats formula only ("parity
unpinned" against the reference proper).  Only tests/ may import this module.
"""
import numpy as np
from scipy.stats import kurtosis, norm, skew


def comp(r):                      # qs.stats.comp (:513): (1 + r).prod() - 1
    return float(np.prod(1.0 + r) - 1.0)


def cagr(r, index, periods=365):  # qs.stats.cagr(periods=365) (:520): abs(total + 1) ** (1 / years) - 1
    years = (index[-1] - index[0]).days / periods
    return abs(comp(r) + 1.0) ** (1.0 / years) - 1.0


def sharpe(x, periods=252):       # qs.stats.sharpe (:535): mean / std(ddof=1) * sqrt(periods)
    return float(np.mean(x) / np.std(x, ddof=1) * np.sqrt(periods))


def sortino(x, periods=252):      # qs.stats.sortino (:560): mean / sqrt(sum(neg^2) / len) * sqrt(periods)
    downside = np.sqrt(np.sum(x[x < 0] ** 2) / len(x))
    return float(np.mean(x) / downside * np.sqrt(periods))


def max_drawdown(r):              # qs.stats.max_drawdown (:569,:577): prices = 1e5 * cumprod(1 + r); (p / expanding max).min() - 1
    p = 1e5 + 1e5 * (np.cumprod(1.0 + r) - 1.0)
    return float(np.min(p / np.maximum.accumulate(p)) - 1.0)


def avg_loss(r):                  # :584
    return float(np.mean(r[r < 0]))


def avg_return(r):                # :598
    return float(np.mean(r[r != 0]))


def avg_win(r):                   # :615
    return float(np.mean(r[r > 0]))


def volatility(r, periods=252):   # :646
    return float(np.std(r, ddof=1) * np.sqrt(periods))


def value_at_risk(r, confidence=0.95):   # :664: norm.ppf(1 - confidence, mean, std(ddof=1))
    return float(norm.ppf(1 - confidence, np.mean(r), np.std(r, ddof=1)))


def path_row(r, x, index):
    """The 16 statistics of one series in the order of incorporating_different_sources_b200.evaluation.METRIC_ROWS."""
    c = cagr(r, index)
    mdd = max_drawdown(r)
    with np.errstate(all="ignore"):
        calmar = float(np.float64(c) / np.float64(abs(mdd)))          # no drawdown: inf, as the reference's division gives
        return np.array([comp(r), c, sharpe(x), sortino(x), mdd, calmar, _mean(r[r < 0]), _mean(r[r != 0]), _mean(r[r > 0]),
                         float(np.max(r)), float(np.min(r)), volatility(r), value_at_risk(r), float(skew(x)),
                         float(kurtosis(x, fisher=False)), sharpe(x, periods=1)])


def _mean(v):
    return float(np.mean(v)) if len(v) else float("nan")               # pandas: mean of an empty selection is NaN


def prob_sharpe(x, xb):           # :78-120
    sr, srb = sharpe(x, 1), sharpe(xb, 1)
    var = (1 - skew(x) * sr + ((kurtosis(x, fisher=False) - 1) / 4) * sr ** 2) / (len(x) - 1)
    return float(norm.cdf((sr - srb) / np.sqrt(var)))
