"""Loop level: ``backtest_portfolio`` (reference signature, three output containers) on the batched CUDA
engine + loop-body kernel against the unmodified reference's own backtest (tests/golden/bt_*.npz)."""
import glob
import json
import os

import numpy as np
import pandas as pd
import pytest

from incorporating_different_sources_b200.synthetic import generate_market

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "bt_*.npz")))
TOL = 1e-9


def _close(a, b, tol=TOL, atol=1e-13):
    """max|a-b| <= tol * max|b| + atol; atol covers quantities that are exactly zero in exact arithmetic
    (the turnover of a daily rebalanced value-weighted portfolio is pure rounding noise, ~1e-16)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    if m.any():
        assert np.max(np.abs(a[m] - b[m])) <= tol * np.max(np.abs(b[m])) + atol


@pytest.mark.parametrize("name", NAMES)
def test_backtest_matches_reference(name):
    from incorporating_different_sources_b200 import portfolio_calculations as pc
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    mkt = generate_market(**meta["market"])
    md = mkt.market_data()
    if meta.get("damage"):
        from tests.golden.make_backtest_golden import apply_damage
        md = apply_damage(md, meta["damage"])
    res = pc.backtest_portfolio(meta["spec"], pd.Timestamp(meta["start"]), pd.Timestamp(meta["end"]), md)
    r, t, m = (res["portfolio_simple_returns_series"], res["portfolio_turnover_series"],
               res["portfolio_weights_metrics_df"])
    assert r.name == meta["series_name"] and t.name == meta["series_name"]
    assert list(m.columns) == meta["metrics_columns"]
    assert np.array_equal(r.index.values.astype("int64"), z["returns_idx"])      # window / date indexing bit-exact
    assert np.array_equal(t.index.values.astype("int64"), z["turnover_idx"])
    assert np.array_equal(m.index.values.astype("int64"), z["metrics_idx"])
    _close(r.to_numpy(), z["returns"])
    _close(t.to_numpy(), z["turnover"])
    _close(m.to_numpy(), z["metrics"])


def test_dispatcher_matches_per_window_reference_order():
    """calculate_portfolio_weights keeps the cap-descending index order of the reference (F7)."""
    from incorporating_different_sources_b200 import portfolio_calculations as pc
    z = np.load(os.path.join(GOLD, "n12of30_conj_vix_vw_topk.npz"))
    meta = json.loads(str(z["meta"]))
    mkt = generate_market(**meta["market"])
    md = mkt.market_data()
    for wi, w in enumerate(meta["windows"]):
        d = pd.Timestamp(mkt.dates[w["d_idx"]])
        wdf = pc.calculate_portfolio_weights(d, meta["spec"], md)
        names = [mkt.tickers[c] for c in z[f"w{wi}_cols"]]
        assert list(wdf.index) == names
        ref = z[f"w{wi}_weights"]
        assert np.max(np.abs(wdf["Weight"].to_numpy() - ref)) / np.max(np.abs(ref)) <= TOL
