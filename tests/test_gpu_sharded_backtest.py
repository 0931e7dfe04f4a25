"""The north-star partition: ONE daily-rebalance backtest split by contiguous date range (+ window halo + the
one-window loop halo) over `world` ranks (incorporating_different_sources_b200.sharding.ShardedBacktest).  The ranks
are evaluated one after the other on one GPU (nothing guarantees co-scheduling of several ranks on one device); the
concatenated per-rank rows must equal the unsharded backtest: weights of both priors, the daily portfolio returns
and the turnover of the loop body (:1127-1219).  Every 9th window is also checked against the CPU oracle."""
import numpy as np
import pytest

from oracle import bayes_oracle as bo
from tests._golden import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from incorporating_different_sources_b200.engine import BayesEngine
    eng = BayesEngine(0)
    yield eng
    eng.close()


def _run(engine, mkt, conj, jeff, d_idx, world):
    import torch
    from incorporating_different_sources_b200.sharding import ShardedBacktest
    N = mkt.n_assets
    per_rank = []
    for rank in range(world):
        sb = ShardedBacktest(engine, mkt, conj, jeff, d_idx, rank, world, hf_lookback_days=7)
        sb.upload()
        out_c = {"weights": torch.empty((sb.n_ext, N), dtype=torch.float64, device="cuda:0"),
                 "status": torch.empty((sb.n_ext,), dtype=torch.int32, device="cuda:0")}
        out_j = {"weights": torch.empty((sb.n_ext, N), dtype=torch.float64, device="cuda:0"),
                 "status": torch.empty((sb.n_ext,), dtype=torch.int32, device="cuda:0")}
        rows = sb.compute(out_c, out_j, loop=True)
        assert not out_c["status"].any() and not out_j["status"].any()
        per_rank.append([x.cpu().numpy().copy() for x in rows])
        assert per_rank[-1][0].shape[0] == sb.hi - sb.lo
        assert per_rank[-1][2].shape[0] == sb.hi - sb.lo - (1 if rank == 0 else 0)
    return [np.concatenate([p[k] for p in per_rank], axis=0) for k in range(6)]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_date_range_shards_equal_unsharded_backtest(engine, world):
    from incorporating_different_sources_b200.synthetic import generate_market
    N, n_c, n_j, W = 24, 60, 100, 83
    mkt = generate_market(N, n_j + W + 2, seed=41)
    conj = dict(weighting_strategy="conjugate_hf_vix_vw", size=N, risk_aversion=5, turnover_cost=15,
                rebalancing_frequency="daily", rolling_window=n_c, rolling_window_frequency="daily", mcm_scaling=1,
                display_name="conj")
    jeff = dict(conj, weighting_strategy="jeffreys", mcm_scaling=None, rolling_window=n_j, display_name="jeff")
    d_idx = np.arange(n_j + 1, n_j + 1 + W)
    full = _run(engine, mkt, conj, jeff, d_idx, 1)
    got = _run(engine, mkt, conj, jeff, d_idx, world)
    names = ("weights_c", "weights_j", "returns_c", "returns_j", "turnover_c", "turnover_j")
    for k, name in enumerate(names):
        assert got[k].shape == full[k].shape, name
        assert relerr(got[k], full[k]) <= 1e-11, name
    assert full[2].shape == (W - 1,) and full[4].shape == (W - 1,)
    cols = np.arange(N)
    for i in range(0, W, 9):
        ref = bo.conjugate_window(conj, mkt, int(d_idx[i]), cols, hf_lookback_days=7)
        assert relerr(got[0][i], ref["weights"]) <= 1e-9
        ref = bo.jeffreys_window(jeff, mkt, int(d_idx[i]), cols)
        assert relerr(got[1][i], ref["weights"]) <= 1e-9


def test_pool_gather_equals_direct_upload(engine):
    """bp_upload_pool + bp_select_market (device gather of a column subset in a given order and of a row range) must
    leave exactly the working market a direct upload of the same slice leaves: bit-identical weights; and the error
    paths of the two entry points."""
    from incorporating_different_sources_b200.engine import BayesPortfolioError
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import ffill_rows, plan_daily_windows
    N_all, n, W = 30, 50, 40
    mkt = generate_market(N_all, n + W + 20, seed=8)
    spec = dict(weighting_strategy="conjugate_hf_epu_vw", size=12, risk_aversion=4, rolling_window=n,
                rolling_window_frequency="daily", mcm_scaling=1)
    cols = np.random.default_rng(1).permutation(N_all)[:12]
    d_idx = np.arange(n + 10, n + 10 + W)
    day_lo, day_hi = int(d_idx.min()) - (n - 1), int(d_idx.max()) + 1
    day = np.timedelta64(1, "D")
    hf_lo = int(np.searchsorted(mkt.hf_ts, mkt.dates[d_idx.min()] - 7 * day + day, side="right")) - 1
    hf_hi = int(np.searchsorted(mkt.hf_ts, mkt.dates[d_idx.max()] + day, side="right"))
    rf_row = ffill_rows(mkt.dates, mkt.dates, mkt.rf)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7, row_offset=day_lo, hf_row_offset=hf_lo)
    batch.prior_n = np.full(W, 61.5)                     # no MCM series in the pool: n0 is injected (as backtest.py does)
    outs = ("weights", "w1", "scalars", "status")
    engine.upload_market(prices=mkt.prices[day_lo:day_hi][:, cols], rf_row=rf_row[day_lo:day_hi],
                         caps=mkt.caps[day_lo:day_hi][:, cols], hf_prices=mkt.hf_prices[hf_lo:hf_hi][:, cols])
    ref = engine.conjugate(batch, outputs=outs)
    with pytest.raises(BayesPortfolioError):
        engine.pool_shape = (mkt.n_days, N_all, len(mkt.hf_ts))
        engine.select_market(cols, day_lo, day_hi, hf_lo, hf_hi)          # no pool resident yet
    engine.upload_pool(prices=mkt.prices, rf_row=rf_row, caps=mkt.caps, hf_prices=mkt.hf_prices)
    engine.select_market(cols, day_lo, day_hi, hf_lo, hf_hi)
    got = engine.conjugate(batch, outputs=outs)
    assert not ref["status"].any()
    for k in outs:
        assert np.array_equal(got[k], ref[k]), k
    # a second selection from the same pool (other columns, other rows) and back: still identical
    engine.select_market(np.arange(5), 0, 60, 0, 200)
    engine.select_market(cols, day_lo, day_hi, hf_lo, hf_hi)
    assert np.array_equal(engine.conjugate(batch, outputs=("weights",))["weights"], ref["weights"])
    for bad in (dict(cols=[N_all]), dict(day_lo=-1), dict(day_hi=mkt.n_days + 1), dict(hf_hi=len(mkt.hf_ts) + 1), dict(cols=[])):
        kw = dict(cols=cols, day_lo=day_lo, day_hi=day_hi, hf_lo=hf_lo, hf_hi=hf_hi)
        kw.update(bad)
        with pytest.raises(ValueError):                    # BP_ERR_INVALID maps to ValueError (engine._raise)
            engine.select_market(**kw)
    engine.upload_pool(None, None)
