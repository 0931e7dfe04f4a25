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
