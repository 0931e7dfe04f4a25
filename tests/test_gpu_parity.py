"""CUDA path (through the C-ABI) against the reference-generated golden vectors and the oracle.

Tolerance (BASELINE.json north_star): weights and posterior moments within 1e-9 relative in FP64,
measured as max|x_gpu - x_ref| / max|x_ref|; window and asset indexing bit-exact.
"""
import numpy as np
import pytest

from oracle import bayes_oracle as bo
from tests._golden import check_matrix, golden_names, load_golden, market_for, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-9

DAILY = [n for n in golden_names()]


@pytest.fixture(scope="module")
def engine():
    from incorporating_different_sources_b200.engine import BayesEngine
    eng = BayesEngine(0)
    yield eng
    eng.close()


def _daily(meta):
    return meta["spec"]["rolling_window_frequency"] == "daily"


@pytest.mark.parametrize("name", DAILY)
def test_gpu_matches_reference_golden(engine, name):
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.windows import plan_daily_windows
    from incorporating_different_sources_b200.windows import plan_weekly_windows
    z, meta = load_golden(name)
    mkt = market_for(meta)
    spec = meta["spec"]
    conj = spec["weighting_strategy"].startswith("conjugate")
    for wi, w in enumerate(meta["windows"]):
        pre = f"w{wi}_"
        d_idx = w["d_idx"]
        cols = z[pre + "cols"]
        upload_synthetic(engine, mkt, cols=cols)
        if _daily(meta):
            batch = plan_daily_windows(spec, mkt.dates, [d_idx], mkt.hf_ts, hf_lookback_days=meta["hf_days"],
                                       need_hf=conj)
        else:
            rows, batch = plan_weekly_windows(spec, mkt.dates, [d_idx], mkt.dates, mkt.rf,
                                              np.stack([mkt.vix, mkt.epu]), mkt.hf_ts,
                                              hf_lookback_days=meta["hf_days"], need_hf=conj)
            engine.set_resampled(rows)
        if conj:
            got = engine.conjugate(batch, outputs=("weights", "nu", "w1", "t", "w0", "scalars", "status", "T", "S0", "S1"))
            assert got["status"][0] == 0
            s = got["scalars"][0]
            for k, i in (("n0", 0), ("n1", 1), ("c", 4), ("v1", 8)):
                assert abs(s[i] - float(z[pre + k])) <= TOL * abs(float(z[pre + k])), k
            for k in ("t", "w0", "w1", "nu", "weights"):
                assert relerr(got[k][0], z[pre + k]) <= TOL, k
            for k in ("T", "S0", "S1"):
                check_matrix(k, got[k][0], z, pre, TOL)
        else:
            got = engine.jeffreys(batch, outputs=("weights", "nu", "t", "status", "T"))
            assert got["status"][0] == 0
            for k in ("t", "nu", "weights"):
                assert relerr(got[k][0], z[pre + k]) <= TOL, k
            check_matrix("T", got["T"][0], z, pre, TOL)


@pytest.mark.parametrize("n_assets,hf_days,n_windows", [(10, None, 40), (50, 7, 25), (130, 7, 12), (100, 31, 6)])
def test_gpu_batched_matches_oracle(engine, n_assets, hf_days, n_windows):
    """Many overlapping windows in one call vs the oracle evaluated window by window."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(n_assets, 300, seed=4000 + n_assets)
    spec = dict(weighting_strategy="conjugate_hf_epu_vw", size=n_assets, risk_aversion=4, turnover_cost=15,
                rebalancing_frequency="daily", rolling_window=252, rolling_window_frequency="daily",
                mcm_scaling=2, display_name="x")
    d_idx = list(range(300 - n_windows, 300))
    cols = np.arange(n_assets)
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=hf_days)
    got = engine.conjugate(batch, outputs=("weights", "w1", "scalars", "status"))
    jspec = dict(spec, weighting_strategy="jeffreys")
    jb = plan_daily_windows(jspec, mkt.dates, d_idx, need_hf=False)
    gotj = engine.jeffreys(jb, outputs=("weights", "status"))
    assert not got["status"].any() and not gotj["status"].any()
    for i, d in enumerate(d_idx):
        ref = bo.conjugate_window(spec, mkt, d, cols, hf_lookback_days=hf_days)
        assert relerr(got["weights"][i], ref["weights"]) <= TOL
        assert relerr(got["w1"][i], ref["w1"]) <= TOL
        assert abs(got["scalars"][i][4] - ref["c"]) <= TOL * abs(ref["c"])
        refj = bo.jeffreys_window(jspec, mkt, d, cols)
        assert relerr(gotj["weights"][i], refj["weights"]) <= TOL


def test_gpu_stats_and_hf_cov_entry_points(engine):
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(37, 280, seed=77)        # odd N: exercises every padding path
    spec = dict(weighting_strategy="conjugate_hf_vix_ew", size=37, risk_aversion=5, rolling_window=252,
                rolling_window_frequency="daily", mcm_scaling=1)
    d_idx = [270, 279]
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts)
    t, T = engine.stats(batch)
    n0, S0 = engine.hf_cov(batch)
    cols = np.arange(37)
    for i, d in enumerate(d_idx):
        ref = bo.conjugate_window(spec, mkt, d, cols)
        assert relerr(t[i], ref["t"]) <= TOL and relerr(T[i], ref["T"]) <= TOL
        assert abs(n0[i] - ref["n0"]) <= TOL * ref["n0"] and relerr(S0[i], ref["S0"]) <= TOL
        assert np.array_equal(T[i], T[i].T)


def test_gpu_status_flags_singular_window(engine):
    """N >= n-1 with a 1-day HF prior is singular in the reference: the CUDA path must
    flag the window instead of returning silent garbage."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(64, 70, seed=5)
    spec = dict(weighting_strategy="jeffreys", size=64, risk_aversion=5, rolling_window=40,
                rolling_window_frequency="daily", mcm_scaling=None)
    upload_synthetic(engine, mkt)
    jb = plan_daily_windows(spec, mkt.dates, [69], need_hf=False)
    got = engine.jeffreys(jb, outputs=("weights", "status"))
    assert got["status"][0] != 0


def test_gpu_invalid_arguments_raise(engine):
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import WindowBatch
    mkt = generate_market(8, 60, seed=5)
    upload_synthetic(engine, mkt)
    bad = WindowBatch(rolling_window=252, day_row=np.array([59], dtype=np.int32),
                      span_days=np.array([350], dtype=np.int32), hf_lo=None, hf_hi=None)
    with pytest.raises(ValueError):
        engine.jeffreys(bad)


@pytest.mark.parametrize("world", [3])
def test_gpu_shard_slices_equal_unsharded(engine, world):
    """Date-range shards (own days + halo resident only) reproduce the unsharded weights: the per-rank row
    offsets of sharding.engine_compute are exact.  Values agree to rounding only (1e-12), not bit for bit:
    the Gram kernel's block grid (window-overlap reuse) is anchored to the resident slice, so the same
    products are summed in a different association."""
    import torch
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.sharding import engine_compute, make_shard
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(40, 120, seed=91)
    spec = dict(weighting_strategy="conjugate_hf_vix_vw", size=40, risk_aversion=5, rolling_window=60,
                rolling_window_frequency="daily", mcm_scaling=1)
    d_idx = list(range(70, 120))
    upload_synthetic(engine, mkt)
    full = engine.conjugate(plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7),
                            outputs=("weights",))["weights"]
    compute = engine_compute(engine, mkt, spec, hf_lookback_days=7)
    parts = []
    for r in range(world):
        sh = make_shard(d_idx, spec["rolling_window"], r, world, mkt.hf_ts, mkt.dates, 7)
        parts.append(compute(sh).cpu().numpy())
    got = np.concatenate(parts, axis=0)
    assert got.shape == full.shape and relerr(got, full) <= 1e-12
    jspec = dict(spec, weighting_strategy="jeffreys")
    upload_synthetic(engine, mkt)
    fullj = engine.jeffreys(plan_daily_windows(jspec, mkt.dates, d_idx, need_hf=False), outputs=("weights",))["weights"]
    cj = engine_compute(engine, mkt, jspec)
    gotj = np.concatenate([cj(make_shard(d_idx, 60, r, world)).cpu().numpy() for r in range(world)], axis=0)
    assert gotj.shape == fullj.shape and relerr(gotj, fullj) <= 1e-12


def test_gpu_async_upload_matches_blocking(engine):
    import torch
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import ffill_rows, plan_daily_windows
    mkt = generate_market(24, 90, seed=17)
    spec = dict(weighting_strategy="conjugate_hf_epu_vw", size=24, risk_aversion=5, rolling_window=60,
                rolling_window_frequency="daily", mcm_scaling=1)
    d_idx = list(range(60, 90))
    arrays = dict(prices=mkt.prices, caps=mkt.caps, hf_prices=mkt.hf_prices, mcm=np.stack([mkt.vix, mkt.epu]),
                  rf_row=ffill_rows(mkt.dates, mkt.dates, mkt.rf))
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts)
    engine.upload_market(**arrays)
    ref = engine.conjugate(batch, outputs=("weights",))["weights"]
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in arrays.items()}
    for _ in range(3):       # buffer reuse + copy-stream ordering across repeated uploads
        engine.upload_market(**{k: v.numpy() for k, v in pinned.items()}, async_copy=True)
        jb = plan_daily_windows(dict(spec, weighting_strategy="jeffreys"), mkt.dates, d_idx, need_hf=False)
        engine.jeffreys(jb, outputs=("weights",))
        got = engine.conjugate(batch, outputs=("weights",))["weights"]
        engine.synchronize()
        assert np.array_equal(got, ref)

@pytest.mark.parametrize("segments,order,n_days", [(2, "sorted", 200), (5, "sorted", 200), (8, "sorted", 200),
                                                   (8, "shuffled", 200), (8, "sorted", 1100), ("waves", "sorted", 2000),
                                                   ([0.1, 0.1, 0.55, 0.9], "sorted", 200)])
def test_gpu_pipelined_upload_matches_blocking(engine, segments, order, n_days):
    """Segmented asynchronous intraday upload: the conjugate statistics / Gram stages of the windows whose bars
    have arrived run while later segments are still being copied.  Same kernels, same descriptors: bit-identical
    to the blocking upload, for date-sorted batches (one chunk of windows per segment) and for shuffled ones
    (everything waits for the last segment).  The 1,100-day case has more windows than the solver works on at a
    time (6 per SM), so the full waves that are ready are solved between the segments."""
    import torch
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import ffill_rows, plan_daily_windows
    mkt = generate_market(40, n_days, seed=23)
    spec = dict(weighting_strategy="conjugate_hf_vix_vw", size=40, risk_aversion=5, rolling_window=60,
                rolling_window_frequency="daily", mcm_scaling=1)
    d_idx = list(range(60, n_days))
    if order == "shuffled":
        d_idx = list(np.random.default_rng(3).permutation(d_idx))
    arrays = dict(prices=mkt.prices, caps=mkt.caps, hf_prices=mkt.hf_prices, mcm=np.stack([mkt.vix, mkt.epu]),
                  rf_row=ffill_rows(mkt.dates, mkt.dates, mkt.rf))
    batch = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
    outs = ("weights", "scalars", "status", "w1")
    engine.upload_market(**arrays)
    ref = engine.conjugate(batch, outputs=outs)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in arrays.items()}
    if isinstance(segments, int):
        engine.set_upload_pipeline(segments, 0)
    else:
        # explicit boundaries: wave-aligned (one solver wave of ready windows per segment, then halves) or given
        engine.set_upload_pipeline(8, 0)
        fr = engine.plan_upload_fractions(batch, mkt.hf_prices.shape[0]) if segments == "waves" else segments
        assert fr is not None and len(fr) <= 8 and all(b >= a for a, b in zip(fr, fr[1:]))
        engine.set_upload_fractions(fr)
    try:
        for _ in range(2):
            engine.upload_market(**{k: v.numpy() for k, v in pinned.items()}, async_copy=True)
            got = engine.conjugate(batch, outputs=outs)
            engine.synchronize()
            for k in outs:
                assert np.array_equal(got[k], ref[k]), k
            # a second batch on the same (now complete) upload takes the ordinary path
            got2 = engine.conjugate(batch, outputs=("weights",))
            assert np.array_equal(got2["weights"], ref["weights"])
        # outputs that the pipelined path does not serve (T, S0) fall back to waiting for the whole block
        engine.upload_market(**{k: v.numpy() for k, v in pinned.items()}, async_copy=True)
        n0a, s0a = engine.hf_cov(batch)
        engine.upload_market(**arrays)
        n0b, s0b = engine.hf_cov(batch)
        assert np.array_equal(n0a, n0b) and np.array_equal(s0a, s0b)
    finally:
        engine.set_upload_pipeline()
        engine.set_upload_fractions(None)

def test_gpu_async_outputs_match_blocking(engine):
    """bp_set_async_outputs: batched calls with page-locked host outputs return once queued (so the host plans the
    conjugate batch while the GPU runs Jeffreys); after synchronize() the buffers hold the blocking results.
    Three batches in a row cycle both descriptor slots."""
    import torch
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(30, 160, seed=29)
    upload_synthetic(engine, mkt)
    spec = dict(weighting_strategy="conjugate_hf_vix_vw", size=30, risk_aversion=5, rolling_window=60,
                rolling_window_frequency="daily", mcm_scaling=1)
    jspec = dict(spec, weighting_strategy="jeffreys")
    d_idx = list(range(60, 160))
    cb = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
    jb = plan_daily_windows(jspec, mkt.dates, d_idx, need_hf=False)
    jb2 = plan_daily_windows(dict(jspec, rolling_window=40), mkt.dates, d_idx, need_hf=False)
    ref = [engine.jeffreys(jb, outputs=("weights", "status")), engine.conjugate(cb, outputs=("weights", "status")),
           engine.jeffreys(jb2, outputs=("weights", "status"))]
    W, N = len(d_idx), 30
    bufs = [{"weights": torch.zeros(W, N, dtype=torch.float64).pin_memory().numpy(),
             "status": torch.ones(W, dtype=torch.int32).pin_memory().numpy()} for _ in range(3)]
    engine.set_async_outputs(True)
    try:
        engine.jeffreys(jb, outputs=("weights", "status"), into=bufs[0])
        engine.conjugate(cb, outputs=("weights", "status"), into=bufs[1])
        engine.jeffreys(jb2, outputs=("weights", "status"), into=bufs[2])
        engine.synchronize()
    finally:
        engine.set_async_outputs(False)
    for got, want in zip(bufs, ref):
        assert np.array_equal(got["weights"], want["weights"]) and np.array_equal(got["status"], want["status"])


@pytest.mark.parametrize("dates", ["consecutive", "every3rd", "random"])
def test_gpu_window_overlap_reuse_matches_oracle(engine, dates):
    """Batches of >= 32 windows take the block-reuse path of the Gram kernel (whole blocks of rows are added
    from precomputed tiles, only head/tail rows are contracted): consecutive dates (regular one-day intraday
    blocks), strided dates, and random dates (generic 64-row intraday blocks), for the conjugate prior with a
    7-day HF window and for Jeffreys with a long daily window (128-row blocks)."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    n_assets = 40
    mkt = generate_market(n_assets, 760, seed=777)
    rng = np.random.default_rng(5)
    if dates == "consecutive":
        d_idx = list(range(700, 760))
    elif dates == "every3rd":
        d_idx = list(range(640, 760, 3))
    else:
        d_idx = sorted(rng.choice(np.arange(610, 760), size=48, replace=False).tolist())
    cols = np.arange(n_assets)
    spec = dict(weighting_strategy="conjugate_hf_vix_vw", size=n_assets, risk_aversion=5, rolling_window=252,
                rolling_window_frequency="daily", mcm_scaling=1)
    jspec = dict(spec, weighting_strategy="jeffreys", rolling_window=600)
    upload_synthetic(engine, mkt)
    cb = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
    got = engine.conjugate(cb, outputs=("weights", "status", "S1", "T", "S0"))
    jb = plan_daily_windows(jspec, mkt.dates, d_idx, need_hf=False)
    gotj = engine.jeffreys(jb, outputs=("weights", "status", "T"))
    assert not got["status"].any() and not gotj["status"].any()
    for i in range(0, len(d_idx), 5):
        ref = bo.conjugate_window(spec, mkt, d_idx[i], cols, hf_lookback_days=7)
        for k in ("T", "S0", "S1", "weights"):
            assert relerr(got[k][i], ref[k]) <= TOL, (k, i)
        refj = bo.jeffreys_window(jspec, mkt, d_idx[i], cols)
        assert relerr(gotj["T"][i], refj["T"]) <= TOL
        assert relerr(gotj["weights"][i], refj["weights"]) <= TOL


@pytest.mark.parametrize("n_windows", [6, 40])
def test_gpu_weekly_batched_matches_oracle(engine, n_windows):
    """Weekly windows (the reference's shipped configuration: resample('W').last(), :151-153) in the batched
    engine: shared weekly return rows + one per-date row, for conjugate and Jeffreys."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_weekly_windows
    n = 20
    mkt = generate_market(n, 460, seed=1234)
    spec = dict(weighting_strategy="conjugate_hf_epu_vw", size=n, risk_aversion=5, rolling_window=70,
                rolling_window_frequency="weekly", mcm_scaling=2)
    jspec = dict(spec, weighting_strategy="jeffreys")
    d_idx = list(range(460 - n_windows, 460))
    cols = np.arange(n)
    upload_synthetic(engine, mkt)
    rows, cb = plan_weekly_windows(spec, mkt.dates, d_idx, mkt.dates, mkt.rf, np.stack([mkt.vix, mkt.epu]), mkt.hf_ts)
    engine.set_resampled(rows)
    got = engine.conjugate(cb, outputs=("weights", "t", "scalars", "status", "T", "S1"))
    _, jb = plan_weekly_windows(jspec, mkt.dates, d_idx, mkt.dates, mkt.rf, None, need_hf=False)
    gotj = engine.jeffreys(jb, outputs=("weights", "status"))
    assert not got["status"].any() and not gotj["status"].any()
    for i, d in enumerate(d_idx):
        ref = bo.conjugate_window(spec, mkt, d, cols)
        assert relerr(got["t"][i], ref["t"]) <= TOL and relerr(got["T"][i], ref["T"]) <= TOL
        assert abs(got["scalars"][i][0] - ref["n0"]) <= TOL * ref["n0"]
        assert relerr(got["S1"][i], ref["S1"]) <= TOL
        assert relerr(got["weights"][i], ref["weights"]) <= TOL
        refj = bo.jeffreys_window(jspec, mkt, d, cols)
        assert relerr(gotj["weights"][i], refj["weights"]) <= TOL


@pytest.mark.parametrize("n_assets,rolling_window,n_windows,group", [
    (10, 60, 37, 8), (33, 120, 50, 8), (50, 252, 64, 8), (130, 252, 21, 5), (200, 400, 35, 8), (96, 300, 17, 2)])
def test_gpu_jeffreys_chain_matches_oracle_and_per_window_path(engine, n_assets, rolling_window, n_windows, group):
    """Jeffreys windows of consecutive dates solved relative to a factorised base window (Woodbury, rank 2k+4)
    against (i) the oracle, 1e-9, and (ii) the per-window factorisation path of the same library, 1e-11; group sizes
    that do not divide the batch, N below / not a multiple of the 32-row panel, a group of two."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    D = rolling_window + n_windows + 3
    mkt = generate_market(n_assets, D, seed=8100 + n_assets, bars_per_day=2)
    spec = dict(weighting_strategy="jeffreys", size=n_assets, risk_aversion=3, turnover_cost=15,
                rebalancing_frequency="daily", rolling_window=rolling_window, rolling_window_frequency="daily",
                mcm_scaling=None, display_name="x")
    d_idx = list(range(D - n_windows, D))
    cols = np.arange(n_assets)
    upload_synthetic(engine, mkt)
    batch = plan_daily_windows(spec, mkt.dates, d_idx, need_hf=False)
    try:
        engine.set_jeffreys_chain(0)
        engine.solve_work()
        plain = engine.jeffreys(batch, outputs=("weights", "nu", "w1", "scalars", "status"))
        assert engine.solve_work() == {"factored": float(n_windows), "chained": 0.0}
        engine.set_jeffreys_chain(group)
        got = engine.jeffreys(batch, outputs=("weights", "nu", "w1", "scalars", "status"))
        work = engine.solve_work()
    finally:
        engine.set_jeffreys_chain(8)
    n_base = -(-n_windows // group)
    assert work == {"factored": float(n_base), "chained": float(n_windows - n_base)}
    assert not got["status"].any() and not plain["status"].any()
    for k in ("weights", "nu", "w1"):
        assert relerr(got[k], plain[k]) <= 1e-11, k
    assert np.max(np.abs(got["scalars"][:, 8] - plain["scalars"][:, 8]) / np.abs(plain["scalars"][:, 8])) <= 1e-10   # v1
    for i in (0, 1, group - 1, group, n_windows // 2, n_windows - 2, n_windows - 1):
        ref = bo.jeffreys_window(spec, mkt, d_idx[i], cols)["weights"]
        assert relerr(got["weights"][i], ref) <= TOL, f"window {i}"


def test_gpu_jeffreys_chain_guard_near_rank_limit(engine):
    """n - 1 close to N: the down-dated matrices of a chain group approach rank deficiency (leverage -> 1) and the
    Woodbury pivots lose digits without ever being exactly zero, so such batches are NOT chained
    (rolling_window - 1 < N + 2*group + 8): every window takes the per-window factorisation the parity is pinned on."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    N, n, W = 60, 80, 40
    mkt = generate_market(N, n + W + 3, seed=8177, bars_per_day=2)
    spec = dict(weighting_strategy="jeffreys", size=N, risk_aversion=3, turnover_cost=15, rebalancing_frequency="daily",
                rolling_window=n, rolling_window_frequency="daily", mcm_scaling=None, display_name="x")
    d_idx = list(range(n + 3, n + W + 3))
    upload_synthetic(engine, mkt)
    engine.solve_work()
    got = engine.jeffreys(plan_daily_windows(spec, mkt.dates, d_idx, need_hf=False), outputs=("weights", "status"))
    assert engine.solve_work() == {"factored": float(W), "chained": 0.0}
    assert not got["status"].any()
    for i in (0, 9, W - 1):
        ref = bo.jeffreys_window(spec, mkt, d_idx[i], np.arange(N))["weights"]
        assert relerr(got["weights"][i], ref) <= TOL, f"window {i}"


def test_gpu_jeffreys_chain_only_for_consecutive_dates(engine):
    """Every other date: not consecutive -> per-window path (no window is chained), results still match the oracle."""
    from incorporating_different_sources_b200.engine import upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    mkt = generate_market(24, 200, seed=812, bars_per_day=2)
    spec = dict(weighting_strategy="jeffreys", size=24, risk_aversion=3, turnover_cost=15, rebalancing_frequency="daily",
                rolling_window=100, rolling_window_frequency="daily", mcm_scaling=None, display_name="x")
    d_idx = list(range(110, 200, 2))
    upload_synthetic(engine, mkt)
    engine.solve_work()
    got = engine.jeffreys(plan_daily_windows(spec, mkt.dates, d_idx, need_hf=False), outputs=("weights", "status"))
    assert engine.solve_work()["chained"] == 0.0
    for i in (0, 7, len(d_idx) - 1):
        assert relerr(got["weights"][i], bo.jeffreys_window(spec, mkt, d_idx[i], np.arange(24))["weights"]) <= TOL
