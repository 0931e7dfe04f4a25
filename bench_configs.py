#!/usr/bin/env python
"""The bench lines of BASELINE.json's other configurations and of the north-star partition (bench.py dispatches here):

    python bench.py --config C1|C3|C4|C5 [--gpus N]       one JSON line, same contract as the headline (C2) line
    python bench.py --scaling strong --gpus N            ONE 4,150-date C2 backtest split by date range over N GPUs

C1  single rebalance window, N=10, 252-day daily + 1-day 5-minute window: latency of one drop-in call
    (``calculate_conjugate_hf_mcm_portfolio`` with pandas frames) and of one batched C-ABI call with W=1.
C3  HF-heavy: N=100, 252-trading-day intraday look-back (366 calendar days, ~19.6k five-minute returns per window),
    4,150 consecutive rebalance dates, conjugate prior (the reference reaches this look-back through
    ``conjugate_prior_S_df=``, :299-318).
C4  strategy sweep: N in {5,10,25,50,100,500} x {conjugate with constant MCM, conjugate+VIX, conjugate+EPU, Jeffreys}
    x two date splits (2007-2015, 2015-2023).
C5  64 independent synthetic paths x 4,150 dates x N=500 (conjugate + Jeffreys), 64 / n_gpus paths per GPU, two
    engines per GPU so that path k+1 uploads while path k computes; weights all-gathered per path with NCCL.
strong  each rank owns a contiguous date range plus its halo (sharding.ShardedBacktest) and evaluates weights AND the
    loop body of its range; weights, returns and turnover are all-gathered with NCCL.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

import bench as B


# ----------------------------------------------------------------------------- shared plumbing
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench: no CUDA device; the CUDA path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            with B._StdoutToStderr():
                dist.init_process_group("nccl", device_id=self.dev)
                warm = torch.zeros(1, device=self.dev)
                dist.all_reduce(warm)
                torch.cuda.synchronize()
        self.pinned = []

    def pin(self, a):
        t = self.torch.empty(a.shape, dtype=self.torch.float64).pin_memory()
        v = t.numpy()
        v[...] = a
        self.pinned.append(t)
        return v

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def time_device(self, step, steps, warmup):
        """W untimed steps, then K steps between CUDA events with a barrier + synchronize on both sides; max over
        ranks; clocks sampled during the timed region."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        sampler = B.ClockSampler(self.local)
        sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        self.barrier()
        clocks = sampler.stop()
        return self.max_over_ranks(e0.elapsed_time(e1) / steps), clocks

    def time_wall(self, step, steps):
        step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        return self.max_over_ranks((time.perf_counter() - t0) / steps)

    def h2d_ceiling(self, nbytes, reps=3):
        """Plain cudaMemcpyAsync of `nbytes` of pinned memory per rank, all ranks at once: the platform's H2D limit."""
        torch = self.torch
        n = int(nbytes // 8)
        src = torch.empty(n, dtype=torch.float64).pin_memory()
        dst = torch.empty(n, dtype=torch.float64, device=self.dev)
        dst.copy_(src, non_blocking=True)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = self.max_over_ranks((time.perf_counter() - t0) / reps)
        del src, dst
        return {"bytes_per_rank": n * 8, "ranks": self.world, "ms": dt * 1e3,
                "aggregate_gbs": self.world * n * 8 / dt / 1e9, "per_rank_gbs": n * 8 / dt / 1e9,
                "what": "torch copy_(non_blocking) of pinned host memory = cudaMemcpyAsync, all ranks at once, max over ranks"}

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def peaks(ctx):
    pk = {}
    try:
        with open(os.path.join(B.ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
    except Exception:
        pass
    hbm = float(pk.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in pk else "fallback 6650 GB/s (B200_PROFILING.md)"
    return hbm, src, B.measure_dgemm_peak(ctx.torch, ctx.dev)


def dev_outputs(ctx, W, N):
    t = ctx.torch
    return {"weights": t.empty((W, N), dtype=t.float64, device=ctx.dev), "status": t.empty((W,), dtype=t.int32, device=ctx.dev)}


def stage_ms(eng, k):
    st = eng.stage_times()
    return {name: st[name]["ms"] / k for name in ("logret", "prep", "gram", "solve", "chain")}


def chol_flops(N):
    return N ** 3 / 3.0 + 4.0 * N * N


def roofline_of(stages, ms, N, gwork, swork, prep_bytes, dgemm_tf, hbm_peak, conv_gram_flops=None):
    """The dominant stage of the step.  Gram: the DMMA work REALLY ISSUED (k-rows contracted by the window launches +
    by the block precompute, lower-triangular 128x128 tile pairs) over the stage time; solve: N^3/3 + 4N^2 per window the
    kernel factorises (windows solved relative to a base by the Jeffreys chain are not charged); prep: algorithmic bytes."""
    nt = (N + 127) // 128
    gram_exec = 2.0 * (nt * (nt + 1) // 2) * 128 * 128 * (gwork["k_rows"] + gwork["precompute_rows"])
    solve = swork["factored"] * chol_flops(N)
    cand = {
        "gram": ("gram_dmma_kernel (block precompute + window launches) + scans", "tensor", gram_exec / 1e12, dgemm_tf, "TFLOP/s"),
        "solve": ("chol_solve_kernel / chol_cluster_kernel", "tensor", solve / 1e12, dgemm_tf, "TFLOP/s"),
        "prep": ("window_prep / band kernels", "hbm", prep_bytes / 1e9, hbm_peak, "GB/s"),
    }
    name = max(cand, key=lambda k: stages[k])
    kernel, bound, work, peak, unit = cand[name]
    t = stages[name] * 1e-3
    ach = work / t if t > 0 else 0.0
    r = {"kernel": kernel, "stage": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
         "frac": ach / peak if peak else None, "traffic": None, "ms_per_step": stages[name],
         "share_of_step": stages[name] / ms if ms else None,
         "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run" if bound == "tensor" else "MEASURED_PEAKS.json hbm_gbs",
         "windows_factorised": swork["factored"], "windows_chained": swork["chained"],
         "gram_executed_gflop": gram_exec / 1e9}
    if conv_gram_flops:
        r["gram_conventional_gflop"] = conv_gram_flops / 1e9
        r["gram_conventional_tflops"] = conv_gram_flops / (stages["gram"] * 1e-3) / 1e12 if stages["gram"] > 0 else None
        r["note"] = ("conventional = SURVEY 8(d), every window contracted from scratch (2 N^2 K flops); the executed work is smaller "
                     "because overlapping windows share block tiles, so the conventional rate is not a utilisation")
    return r


# ----------------------------------------------------------------------------- strong scaling (north-star partition)
def run_strong(args):
    from incorporating_different_sources_b200.engine import BayesEngine
    from incorporating_different_sources_b200.sharding import ShardedBacktest
    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    hbm_peak, hbm_src, dgemm_tf = peaks(ctx)
    mkt, conj, jeff, d_idx = B.make_workload(args, 0)          # the SAME path on every rank
    N, W = args.n_assets, len(d_idx)
    eng = BayesEngine(ctx.local)
    sb = ShardedBacktest(eng, mkt, conj, jeff, d_idx, ctx.rank, ctx.world, hf_lookback_days=args.hf_days, pin=ctx.pin)
    sb.upload()
    out_c, out_j = dev_outputs(ctx, sb.n_ext, N), dev_outputs(ctx, sb.n_ext, N)
    state = {}

    def step_device():
        eng.prepare_market()
        rows = sb.compute(out_c, out_j, loop=True)
        state["rows"] = sb.gather(rows, dist) if ctx.world > 1 else rows

    ms, clocks = None, None
    eng.set_stage_timing(False)
    ms, clocks = ctx.time_device(step_device, args.steps, args.warmup)
    # per-stage times of this rank's share (second, untimed pass with the stage timers on)
    eng.set_stage_timing(True)
    eng.stage_times()
    eng.gram_work()
    eng.solve_work()
    l0 = eng.launch_count
    step_device()
    torch.cuda.synchronize()
    stages = stage_ms(eng, 1)
    gwork, swork = eng.gram_work(), eng.solve_work()
    launches = eng.launch_count - l0 + (1 if ctx.world > 1 else 0)      # + the one all-gather
    eng.set_stage_timing(False)
    bad = ctx.sum_over_ranks(float((out_c["status"] != 0).sum().item() + (out_j["status"] != 0).sum().item()))

    host_out = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in state["rows"]] if ctx.rank == 0 else None

    def step_e2e():
        sb.upload(async_copy=True)
        rows = sb.compute(out_c, out_j, loop=True)
        full = sb.gather(rows, dist) if ctx.world > 1 else rows
        if ctx.rank == 0:
            for h, x in zip(host_out, full):
                h.copy_(x, non_blocking=True)
        eng.synchronize()
        torch.cuda.synchronize()

    e2e_steps = max(1, min(args.steps, 3))
    e2e_s = ctx.time_wall(step_e2e, e2e_steps)
    h2d_total = ctx.sum_over_ranks(float(sb.h2d_bytes))
    ceiling = ctx.h2d_ceiling(sb.h2d_bytes)
    value = 2 * W / (ms * 1e-3)
    if ctx.rank == 0:
        work = B.algorithmic_work(N, W, conj["rolling_window"], int((sb.cb.hf_hi - sb.cb.hf_lo - 1).max()), jeff["rolling_window"])
        share = 1.0 / ctx.world
        roof = roofline_of(stages, ms, N, gwork, swork, work["prep_bytes"] * share, dgemm_tf, hbm_peak, work["gram_flops"] * share)
        roof["rank_note"] = "rank 0's share of the backtest (1 / n_gpus of the windows, halo included) over rank 0's stage time"
        full = state["rows"]
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(B.config_dict(args, ctx.world, conj, jeff, args.hf_days),
                           sharding=f"ONE backtest, {ctx.world} contiguous date ranges (+ {jeff['rolling_window'] - 1} daily rows and "
                                    f"the intraday look-back as halo, + the predecessor window for the loop body); weights, "
                                    f"portfolio returns and turnover all-gathered (NCCL)",
                           windows_per_rank=[2 * c for c in sb.counts], loop_body=True),
            "clocks": clocks,
            "e2e": {"value": 2 * W / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": int(h2d_total),
                    "d2h_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host_out)), "ms_per_step": e2e_s * 1e3,
                    "steps": e2e_steps, "inputs": "pre-pinned host buffers (pinning outside the timed region)",
                    "h2d_ceiling": ceiling,
                    "h2d_ms_at_ceiling": sb.h2d_bytes / (ceiling["per_rank_gbs"] * 1e9) * 1e3},
            "gpu_launches": int(launches),
            "roofline": roof,
            "stages_rank0_ms": stages,
            "windows_flagged_singular": int(bad),
            "gathered": {"weights": list(full[0].shape), "returns": list(full[2].shape), "turnover": list(full[4].shape)},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    ctx.finish()


# ----------------------------------------------------------------------------- C5: 64 independent paths
def run_c5(args):
    from incorporating_different_sources_b200.engine import BayesEngine
    from incorporating_different_sources_b200.windows import ffill_rows, plan_daily_windows, trim_intraday
    ctx = Ctx()
    torch, dist = ctx.torch, ctx.dist
    hbm_peak, hbm_src, dgemm_tf = peaks(ctx)
    n_paths = args.paths
    if n_paths % ctx.world:
        raise SystemExit("--paths must be a multiple of the number of GPUs")
    per_gpu = n_paths // ctx.world
    pool = max(1, min(args.path_pool, per_gpu))
    N = args.n_assets
    hosts, batches = [], []
    for k in range(pool):
        mkt, conj, jeff, d_idx = B.make_workload(args, 1000 + ctx.rank * pool + k)
        cb = plan_daily_windows(conj, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=args.hf_days)
        jb = plan_daily_windows(jeff, mkt.dates, d_idx, need_hf=False)
        lo, hi = trim_intraday(cb)
        hosts.append(dict(prices=ctx.pin(mkt.prices), caps=ctx.pin(mkt.caps), hf_prices=ctx.pin(mkt.hf_prices[lo:hi]),
                          mcm=ctx.pin(np.stack([mkt.vix, mkt.epu])), rf_row=ctx.pin(ffill_rows(mkt.dates, mkt.dates, mkt.rf))))
        batches.append((cb, jb))
        n_hf = hi - lo
        if k == 0 and ctx.rank == 0:
            parity_ref = (mkt, conj, jeff)           # the first market of rank 0 is re-checked against the oracle below
        del mkt
    W = len(d_idx)
    h2d_path = int(sum(v.nbytes for v in hosts[0].values()))
    # each engine computes on a stream of its own: on a shared stream the upload of path k+1 would queue behind the
    # kernels of path k (the copy stream waits for everything queued on the compute stream before it overwrites the market)
    streams = [torch.cuda.Stream(device=ctx.dev) for _ in range(2)]
    engs = [BayesEngine(ctx.local, stream=streams[s]) for s in range(2)]
    outs = [(dev_outputs(ctx, W, N), dev_outputs(ctx, W, N)) for _ in range(2)]
    mine = [torch.empty((2, W, N), dtype=torch.float64, device=ctx.dev) for _ in range(2)]
    gathered = torch.empty((ctx.world, 2, W, N), dtype=torch.float64, device=ctx.dev) if ctx.world > 1 else None
    host_w = [torch.empty((2, W, N), dtype=torch.float64).pin_memory() for _ in range(2)]
    fr = engs[0].plan_upload_fractions(batches[0][0], n_hf)

    def run_path(k, slot, upload):
        eng, (oc, oj) = engs[slot], outs[slot]
        cb, jb = batches[k % pool]
        if upload:
            eng.set_upload_fractions(fr)
            eng.upload_market(**hosts[k % pool], async_copy=True)
        else:
            eng.prepare_market()
        eng.jeffreys(jb, outputs=("weights", "status"), into=oj)
        eng.conjugate(cb, outputs=("weights", "status"), into=oc)

    def collect(slot, to_host):
        engs[slot].synchronize()                     # host-side: the outputs of this path are complete
        mine[slot][0].copy_(outs[slot][0]["weights"])
        mine[slot][1].copy_(outs[slot][1]["weights"])
        if ctx.world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), mine[slot].view(-1))
        if to_host:
            host_w[slot].copy_(mine[slot], non_blocking=True)

    # device-resident: the two engines keep the two pool markets resident (path k uses market k mod pool)
    for s in range(2):
        engs[s].upload_market(**hosts[s % pool])

    def step_device():
        for k in range(per_gpu):
            run_path(k % 2 if pool > 1 else 0, k % 2, upload=False)
            if k:
                collect((k - 1) % 2, False)
        collect((per_gpu - 1) % 2, False)

    def step_e2e():
        # path k+1 is queued (upload + compute on its engine's streams) before path k is collected
        run_path(0, 0, upload=True)
        for k in range(1, per_gpu):
            engs[(k - 1) % 2].wait_upload()          # the link is free: do not share it between two transfers
            run_path(k, k % 2, upload=True)
            collect((k - 1) % 2, True)
        collect((per_gpu - 1) % 2, True)
        torch.cuda.synchronize()

    ms, clocks = ctx.time_device(step_device, args.steps, args.warmup)
    launches0 = sum(e.launch_count for e in engs)
    step_device()
    torch.cuda.synchronize()
    launches = sum(e.launch_count for e in engs) - launches0 + per_gpu * (3 if ctx.world > 1 else 2)
    bad = ctx.sum_over_ranks(float(sum((o["status"] != 0).sum().item() for pair in outs for o in pair)))
    parity = None
    if ctx.rank == 0 and not args.no_cpu:
        # engine 0 holds rank 0's first market: a few of its windows against the CPU oracle (1e-9 bar)
        from oracle import bayes_oracle as bo
        mkt0, conj0, jeff0 = parity_ref
        cols = np.arange(N)
        wc, wj = outs[0][0]["weights"].cpu().numpy(), outs[0][1]["weights"].cpu().numpy()
        parity = 0.0
        for i in np.linspace(0, W - 1, 4).round().astype(int):
            for got, ref in ((wc[i], bo.conjugate_window(conj0, mkt0, int(d_idx[i]), cols, hf_lookback_days=args.hf_days)["weights"]),
                             (wj[i], bo.jeffreys_window(jeff0, mkt0, int(d_idx[i]), cols)["weights"])):
                parity = max(parity, float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
    e2e_steps = max(1, min(args.steps, 2))
    e2e_s = ctx.time_wall(step_e2e, e2e_steps)
    ceiling = ctx.h2d_ceiling(h2d_path)
    total_windows = n_paths * 2 * W
    if ctx.rank == 0:
        conj, jeff = B.make_specs(N)
        work = B.algorithmic_work(N, W, conj["rolling_window"], int((batches[0][0].hf_hi - batches[0][0].hf_lo - 1).max()),
                                  jeff["rolling_window"])
        solve_tf = per_gpu * work["solve_flops"] / (ms * 1e-3) / 1e12
        line = {
            "metric": B.METRIC, "value": total_windows / (ms * 1e-3), "unit": B.UNIT, "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(B.config_dict(args, ctx.world, conj, jeff, args.hf_days),
                           workload=f"C5: {n_paths} independent synthetic paths x {W} rebalance dates x N={N}, conjugate(VIX)+Jeffreys",
                           sharding=f"{per_gpu} paths per GPU, no halo; two engines per GPU (path k+1 uploads while path k "
                                    f"computes); weights of every path all-gathered (NCCL)",
                           windows_per_step=total_windows, distinct_markets_per_gpu=pool,
                           note="the paths of one GPU cycle through `distinct_markets_per_gpu` generated markets (host "
                                "memory: 1.3 GB pinned per market); every path is uploaded and computed in full"),
            "clocks": clocks,
            "e2e": {"value": total_windows / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": int(h2d_path * n_paths),
                    "d2h_bytes_per_step": int(host_w[0].numel() * 8 * n_paths), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "inputs": "pre-pinned host buffers", "h2d_ceiling": ceiling,
                    "h2d_ms_at_ceiling": per_gpu * h2d_path / (ceiling["per_rank_gbs"] * 1e9) * 1e3},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "whole step (Gram + solve of every path)", "bound": "tensor",
                         "achieved": per_gpu * (work["gram_flops"] + work["solve_flops"]) / (ms * 1e-3) / 1e12, "peak": dgemm_tf,
                         "unit": "TFLOP/s", "frac": per_gpu * (work["gram_flops"] + work["solve_flops"]) / (ms * 1e-3) / 1e12 / dgemm_tf,
                         "traffic": None, "solve_conventional_tflops": solve_tf,
                         "note": "conventional flops per GPU (SURVEY 8(d)); > 1 reflects block reuse and the Jeffreys chain, "
                                 "see the C2 line for the per-kernel rooflines"},
            "windows_flagged_singular": int(bad), "parity_max_rel_err": parity,
        }
        print(json.dumps(line), flush=True)
    for e in engs:
        e.close()
    ctx.finish()


# ----------------------------------------------------------------------------- C1 / C3 / C4: one strategy, one GPU
def _spec(strategy, n, rolling_window=252, **kw):
    s = dict(weighting_strategy=strategy, size=n, risk_aversion=5, turnover_cost=15, rebalancing_frequency="daily",
             rolling_window=rolling_window, rolling_window_frequency="daily", mcm_scaling=1, display_name=strategy)
    s.update(kw)
    return s


def _time_batch(ctx, eng, run, batch, W, N, steps, warmup):
    out = dev_outputs(ctx, W, N)

    def step():
        eng.prepare_market()
        run(batch, outputs=("weights", "status"), into=out)

    eng.set_stage_timing(False)
    ms, clocks = ctx.time_device(step, steps, warmup)
    eng.set_stage_timing(True)
    eng.stage_times()
    eng.gram_work()
    eng.solve_work()
    l0 = eng.launch_count
    step()
    ctx.torch.cuda.synchronize()
    st = stage_ms(eng, 1)
    st["_gram_work"], st["_solve_work"] = eng.gram_work(), eng.solve_work()
    launches = eng.launch_count - l0
    eng.set_stage_timing(False)
    return ms, clocks, st, launches, out


def run_c3(args):
    from incorporating_different_sources_b200.engine import BayesEngine
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import ffill_rows, plan_daily_windows, trim_intraday
    from oracle import bayes_oracle as bo
    ctx = Ctx()
    torch = ctx.torch
    hbm_peak, hbm_src, dgemm_tf = peaks(ctx)
    N, W, look = 100, args.windows, 366
    n = 252
    mkt = generate_market(N, 262 + n + W, seed=3, start="2005-12-01")
    spec = _spec("conjugate_hf_vix_vw", N)
    d_idx = np.arange(mkt.n_days - W, mkt.n_days)
    cb = plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=look)
    lo, hi = trim_intraday(cb)
    host = dict(prices=ctx.pin(mkt.prices), caps=ctx.pin(mkt.caps), hf_prices=ctx.pin(mkt.hf_prices[lo:hi]),
                mcm=ctx.pin(np.stack([mkt.vix, mkt.epu])), rf_row=ctx.pin(ffill_rows(mkt.dates, mkt.dates, mkt.rf)))
    eng = BayesEngine(ctx.local)
    eng.upload_market(**host)
    ms, clocks, st, launches, out = _time_batch(ctx, eng, eng.conjugate, cb, W, N, args.steps, args.warmup)
    m_hf = int((cb.hf_hi - cb.hf_lo - 1).max())
    hw = torch.empty((W, N), dtype=torch.float64).pin_memory()
    hs = torch.empty((W,), dtype=torch.int32).pin_memory()

    # (cutting the batch along upload segments, as bench.py does for C2 with long look-backs, was measured here too:
    # 8.35 instead of 7.95 ms -- at N = 100 every sub-batch is latency bound and repeats the 252-day halo)
    subs = [(0, W, cb)]
    hwv, hsv = hw.numpy(), hs.numpy()

    def step_e2e():
        eng.set_async_outputs(True)
        eng.upload_market(**host, async_copy=True)
        eng.conjugate(cb, outputs=("weights", "status"), into={"weights": hwv, "status": hsv})
        eng.synchronize()
        eng.set_async_outputs(False)

    e2e_s = ctx.time_wall(step_e2e, max(1, min(args.steps, 3)))
    h2d = int(sum(v.nbytes for v in host.values()))
    ceiling = ctx.h2d_ceiling(h2d)
    e2e_err = float(np.max(np.abs(hwv - out["weights"].cpu().numpy())) / np.max(np.abs(hwv)))
    # parity + CPU port on a bounded sample (each oracle window contracts 19.6k x 100 intraday returns)
    cols = np.arange(N)
    sel = np.linspace(0, W - 1, max(2, args.cpu_sample // 32)).round().astype(int)
    w = out["weights"].cpu().numpy()
    worst, t0 = 0.0, time.perf_counter()
    for i in sel:
        ref = bo.conjugate_window(spec, mkt, int(d_idx[i]), cols, hf_lookback_days=look)["weights"]
        worst = max(worst, float(np.max(np.abs(w[i] - ref)) / np.max(np.abs(ref))))
    cpu_dt = time.perf_counter() - t0
    gram_flops = W * 2.0 * N * N * ((n - 1) + m_hf)
    prep_bytes = W * 8.0 * N * ((n - 1) + 2 * m_hf)
    gwork, swork = st.pop("_gram_work"), st.pop("_solve_work")
    roof = roofline_of(st, ms, N, gwork, swork, prep_bytes, dgemm_tf, hbm_peak, gram_flops)
    roof["unique_work_note"] = ("each of the %d distinct intraday return rows is contracted ONCE (day-block tiles, then suffix / "
                                "prefix scans over the tiles); a window then adds <= 3 intraday tiles + 1 daily tile" % (hi - lo))
    roof["hbm_view"] = {"unique_input_bytes": 8.0 * N * (hi - lo + mkt.n_days), "ms_at_hbm_peak": 8.0 * N * (hi - lo + mkt.n_days) / (hbm_peak * 1e9) * 1e3}
    line = {
        "metric": "posterior tangency-weight windows/sec at N=100 (HF-heavy)", "value": W / (ms * 1e-3), "unit": B.UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: HF-heavy, N=100, 252-trading-day intraday look-back (366 calendar days), conjugate(VIX)",
                   "n_assets": N, "rebalance_dates": W, "rolling_window": n, "hf_returns_per_window": m_hf,
                   "hf_lookback_calendar_days": look, "intraday_rows_uploaded": int(hi - lo),
                   "cache": "intraday block (%.2f GB) larger than L2" % ((hi - lo) * N * 8 / 1e9)},
        "clocks": clocks,
        "e2e": {"value": W / e2e_s, "unit": B.UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(hw.numel() * 8 + hs.numel() * 4),
                "ms_per_step": e2e_s * 1e3, "inputs": "pre-pinned host buffers", "conjugate_sub_batches": len(subs),
                "h2d_ceiling": ceiling, "h2d_ms_at_ceiling": h2d / (ceiling["per_rank_gbs"] * 1e9) * 1e3},
        "gpu_launches": int(launches), "roofline": roof, "stages_ms": st,
        "cpu_baseline": {"value": len(sel) / cpu_dt, "unit": B.UNIT, "cores": B.blas_threads(), "kind": "port",
                         "sample": f"{len(sel)} windows, oracle/bayes_oracle.py, one process, NumPy default BLAS threads"},
        "parity_max_rel_err": worst, "e2e_vs_device_path_max_rel_diff": e2e_err,
        "windows_flagged_singular": int((out["status"] != 0).sum().item()),
    }
    print(json.dumps(line), flush=True)
    eng.close()
    ctx.finish()


def run_c1(args):
    """One window, N=10: latency of the drop-in façade call and of the batched C-ABI call with W=1."""
    from incorporating_different_sources_b200 import portfolio_calculations as pcg
    from incorporating_different_sources_b200.engine import BayesEngine, upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    from oracle import bayes_oracle as bo
    import pandas as pd
    ctx = Ctx()
    torch = ctx.torch
    hbm_peak, hbm_src, dgemm_tf = peaks(ctx)
    N, n = 10, 252
    mkt = generate_market(N, n + 8, seed=1)
    spec = _spec("conjugate_hf_vix_vw", N)
    d = mkt.n_days - 1
    eng = BayesEngine(ctx.local)
    upload_synthetic(eng, mkt)
    cb = plan_daily_windows(spec, mkt.dates, [d], mkt.hf_ts)
    reps = 50
    ms, clocks, st, launches, out = _time_batch(ctx, eng, eng.conjugate, cb, 1, N, max(args.steps, reps), args.warmup)
    st.pop("_gram_work"), st.pop("_solve_work")
    ref = bo.conjugate_window(spec, mkt, d, np.arange(N))
    err = float(np.max(np.abs(out["weights"].cpu().numpy()[0] - ref["weights"])) / np.max(np.abs(ref["weights"])))
    # the drop-in call with pandas frames (upload of the frames + kernels + DataFrame out)
    md = mkt.market_data()
    ts = pd.Timestamp(mkt.dates[d])
    names = list(mkt.tickers)
    prices = md["stock_prices_df"].loc[:ts, names]
    caps = md["stock_market_caps_df"].loc[:ts, names]
    intr = md["stock_intraday_prices_df"]
    intr = intr.loc[:ts + pd.Timedelta(hours=23, minutes=59, seconds=59), names]
    key = "vix_prices_df" if "vix_prices_df" in md else [k for k in md if "vix" in k.lower()][0]
    mcm = md[key].loc[:ts]
    rf = md["risk_free_rate_df"]
    call = lambda: pcg.calculate_conjugate_hf_mcm_portfolio(spec, ts, caps, prices, intr, mcm, rf)
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(20):
        got = call()
    facade_s = (time.perf_counter() - t0) / 20
    ferr = float(np.max(np.abs(got["Weight"].reindex(names).to_numpy() - ref["weights"])) / np.max(np.abs(ref["weights"])))
    t0 = time.perf_counter()
    for _ in range(20):
        bo.conjugate_window(spec, mkt, d, np.arange(N))
    cpu_dt = (time.perf_counter() - t0) / 20
    h2d = int(prices.size + caps.size + intr.size + len(mcm) + len(rf)) * 8
    line = {
        "metric": "posterior tangency-weight windows/sec at N=10 (single window)", "value": 1.0 / (ms * 1e-3), "unit": B.UNIT,
        "n_gpus": 1, "steps": max(args.steps, reps), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C1: single rebalance window, conjugate(VIX), N=10, 252-day daily + 1-day 5-minute window",
                   "n_assets": N, "rolling_window": n, "hf_returns": int(ref["hf_returns"]),
                   "cache": "latency case: 60 KB of inputs, launch-bound (no L2 flush applies)"},
        "clocks": clocks,
        "e2e": {"value": 1.0 / facade_s, "unit": B.UNIT, "ms_per_step": facade_s * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": N * 8, "what": "calculate_conjugate_hf_mcm_portfolio(spec, d, caps, prices, intraday, mcm, rf) with "
                "pandas frames: frame -> array conversion, upload, kernels, DataFrame out"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "whole call (launch-latency bound: %d launches of a few microseconds)" % launches, "bound": "hbm",
                     "achieved": 8.0 * N * (n + ref["hf_returns"] + 1) / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": 8.0 * N * (n + ref["hf_returns"] + 1) / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                     "note": "one window of N=10 cannot load a B200; the figure of merit is the latency"},
        "stages_ms": st,
        "cpu_baseline": {"value": 1.0 / cpu_dt, "unit": B.UNIT, "cores": B.blas_threads(), "kind": "port",
                         "sample": "20 repetitions of the window, oracle/bayes_oracle.py"},
        "parity_max_rel_err": max(err, ferr),
    }
    print(json.dumps(line), flush=True)
    eng.close()
    ctx.finish()


def run_c4(args):
    """Strategy sweep: every cell is one batched call over the split's rebalance dates, device resident."""
    from incorporating_different_sources_b200.engine import BayesEngine, upload_synthetic
    from incorporating_different_sources_b200.synthetic import generate_market
    from incorporating_different_sources_b200.windows import plan_daily_windows
    from oracle import bayes_oracle as bo
    import dataclasses
    ctx = Ctx()
    torch = ctx.torch
    hbm_peak, hbm_src, dgemm_tf = peaks(ctx)
    W_half = args.windows // 2
    cells, total_w, total_ms, worst, e2e_total = [], 0, 0.0, 0.0, 0.0
    clocks_all = None
    launches_all = 0
    eng = BayesEngine(ctx.local)
    for N in (5, 10, 25, 50, 100, 500):
        n_j = 252 if N <= 100 else 1008                    # Jeffreys needs n - 1 >= N with margin
        look = None if N <= 50 else 7                      # conjugate: (m-1) + (n-1) >= N with margin
        n_days = n_j + 2 * W_half
        mkt = generate_market(N, n_days, seed=4000 + N)
        const = dataclasses.replace(mkt, vix=np.full_like(mkt.vix, 20.0))
        for strat, market in (("conjugate (constant MCM)", const), ("conjugate_hf_vix_vw", mkt), ("conjugate_hf_epu_vw", mkt),
                              ("jeffreys", mkt)):
            jeff = strat == "jeffreys"
            name = "jeffreys" if jeff else ("conjugate_hf_vix_vw" if strat.startswith("conjugate (") else strat)
            spec = _spec(name, N, rolling_window=n_j if jeff else 252, mcm_scaling=None if jeff else 1)
            upload_synthetic(eng, market)
            for split, lo in (("2007-2015", n_days - 2 * W_half), ("2015-2023", n_days - W_half)):
                d_idx = np.arange(lo, lo + W_half)
                batch = plan_daily_windows(spec, market.dates, d_idx, None if jeff else market.hf_ts,
                                           hf_lookback_days=look, need_hf=not jeff)
                run = eng.jeffreys if jeff else eng.conjugate
                ms, clocks, st, launches, out = _time_batch(ctx, eng, run, batch, W_half, N, args.steps, args.warmup)
                st.pop("_gram_work"), st.pop("_solve_work")
                clocks_all = clocks
                launches_all += launches
                cols = np.arange(N)
                w = out["weights"].cpu().numpy()
                err = 0.0
                for i in (0, W_half // 2, W_half - 1):
                    ref = (bo.jeffreys_window(spec, market, int(d_idx[i]), cols) if jeff else
                           bo.conjugate_window(spec, market, int(d_idx[i]), cols, hf_lookback_days=look))["weights"]
                    err = max(err, float(np.max(np.abs(w[i] - ref)) / np.max(np.abs(ref))))
                worst = max(worst, err)
                # end to end for the cell: host arrays (pageable, as a pandas frame's .to_numpy() gives them) -> weights on the host
                t0 = time.perf_counter()
                upload_synthetic(eng, market)
                run(batch, outputs=("weights", "status"))
                e2e_cell = time.perf_counter() - t0
                e2e_total += e2e_cell
                cells.append({"n_assets": N, "strategy": strat, "split": split, "windows": W_half, "e2e_ms": e2e_cell * 1e3,
                              "rolling_window": spec["rolling_window"], "hf_lookback_days": 1 if look is None and not jeff else look,
                              "ms": ms, "windows_per_s": W_half / (ms * 1e-3), "stages_ms": st,
                              "flagged": int((out["status"] != 0).sum().item()), "parity_max_rel_err": err})
                total_w += W_half
                total_ms += ms
    big = [c for c in cells if c["n_assets"] == 500]
    line = {
        "metric": "posterior tangency-weight windows/sec (strategy sweep)", "value": total_w / (total_ms * 1e-3), "unit": B.UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4: N in {5,10,25,50,100,500} x {conjugate const MCM, +VIX, +EPU, Jeffreys} x 2 date splits, "
                               "one batched call per cell, device resident", "cells": len(cells), "windows_per_cell": W_half,
                   "cache": "N=500 cells exceed L2; the small-N cells are L2 resident by nature (whole market < 126 MB)"},
        "clocks": clocks_all, "gpu_launches": int(launches_all),
        "e2e": {"value": total_w / e2e_total, "unit": B.UNIT, "ms_per_step": e2e_total * 1e3,
                "what": "per cell: blocking upload of the market from pageable host arrays + one batched call with host outputs"},
        "roofline": {"kernel": "N=500 cells: see the C2 line for the per-kernel rooflines", "bound": "tensor",
                     "achieved": None, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": None, "traffic": None,
                     "n500_windows_per_s": float(np.mean([c["windows_per_s"] for c in big])) if big else None},
        "parity_max_rel_err": worst, "cells": cells,
    }
    print(json.dumps(line), flush=True)
    eng.close()
    ctx.finish()


def dispatch(args):
    if args.scaling == "strong" and args.config == "C2":
        return run_strong(args)
    return {"C1": run_c1, "C3": run_c3, "C4": run_c4, "C5": run_c5}[args.config](args)
