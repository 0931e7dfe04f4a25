"""Device timeline of ONE end-to-end step (BP_TIMELINE=1): start offset and duration of every timed stage launch
relative to the start of the upload, for the geometric and for the wave-aligned segment plans."""
import os, sys, time
os.environ["BP_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from incorporating_different_sources_b200.engine import BayesEngine
from incorporating_different_sources_b200.windows import ffill_rows, plan_daily_windows

class A: pass
args = A(); args.n_assets = 500; args.windows = 4150; args.hf_days = 7
mkt, conj, jeff, d_idx = bench.make_workload(args, 0)
eng = BayesEngine(0)
def pin(a):
    t = torch.empty(a.shape, dtype=torch.float64).pin_memory(); v = t.numpy(); v[...] = a; return t, v
keep = []; host = {}
for name, arr in (("prices", mkt.prices), ("caps", mkt.caps), ("hf_prices", mkt.hf_prices),
                  ("mcm", np.stack([mkt.vix, mkt.epu])), ("rf_row", ffill_rows(mkt.dates, mkt.dates, mkt.rf))):
    t, v = pin(arr); keep.append(t); host[name] = v
cb = plan_daily_windows(conj, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
jb = plan_daily_windows(jeff, mkt.dates, d_idx, need_hf=False)
W, N = len(d_idx), 500
hw_c, hw_cv = pin(np.zeros((W, N))); hw_j, hw_jv = pin(np.zeros((W, N)))
hs_ct = torch.zeros(W, dtype=torch.int32).pin_memory(); hs_jt = torch.zeros(W, dtype=torch.int32).pin_memory()
hs_c, hs_j = hs_ct.numpy(), hs_jt.numpy()
eng.set_async_outputs(True)
fr = eng.plan_upload_fractions(cb, mkt.hf_prices.shape[0])
print("fractions", fr)
def step():
    eng.upload_market(**host, async_copy=True)
    eng.jeffreys(jb, outputs=("weights", "status"), into={"weights": hw_jv, "status": hs_j})
    eng.conjugate(cb, outputs=("weights", "status"), into={"weights": hw_cv, "status": hs_c})
    eng.synchronize()
for plan in (None, fr):
    eng.set_upload_fractions(plan)
    for _ in range(2):
        step()
    eng.set_stage_timing(True); eng.stage_times()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    print(f"=== plan {'geometric' if plan is None else 'waves'}: wall {1e3*(t1-t0):.2f} ms", file=sys.stderr)
    eng.stage_times()
    eng.set_stage_timing(False)
