"""Where does the end-to-end step spend its time?  (host wall clock per call, pinned inputs)"""
import sys, time, os
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
eng.set_async_outputs(len(sys.argv) > 1 and sys.argv[1] == 'async')
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.upload_market(**host, async_copy=True); t1 = time.perf_counter()
    eng.jeffreys(jb, outputs=("weights", "status"), into={"weights": hw_jv, "status": hs_j}); t2 = time.perf_counter()
    eng.conjugate(cb, outputs=("weights", "status"), into={"weights": hw_cv, "status": hs_c}); t3 = time.perf_counter()
    eng.synchronize(); t4 = time.perf_counter()
    print(f"iter {it}: upload call {1e3*(t1-t0):.1f} ms, jeffreys {1e3*(t2-t1):.1f}, conjugate {1e3*(t3-t2):.1f}, sync {1e3*(t4-t3):.1f}, total {1e3*(t4-t0):.1f}")
# pure copy bandwidth of the intraday block
x = torch.empty(host["hf_prices"].shape, dtype=torch.float64, device="cuda")
src = keep[2]
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); x.copy_(src, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"pinned H2D {src.numel()*8/1e9:.2f} GB in {1e3*dt:.1f} ms = {src.numel()*8/1e9/dt:.1f} GB/s")
# stage launches / device time of one more end-to-end iteration (shows whether the pipelined path split the solve)
eng.set_stage_timing(True); eng.stage_times()
torch.cuda.synchronize(); t0 = time.perf_counter()
eng.upload_market(**host, async_copy=True)
eng.jeffreys(jb, outputs=("weights", "status"), into={"weights": hw_jv, "status": hs_j})
eng.conjugate(cb, outputs=("weights", "status"), into={"weights": hw_cv, "status": hs_c})
eng.synchronize(); t1 = time.perf_counter()
print(f"timed iteration {1e3*(t1-t0):.1f} ms; stages:", eng.stage_times())
