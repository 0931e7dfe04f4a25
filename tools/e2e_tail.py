"""End-to-end step (pinned host buffers -> weights in host memory) for different tail plans of the segmented upload:
wall clock of 5 steps each and, with BP_TIMELINE=1, the device timeline of one step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from incorporating_different_sources_b200.engine import BayesEngine
from incorporating_different_sources_b200.windows import ffill_rows, plan_daily_windows, plan_wave_fractions, trim_intraday

class A: pass
args = A(); args.n_assets = 500; args.windows = 4150; args.hf_days = 7
mkt, conj, jeff, d_idx = bench.make_workload(args, 0)
eng = BayesEngine(0)
def pin(a):
    t = torch.empty(a.shape, dtype=torch.float64).pin_memory(); v = t.numpy(); v[...] = a; return t, v
cb = plan_daily_windows(conj, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7)
jb = plan_daily_windows(jeff, mkt.dates, d_idx, need_hf=False)
lo, hi = trim_intraday(cb)
keep = []; host = {}
for name, arr in (("prices", mkt.prices), ("caps", mkt.caps), ("hf_prices", mkt.hf_prices[lo:hi]),
                  ("mcm", np.stack([mkt.vix, mkt.epu])), ("rf_row", ffill_rows(mkt.dates, mkt.dates, mkt.rf))):
    t, v = pin(arr); keep.append(t); host[name] = v
W, N = len(d_idx), 500
hw_c, hw_cv = pin(np.zeros((W, N))); hw_j, hw_jv = pin(np.zeros((W, N)))
hs_ct = torch.zeros(W, dtype=torch.int32).pin_memory(); hs_jt = torch.zeros(W, dtype=torch.int32).pin_memory()
eng.set_async_outputs(True)
def step():
    eng.upload_market(**host, async_copy=True)
    eng.jeffreys(jb, outputs=("weights", "status"), into={"weights": hw_jv, "status": hs_jt.numpy()})
    eng.conjugate(cb, outputs=("weights", "status"), into={"weights": hw_cv, "status": hs_ct.numpy()})
    eng.synchronize()
wave = eng.solve_wave_windows()
for split in (1, 2, 3):
    fr = plan_wave_fractions(cb.hf_hi, hi - lo, wave, tail_split=split)
    eng.set_upload_fractions(fr)
    for _ in range(2):
        step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        step()
    dt = (time.perf_counter() - t0) / 5
    print(f"tail_split={split}: {len(fr)} segments, e2e {1e3 * dt:.2f} ms  fractions {[round(f, 4) for f in fr]}", flush=True)
