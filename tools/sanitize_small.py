"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck): conjugate, Jeffreys (per-window and
chain paths), Jorion and shrinkage batches at N = 33 (not a multiple of the 32-row panels) plus one wider case.
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from incorporating_different_sources_b200.engine import BayesEngine, upload_synthetic
from incorporating_different_sources_b200.synthetic import generate_market
from incorporating_different_sources_b200.windows import plan_daily_windows
from oracle import bayes_oracle as bo

eng = BayesEngine(0, use_torch_stream=False)
for N, n, W in ((33, 90, 40), (70, 160, 19)):
    mkt = generate_market(N, n + W + 3, seed=77 + N)
    spec = dict(weighting_strategy="conjugate_hf_vix_vw", size=N, risk_aversion=4, turnover_cost=15,
                rebalancing_frequency="daily", rolling_window=n, rolling_window_frequency="daily", mcm_scaling=1,
                display_name="x")
    d_idx = list(range(n + 2, n + 2 + W))
    upload_synthetic(eng, mkt)
    cols = np.arange(N)
    c = eng.conjugate(plan_daily_windows(spec, mkt.dates, d_idx, mkt.hf_ts, hf_lookback_days=7), outputs=("weights", "status"))
    js = dict(spec, weighting_strategy="jeffreys", mcm_scaling=None)
    jb = plan_daily_windows(js, mkt.dates, d_idx, need_hf=False)
    j = eng.jeffreys(jb, outputs=("weights", "status"))
    eng.set_jeffreys_chain(0)
    j0 = eng.jeffreys(jb, outputs=("weights", "status"))
    eng.set_jeffreys_chain(8)
    jo = eng.jorion(jb, outputs=("weights", "status"))
    sh = eng.shrinkage(jb, outputs=("weights", "status"))
    i = W - 3
    e = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    print(N, "conj", e(c["weights"][i], bo.conjugate_window(spec, mkt, d_idx[i], cols, hf_lookback_days=7)["weights"]),
          "jeff chain", e(j["weights"][i], bo.jeffreys_window(js, mkt, d_idx[i], cols)["weights"]),
          "chain vs plain", e(j["weights"], j0["weights"]),
          "jorion", e(jo["weights"][i], bo.jorion_window(js, mkt, d_idx[i], cols)["weights"]),
          "shrink", e(sh["weights"][i], bo.shrinkage_window(js, mkt, d_idx[i], cols)["weights"]),
          "status", int(c["status"].any() or j["status"].any() or jo["status"].any() or sh["status"].any()))
eng.close()
