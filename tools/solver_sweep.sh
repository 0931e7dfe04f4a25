#!/bin/bash
# compile-time sweep of the solver's shape (run on the GPU box): TPW / stages / CTAs per SM
for cfg in "2 2 6" "1 3 6" "1 2 8" "1 4 5" "2 3 4"; do
  set -- $cfg
  BP_NVCC_EXTRA="-DCH_TPW=$1 -DCH_NSTAGES=$2 -DCH_CTAS_PER_SM=$3" python -m incorporating_different_sources_b200.build --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  grep -A2 "chol_solve_kernelILi1" incorporating_different_sources_b200/build/ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '
  BP_CHOL_CLUSTER=0 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-widened --no-loop > gpurun_out/sweep.json 2> gpurun_out/sweep.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sweep.json'))
    print("cfg TPW/stages/CTAs = $cfg : solve %.2f ms step %.2f flagged %d" % (d['stages']['solve']['ms'], d['ms_per_step'], d['windows_flagged_singular']))
except Exception as e:
    print("cfg $cfg failed", e)
PY
done
python -m incorporating_different_sources_b200.build --force > /dev/null 2>&1
