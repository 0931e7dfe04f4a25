#!/bin/bash
# Profiling pass of one round on the GPU box (run through gpurun; outputs land in gpurun_out/).
#   1. plain bench run (the program must exit 0 without ncu first)
#   2. ncu launch list with per-launch time / DRAM bytes / L2 hit rate / DMMA activity of the same command, full size
#   3. ncu --set full of the estimator kernels on a 600-window workload
# usage: tools/profile_round.sh r1i
set -u
TAG=${1:-rXX}
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active
python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/${TAG}_bench_short.json 2> $OUT/${TAG}_bench_short.err || exit 1
timeout 700 ncu --metrics $M --clock-control none -c 260 --csv --log-file $OUT/${TAG}_launches_full_size_metrics.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:'lw_shrink_kernel|chol_solve_kernel<2>' -c 2 -o $OUT/${TAG}_estimators_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --windows 600 > $OUT/${TAG}_ncu_full.log 2>&1
echo "set full rc=$?"
ncu -i $OUT/${TAG}_estimators_full.ncu-rep --page raw --csv > $OUT/${TAG}_estimators_full_raw.csv 2>/dev/null
python profiles/summarize_ncu.py $OUT/${TAG}_estimators_full_raw.csv > $OUT/${TAG}_estimators_ncu_full_summary.txt 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
