#!/bin/bash
# Profiling pass of one round on the GPU box (run through gpurun; outputs land in gpurun_out/).
#   1. plain bench run (the program must exit 0 without ncu first)
#   2. ncu launch list with per-launch time / DRAM bytes / L2 hit rate / DMMA activity of the same command at full
#      size, limited to the warm-up + timed device-resident step (-c), about 2.7 s per profiled launch
#   3. ncu --set full of one kernel (regex) on a 600-window workload
# usage: tools/profile_round.sh r1j 'jeffreys_chain_kernel' [launch-count]
set -u
TAG=${1:-rXX}
KERNEL=${2:-jeffreys_chain_kernel}
COUNT=${3:-90}
OUT=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
python bench.py --steps 1 --warmup 1 --no-cpu --no-widened > $OUT/${TAG}_bench_short.json 2> $OUT/${TAG}_bench_short.err || exit 1
timeout 420 ncu --metrics $M --clock-control none -c $COUNT --csv --log-file $OUT/${TAG}_launches_full_size_metrics.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-widened > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$KERNEL" -c 1 -o $OUT/${TAG}_kernel_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-widened --windows 600 > $OUT/${TAG}_ncu_full.log 2>&1
echo "set full rc=$?"
ncu -i $OUT/${TAG}_kernel_full.ncu-rep --page raw --csv > $OUT/${TAG}_kernel_full_raw.csv 2>/dev/null
python profiles/summarize_ncu.py $OUT/${TAG}_kernel_full_raw.csv > $OUT/${TAG}_kernel_ncu_full_summary.txt 2>&1
