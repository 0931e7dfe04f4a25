#!/bin/bash
# Run on the GPU box (one GPU): plain bench (must exit 0), then the ncu launch list of the same command and one
# --set full capture of the two top kernels.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 0 --no-cpu --no-widened --no-loop"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'chol_solve|gram_dmma' -s 1 -c 12 -o gpurun_out/${TAG}_kernels_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
