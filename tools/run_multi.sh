#!/bin/bash
# usage: tools/run_multi.sh N tag  -- strong-scaling, C5 and weak lines at N GPUs (torchrun for N > 1, as the driver launches it)
N=$1; TAG=$2
if [ "$N" -gt 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; else L="python"; fi
timeout 400 $L bench.py --gpus $N --scaling strong --steps 3 --warmup 3 > gpurun_out/${TAG}_strong_n$N.json 2> gpurun_out/${TAG}_strong_n$N.err; tail -2 gpurun_out/${TAG}_strong_n$N.err
timeout 600 $L bench.py --gpus $N --config C5 --steps 2 --warmup 3 ${PATHS:+--paths $PATHS} > gpurun_out/${TAG}_c5_n$N.json 2> gpurun_out/${TAG}_c5_n$N.err; tail -2 gpurun_out/${TAG}_c5_n$N.err
timeout 400 $L bench.py --gpus $N --steps 3 --warmup 3 --no-widened --no-loop > gpurun_out/${TAG}_weak_n$N.json 2> gpurun_out/${TAG}_weak_n$N.err; tail -2 gpurun_out/${TAG}_weak_n$N.err
