for c in 1 2 3 6; do
BP_CHOL_CTAS_PER_SM=$c ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_fp64.sum --clock-control none -k regex:chol_solve -c 2 --csv --log-file gpurun_out/r2a_l2probe_$c.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-widened > /dev/null 2>&1
done
