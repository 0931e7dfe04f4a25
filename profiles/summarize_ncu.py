#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: one block per launch with the metrics DESIGN.md quotes.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv && python profiles/summarize_ncu.py raw.csv
"""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.sum.per_second",
    "sm__ops_path_tensor_src_fp64.sum.peak_sustained_elapsed.per_second",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("kernel:", r[idx["Kernel Name"]][:70])
        for m in METRICS:
            if m in idx:
                print(f"  {m:85s} {r[idx[m]]:>18s} {units[idx[m]]}")


if __name__ == "__main__":
    main(sys.argv[1])
