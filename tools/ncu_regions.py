"""Summarise one ncu report of one kernel: headline metrics, stall reasons, and instruction / stall-sample shares
per source line (needs -lineinfo and --import-source on).

    python tools/ncu_regions.py gpurun_out/chol.ncu-rep [top_lines]
"""
import csv, subprocess, sys, io

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1], *args], capture_output=True, text=True).stdout

def main():
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'launch__registers_per_thread', 'launch__grid_size']
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h:88s} {v} {u}")
        elif 'average_warps_issue_stalled' in h and 'not_issued' not in h:
            try:
                if float(v.replace(',', '')) > 0.3: print(f"{h:88s} {v}")
            except ValueError:
                pass
    rows = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur = None; hdr = None; data = {}
    for r in rows:
        if r and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
        if r and r[0] == "Line No":
            hdr = r; i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); continue
        if not r or r[0] == "" or hdr is None: continue
        try: ln = int(r[0]); inst = int(r[i_i]); smp = int(r[i_s])
        except ValueError: continue
        data.setdefault(cur, []).append((ln, inst, smp, r[1].strip()[:100]))
    ti = sum(x[1] for v in data.values() for x in v) or 1; ts = sum(x[2] for v in data.values() for x in v) or 1
    print(f"\ninstructions {ti}, samples {ts}")
    allv = [(f,) + x for f, v in data.items() for x in v]
    print("--- by stall samples")
    for x in sorted(allv, key=lambda x: -x[3])[:top]:
        print(f"{x[0]:16s}:{x[1]:4d} inst {100*x[2]/ti:5.1f}%  samples {100*x[3]/ts:5.1f}%  {x[4]}")
    print("--- by instructions")
    for x in sorted(allv, key=lambda x: -x[2])[:top]:
        print(f"{x[0]:16s}:{x[1]:4d} inst {100*x[2]/ti:5.1f}%  samples {100*x[3]/ts:5.1f}%  {x[4]}")

if __name__ == "__main__":
    main()
