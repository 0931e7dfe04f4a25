#!/usr/bin/env python
"""profiles/kernel_traffic.json from an ncu launch list of bench.py (same commit):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-widened --no-loop
    python tools/kernel_traffic.py gpurun_out/launches.csv profiles/kernel_traffic.json

Per (kernel, grid): launches, median duration, median DRAM bytes read / written per launch.  bench.py reports
``roofline.traffic`` from this file and says whether its ``csrc_sha16`` (hash of the kernel sources the profiled library
was built from) matches the tree it runs in."""
import collections
import csv
import json
import os
import re
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_summary import csrc_sha16  # noqa: E402


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").strip()
        if not name.startswith("bp::"):
            continue
        key = (name, r[ix["Grid Size"]], r[ix["Block Size"]])
        d = per.setdefault(key, collections.defaultdict(dict))
        d[r[ix["ID"]]][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    kernels = []
    single = []          # every launch on its own (the Gram kernel runs launches of very different sizes on one grid)
    for (name, grid, block), launches in per.items():
        for lid, v in launches.items():
            single.append({"kernel": name, "grid": grid, "block": block, "launch_id": int(lid),
                           "ms": v.get("gpu__time_duration.sum", 0.0) / 1e6, "dram_read_bytes": v.get("dram__bytes_read.sum", 0.0),
                           "dram_write_bytes": v.get("dram__bytes_write.sum", 0.0)})
    for (name, grid, block), launches in per.items():
        t = [v.get("gpu__time_duration.sum", 0.0) for v in launches.values()]
        rd = [v.get("dram__bytes_read.sum", 0.0) for v in launches.values()]
        wr = [v.get("dram__bytes_write.sum", 0.0) for v in launches.values()]
        kernels.append({"kernel": name, "grid": grid, "block": block, "launches": len(t), "ms_median": statistics.median(t) / 1e6,
                        "dram_read_bytes": statistics.median(rd), "dram_write_bytes": statistics.median(wr)})
    solve = max((k for k in single if "chol_solve_kernel" in k["kernel"]), key=lambda k: k["ms"], default=None)
    gram = max((k for k in single if "gram_dmma_kernel" in k["kernel"]), key=lambda k: k["ms"], default=None)
    out = {
        "csrc_sha16": csrc_sha16(),
        "source": os.path.basename(src),
        "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                   "python bench.py --steps 1 --warmup 0 --no-cpu --no-widened --no-loop",
        "note": "solve_launch / gram_launch: the single longest launch of the kernel (the 4,150-window conjugate launches); "
                "kernels: median per launch shape over one bench run; cold-cache, serialised (compare shares, not absolutes)",
        "solve_dram_bytes_per_launch": (solve["dram_read_bytes"] + solve["dram_write_bytes"]) if solve else None,
        "solve_launch": solve,
        "gram_dram_bytes_per_launch": (gram["dram_read_bytes"] + gram["dram_write_bytes"]) if gram else None,
        "gram_launch": gram,
        "kernels": kernels,
    }
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{dst}: {len(kernels)} kernel shapes, solver {out['solve_dram_bytes_per_launch']:.3e} B/launch, "
          f"gram {out['gram_dram_bytes_per_launch']:.3e} B/launch, csrc_sha16 {out['csrc_sha16']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
