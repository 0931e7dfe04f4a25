#!/usr/bin/env python
"""SASS / ptxas evidence of the built library (runs without a GPU):
    python tools/sass_summary.py > profiles/sass_summary.txt
Per kernel: counts of the mnemonics that prove the sm_100a data path (DMMA = FP64 tensor cores, UTMALDG = TMA tensor
loads, UBLKCP = bulk copies, SYNCS = mbarrier, UCGABAR = cluster barrier; tcgen05 has no FP64 kind, so no UTC*MMA / LDTM
is expected), and registers / spills / shared memory from `nvcc -Xptxas -v` (build/ptxas.log)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "incorporating_different_sources_b200")
LIB = os.path.join(PKG, "libbayes_portfolio.so")
MNEMONICS = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "UCGABAR", "LDS", "STS", "LDG", "STG", "DFMA", "DADD", "MUFU", "UTCHMMA", "LDTM"]


def csrc_sha16():
    h = hashlib.sha256()
    d = os.path.join(PKG, "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "bayes_portfolio.h"), "rb").read())
    return h.hexdigest()[:16]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur, arch = collections.OrderedDict(), None, set()
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        if cur:
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                counts[cur]["_total"] += 1
                for k in MNEMONICS:
                    if op.startswith(k):
                        counts[cur][k] += 1
    names = demangle(list(counts))
    regs = {}
    log = open(os.path.join(PKG, "build", "ptxas.log")).read()
    for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", log):
        regs[m.group(1)] = dict(stack=m.group(2), spill_st=m.group(3), spill_ld=m.group(4), regs=m.group(5), smem=m.group(8) or "0")
    print(f"library: incorporating_different_sources_b200/libbayes_portfolio.so   arch: {', '.join(sorted(arch))}   csrc_sha16: {csrc_sha16()}")
    print("flags: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v")
    hdr = f"{'kernel':62s} {'regs':>4s} {'spill st/ld':>11s} {'smem':>6s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in MNEMONICS)
    print(hdr)
    tot = collections.Counter()
    for fn, c in counts.items():
        short = re.sub(r"\(.*", "", names.get(fn, fn)).replace("void ", "").replace("bp::", "")
        r = regs.get(fn, {})
        print(f"{short[:62]:62s} {r.get('regs', '?'):>4s} {r.get('spill_st', '?') + '/' + r.get('spill_ld', '?'):>11s} {r.get('smem', '?'):>6s} "
              f"{c['_total']:>6d} " + " ".join(f"{c[k]:>7d}" for k in MNEMONICS))
        tot.update(c)
    print(f"{'TOTAL':62s} {'':>4s} {'':>11s} {'':>6s} {tot['_total']:>6d} " + " ".join(f"{tot[k]:>7d}" for k in MNEMONICS))
    print("\nNo UTC*MMA / LDTM (tcgen05): the path is FP64, and tcgen05 has no FP64 kind (SURVEY F11); FP64 tensor work is DMMA.8x8x4\n"
          "fed by TMA (UTMALDG / UBLKCP + SYNCS mbarriers).")


if __name__ == "__main__":
    main()
