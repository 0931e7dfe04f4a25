#!/usr/bin/env python
"""Top stall locations of an `ncu --page source --csv` dump (SASS view with -lineinfo): share of warp-stall samples
per instruction with its dominant stall reasons.   python tools/stalls_by_line.py source.csv [top]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith("stall_") and "(" not in h]
    agg, tot, seen = [], 0.0, set()
    for r in rows[2:]:
        if r and r[0] in seen:          # a report section repeated (two metric sections of one launch)
            continue
        seen.add(r[0] if r else None)
        try:
            v = float(r[ix["# Samples"]])
        except Exception:
            continue
        tot += v
        rs = sorted(((float(r[ix[k]] or 0), k[6:]) for k in reasons), reverse=True)[:3]
        agg.append((v, r[ix["Source"]].strip()[:100], ", ".join(f"{k} {int(x)}" for x, k in rs if x > 0)))
    agg.sort(reverse=True)
    print(f"total samples {tot:.0f}")
    uniq = {}
    for r in rows[2:]:
        if r and r[0] not in uniq:
            uniq[r[0]] = r
    by_reason = {k: sum(float(r[ix[k]] or 0) for r in uniq.values() if len(r) > ix[k] and r[ix[k]].replace('.', '').isdigit()) for k in reasons}
    print("by reason:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
    for v, s, why in agg[:top]:
        print(f"{100 * v / tot:5.1f}%  {s:100s}  [{why}]")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
