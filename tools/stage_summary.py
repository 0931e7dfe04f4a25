import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"]), round(d["ms_per_step"],2), {k:round(v["ms"],2) for k,v in d["stages"].items() if "ms" in v}, "dominant", d["roofline"]["kernel"][:12], round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["ms_per_step"],1))
    else: print(l.strip()[:300])
