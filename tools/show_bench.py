#!/usr/bin/env python
"""One-line summaries of bench JSON lines: python tools/show_bench.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "ERR", e)
        continue
    e2e = d.get("e2e") or {}
    r = d.get("roofline") or {}
    st = d.get("stages_rank0_ms") or d.get("stages_ms") or {k: v.get("ms") for k, v in (d.get("stages") or {}).items() if isinstance(v, dict) and "ms" in v}
    print(f"{f}: n={d.get('n_gpus')} {d.get('scaling')} value={d['value']:.0f} ms={d['ms_per_step']:.2f} | e2e={e2e.get('value', 0):.0f} "
          f"ms={e2e.get('ms_per_step', 0):.2f} h2d@ceil={e2e.get('h2d_ms_at_ceiling', 0):.1f}ms ceil={((e2e.get('h2d_ceiling') or {}).get('per_rank_gbs') or 0):.1f}GB/s/rank | "
          f"roof[{r.get('stage', '')}] {r.get('achieved') or 0:.2f}/{r.get('peak') or 0:.2f} frac={r.get('frac') or 0:.3f} | "
          f"stages={ {k: round(v, 2) for k, v in st.items() if v is not None} } flag={d.get('windows_flagged_singular')} par={d.get('parity_max_rel_err')}")
