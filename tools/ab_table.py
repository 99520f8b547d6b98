#!/usr/bin/env python3
"""Markdown table of bench.py JSON lines: python tools/ab_table.py label=file.json ...   (value / e2e medians and passes, roofline, stage busy times)"""
import json, sys
print("| run | value (frames/s) | passes | e2e (frames/s) | passes | `k_integrate` µs live / frac | stage busy ms per step (extract / track / tri / map) | clocks MHz |")
print("|---|---:|---|---:|---|---|---|---|")
for arg in sys.argv[1:]:
    label, f = arg.split("=", 1)
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f"| {label} | no line ({e}) | | | | | | |")
        continue
    sb = d.get("stage_busy_ms_per_step", {})
    vp = ", ".join(str(round(x)) for x in d.get("value_passes", []))
    ep = ", ".join(str(round(x)) for x in d["e2e"].get("passes", []))
    r = d["roofline"]
    print(f"| {label} | **{d['value']:.0f}** | {vp} | **{d['e2e']['value']:.0f}** | {ep} | {1e3 * r.get('avg_launch_ms', 0):.0f} / {r['frac']:.3f} | "
          f"{sb.get('extract', 0):.2f} / {sb.get('track', 0):.2f} / {sb.get('tri', 0):.2f} / {sb.get('map', 0):.2f} | {d.get('clocks', {}).get('sm_mhz', '')} |")
