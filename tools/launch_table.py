#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid, block) count, mean and total duration."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); gs = h.index("Grid Size"); bs = h.index("Block Size"); mu = h.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[kn])).replace("<unnamed>::", "").replace("plvs::orb::", "").replace("plvs::", "")
    try:
        v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    except ValueError:
        continue
    d.setdefault((name, r[gs], r[bs]), []).append(v)
md = "--md" in sys.argv
if md:
    print("| kernel | grid | block | launches | mean µs | total µs |\n|---|---|---|---:|---:|---:|")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    if md:
        print(f"| `{k[0]}` | {k[1]} | {k[2]} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.0f} |")
    else:
        print(f"{k[0][:44]:44s} {k[1]:>16s} {k[2]:>14s} n={len(v):3d} mean {sum(v) / len(v):8.1f} us  total {sum(v):9.1f}")
