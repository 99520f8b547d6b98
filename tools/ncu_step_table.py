#!/usr/bin/env python3
"""profiles/<name>.md from the raw CSV of one `ncu --set full` capture of ~one pipeline step (tools/gpu_profile_all.sh):
per kernel type the mean duration, DRAM bytes, achieved DRAM GB/s against the measured HBM peak, and the issue / SM figures.
    python tools/ncu_step_table.py gpurun_out/full_step_raw.csv profiles/r01_step_ncu_full.md"""
import collections, csv, json, pathlib, sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
col = {n: i for i, n in enumerate(hdr)}
peak = 6581.6
try:
    peak = float(json.loads((pathlib.Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def val(r, name, scale=None):
    v = float(r[col[name]].replace(",", "") or 0)
    u = units[col[name]]
    if scale == "us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    if scale == "MB":
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    return v


agg = collections.OrderedDict()
for r in data:
    name = r[col['Kernel Name']].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, sm=0.0, issue=0.0, dram=0.0, warps=0.0, l1=0.0, l2=0.0, inst=0.0,
                                  grid=r[col['launch__grid_size']], block=r[col['launch__block_size']], regs=r[col['launch__registers_per_thread']]))
    a["n"] += 1
    a["us"] += val(r, 'gpu__time_duration.sum', "us")
    a["rd"] += val(r, 'dram__bytes_read.sum', "MB"); a["wr"] += val(r, 'dram__bytes_write.sum', "MB")
    a["sm"] += val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'); a["issue"] += val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')
    a["dram"] += val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'); a["warps"] += val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')
    a["l1"] += val(r, 'l1tex__t_sector_hit_rate.pct'); a["l2"] += val(r, 'lts__t_sector_hit_rate.pct'); a["inst"] += val(r, 'smsp__inst_executed.sum')
tot = sum(a["us"] for a in agg.values())
with open(dst, "w") as f:
    f.write("# `ncu --set full --clock-control none`: every kernel of one pipeline step\n\n"
            "Command: `ncu --set full --clock-control none -s 700 -c 130 python bench.py --steps 2 --warmup 1 --no-cpu-baseline` "
            "(tools/gpu_profile_all.sh; 130 consecutive launches = a little more than one 8-frame step of the 4-stage pipeline; the 149 MB report stays on the box, "
            "this table is made from its `--page raw --csv` export by tools/ncu_step_table.py).\n\n"
            "Per-launch means. ncu serialises the kernels and flushes caches between replay passes, so durations are cold-cache and **isolated** (no SM sharing with the "
            f"other pipeline stages); compare shares. `DRAM GB/s` = (dram read + write bytes) / duration; `of peak` against the measured HBM copy bandwidth {peak:.1f} GB/s "
            "(MEASURED_PEAKS.json). The small kernels of the matcher and the extractor move a few hundred KB that stay in the 126 MB L2: they are latency-bound, "
            "their DRAM column is for the record.\n\n"
            f"Total: {sum(a['n'] for a in agg.values())} launches, {tot / 1e3:.3f} ms of kernel time.\n\n"
            "| kernel | n | avg us | share | DRAM rd MB | DRAM wr MB | DRAM GB/s | of peak | ncu dram % | SM thr % | issue act % | warps act % | L1 hit % | L2 hit % | Minst | grid | block | regs |\n"
            "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|---:|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        gbs = (a["rd"] + a["wr"]) * 1e6 / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        f.write(f"| `{k}` | {n} | {a['us'] / n:.2f} | {100 * a['us'] / tot:.1f}% | {a['rd'] / n:.3f} | {a['wr'] / n:.3f} | {gbs:.0f} | {gbs / peak:.3f} | {a['dram'] / n:.1f} | "
                f"{a['sm'] / n:.1f} | {a['issue'] / n:.1f} | {a['warps'] / n:.1f} | {a['l1'] / n:.1f} | {a['l2'] / n:.1f} | {a['inst'] / n / 1e6:.2f} | {a['grid']} | {a['block']} | {a['regs']} |\n")
print("wrote", dst)
