#!/usr/bin/env python3
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r1_c.csv profiles/r01_launches.md "<command>"
    python tools/summarize_profiles.py full gpurun_out/prof_integrate_r1_b.ncu-rep profiles/r01_k_integrate_ncu.md
"""
import collections, csv, subprocess, sys


def launches(src, dst, cmd):
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for row in r:
        if len(row) <= vi:
            continue
        name = row[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0, row[gi], row[bi]])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\nCommand: `{cmd}`\n\n"
                "Per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes.\n\n"
                f"Total: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms of kernel time in the captured window.\n\n"
                "| kernel | launches | avg us | total us | share | example grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / a[0]:.2f} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |\n")
    print("wrote", dst)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {src.split('/')[-1]}\n\n")
        for r in rows[2:]:
            f.write(f"## launch {r[hdr.index('ID')]}: `{r[hdr.index('Kernel Name')][:80]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
            rd, wr, t = (float(r[hdr.index(k)].replace(",", "")) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
            f.write(f"\ntraffic = dram read + write = {rd + wr:.1f} {units[hdr.index('dram__bytes_read.sum')]} per launch\n\n")
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
