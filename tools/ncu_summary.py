"""Print the metrics we track from an .ncu-rep (raw page) -- development helper."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_active.avg', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, vals):
        if h in want or ('issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and float(v or 0) > 0.15):
            print(f"  {h:85s} {v} {u}")
