#!/usr/bin/env python3
"""Print selected raw metrics of the first kernel in an ncu report: ncu_raw.py X.ncu-rep [metric ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines()) if r]
hdr = next(r for r in rows if r[0] == "ID")
i0 = rows.index(hdr)
units, vals = rows[i0 + 1], rows[i0 + 2]
for i, h in enumerate(hdr):
    if any(h == w or (w.endswith("*") and h.startswith(w[:-1])) for w in want):
        print(f"{h} [{units[i]}] = {vals[i]}")
