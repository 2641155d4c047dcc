#!/usr/bin/env python3
"""ncu_summary.py X.ncu-rep [...]: one block of headline metrics per captured kernel (raw page of the report)."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if r]
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {rep}: {r[hdr.index('Kernel Name')][:70]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w} [{units[i]}] = {r[i]}")
        st = [(float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        st.sort(reverse=True)
        print("   stalls (warps per issue slot): " + ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
