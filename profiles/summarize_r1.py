#!/usr/bin/env python3
"""Turn the scratch output of profiles/run_r1_suite.sh (gpurun_out/) into the tracked round-1 summaries:
bench lines, ncu launch lists, headline ncu metrics per kernel, per-source-line stall profiles, and
profiles/traffic.json (DRAM bytes per launch of the dominant kernels, read by bench.py for roofline.traffic)."""
import csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, P)

def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if r]
    return rows[0], rows[1], rows[2:]

for w in ("deflate1", "deflate2", "checksum", "inflate", "reference"):
    src = os.path.join(O, f"r1_bench_{w}.json")
    if os.path.exists(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(P, f"r1_bench_{w}.json"))
for f in ("r1_host_api_roundtrip.txt",):
    if os.path.exists(os.path.join(O, f)):
        shutil.copy(os.path.join(O, f), os.path.join(P, f))
for w in ("deflate1", "deflate2", "checksum", "inflate"):
    src = os.path.join(O, f"r1_launches_{w}.csv")
    if not os.path.exists(src):
        continue
    rows = [r for r in csv.reader(open(src, errors="replace")) if r and (r[0] == "ID" or r[0].isdigit())]
    with open(os.path.join(P, f"r1_launches_{w}.csv"), "w", newline="") as fo:
        cw = csv.writer(fo)
        for r in rows:
            if r[0] == "ID":
                cw.writerow(["ID", "Kernel", "Block", "Grid", "gpu__time_duration.sum [ns]"])
            else:
                cw.writerow([r[0], re.sub(r"\(.*", "", r[4]), r[7], r[8], r[-1]])

traffic = {}
summ = []
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
for w in ("deflate1", "deflate2", "checksum", "inflate"):
    rep = os.path.join(O, f"r1_full_{w}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw_rows(rep)
    tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])
        summ.append(f"== {w}: {name}   (ncu --set full --clock-control none, r1_full_{w}.ncu-rep)")
        for m in want:
            if m in hdr:
                i = hdr.index(m); summ.append(f"   {m} [{units[i]}] = {r[i]}")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try: st.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError: pass
        st = [x for x in st if x[0] == x[0]]
        st.sort(reverse=True)
        summ.append("   stalls (warps per issue slot): " + ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
        try:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            b = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            if b == b: tot += b
        except (ValueError, IndexError):
            pass
    if tot:
        traffic[w] = {"dram_bytes_per_launch": tot, "source": f"profiles/r1_ncu_summary.txt ({w}: dram__bytes_read.sum + dram__bytes_write.sum over the step's hot kernels, one ncu --set full capture)"}
open(os.path.join(P, "r1_ncu_summary.txt"), "w").write("\n".join(summ) + "\n")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
