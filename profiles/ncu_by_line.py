#!/usr/bin/env python3
"""Aggregate an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) by CUDA source
line, using nvdisasm -g line markers of the matching cubin.  Usage:
    ncu_by_line.py <source.csv> <nvdisasm -g -c output> <kernel-name-substring> [top N]
"""
import csv, re, sys
src_csv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# offset -> (file, line) from nvdisasm
off2line = {}
cur = None
infn = False
for ln in open(sass, errors="replace"):
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infn = kname in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infn:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
base = None
agg = {}
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    key = off2line.get(a - base, ("?", 0))
    s = int(r[col["# Samples"]] or 0)
    ins = int(r[col["Instructions Executed"]] or 0)
    d = agg.setdefault(key, {"s": 0, "i": 0, "st": {}})
    d["s"] += s; d["i"] += ins
    tot_s += s; tot_i += ins
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            v = int(r[col[h]] or 0)
            if v:
                d["st"][h] = d["st"].get(h, 0) + v
print(f"total samples {tot_s}, total warp-instructions {tot_i}")
for key, d in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(d["st"].items(), key=lambda kv: -kv[1])[:4])
    print(f"{key[0]}:{key[1]:<5} samples {d['s']:>7} ({100*d['s']/max(tot_s,1):5.1f}%)  inst {d['i']:>11} ({100*d['i']/max(tot_i,1):5.1f}%)  {st}")
