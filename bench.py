#!/usr/bin/env python3
"""bench.py -- chunked-DEFLATE hot path on B200, one JSON line per run.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (BASELINE metric: level 1)
    python bench.py --impl reference --gpus N --steps K ...  # zlib-ng's own CPU path, all host cores
    python bench.py --workload deflate2..deflate6|checksum|inflate     # the other BASELINE configs (same JSON contract)
    python bench.py --config 0..4 [--gpus N]                 # BASELINE.json configs[k] at its full size (see CONFIGS below)

Default workload (BASELINE.json configs[0], the configuration the metric is quoted on): per GPU a 1 GiB synthetic
mixed text/binary buffer = 16,384 independent 64 KiB chunks, deflate_quick (level 1) raw deflate with Z_FULL_FLUSH
chunk ends + per-chunk CRC-32, then the output-offset scan, the gather into one contiguous raw-deflate stream and
the crc32_combine fold.  With N > 1 each rank owns its own 1 GiB shard (weak scaling, contiguous chunk ranges) and
the only collective is an NCCL allgather of the per-chunk (size, crc32) pairs.

A "step" is one pass over the batch.  `value` is device-resident throughput (inputs already in HBM, CUDA events,
max over ranks); `e2e` is the same work through the host-buffer C-ABI call (what zng_deflate / zng_inflate /
zng_crc32 of the host library call) with pinned host buffers, H2D and D2H inside the timing.  `roofline` is the
dominant kernel's algorithmic bytes / its own CUDA-event time against the measured HBM peak.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import struct
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 65536
MEMBER = 4096
SEED = 0x9E3779B97F4A7C15
UNIT = "GB/s"
METRICS = {"deflate1": "level1_deflate_input_throughput", "deflate2": "level2_deflate_input_throughput",
           "deflate3": "level3_deflate_input_throughput", "deflate4": "level4_deflate_input_throughput",
           "deflate5": "level5_deflate_input_throughput", "deflate6": "level6_deflate_input_throughput",
           "checksum": "crc32_adler32_input_throughput", "inflate": "inflate_output_throughput"}


# BASELINE.json configs[k] -> (workload, total MiB over all GPUs or None = 1 GiB per GPU, scaling).  configs[4] (64 GiB on
# 8 GPUs) is 8 GiB per GPU: with fewer GPUs the per-GPU size stays 8 GiB (input + slots + packed stream of 64 GiB do not
# fit one GPU's 180 GB), so it weak-scales; configs[1] defaults to its largest size, --total-gib 1|4 selects the others;
# configs[2] runs level 2, --level 3 selects deflate_medium.
CONFIGS = {0: ("deflate1", None, "weak"), 1: ("checksum", 16384, "strong"), 2: ("deflate2", 4096, "strong"),
           3: ("inflate", 4096, "strong"), 4: ("deflate1", None, "weak")}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=None, choices=sorted(CONFIGS), help="BASELINE.json configs[k] at full size")
    ap.add_argument("--total-gib", type=int, default=None, help="with --config 1/2/3: total GiB over all GPUs")
    ap.add_argument("--level", type=int, default=None, choices=[1, 2, 3, 4, 5, 6], help="with --config 2: deflate level (default 2)")
    ap.add_argument("--no-parity", action="store_true", help="skip the every-chunk comparison with the reference")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="deflate1", choices=list(METRICS))
    ap.add_argument("--mib-per-gpu", type=int, default=1024, help="input (inflate: output) MiB per GPU; default = the 1 GiB BASELINE config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.scaling = "weak"
    args.config_name = None
    if args.config is not None:
        world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
        wl, total, scaling = CONFIGS[args.config]
        if args.config == 2 and args.level:
            wl = f"deflate{args.level}"
        if args.total_gib and total is not None:
            total = args.total_gib << 10
        args.workload = wl
        args.scaling = scaling
        if args.config == 4:
            args.mib_per_gpu = 8192
        elif total is not None:
            args.mib_per_gpu = max(64, total // max(1, world))
        args.config_name = f"BASELINE configs[{args.config}]"
    return args


def workload_text(w, mib, ngpu):
    n = mib << 20
    if w.startswith("deflate"):
        lvl, fn, cfg = {"deflate1": (1, "deflate_quick", "configs[0]"), "deflate2": (2, "deflate_fast", "configs[2] at 1 GiB per GPU"),
                        "deflate3": (3, "deflate_medium", "configs[2] at 1 GiB per GPU"),
                        "deflate4": (4, "deflate_medium", "configs[2] at 1 GiB per GPU, higher level"),
                        "deflate5": (5, "deflate_medium", "configs[2] at 1 GiB per GPU, higher level"),
                        "deflate6": (6, "deflate_medium", "configs[2] at 1 GiB per GPU, zlib-ng's default level")}[w]
        return (f"{fn} level {lvl}, {mib} MiB per GPU as {n // CHUNK} x 64 KiB raw-deflate chunks + per-chunk crc32, offset scan + gather + "
                f"crc32_combine fold (BASELINE {cfg})")
    if w == "checksum":
        return f"zng_crc32 + zng_adler32 of a flat {mib} MiB buffer per GPU, 64 KiB tiles + combine folds (BASELINE configs[1])"
    return (f"batched inflate of {n // MEMBER} independent 4 KiB gzip members per GPU (bodies = level-1 Z_FINISH streams, minigzip -1 framing), "
            f"output + CRC-32 + ISIZE checked on device (BASELINE configs[3] at {mib} MiB of output per GPU)")


def config_dict(args, ngpu):
    """`config` of the JSON line: the same dict from both arms (the driver compares them)."""
    wl = args.workload
    text = workload_text(wl, args.mib_per_gpu, ngpu)
    if args.config_name:
        text += f" [{args.config_name} at full size: {args.mib_per_gpu * ngpu} MiB over {ngpu} GPU(s)]"
    return {"workload": text, "chunk_bytes": CHUNK, "sharding": f"contiguous chunk ranges x{ngpu}",
            "collective": ("nccl allgather of (size, crc32) per chunk" if wl.startswith("deflate") else
                           ("nccl allgather of (crc32, adler32) per rank" if wl == "checksum" else "none")) if ngpu > 1 else "none",
            "l2": f"input {args.mib_per_gpu} MiB per step >> 126 MB L2, no flush needed"}


ROOFLINE_NOTES = {
    "deflate1": "serial-per-chunk LZ77 parse; the hash-head tables of the resident chains exceed L2, so the parser is bound by DRAM "
                "sector (random access) rate, not by streaming bandwidth (DESIGN.md section 7)",
    "deflate2": "hash-chain walk (<= 4 links) per position + per-block Huffman tree build: latency bound",
    "deflate3": "deflate_medium, chain 6 / nice 16, every position of a match inserted: latency bound",
    "deflate4": "deflate_medium, chain 24 / nice 32: dependent prev[] link loads dominate",
    "deflate5": "deflate_medium with look-ahead, chain 32: dependent prev[] link loads dominate",
    "deflate6": "deflate_medium with look-ahead, chain 128 (zlib-ng's default level): dependent prev[] link loads dominate",
}


def synth_mod():
    from __graft_entry__ import load_synth
    return load_synth()


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ synthetic members (inflate workload)
def make_members(pkg, n_bytes, seed_offset=0):
    """config 3: every 4096-byte slice of the synthetic stream as one gzip member whose body is the level-1 Z_FINISH
    stream of the slice (what minigzip -1 writes: header 1f 8b 08 00 00000000 04 03, raw deflate, CRC32, ISIZE).
    Built with this repo's own GPU compressor (bit-exact with the reference, tests/test_gpu_deflate_quick.py) so that
    setup stays fast; returns (members bytes, in_off u64[n+1], raw data)."""
    import numpy as np
    import torch
    ctx = pkg.Context(torch.cuda.current_device())
    n_members = n_bytes // MEMBER
    data = np.empty(n_bytes, dtype=np.uint8)
    synth_mod().fill(data.ctypes.data, n_bytes, SEED, seed_offset)
    stride = pkg.deflate_bound(MEMBER)
    parts_sizes = np.zeros(n_members, dtype=np.int64)
    bodies = []
    batch = 65536
    for m0 in range(0, n_members, batch):
        m1 = min(m0 + batch, n_members)
        d_in = torch.from_numpy(data[m0 * MEMBER:m1 * MEMBER]).cuda()
        slots = torch.empty((m1 - m0) * stride, dtype=torch.uint8, device="cuda")
        sizes = torch.zeros(m1 - m0, dtype=torch.int32, device="cuda")
        crcs = torch.zeros(m1 - m0, dtype=torch.int32, device="cuda")
        ctx.deflate_chunks(d_in, (m1 - m0) * MEMBER, MEMBER, 1, pkg.Z_FINISH, slots, stride, sizes, crcs, None)
        torch.cuda.synchronize()
        bodies.append((slots.cpu().numpy().reshape(-1, stride), sizes.cpu().numpy().astype(np.int64), crcs.cpu().numpy().view(np.uint32)))
        parts_sizes[m0:m1] = bodies[-1][1] + 18
    ctx.close()
    in_off = np.zeros(n_members + 1, dtype=np.uint64)
    in_off[1:] = np.cumsum(parts_sizes).astype(np.uint64)
    total = int(in_off[-1])
    members = np.zeros(total + 16, dtype=np.uint8)
    hdr = np.frombuffer(bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4, 3]), dtype=np.uint8)
    isize = np.frombuffer(struct.pack("<I", MEMBER), dtype=np.uint8)
    m = 0
    for rows, sizes, crcs in bodies:
        for i in range(len(sizes)):
            o = int(in_off[m]); s = int(sizes[i])
            members[o:o + 10] = hdr
            members[o + 10:o + 10 + s] = rows[i, :s]
            members[o + 10 + s:o + 14 + s] = np.frombuffer(struct.pack("<I", int(crcs[i])), dtype=np.uint8)
            members[o + 14 + s:o + 18 + s] = isize
            m += 1
    return members, in_off, data


def make_members_cpu(zo, n_bytes):
    """Same members from the CPU side (reference arm: no GPU needed): refdrv_gzip_members / oracle port."""
    import numpy as np
    n_members = n_bytes // MEMBER
    data = synth_mod().synth(n_bytes, SEED)
    out, sizes, crcs, _ = (zo.ref_deflate_chunks if zo.have_ref() else zo.port_deflate_chunks)(data, MEMBER, 1, 4)
    in_off = np.zeros(n_members + 1, dtype=np.uint64)
    in_off[1:] = np.cumsum(sizes.astype(np.int64) + 18).astype(np.uint64)
    members = np.zeros(int(in_off[-1]) + 16, dtype=np.uint8)
    hdr = np.frombuffer(bytes([0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 4, 3]), dtype=np.uint8)
    for i in range(n_members):
        o = int(in_off[i]); s = int(sizes[i])
        members[o:o + 10] = hdr
        members[o + 10:o + 10 + s] = out[i, :s]
        members[o + 10 + s:o + 18 + s] = np.frombuffer(struct.pack("<II", int(crcs[i]), MEMBER), dtype=np.uint8)
    return members, in_off, data


# ------------------------------------------------------------------------------------ reference arm
def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_reference_run(workload: str, n_bytes: int, steps: int, warmup: int):
    """Time the reference's CPU implementation (oracle/_ref when it was built from /root/reference, else the oracle
    port) on the same workload with every host core: one zng_stream per worker thread, one chunk / member per task."""
    import numpy as np
    from __graft_entry__ import load_oracle
    zo = load_oracle()          # the reference arm loads oracle/ and the synthetic generator only, never the product library
    sd = synth_mod()
    cores = host_cores()
    kind = "reference" if zo.have_ref() else "port"
    if kind == "port":
        cores = min(cores, 256)
    ratio = None
    if workload.startswith("deflate"):
        level = int(workload[-1])
        fn = zo.ref().refdrv_deflate_chunks if kind == "reference" else zo.port().zo_deflate_chunks
        data = sd.synth(n_bytes, SEED)
        nch = (n_bytes + CHUNK - 1) // CHUNK
        stride = int(zo.port().zo_deflate_bound(CHUNK))
        out = np.empty(nch * stride, dtype=np.uint8)
        sizes = np.zeros(nch, dtype=np.uint32)
        crcs = np.zeros(nch, dtype=np.uint32)
        call = lambda: fn(data.ctypes.data, n_bytes, CHUNK, level, 3, out.ctypes.data, stride, sizes.ctypes.data, crcs.ctypes.data, None, cores)
        what = f"{n_bytes >> 20} MiB of the same synthetic workload per step ({nch} chunks; zng_deflateReset + zng_deflate(Z_FULL_FLUSH) + zng_crc32 per chunk)"
    elif workload == "checksum":
        data = sd.synth(n_bytes, SEED)
        if kind == "reference":
            c, a = ctypes_u32(), ctypes_u32()
            call = lambda: zo.ref().refdrv_checksum_flat(data.ctypes.data, n_bytes, 1 << 20, cores, c, a)
        else:
            call = lambda: (zo.port().zo_crc32(0, data.ctypes.data, n_bytes), zo.port().zo_adler32(1, data.ctypes.data, n_bytes)) and 0
            cores = 1
        what = f"{n_bytes >> 20} MiB flat buffer per step, 1 MiB pieces per thread + combine (zng_crc32_z + zng_adler32_z + *_combine)"
    else:
        members, in_off, data = make_members_cpu(zo, n_bytes)
        nm = len(in_off) - 1
        out = np.empty(n_bytes, dtype=np.uint8)
        out_off = (np.arange(nm + 1, dtype=np.uint64) * MEMBER)
        sizes = np.zeros(nm, dtype=np.uint32); crcs = np.zeros(nm, dtype=np.uint32); status = np.zeros(nm, dtype=np.int32)
        fn = zo.ref().refdrv_inflate_members if kind == "reference" else zo.port().zo_inflate_members
        call = lambda: fn(members.ctypes.data, in_off.ctypes.data, nm, out.ctypes.data, out_off.ctypes.data, sizes.ctypes.data, crcs.ctypes.data, status.ctypes.data, cores)
        what = f"{nm} gzip members of 4 KiB per step (zng_inflateInit2(31) / zng_inflateReset + zng_inflate(Z_FINISH) per member)"
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = call()
        t1 = time.perf_counter()
        if r not in (0, None):
            raise RuntimeError(f"reference {workload} failed: {r}")
        if it >= warmup:
            times.append(t1 - t0)
    if workload.startswith("deflate"):
        ratio = float(sizes.sum()) / n_bytes
    if workload == "inflate" and not (status == 1).all():
        raise RuntimeError("reference inflate: a member did not reach Z_STREAM_END")
    total = sum(times)
    return {"value": n_bytes * len(times) / total / 1e9, "ms_per_step": 1e3 * total / len(times), "cores": cores, "kind": kind, "ratio": ratio,
            "sample": f"{what}, {len(times)} timed steps, wall clock, {cores} threads"}


def ctypes_u32():
    import ctypes
    return ctypes.pointer(ctypes.c_uint32(0))


def cpu_sample_bytes(workload, n, seconds, reps):
    """Bounded sample of the workload for the CPU arm: about `seconds` of wall clock over `reps` passes."""
    cores = host_cores()
    per_core = {"deflate1": 0.10e9, "deflate2": 0.06e9, "deflate3": 0.05e9, "deflate4": 0.04e9, "deflate5": 0.035e9, "deflate6": 0.02e9, "checksum": 5e9, "inflate": 0.5e9}[workload]
    budget = int(per_core * cores * seconds / max(1, reps))
    return min(n, max(64 << 20, (budget >> 26) << 26))


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_bytes = args.mib_per_gpu << 20
    sample = cpu_sample_bytes(args.workload, n_bytes, 120, args.steps + args.warmup)       # ~2 min in total
    res = cpu_reference_run(args.workload, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRICS[args.workload], "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "compression_ratio": res["ratio"],
    }
    emit(line)


# ------------------------------------------------------------------------------------ B200 arm
def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-rank runs: keep this rank's threads (and so its pinned host buffers, by first touch) on the NUMA node its GPU
    hangs off -- what a pigz-style tool would do with numactl.  Best effort: any missing piece of /sys leaves things as
    they are.  Returns the node or None."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        if hasattr(props, "pci_bus_id"):
            path = f"/sys/bus/pci/devices/{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{getattr(props, 'pci_device_id', 0):02x}.0/numa_node"
        else:
            import pynvml
            pynvml.nvmlInit()
            want = str(getattr(props, "uuid", ""))
            path = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if want and want in u:
                    b = pynvml.nvmlDeviceGetPciInfo(h).busId
                    b = b.decode() if isinstance(b, bytes) else b
                    path = f"/sys/bus/pci/devices/{b[-12:].lower()}/numa_node"
            if path is None:
                return None
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def host_memory_allows(n, ngpu):
    """True when MemAvailable covers the pinned input (already allocated) + output + checker buffers of every rank with 25 % to spare."""
    try:
        avail = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable:")) * 1024
    except Exception:
        return True
    need = ngpu * (int(n * 1.2) + (4 << 30))
    return avail > need * 1.25


def copy_ceiling(torch, dev, h_src, h2d_bytes, d2h_bytes, barrier, reps):
    """Seconds (best of reps, best of two transfer shapes) to move h2d_bytes host->device and d2h_bytes device->host concurrently
    on two streams between pinned host memory and HBM -- the transfers of one e2e step without any kernel.  Shapes: the pieces the
    streamed host path uses (16 MiB up, 32 MiB down) and one transfer per direction.  Returns (best seconds, per-shape seconds)."""
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    h2d_bytes = min(int(h2d_bytes), h_src.numel())
    d2h_bytes = int(d2h_bytes)
    d_up = torch.empty(max(h2d_bytes, 16), dtype=torch.uint8, device=dev)
    d_dn = torch.empty(max(d2h_bytes, 16), dtype=torch.uint8, device=dev)
    h_dn = torch.empty(max(d2h_bytes, 16), dtype=torch.uint8, pin_memory=True)
    shapes = {"pieces_16MiB_up_32MiB_down": (16 << 20, 32 << 20), "one_transfer_per_direction": (1 << 62, 1 << 62)}
    out = {}
    for name, (pu, pd) in shapes.items():
        best = float("inf")
        for _ in range(reps + 1):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                for o in range(0, h2d_bytes, pu):
                    e = min(h2d_bytes, o + pu)
                    d_up[o:e].copy_(h_src[o:e], non_blocking=True)
            with torch.cuda.stream(s_dn):
                for o in range(0, d2h_bytes, pd):
                    e = min(d2h_bytes, o + pd)
                    h_dn[o:e].copy_(d_dn[o:e], non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[name] = best
    return min(out.values()), out


def cross_rank_parity(allp, wl, gf):
    """Rank 0: every rank's verdict plus what only exists across ranks -- the folded CRC-32 of the whole stream and its total
    packed length (deflate), the combined crc32 / adler32 (checksum) -- against the reference's own combine functions."""
    from __graft_entry__ import load_oracle
    zo = load_oracle()
    bad = [r for r, (_, ok, _) in enumerate(allp) if ok is not True]
    if bad:
        return f"MISMATCH or unchecked on ranks {bad}: " + " | ".join(str(allp[r][0]) for r in bad)
    text = f"{len(allp)} ranks: " + allp[0][0]
    if gf is None:
        return text
    L = zo.ref() if zo.have_ref() else zo.port()
    ccomb = L.zng_crc32_combine if zo.have_ref() else L.zo_crc32_combine
    acomb = L.zng_adler32_combine if zo.have_ref() else L.zo_adler32_combine
    if wl.startswith("deflate"):
        fold, total = 0, 0
        for _, _, f in allp:
            fold = ccomb(fold, f["crc_fold"], f["n"]); total += f["packed"]
        good = fold == gf["crc_fold"] and total == gf["packed"]
        return text + (f"; stream crc32 {fold:08x} (crc32_combine over all ranks) and total length {total} equal the device fold / scan" if good
                       else f"; MISMATCH across ranks: fold {fold:08x} vs {gf['crc_fold']:08x}, length {total} vs {gf['packed']}")
    if wl == "checksum":
        c, a = 0, 1
        for _, _, f in allp:
            c = ccomb(c, f["crc"], f["n"]); a = acomb(a, f["adler"], f["n"])
        gc, ga = 0, 1
        for (rc, ra) in gf["per_rank"]:
            gc = ccomb(gc, rc, allp[0][2]["n"]); ga = acomb(ga, ra, allp[0][2]["n"])
        good = (c, a) == (gc, ga)
        return text + (f"; combined crc32 {c:08x} / adler32 {a:08x} over all ranks equal the allgathered device values" if good else "; MISMATCH of the combined checksums")
    return text


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    pkg = load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    ngpu = world
    wl = args.workload
    n = args.mib_per_gpu << 20
    ctx = pkg.Context(local_rank)
    steps = args.steps
    k_start = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    k_stop = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    launches = {"n": 0}
    parity = None
    out_bytes = 0
    e2e_fn = None

    def barrier():
        if ngpu > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if wl.startswith("deflate"):
        level = int(wl[-1])
        nch = n // CHUNK
        stride = pkg.deflate_bound(CHUNK)
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)        # this rank's shard of the synthetic stream
        synth_mod().fill(h_in.data_ptr(), n, SEED, rank * n)
        d_in = h_in.to(dev, non_blocking=True)
        slots = torch.empty(nch * stride, dtype=torch.uint8, device=dev)
        meta = torch.zeros(2, nch, dtype=torch.int32, device=dev)          # row 0 sizes, row 1 crc32
        sizes, crcs = meta[0], meta[1]
        offsets = torch.zeros(nch + 1, dtype=torch.int64, device=dev)
        packed = torch.empty(nch * stride, dtype=torch.uint8, device=dev)
        res = torch.zeros(4, dtype=torch.int32, device=dev)
        comm = pkg.Comm.from_torch(ctx) if ngpu > 1 else None               # zng_b200_comm: NCCL communicator behind the C ABI
        offsets_g = torch.zeros(nch + 1, dtype=torch.int64, device=dev)     # global stream offsets of this rank's chunks
        multi = {"end": 0, "crc": 0, "n": 0}
        k_per_step = 3                                                      # parse, emit, checksum tiles

        def step(i_timed=None):
            if i_timed is not None:
                k_start[i_timed].record()
            ctx.deflate_chunks(d_in, n, CHUNK, level, pkg.Z_FULL_FLUSH, slots, stride, sizes, crcs, None)
            if i_timed is not None:
                k_stop[i_timed].record()
            launches["n"] += k_per_step
            if ngpu > 1:
                # zng_b200_stream_index_multi: ONE ncclAllGather of the (size, crc32) pairs, then scan + crc32_combine fold on every rank
                multi["end"], multi["crc"], multi["n"] = comm.stream_index(sizes, crcs, nch, CHUNK, n, 10, offsets_g)
                launches["n"] += 3
            else:
                ctx.crc32_fold(crcs, nch, CHUNK, n, 0, res[0:1])
                launches["n"] += 1
            ctx.chunk_offsets(sizes, nch, 0, offsets)                      # local packing offsets
            ctx.gather_chunks(slots, stride, sizes, offsets, nch, packed)
            launches["n"] += 2

        def parity_check(e2e_out=None):
            """EVERY chunk of this rank's shard against the unmodified reference (oracle/_ref; the port if it is absent):
            size and bytes of the device-resident slots, the per-chunk CRC-32, and -- when the e2e leg ran -- the packed
            host output of zng_b200_deflate_host.  Returns (text, ok, shard facts for the cross-rank check)."""
            from __graft_entry__ import load_oracle
            zo = load_oracle()
            cores = host_cores() if ngpu == 1 else max(1, host_cores() // ngpu)
            hs = sizes.cpu().numpy().view(np.uint32)
            hc = crcs.cpu().numpy().view(np.uint32)
            hin = h_in.numpy()
            piece = 16384
            bad = bad_crc = bad_e2e = 0
            first = None
            fold, eo, kind = 0, 0, "?"
            comb = zo.ref().zng_crc32_combine if zo.have_ref() else zo.port().zo_crc32_combine
            for c0 in range(0, nch, piece):
                c1 = min(nch, c0 + piece)
                kind, exp, es, ec, _ = zo.best_deflate_chunks(hin[c0 * CHUNK:c1 * CHUNK], CHUNK, level, 3, stride, nthreads=cores)
                got = slots[c0 * stride:c1 * stride].cpu().numpy()
                b, f = zo.compare_chunks(got, stride, hs[c0:c1], exp, stride, es)
                if b and first is None:
                    first = c0 + f
                bad += b
                bad_crc += int((hc[c0:c1] != ec).sum())
                if e2e_out is not None:
                    seg = int(es.astype(np.int64).sum())
                    b2, _ = zo.compare_chunks(e2e_out[eo:eo + seg], 0, es, exp, stride, es)
                    bad_e2e += b2
                    eo += seg
                for k in range(c1 - c0):
                    fold = comb(fold, int(ec[k]), CHUNK)
            ok = bad == 0 and bad_crc == 0 and bad_e2e == 0
            text = (f"all {nch} of {nch} chunks equal {kind}: compressed size + bytes and crc32" if ok else
                    f"MISMATCH vs {kind}: {bad} of {nch} chunks differ (first {first}), {bad_crc} crc32 differ, {bad_e2e} e2e chunks differ")
            if ok and e2e_out is not None:
                text += f"; e2e host output ({eo} bytes) equals the concatenated reference chunks"
            return text, ok, {"crc_fold": int(fold), "packed": int(hs.astype(np.int64).sum()), "n": n}

        def finish():
            return int(offsets[nch].item())

        def global_facts():
            """What the collective produced on this rank: the fold of every rank's chunk CRCs and the end of the global scan."""
            u = lambda v: int(v) & 0xffffffff
            if ngpu > 1:
                return {"crc_fold": u(multi["crc"]), "packed": multi["end"] - 10}
            return {"crc_fold": u(res[0].item()), "packed": int(offsets[nch].item())}

        cap = nch * stride
        if n > (2 << 30):
            # multi-GiB shards (configs[4]: 8 GiB per GPU): the pinned output buffer is sized for this workload's ratio (0.56) with
            # margin, not for the 1.126 n worst case -- 8 ranks x (8 GiB in + 9.2 GiB out) of pinned memory is what a box may not have
            cap = int(n * 0.75) + (1 << 20)

        def make_e2e():
            h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            state = {}

            def call():
                state["out_len"], _, _ = ctx.deflate_host(h_in, n, CHUNK, level, False, h_out, cap)
            api = ("zng_b200_deflate_host (pinned host in/out; streamed: one persistent parse kernel fed by the copy engine in 16 MiB pieces, emit / gather / D2H per 32 MiB slab)"
                   if level == 1 else "zng_b200_deflate_host (pinned host in/out, 4 x 128 MiB slabs in flight on separate streams)")
            return call, (lambda: (n, int(state["out_len"]))), api, (lambda: h_out.numpy()[: int(state["out_len"])])
        e2e_fn = make_e2e
        alg_bytes = lambda ob: n + ob                                     # in + out per chunk, x chunks per launch
        kernel_name = "quick_parse_kernel + static_emit_kernel (+ checksum_tiles_kernel)" if level == 1 else f"fast_parse_kernel<{level}> + block_emit_kernel (+ checksum_tiles_kernel)"
        note = ROOFLINE_NOTES[wl]

    elif wl == "checksum":
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        synth_mod().fill(h_in.data_ptr(), n, SEED, rank * n)
        d_in = h_in.to(dev, non_blocking=True)
        res = torch.zeros(4, dtype=torch.int32, device=dev)
        if ngpu > 1:
            gath = torch.zeros(ngpu * 4, dtype=torch.int32, device=dev)

        ntile = (n + CHUNK - 1) // CHUNK
        t_crc = torch.zeros(ntile, dtype=torch.int32, device=dev)
        t_adl = torch.zeros(ntile, dtype=torch.int32, device=dev)

        def step(i_timed=None):
            if i_timed is not None:
                k_start[i_timed].record()
            ctx.checksum_chunks(d_in, n, CHUNK, t_crc, t_adl)             # ONE pass over the buffer: both checksums per 64 KiB tile
            ctx.crc32_fold(t_crc, ntile, CHUNK, n, 0, res[0:1])           # crc32_combine / adler32_combine algebra over the tiles
            ctx.adler32_fold(t_adl, ntile, CHUNK, n, 1, res[1:2])
            if i_timed is not None:
                k_stop[i_timed].record()
            launches["n"] += 3
            if ngpu > 1:
                dist.all_gather_into_tensor(gath, res)                    # G x (crc, adler) -> combined on the host (G values)

        def parity_check(e2e_out=None):
            from __graft_entry__ import load_oracle
            zo = load_oracle()
            r = res.cpu().numpy().view(np.uint32)
            hin = h_in.numpy()
            if zo.have_ref():
                kind, ec, ea = "oracle/_ref (unmodified zlib-ng 2.2.2)", zo.ref_crc32(hin), zo.ref_adler32(hin)
            else:
                kind, ec, ea = "oracle port", zo.port_crc32(hin), zo.port_adler32(hin)
            ok = int(r[0]) == ec and int(r[1]) == ea
            if ok and e2e_out is not None:
                ok = e2e_out == (ec, ea)
            text = (f"crc32 {ec:08x} and adler32 {ea:08x} of the {n >> 20} MiB shard equal {kind}" + ("; e2e host calls return the same values" if e2e_out is not None else "")) if ok else f"MISMATCH vs {kind}"
            return text, ok, {"crc": ec, "adler": ea, "n": n}

        def finish():
            return 0

        def global_facts():
            g = gath.cpu().numpy().view(np.uint32).reshape(ngpu, 4) if ngpu > 1 else res.cpu().numpy().view(np.uint32).reshape(1, 4)
            return {"per_rank": [(int(r[0]), int(r[1])) for r in g]}

        def make_e2e():
            state = {}

            def call():
                state["r"] = (ctx.crc32_host(h_in, n, 0), ctx.adler32_host(h_in, n, 1))
            return call, (lambda: (2 * n, 8)), "zng_b200_crc32_host + zng_b200_adler32_host (what zng_crc32_z / zng_adler32_z call)", (lambda: state["r"])
        e2e_fn = make_e2e
        alg_bytes = lambda ob: n                                          # one fused pass reads the buffer once
        kernel_name = "checksum_tiles_kernel<crc32 + adler32> (+ folds)"
        note = "table-driven CRC-32 without carry-less multiply: shared-memory lookup bound"

    else:  # inflate
        members, in_off, data = make_members(pkg, n, rank * n)
        nm = len(in_off) - 1
        h_members = torch.from_numpy(members).pin_memory()
        d_members = h_members.to(dev)
        d_io = torch.from_numpy(in_off.astype(np.int64)).to(dev)
        out_off = np.arange(nm + 1, dtype=np.int64) * MEMBER
        d_oo = torch.from_numpy(out_off).to(dev)
        d_out = torch.empty(n, dtype=torch.uint8, device=dev)
        sizes = torch.zeros(nm, dtype=torch.int32, device=dev); checks = torch.zeros_like(sizes)
        status = torch.zeros_like(sizes)
        comp_bytes = int(in_off[-1])

        def step(i_timed=None):
            if i_timed is not None:
                k_start[i_timed].record()
            ctx.inflate_members(d_members, d_io, nm, 31, d_out, d_oo, sizes, checks, status, None, None)
            if i_timed is not None:
                k_stop[i_timed].record()
            launches["n"] += 1

        def parity_check(e2e_out=None):
            """EVERY member against the unmodified reference's zng_inflate(Z_FINISH) (oracle/_ref; the port if absent): return
            code, output size, CRC-32, output bytes (= the original buffer)."""
            from __graft_entry__ import load_oracle
            zo = load_oracle()
            kind = "oracle/_ref (unmodified zlib-ng 2.2.2)" if zo.have_ref() else "oracle port"
            fn = zo.ref().refdrv_inflate_members if zo.have_ref() else zo.port().zo_inflate_members
            r_out = np.empty(n, dtype=np.uint8)
            r_sizes = np.zeros(nm, dtype=np.uint32); r_crcs = np.zeros(nm, dtype=np.uint32); r_status = np.zeros(nm, dtype=np.int32)
            u_in = in_off.astype(np.uint64); u_out = out_off.astype(np.uint64)
            rr = fn(members.ctypes.data, u_in.ctypes.data, nm, r_out.ctypes.data, u_out.ctypes.data, r_sizes.ctypes.data, r_crcs.ctypes.data,
                    r_status.ctypes.data, host_cores() if ngpu == 1 else max(1, host_cores() // ngpu))
            g_status = status.cpu().numpy(); g_sizes = sizes.cpu().numpy().view(np.uint32); g_checks = checks.cpu().numpy().view(np.uint32)
            g_out = d_out.cpu().numpy()
            bad = int((g_status != r_status).sum()) + int((g_sizes != r_sizes).sum()) + int((g_checks != r_crcs).sum())
            ok = rr == 0 and bad == 0 and np.array_equal(g_out, r_out) and np.array_equal(r_out, data)
            if ok and e2e_out is not None:
                ok = np.array_equal(e2e_out, r_out)
            text = (f"all {nm} of {nm} members equal {kind}: return code, size, crc32 and output bytes" + ("; e2e host output identical" if e2e_out is not None else "")) if ok else f"MISMATCH vs {kind} ({bad} member fields differ)"
            return text, ok, {"n": n}

        def finish():
            return comp_bytes

        def make_e2e():
            h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            uin_off = in_off.astype(np.uint64); uout_off = out_off.astype(np.uint64)

            def call():
                _, _, st, _, _ = ctx.inflate_members_host(h_members, uin_off, 31, h_out, uout_off)
                assert int(st.min()) == 1
            return call, (lambda: (comp_bytes, n)), "zng_b200_inflate_members_host (pinned host in/out, 3 slabs of 32768 members in flight)", (lambda: h_out.numpy())
        e2e_fn = make_e2e
        alg_bytes = lambda ob: ob + n                                     # compressed in + raw out
        kernel_name = "inflate_members_kernel"
        note = "serial symbol decode per member (one warp each): latency/issue bound"

    torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches["n"] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(steps):
        step(i)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    k_ms = sum(a.elapsed_time(b) for a, b in zip(k_start, k_stop)) / steps
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if ngpu > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / steps
    value = ngpu * n / (ms_step * 1e-3) / 1e9
    out_bytes = finish()

    # ---- e2e: host buffers through the C-ABI host call, H2D/D2H inside the timed region (every rank at once, max over ranks)
    e2e = None
    e2e_out = None
    if not args.no_e2e and not host_memory_allows(n, ngpu):
        args.no_e2e = True
        e2e = {"value": None, "unit": UNIT, "skipped": "not enough free host memory for pinned input + output buffers of every rank"}
    if not args.no_e2e:
        call, bytes_fn, api, out_fn = e2e_fn()
        for _ in range(2):
            call()
        e_steps = max(1, steps)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            call()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if ngpu > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d, d2h = bytes_fn()
        e2e = {"value": ngpu * n * e_steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps, "api": api}
        e2e_out = out_fn()
        # the same bytes over the same pinned buffers with no kernels at all: what the box's host<->device paths allow
        passes = 2 if wl == "checksum" else 1                      # crc32_host and adler32_host each move the buffer
        up = h2d // passes
        scale = min(1.0, float(1 << 30) / max(1, up))               # a bounded sample of the step's transfers (<= 1 GiB up), same ratio
        ceil, shapes = copy_ceiling(torch, dev, h_in if wl != "inflate" else h_members, int(up * scale), int(d2h * scale), barrier, 3)
        names = sorted(shapes)
        tc = torch.tensor([shapes[k] for k in names], dtype=torch.float64, device=dev)
        if ngpu > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        rate = lambda sec: ngpu * n * scale / (passes * sec) / 1e9
        per_shape = {k: rate(float(v)) for k, v in zip(names, tc.tolist())}
        e2e["copy_ceiling"] = {"value": max(per_shape.values()), "unit": UNIT, "per_shape": per_shape,
                               "what": "H2D of the step's input and D2H of its output on two streams, same pinned input buffer, no kernels; every rank at once, max over ranks; best of the transfer shapes"}
        e2e["frac_of_copy_ceiling"] = e2e["value"] / e2e["copy_ceiling"]["value"]

    # ---- parity: EVERY unit of every rank's shard against the unmodified reference, then the cross-rank facts
    ok, facts = None, None
    if args.no_parity:
        parity = "skipped (--no-parity)"
    else:
        try:
            parity, ok, facts = parity_check(e2e_out)
        except Exception as e:  # oracle not built: say so, do not guess
            parity, ok, facts = f"not checked ({type(e).__name__}: {e})", None, None
        if ngpu > 1:
            allp = [None] * ngpu
            dist.all_gather_object(allp, (parity, ok, facts))
            if rank == 0:
                parity = cross_rank_parity(allp, wl, global_facts() if wl != "inflate" else None)

    if rank != 0:
        if ngpu > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "6650 GB/s (B200_PROFILING.md fallback)"
    ab = alg_bytes(out_bytes)
    achieved = ab / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": ab, "note": note}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        roofline["traffic"] = tr.get(wl, {}).get("dram_bytes_per_launch")
        roofline["traffic_source"] = tr.get(wl, {}).get("source")
    except Exception:
        pass

    cpu = None
    if not args.no_cpu_baseline and ngpu == 1:
        try:
            rr = cpu_reference_run(wl, cpu_sample_bytes(wl, n, 20, 4), 3, 1)     # ~20 s of wall clock
            cpu = {"value": rr["value"], "unit": UNIT, "cores": rr["cores"], "kind": rr["kind"], "sample": rr["sample"]}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}

    line = {
        "metric": METRICS[wl], "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args, ngpu),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches["n"], "clocks": clocks, "numa_node": numa,
        "compression_ratio": (out_bytes / n) if wl != "checksum" else None, "parity": parity,
        "pct_hbm_peak_input_only": 100.0 * (value / ngpu) / peak,
    }
    emit(line)
    if ngpu > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of this run, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # libraries (NCCL's version banner, torchrun notices) must not share stdout with the JSON line: everything
    # else this process or its libraries print goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
