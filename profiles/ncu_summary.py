#!/usr/bin/env python
"""Summarise an .ncu-rep capture: headline counters + instruction share per SASS region.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [chunk]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 200
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_register_spilling", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"  {k:70s} {r[i]} {units[i]}")
    st = [(h, r[i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    st = sorted(((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v)) for h, v in st if v), key=lambda x: -x[1])[:8]
    print("  stall (warps per issue):", ", ".join(f"{h}={v:.2f}" for h, v in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, s0 in enumerate(start[:1]):
    hdr = rows[s0]; body = rows[s0 + 1:(start[si + 1] - 1 if si + 1 < len(start) else len(rows))]
    ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = sum(int(r[ia]) for r in body if r[ia].isdigit()); tots = sum(int(r[ismp]) for r in body if r[ismp].isdigit()) or 1
    print(f"-- SASS regions ({len(body)} instructions, {tot} warp-instructions executed)")
    for c0 in range(0, len(body), chunk):
        ch = body[c0:c0 + chunk]
        n = sum(int(r[ia]) for r in ch if r[ia].isdigit()); sm = sum(int(r[ismp]) for r in ch if r[ismp].isdigit())
        if n == 0: continue
        ops = {}
        for r in ch:
            t = r[isrc].split(); op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
            ops[op] = ops.get(op, 0) + (int(r[ia]) if r[ia].isdigit() else 0)
        top = ", ".join(f"{k}:{v * 100 // tot}%" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:6])
        print(f"   [{c0:5d}-{c0 + len(ch) - 1:5d}] inst {n / tot * 100:5.1f}%  samples {sm / tots * 100:5.1f}%   {top}")
