#!/usr/bin/env python
"""Per CUDA-source-line sample share from an .ncu-rep (needs -lineinfo and --import-source on).
usage: python profiles/ncu_lines.py rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
files = {}
cur = None
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or cur is None: continue
    if r[0].isdigit():
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        files.setdefault(cur, []).append((int(r[0]), r[1].strip(), int(r[i_s]) if r[i_s].isdigit() else 0, int(r[i_i]) if r[i_i].isdigit() else 0))
allrows = [(f, *x) for f, v in files.items() for x in v]
ts = sum(x[3] for x in allrows) or 1; ti = sum(x[4] for x in allrows) or 1
print(f"total samples {ts}, instructions {ti}")
for f, ln, src, smp, ins in sorted(allrows, key=lambda x: -x[3])[:top]:
    print(f"{smp*100/ts:5.1f}% smp {ins*100/ti:5.1f}% ins  {f.split('/')[-1]}:{ln}  {src[:100]}")
