#!/usr/bin/env python
"""Per-source-line totals of one kernel launch from an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [launch_index=0] [top=40] [by=instr|stall]
Prints the source lines of that kernel's launch ranked by warp instructions executed (or stall samples).
"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
kern = sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
by = 1 if len(sys.argv) > 5 and sys.argv[5] == "stall" else 0
out = subprocess.run(["ncu", "-i", rep, "-k", "regex:" + kern, "-s", str(launch), "-c", "1", "--page", "source", "--csv",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
launch = 0
# the report lists launches one after another; a launch starts at a "Kernel Name"/"Function Name" header pair
blocks, cur = [], []
for row in csv.reader(io.StringIO(out)):
    if row and row[0] in ("Kernel Name",):
        if cur: blocks.append(cur)
        cur = []
    cur.append(row)
if cur: blocks.append(cur)
if not blocks:
    sys.exit("no launches in report")
rows = blocks[min(launch, len(blocks) - 1)]
hdr = None
per = collections.defaultdict(lambda: [0, 0, ""])
fname, line, src = "", "", ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].rsplit("/", 1)[-1]; continue
    if r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ci: continue
    if r[0]: line, src = r[0], r[1]
    try:
        ins, smp = int(r[ci]), int(r[cs])
    except ValueError:
        continue
    if r[2]:     # a SASS row
        k = (fname, int(line) if line else 0)
        per[k][0] += ins; per[k][1] += smp; per[k][2] = src.strip()
tot_i = sum(v[0] for v in per.values()); tot_s = sum(v[1] for v in per.values())
print(f"launch {launch}: {tot_i} warp instructions, {tot_s} stall samples, {len(blocks)} launches in report")
for (f, l), (i, s, t) in sorted(per.items(), key=lambda kv: -kv[1][by])[:top]:
    print(f"{i:9d} {100.0*i/max(tot_i,1):5.1f}%  smp {s:5d} {100.0*s/max(tot_s,1):5.1f}%  {f}:{l}  {t[:90]}")
