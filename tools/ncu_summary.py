#!/usr/bin/env python3
"""Summarise ncu output into the small text files kept under profiles/.

  ncu_summary.py launches <launches.csv>             per-kernel table of a `--metrics gpu__time_duration.sum` pass
  ncu_summary.py full <report.ncu-rep> [out.json]    per-kernel key metrics of a `--set full` capture
                                                     (out.json: {"kernel": {"dram_bytes_per_launch": ...}})
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", lambda m: m.group(0) if len(m.group(0)) < 12 else "<...>", name)
    return name.split("::")[-1] if "cub" not in name else "cub::" + name.split("::")[-1]


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = collections.OrderedDict()
    total = 0.0
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += ns; a[2] = min(a[2], ns); a[3] = max(a[3], ns)
        total += ns
    print(f"{'kernel':44s} {'grid':>14s} {'block':>12s} {'n':>5s} {'avg us':>9s} {'min us':>9s} {'max us':>9s} {'share':>7s}")
    for (k, g, b), (n, s, lo, hi) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {g:>14s} {b:>12s} {n:5d} {s / n / 1e3:9.2f} {lo / 1e3:9.2f} {hi / 1e3:9.2f} {100 * s / total:6.1f}%")
    print(f"total {total / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised under ncu)")


KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_sectors_op_read.sum", "L2 read sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global load requests"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occ %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM busy %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem pipe %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads/instr"),
]
STALL = "smsp__average_warps_issue_stalled_"


def full(path, out_json=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows:
        name = short(r[col["Kernel Name"]])
        print(f"== {name}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for key, label in KEYS:
            if key in col:
                print(f"   {label:28s} {r[col[key]]:>16s} {units[col[key]]}")
        stalls = []
        for h, i in col.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   top stalls (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))

        def num(k):
            return float(r[col[k]].replace(",", "")) if k in col and r[col[k]] else 0.0

        def to_bytes(k):
            u = units[col[k]].lower() if k in col else "byte"
            return num(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

        d = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0})
        d["launches"] += 1
        d["dram_bytes"] += to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    if out_json:
        res = {k: {"dram_bytes_per_launch": v["dram_bytes"] / v["launches"], "launches_captured": v["launches"]} for k, v in out.items()}
        json.dump(res, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
