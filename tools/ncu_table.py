#!/usr/bin/env python3
"""ncu --set full report (.ncu-rep) -> the handful of per-launch metrics kept under profiles/ (text table).
    python tools/ncu_table.py <report.ncu-rep> [...]"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 gld sectors"), ("lts__t_sectors_op_read.sum", "L2 rd sectors"),
        ("smsp__inst_executed.sum", "warp instr"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_static", "smem/blk"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak")]
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("== %s (ncu --set full --clock-control none; one row per captured launch; cold caches, kernels serialised)" % path.split("/")[-1])
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        grid = r[idx["Grid Size"]] if "Grid Size" in idx else ""
        print("%s  grid %s" % (name, grid))
        for key, label in WANT:
            if key in idx and r[idx[key]] != "":
                print("    %-18s %14s %s" % (label, r[idx[key]], units[idx[key]]))
