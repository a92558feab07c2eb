#!/usr/bin/env python
"""Markdown summary of an ncu report (``ncu --set full``): one block per profiled launch with the metrics the roofline
arguments in DESIGN.md use.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [title] > profiles/x.md"""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_size", "cluster"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "warps stalled on long scoreboard / issue"),
]
path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {title}\n\nSource: `{path}` (`ncu --set full --clock-control none`; per-launch values, cold cache, serialised).\n")
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    print(f"## `{name[:140]}`\n\n| metric | value |\n|---|---|")
    for key, label in WANT:
        if key in col and r[col[key]] != "":
            print(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
    print()
