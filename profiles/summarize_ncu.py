#!/usr/bin/env python
"""Print the key raw-page metrics of every kernel launch in an .ncu-rep as a markdown table row set."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("smsp__inst_executed.sum", "warp insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"\n### {rep}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:70]
        cells = []
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{label} {v} {units[i]}")
        print(f"- `{name}`: " + "; ".join(cells))
