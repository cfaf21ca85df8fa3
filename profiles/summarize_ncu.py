#!/usr/bin/env python
"""Condense `ncu --set full` reports into the few numbers DESIGN.md and
bench.py quote:  python profiles/summarize_ncu.py gpurun_out/*.ncu-rep
Writes profiles/<name>.summary.csv next to this script and prints markdown."""
import csv
import io
import os
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"

here = os.path.dirname(os.path.abspath(__file__))
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name = os.path.splitext(os.path.basename(rep))[0]
    out = [("kernel", "", "")]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out = [("kernel", "", d["Kernel Name"])]
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
                out.append((h, units[i], r[i]))
        break  # first captured launch
    with open(os.path.join(here, name + ".summary.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerows(out)
    print(f"### {name}")
    for m, u, v in out:
        print(f"- `{m}` = {v} {u}")
