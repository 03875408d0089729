#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): the metrics DESIGN.md / profiles/ quote.
Usage: scripts/ncu_summary.py gpurun_out/prof_solve_r01.ncu-rep [--stalls]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warp_issue_stalled", "smsp__average_warp"]
for i, h in enumerate(hdr):
    if any(h == w or (w in h and ("stalled" in w or "average_warp" in w)) for w in want):
        vals = [r[i] for r in data]
        if "stalled" in h or "average_warp" in h:
            if "--stalls" not in sys.argv or not h.endswith(".ratio"):
                continue
        print(f"{h} [{units[i]}] = {vals}")
