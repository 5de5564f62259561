#!/usr/bin/env python
"""Summarise an ncu --set full report (read on the CPU box): one row per captured launch with the metrics the
roofline argument needs.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
units = rows[1]
print(f"# ncu --set full summary of {sys.argv[1]}\n")
for r in rows[2:]:
    name = r[col["Kernel Name"]][:70]
    print(f"## {r[col['ID']]} `{name}`  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for m in METRICS:
        if m in col:
            print(f"- {m}: {r[col[m]]} {units[col[m]]}")
    print()
