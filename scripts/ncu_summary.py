#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote:
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]   (needs only the ncu CLI, no GPU)"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("## kernel:", d.get("Kernel Name", "?")[:100], "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:80s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and v),
                        reverse=True)
        print("  stalls per issue:", ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls[:9]))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hdr = rows[0]
        try:
            ci = hdr.index("# Samples") if "# Samples" in hdr else [i for i, h in enumerate(hdr) if "Samples" in h][0]
        except IndexError:
            print("no sample column:", hdr[:12])
            return
        si = hdr.index("Source")
        body = [r for r in rows[1:] if len(r) > ci and r[ci].replace(',', '').isdigit()]
        tot = sum(int(r[ci].replace(',', '')) for r in body) or 1
        for r in sorted(body, key=lambda r: -int(r[ci].replace(',', '')))[:n]:
            print(f"  {int(r[ci].replace(',', '')) / tot:6.2%}  {r[si][:110]}")


if __name__ == "__main__":
    main()
