#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
Usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
agg = defaultdict(lambda: [0, 0.0, "", ""])
for name, ns, grid, block in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)[:90]
    a = agg[short]
    a[0] += 1; a[1] += ns; a[2] = grid; a[3] = block
total = sum(a[1] for a in agg.values())
print(f"# ncu launch list summary: {len(rows)} launches, {total / 1e6:.3f} ms total (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total ms | share | last grid | block |")
print("|---|---:|---:|---:|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / total:.3f} | {a[2]} | {a[3]} |")
