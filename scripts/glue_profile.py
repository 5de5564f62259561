#!/usr/bin/env python
"""Which torch (non-tgt) operators still run in the training step, by operator, input shapes and Python call site:
torch.profiler with record_shapes + with_stack over one 4-layer step; prints the aten ops with the largest CUDA time.
Usage: python scripts/glue_profile.py [--layers 4] > gpurun_out/glue_profile.txt"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tgt_b200.harness.models import TGT_Multi, pretrain_loss  # noqa: E402
from tgt_b200.harness.synthetic import add_scheme_fields  # noqa: E402
from tgt_b200.harness.dist import rank_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--nodes", type=int, default=64)
ap.add_argument("--layers", type=int, default=4)
a = ap.parse_args()
a.fp32_logits = False
dev = torch.device("cuda", 0)
cfg = bench.model_cfg(a)
torch.manual_seed(0)
model = TGT_Multi(**cfg).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-5, fused=True)
raw = {k: v.to(dev) for k, v in rank_batch(a.batch, a.nodes, 0).items()}


def step():
    batch = add_scheme_fields(raw, with_3d=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        gap, logits = model(batch)
        loss = pretrain_loss(gap.float(), logits, batch, cfg["num_dist_bins"])
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    if ev.key.startswith("aten::") and ev.self_device_time_total > 0:
        stack = [s for s in ev.stack if "tgt_b200" in s or "bench" in s or "harness" in s][:2]
        rows.append((ev.self_device_time_total / 1e3, ev.count, ev.key, str(ev.input_shapes)[:90], " <- ".join(s.split("/")[-1][:60] for s in stack)))
rows.sort(reverse=True)
print(f"layers={a.layers}; aten ops with self CUDA time, ms per step")
for r in rows[:45]:
    print(f"{r[0]:8.3f} x{r[1]:<4d} {r[2]:28s} {r[3]:90s} {r[4]}")
