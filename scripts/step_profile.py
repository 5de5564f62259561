#!/usr/bin/env python
"""Kernel-time table of ONE full training step (bench.py's step, 24 layers by default) from torch.profiler (CUPTI):
every kernel -- ours, cuBLAS, torch element-wise -- with launches and summed device time, so the share of the step
that is NOT in our kernels can be attributed.  Usage: python scripts/step_profile.py [--layers 24] > gpurun_out/step_profile.md"""
import argparse
import os
import re
import sys
from collections import defaultdict

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
ap.add_argument("--layers", type=int, default=24)
ap.add_argument("--top", type=int, default=70)
ap.add_argument("--width", type=int, default=100)
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
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1)
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"^void ", "", ev.name)
        name = (re.sub(r"\(.*", "", name) if "tgt::" in name or "nvjet" in name else name)[:a.width]
        agg[name][0] += 1
        agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values()) / 1e3
ours = sum(v[1] for k, v in agg.items() if "tgt::" in k) / 1e3
print(f"# step profile: {a.layers} layers, B={a.batch}, N={a.nodes}: un-profiled step {wall:.1f} ms; "
      f"kernel time {tot:.1f} ms of which tgt:: {ours:.1f} ms ({ours / tot:.3f})\n")
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / 1e3 / tot:.3f} |")
