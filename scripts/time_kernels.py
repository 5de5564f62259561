#!/usr/bin/env python
"""Time each tgt_b200 kernel family at the config-3 shape (B=256, N=64, bf16) with CUDA events (ops.KernelTimer):
quick A/B tool for kernel tuning.  python scripts/time_kernels.py [--what triplet,egt,ffn] [--iters 5]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgt_b200 import layers as L, ops, _C          # noqa: E402
if os.environ.get("TGT_LIB"):          # A/B builds of the library
    _C.LIB_PATH = os.path.abspath(os.environ["TGT_LIB"])
from tgt_b200.harness.synthetic import make_edge_inputs   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--nodes", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--what", default="triplet,egt,ffn")
a = ap.parse_args()
dev = "cuda"
torch.manual_seed(0)
B, N = a.batch, a.nodes
nn_ = [N] + [max(N // 2, N - (i % (N // 2 + 1))) for i in range(1, B)]
e, mask = make_edge_inputs(B, N, 256, nn_, seed=0)
e = e.to(dev).bfloat16().requires_grad_(True)
mask = mask.to(dev)
h = torch.randn(B, N, 768, device=dev).bfloat16().requires_grad_(True)
mods = {}
if "triplet" in a.what:
    mods["triplet"] = L.TripletAttention(256, 16).to(dev)
if "aggregate" in a.what:
    mods["aggregate"] = L.TripletAggregate(256, 16).to(dev)
if "egt" in a.what:
    mods["egt"] = L.EGT_Attention(768, 256, 64).to(dev)
if "ffn" in a.what:
    mods["ffn"] = L.FFN(256).to(dev)


def run():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        for name, m in mods.items():
            t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            t0.record()
            if name == "egt":
                ho, eo = m(h, e, mask)
                out = ho.float().sum() + eo.float().sum()
            elif name == "ffn":
                out = m(e)
            else:
                out = m(e, mask)
            t1.record()
            if name == "egt":
                out.backward()
            else:
                out.backward(torch.ones_like(out))
            t2.record()
            mod_times.setdefault(name, []).append((t0, t1, t2))


mod_times = {}
for _ in range(2):
    run()
torch.cuda.synchronize()
mod_times = {}
ops.KernelTimer.reset(True)
for _ in range(a.iters):
    run()
torch.cuda.synchronize()
for name, evs in mod_times.items():
    f = sum(x.elapsed_time(y) for x, y, _ in evs) / len(evs)
    b = sum(y.elapsed_time(z) for _, y, z in evs) / len(evs)
    print(f"module {name:10s} fwd {f:7.3f} ms   bwd {b:7.3f} ms")
for k, (n, tot) in sorted(ops.KernelTimer.summary().items()):
    print(f"  kernel {k:24s} {tot / n:7.3f} ms  ({n} launches)")
print("peak mem GiB", torch.cuda.max_memory_allocated() / 2 ** 30)
