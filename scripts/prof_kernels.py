#!/usr/bin/env python
"""Run each tgt_b200 kernel family a few times at the BASELINE config-3 shape (B=256, N=64, We=256, Ht=16, Wn=768,
Hn=64, bf16) so that ncu can capture them:  ncu --set full -k regex:<pattern> -c <n> python scripts/prof_kernels.py"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgt_b200 import layers as L          # noqa: E402
from tgt_b200.harness.synthetic import make_edge_inputs   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--nodes", type=int, default=64)
ap.add_argument("--iters", type=int, default=2)
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
for it in range(a.iters):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        if "triplet" in mods:
            mods["triplet"](e, mask).float().sum().backward()
        if "aggregate" in mods:
            mods["aggregate"](e, mask).float().sum().backward()
        if "egt" in mods:
            ho, eo = mods["egt"](h, e, mask)
            (ho.float().sum() + eo.float().sum()).backward()
        if "ffn" in mods:
            mods["ffn"](e).float().sum().backward()
    torch.cuda.synchronize()
print("ok")
