#!/usr/bin/env python
"""Time the triplet-attention CORE kernels alone at the config-3 shape (B=256, N=64, We=256, Ht=16, bf16) under each
kernel policy, with the library's own device-side timer (CUDA events around the main kernel only):
    python scripts/time_triplet_core.py [--policies 4,5] [--iters 5] [--bwd]
Prints one JSON line per policy: ms per launch, achieved GB/s on the algorithmic bytes, fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tgt_b200 import _C          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--nodes", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--policies", default="4,5")
ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
dev = "cuda"
B, N, W, H, d = a.batch, a.nodes, 256, 16, 16
C = 6 * W + 4 * H
R = B * N * N
g = torch.Generator(device=dev).manual_seed(0)
proj = (torch.randn(R, C, device=dev, generator=g) * 1.2).bfloat16()
nn_ = torch.tensor([N] + [max(N // 2, N - (i % (N // 2 + 1))) for i in range(1, B)], device=dev)
m = (torch.arange(N, device=dev)[None, :] < nn_[:, None]).float()
mask = ((1 - m[:, :, None] * m[:, None, :]) * torch.finfo(torch.float32).min).contiguous()
desc = _C.TripletAttnDesc(B, N, H, d, C, (0, 3 * W), (W, 4 * W), (2 * W, 5 * W), (6 * W, 6 * W + 2 * H),
                          (6 * W + H, 6 * W + 3 * H), d ** -0.5, _C.BF16)
va = torch.empty((R, 2 * W), dtype=torch.bfloat16, device=dev)
dva = (torch.randn(R, 2 * W, device=dev, generator=g)).bfloat16()
dproj = torch.empty_like(proj)
stats = torch.empty((B, 2, H, N, N, 2), dtype=torch.float32, device=dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
bytes_fwd = (8 * R * W + 4 * R * H) * 2 + 4 * R
bytes_bwd = (16 * R * W + 8 * R * H) * 2 + 4 * R
lib = _C.lib()
for pol in [int(x) for x in a.policies.split(",")]:
    _C.set_kernel_policy(pol)
    nb = lib.tgt_triplet_attn_workspace_bytes(desc, 0)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    nbb = lib.tgt_triplet_attn_workspace_bytes(desc, 1)
    wsb = torch.empty(nbb, dtype=torch.uint8, device=dev)

    def run():
        _C.check(lib.tgt_triplet_attn_fwd(desc, _C.ptr(proj), _C.ptr(mask), _C.ptr(va), _C.ptr(stats), _C.ptr(ws), nb,
                                          _C.stream_ptr()), "fwd")
        if a.bwd:
            _C.check(lib.tgt_triplet_attn_bwd_tiles(desc, _C.ptr(proj), _C.ptr(mask), _C.ptr(va), _C.ptr(dva), _C.ptr(stats),
                                                    _C.ptr(dproj), _C.ptr(wsb), nbb, _C.ptr(ws), _C.stream_ptr()), "bwd")
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    _C.kernel_timer(True)
    for _ in range(a.iters):
        run()
    torch.cuda.synchronize()
    out = {}
    for name, (n, tot) in _C.kernel_timer_read().items():
        ms = tot / n
        by = bytes_bwd if "bwd" in name else bytes_fwd
        out[name] = dict(ms=round(ms, 4), gbs=round(by / ms / 1e6, 1), hbm_frac=round(by / ms / 1e6 / peaks["hbm_gbs"], 3))
    _C.kernel_timer(False)
    print(json.dumps(dict(policy=pol, B=B, N=N, finite=bool(torch.isfinite(va.float()).all()), kernels=out)))
_C.set_kernel_policy(0)
