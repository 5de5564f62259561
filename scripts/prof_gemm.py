"""One launch of each interesting tgt_gemm_tc configuration at config-3 size (for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from tgt_b200 import ops
R = 256 * 64 * 64
a = torch.randn(R, 256, device="cuda").bfloat16()
mean, rstd = ops.row_stats(a)
for N in (1600, 256):
    w = (torch.randn(N, 256, device="cuda") / 16).bfloat16()
    b = torch.randn(N, device="cuda")
    out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_tc(a, w, bias=b, ln=(mean, rstd, w.float().sum(1)), out=out)
r = torch.randn(R, 256, device="cuda").bfloat16()
sc = torch.ones(256, device="cuda")
ops.gemm_tc(a, w, bias=b, res=r, row_scale=sc, rows_per_scale=4096, out=out)
ops.gemm_tc(a, w, bias=b, gelu=(0.1, 1), out=out)
torch.cuda.synchronize()
