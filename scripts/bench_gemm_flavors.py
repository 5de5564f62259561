"""Times the fused-epilogue flavours of tgt_gemm_tc at the config-3 edge shapes (R = 256*64*64 rows), CUDA events around
10 launches after 3 warm-ups; prints one JSON line.  TGT_GEMM_WQ selects the epilogue warps per TMEM quadrant."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgt_b200 import _C, ops
if os.environ.get("TGT_LIB"):          # A/B builds of the library (scripts only)
    _C.LIB_PATH = os.path.abspath(os.environ["TGT_LIB"])


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters, 4)


R = 256 * 64 * 64
a = torch.randn(R, 256, device="cuda").bfloat16()
mean, rstd = ops.row_stats(a)
res = {"wq": os.environ.get("TGT_GEMM_WQ", "default")}
for N in (1600, 128, 256):
    w = (torch.randn(N, 256, device="cuda") / 16).bfloat16()
    b = torch.randn(N, device="cuda")
    out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
    cs = w.float().sum(1)
    res[f"ln_bias_N{N}"] = timeit(lambda: ops.gemm_tc(a, w, bias=b, ln=(mean, rstd, cs), out=out))
    res[f"plain_N{N}"] = timeit(lambda: ops.gemm_tc(a, w, out=out))
    ref = ((a[:4096].float() - mean[:4096, None]) * rstd[:4096, None]) @ w.float().t() + b
    got = ops.gemm_tc(a, w, bias=b, ln=(mean, rstd, cs), out=out)[:4096].float()
    res[f"err_N{N}"] = float((got - ref).norm() / ref.norm())
w = (torch.randn(256, 256, device="cuda") / 16).bfloat16()
b = torch.randn(256, device="cuda")
r = torch.randn(R, 256, device="cuda").bfloat16()
sc = torch.ones(256, device="cuda")
out = torch.empty_like(r)
res["bias_res"] = timeit(lambda: ops.gemm_tc(a, w, bias=b, res=r, row_scale=sc, rows_per_scale=4096, out=out))
res["ln_gelu_p.1"] = timeit(lambda: ops.gemm_tc(a, w, bias=b, ln=(mean, rstd, w.float().sum(1)), gelu=(0.1, 1), store_gp=True))
a5 = torch.randn(R, 512, device="cuda").bfloat16()
w5 = (torch.randn(256, 512, device="cuda") / 22).bfloat16()
res["bias_res_K512"] = timeit(lambda: ops.gemm_tc(a5, w5, bias=b, res=r, out=out))
res["plain_K512"] = timeit(lambda: ops.gemm_tc(a5, w5, out=out))
a6 = torch.randn(R, 64, device="cuda").bfloat16()
w6 = (torch.randn(256, 64, device="cuda") / 8).bfloat16()
res["bias_res_K64"] = timeit(lambda: ops.gemm_tc(a6, w6, bias=b, res=r, row_scale=sc, rows_per_scale=4096, out=out))
res["lib"] = os.environ.get("TGT_LIB", "in-tree")
print(json.dumps(res))
