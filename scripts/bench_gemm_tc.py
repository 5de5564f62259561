"""Micro-benchmark of tgt_gemm_tc against cuBLAS (torch.addmm) at the config-3 edge GEMM shapes."""
import sys, json
import torch
sys.path.insert(0, ".")
from tgt_b200 import ops

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

R = 256 * 64 * 64
res = {}
for name, (N, K) in {"proj_1600x256": (1600, 256), "eg_128x256": (128, 256), "ffn_256x256": (256, 256),
                     "linO_256x512": (256, 512), "linOe_256x64": (256, 64), "dva_512x256": (512, 256)}.items():
    a = torch.randn(R, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    bb = b.bfloat16()
    out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
    t_tc = timeit(lambda: ops.gemm_tc(a, w, bias=b, out=out))
    t_cb = timeit(lambda: torch.addmm(bb, a, w.t(), out=out))
    fl = 2.0 * R * N * K
    by = 2.0 * (R * K + R * N)
    res[name] = dict(tc_ms=t_tc, cublas_ms=t_cb, tc_tflops=fl / t_tc / 1e9, cublas_tflops=fl / t_cb / 1e9,
                     tc_gbs=by / t_tc / 1e6, cublas_gbs=by / t_cb / 1e6)
    print(name, {k: round(v, 3) for k, v in res[name].items()}, flush=True)
    del a, out
# fused epilogues
a = torch.randn(R, 256, device="cuda").bfloat16()
w = (torch.randn(256, 256, device="cuda") / 16).bfloat16()
b = torch.randn(256, device="cuda")
r = torch.randn(R, 256, device="cuda").bfloat16()
out = torch.empty_like(r)
mean, rstd = ops.row_stats(a)
cs = w.float().sum(1)
sc = torch.ones(256, device="cuda")
print("row_stats", round(timeit(lambda: ops.row_stats(a)), 3))
print("ln+bias", round(timeit(lambda: ops.gemm_tc(a, w, bias=b, ln=(mean, rstd, cs), out=out)), 3))
print("bias+res", round(timeit(lambda: ops.gemm_tc(a, w, bias=b, res=r, row_scale=sc, rows_per_scale=4096, out=out)), 3))
print("gelu p=0", round(timeit(lambda: ops.gemm_tc(a, w, bias=b, gelu=(0.0, 1), out=out)), 3))
print("gelu p=.1", round(timeit(lambda: ops.gemm_tc(a, w, bias=b, gelu=(0.1, 1), out=out)), 3))
json.dump(res, open("gpurun_out/bench_gemm_tc.json", "w"), indent=1)
