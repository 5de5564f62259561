import os, sys, json, torch
sys.path.insert(0, "/root/repo")
from tgt_b200 import ops
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters, 4)
R = 256 * 64 * 64
a = torch.randn(R, 256, device="cuda").bfloat16()
res = {}
for N in (1600,):
    w = (torch.randn(N, 256, device="cuda") / 16).bfloat16()
    out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
    for d in (0, 1, 2, 3, 4, 8, 9, 10, 11, 12, 15):
        os.environ["TGT_GEMM_DEBUG"] = str(d)
        res[f"N{N}_dbg{d}"] = timeit(lambda: ops.gemm_tc(a, w, out=out))
print(json.dumps(res))
