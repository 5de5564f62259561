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
res = {}
for R in (256*64*64, 256*64*16, 256*64*4, 2688*2):
    a = torch.randn(R, 256, device="cuda").bfloat16()
    for N in (1600, 256):
        w = (torch.randn(N, 256, device="cuda") / 16).bfloat16()
        out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
        for d in (0, 15):
            os.environ["TGT_GEMM_DEBUG"] = str(d)
            res[f"R{R}_N{N}_dbg{d}"] = timeit(lambda: ops.gemm_tc(a, w, out=out))
print(json.dumps(res))
