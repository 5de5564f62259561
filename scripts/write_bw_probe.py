"""HBM bandwidth by access mix (torch fill / copy on 3.3 GB buffers): pure write, pure read (sum), 1:1 copy."""
import json, torch
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
n = 256 * 64 * 64 * 1600
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
gb = n * 2 / 1e9
res = dict(write_gbs=gb / timeit(lambda: a.zero_()) * 1e3, fill_gbs=gb / timeit(lambda: a.fill_(1.5)) * 1e3,
           copy_gbs=2 * gb / timeit(lambda: b.copy_(a)) * 1e3,
           read_gbs=gb / timeit(lambda: torch.ops.aten.sum(a.view(torch.int16)[: n // 2 * 2].view(torch.int32))) * 1e3)
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
