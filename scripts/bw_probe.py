"""HBM read / write / copy bandwidth probe (torch ops; CUDA events): is write-only traffic capped below the copy peak?"""
import torch
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
n = 1 << 31                                   # 2 Gi bf16 elements = 4 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
gb = a.numel() * 2 / 1e9
t = timeit(lambda: b.zero_());            print(f"fill   : {t:.3f} ms  write {gb / t:.2f} TB/s")
t = timeit(lambda: b.copy_(a));           print(f"copy   : {t:.3f} ms  total {2 * gb / t:.2f} TB/s")
t = timeit(lambda: a.view(torch.int16).max()); print(f"reduce : {t:.3f} ms  read  {gb / t:.2f} TB/s")
t = timeit(lambda: torch.add(a[: n // 2], a[n // 2:], out=b[: n // 2])); print(f"add 2r1w: {t:.3f} ms total {1.5 * gb / t:.2f} TB/s")
