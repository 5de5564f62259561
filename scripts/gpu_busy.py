"""How busy is the GPU inside one training step?  torch.profiler kernel-time sum vs wall time of the step."""
import sys, os, json, torch
sys.path.insert(0, ".")
from tgt_b200.harness.models import TGT_Multi, pretrain_loss, TGT_AT_CONFIG
from tgt_b200.harness.synthetic import add_scheme_fields
from tgt_b200.harness.dist import rank_batch
dev = torch.device("cuda", 0)
cfg = dict(TGT_AT_CONFIG)
torch.manual_seed(0)
model = TGT_Multi(**cfg).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-5, fused=True)
raw = {k: v.to(dev) for k, v in rank_batch(256, 64, 0).items()}
def step():
    batch = add_scheme_fields(raw, with_3d=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        gap, logits = model(batch)
        loss = pretrain_loss(gap.float(), logits, batch, cfg["num_dist_bins"])
    loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("step ms (no profiler)", e0.elapsed_time(e1))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in evs) / 1e3
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print("kernels", len(evs), "sum kernel ms", round(tot, 2), "span ms", round((t1 - t0) / 1e3, 2))
import collections
agg = collections.Counter()
for e in evs: agg[e.name[:70]] += e.device_time / 1e3
for k, v in agg.most_common(12): print(f"{v:8.2f} {k}")

print("-- copy-like ops by input shape")
rows = [a for a in prof.key_averages(group_by_input_shape=True) if a.self_device_time_total > 0]
rows.sort(key=lambda a: -a.self_device_time_total)
for a in rows[:60]:
    print(f"{a.self_device_time_total / 1e3:8.2f} ms  n={a.count:4d}  {a.key[:28]:28s} {str(a.input_shapes)[:100]}")
