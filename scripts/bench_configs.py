#!/usr/bin/env python
"""Throughput of the BASELINE.json parity configs that are not the bench.py headline, on ONE GPU (they shard over GPUs by
graphs with no collective, like the reference's per-rank prediction shards, dist_pred/scheme.py:301-305):

  config 2: TGT-Agx2 12L x2 gap predictor, forward (no_grad, train mode = MC-dropout inference), B=256, N=32, bf16
  config 5: TGT-At distance predictor (256 bins, no 3-D input) -> bins -> distances -> TGT-At gap predictor, B=512, N=48,
            S Monte-Carlo samples per stage (reference default 50; stated in the output), bf16

Prints one JSON line per config.  Usage: python scripts/bench_configs.py [--samples 2] [--iters 3]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgt_b200 import _C  # noqa: E402
from tgt_b200.harness import inference as INF  # noqa: E402
from tgt_b200.harness.models import TGT_AGX2_CONFIG, TGT_AT_CONFIG, TGT_Distance, TGT_Gap  # noqa: E402
from tgt_b200.harness.synthetic import add_scheme_fields, make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=2)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--which", default="2,5")
a = ap.parse_args()
# under torchrun every rank runs its own graphs (the reference's per-rank prediction shards): no collective on the data path,
# one barrier + max-over-ranks for the timing
RANK, LOCAL, WORLD = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(LOCAL)
dev = torch.device("cuda", LOCAL)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
_C.lib()


def over_ranks(ms):
    if WORLD == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return over_ranks(e0.elapsed_time(e1) / iters)


if "2" in a.which.split(","):
    B, N = 256, 32
    torch.manual_seed(0)
    gm = TGT_Gap(**TGT_AGX2_CONFIG).to(dev).train()
    batch = add_scheme_fields({k: v.to(dev) for k, v in make_batch(B, N, seed=1 + RANK).items()}, with_3d=True)

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return gm(batch)
    n0 = _C.launch_count()
    ms = timed(fwd, a.iters)
    launches = (_C.launch_count() - n0) // (a.iters + 1)
    # the same forward captured once as a CUDA graph and replayed (MC-dropout sampling draws fresh masks per replay)
    gf = INF.GraphedForward(gm, batch)
    outs = [gf().float().clone() for _ in range(2)]
    assert bool(torch.isfinite(outs[0]).all()) and not torch.equal(outs[0], outs[1]), "replays must resample dropout"
    ms_graph = timed(lambda: gf(), max(a.iters, 5))
    if RANK == 0:
        print(json.dumps(dict(config=f"2: TGT-Agx2 12Lx2 gap-predictor forward (train mode = MC-dropout sample, no_grad), "
                                     f"B=256 (per GPU) N=32, bf16, {WORLD}xB200",
                              ms_per_batch_eager=ms, molecules_per_s_eager=WORLD * B / ms * 1e3,
                              ms_per_batch=ms_graph, molecules_per_s=WORLD * B / ms_graph * 1e3,
                              note="ms_per_batch = CUDA-graph replay of the captured forward; *_eager = launched op by op",
                              tgt_b200_launches_per_forward=launches,
                              peak_mem_gib=torch.cuda.max_memory_allocated() / 2 ** 30)))
    del gf
    del gm, batch
    torch.cuda.empty_cache()

if "5" in a.which.split(","):
    B, N, S = 512, 48, a.samples
    torch.manual_seed(0)
    cfg = {k: v for k, v in TGT_AT_CONFIG.items() if k != "num_dist_bins"}
    dm = TGT_Distance(num_dist_bins=256, embed_3d_type="none", **cfg).to(dev).train()
    gm = TGT_Gap(**cfg).to(dev).train()
    batch = {k: v.to(dev) for k, v in make_batch(B, N, seed=1 + RANK).items()}
    out = {}

    def two_stage():
        out["pred"] = INF.two_stage_predict(dm, gm, batch, samples=S)
    torch.cuda.reset_peak_memory_stats()
    n0 = _C.launch_count()
    ms = timed(two_stage, a.iters)
    assert bool(torch.isfinite(out["pred"]).all())
    if RANK == 0:
        print(json.dumps(dict(config=f"5: TGT-At distance-predictor + gap-predictor two-stage inference, B=512 per GPU "
                                     f"(global {512 * WORLD}) N=48, S={S} MC samples per stage (reference default 50, "
                                     f"README.md:81), bf16, {WORLD}xB200, max over ranks", ms_per_batch=ms,
                              molecules_per_s=WORLD * B / ms * 1e3, forward_passes_per_batch=2 * S,
                              molecule_passes_per_s=WORLD * B * 2 * S / ms * 1e3, gpu_launches=_C.launch_count() - n0,
                              peak_mem_gib=torch.cuda.max_memory_allocated() / 2 ** 30)))

if WORLD > 1:
    dist.destroy_process_group()
