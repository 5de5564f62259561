#!/usr/bin/env python
"""bench.py -- molecules/s of a TGT-At 24-layer training step (fwd + loss + bwd + Adam) at batch 256 x N=64
(BASELINE.json configs[2]/[3]) on N x B200, plus the roofline of our dominant kernel and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every field.

  value        : molecules/s, inputs already resident in HBM, K steps timed with CUDA events, max over ranks
  e2e          : same step driven from pinned HOST buffers (H2D copy of the raw batch + D2H read of the loss
                 inside the timed region)
  roofline     : our kernel with the largest share of the step, timed with CUDA events INSIDE the step
  kernels      : the same for every one of our kernels (share of step, achieved vs bound)
  cpu_baseline : the unmodified reference (oracle/_ref, staged by oracle/build_ref.py; kind "reference") on the host
                 cores, on a bounded micro-batch of the same workload (falls back to the oracle port, kind "port")
  gpu_eager_baseline : the same unmodified reference on this B200 under PyTorch-CUDA bf16 autocast at the largest
                 micro-batch that fits (N = 1 only)
  library_calls_per_step : every GEMM that still goes to cuBLAS through torch, by call site
  --impl reference : only the CPU path, K steps of a bounded micro-batch each.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "molecules/sec TGT-At 24L fwd+bwd @ batch256 N=64"
UNIT = "molecules/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tc_burst=float(p["bf16_tflops"]),
                    tc_sustained=float(p["bf16_tflops_sustained"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ model config
def model_cfg(args):
    from tgt_b200.harness.models import TGT_AT_CONFIG
    cfg = dict(TGT_AT_CONFIG)
    cfg["model_height"] = args.layers
    return cfg


# ------------------------------------------------------------------------------------------------ CPU baseline
def _reference_kwargs(cfg):
    """TGT_AT_CONFIG -> kwargs of the reference's lib.models.pcqm.multitask.TGT_Multi (same names)."""
    return dict(cfg)


def reference_step_fn(args, B, device, autocast_dtype=None, fused_adam=False):
    """The UNMODIFIED reference (oracle/_ref = lib.tgt + lib.models.pcqm + commons, staged by oracle/build_ref.py):
    TGT_Multi forward + the pretrain loss of pretrain/scheme.py:78-88 + backward + Adam, train mode, shipped dropouts."""
    from oracle import ref_loader as R                          # checker / baseline only
    from tgt_b200.harness.synthetic import make_batch
    cfg = model_cfg(args)
    stack = R.reference()
    ref = stack.__enter__()                                     # stays imported for the lifetime of the process
    torch.manual_seed(0)
    model = ref.TGT_Multi(**_reference_kwargs(cfg)).to(device).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, **(dict(fused=True) if fused_adam else {}))
    batch = {k: v.to(device) for k, v in make_batch(B, args.nodes, seed=1).items()}
    loss_fn = ref.commons.DiscreteDistLoss(cfg["num_dist_bins"], 8)

    def step():
        if autocast_dtype is not None:
            with torch.autocast("cuda", dtype=autocast_dtype):
                gap, logits = model(batch)
        else:
            gap, logits = model(batch)
        loss = (torch.nn.functional.l1_loss(gap.float(), batch["target"].float())
                + 0.1 * loss_fn(logits.float(), ref.commons.coords2dist(batch["dft_coords"]), batch["edge_mask"]))
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss.detach())
    return step


def port_step_fn(args, B):
    """Fallback when oracle/_ref is not staged: the oracle port of the same step (oracle/tgt_oracle.py)."""
    from oracle import tgt_oracle as O                       # checker / baseline only
    from tgt_b200.harness.models import TGT_Multi
    from tgt_b200.harness.synthetic import make_batch
    cfg = model_cfg(args)
    torch.manual_seed(0)
    model = TGT_Multi(**cfg)                                  # parameter container only (never called on CPU)
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    opt = torch.optim.Adam([v for v in params.values() if v.requires_grad], lr=1e-5)
    batch = make_batch(B, args.nodes, seed=1)
    ocfg = {k: v for k, v in cfg.items() if k in ("num_heads", "triplet_heads", "triplet_type", "activation",
                                                   "scale_degree")}

    def step():
        gap, logits = O.tgt_multi(params, batch, model_height=cfg["model_height"], upto_hop=cfg["upto_hop"], **ocfg)
        loss = O.pretrain_loss(gap, logits, batch, cfg["num_dist_bins"])
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss.detach())
    return step


def cpu_baseline(args, B, steps, warmup):
    from oracle import ref_loader as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "reference" if R.available() else "port"
    step = reference_step_fn(args, B, "cpu") if kind == "reference" else port_step_fn(args, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    what = ("unmodified reference lib.models.pcqm.multitask.TGT_Multi on lib.tgt (oracle/_ref)" if kind == "reference"
            else "oracle port of lib.tgt + lib.models.pcqm")
    return dict(value=B / dt, unit=UNIT, cores=cores, kind=kind,
                sample=f"{what}, PyTorch-CPU fp32 eager, train mode, TGT-At {args.layers}L N={args.nodes}, micro-batch "
                       f"{B} molecules/step (fwd + pretrain loss + bwd + Adam), {steps} timed step(s) after {warmup} "
                       f"warm-up, {dt:.2f} s/step"), dt


def gpu_eager_baseline(args, dev):
    """The reference itself on THIS B200 ('the bar to beat'): unmodified lib.tgt / lib.models.pcqm under
    PyTorch-CUDA bf16 autocast, eager, fused Adam, at the largest micro-batch of the ladder that fits (the reference
    materialises and saves several [B,N,N,N,H] tensors per layer, so batch 256 cannot fit: SURVEY 2.3 note)."""
    from oracle import ref_loader as R
    if not R.available():
        return dict(unavailable="oracle/_ref not staged")
    tried = []
    for mb in (64, 48, 32, 24, 16, 8, 4):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        try:
            step = reference_step_fn(args, mb, dev, torch.bfloat16, fused_adam=True)
            step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 2
            for _ in range(n):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / n
            return dict(value=mb / (ms * 1e-3), unit=UNIT, micro_batch=mb, ms_per_step=ms,
                        peak_mem_gib=torch.cuda.max_memory_allocated(dev) / 2 ** 30, oom_at=tried,
                        what=f"unmodified reference (oracle/_ref) TGT_Multi TGT-At {args.layers}L N={args.nodes}, "
                             f"PyTorch-CUDA eager, bf16 autocast, train mode, fwd + pretrain loss + bwd + fused Adam, "
                             f"micro-batch {mb} (largest of 64/48/32/24/16/8/4 that fits), 1 warm-up + {n} timed steps")
        except torch.OutOfMemoryError:
            tried.append(mb)
            step = None
            import gc
            gc.collect()
    return dict(unavailable=f"out of memory at every micro-batch tried {tried}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch or 2
    base, dt = cpu_baseline(args, B, args.steps, args.warmup)
    out = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
               data="synthetic", impl="reference",
               config=dict(workload=f"TGT-At {args.layers}L training step (fwd+loss+bwd+Adam), N={args.nodes}, "
                                    f"CPU micro-batch {B} (bounded sample of the batch-{args.batch} workload)",
                           inputs="synthetic PCQM-shaped batch, random-init weights"),
               cpu_baseline=base,
               e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(out)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ roofline model
def kernel_models(B, N, We, Ht, Wn, Hn, s=2):
    """Algorithmic bytes / flops per launch of each of our kernels at this shape (DESIGN.md 'Kernels')."""
    R = B * N * N
    d = {}
    d["triplet_attn_fwd"] = dict(bytes=(8 * R * We + 4 * R * Ht) * s + 4 * R, flops=8 * B * N ** 3 * We)
    d["triplet_attn_bwd"] = dict(bytes=(16 * R * We + 8 * R * Ht) * s + 4 * R, flops=16 * B * N ** 3 * We)   # dA, dV, dQ, dK per direction; the S recompute is not algorithmic work
    d["triplet_aggr_fwd"] = dict(bytes=(4 * R * We + 4 * R * Ht) * s + 4 * R, flops=4 * B * N ** 3 * We)
    d["triplet_aggr_bwd"] = dict(bytes=(8 * R * We + 8 * R * Ht) * s + 4 * R, flops=8 * B * N ** 3 * We)
    d["egt_attn_fwd"] = dict(bytes=(3 * R * Hn + 4 * B * N * Wn) * s + 4 * R, flops=4 * B * N * N * Wn)
    d["egt_attn_bwd"] = dict(bytes=(5 * R * Hn + 8 * B * N * Wn) * s + 4 * R, flops=10 * B * N * N * Wn)
    for W in (We, Wn):
        rows = R if W == We else B * N
        d[f"layernorm_fwd_W{W}"] = dict(bytes=2 * rows * W * s, flops=0)
        d[f"layernorm_bwd_W{W}"] = dict(bytes=3 * rows * W * s, flops=0)
    d[f"row_stats_W{We}"] = dict(bytes=R * We * s, flops=0)

    def gemm(N_, K_, extra_cols=0):        # tcgen05 GEMM on the R edge rows: read A, write D (+ extra edge-sized operands)
        return dict(bytes=R * (K_ + N_ + extra_cols) * s, flops=2 * R * N_ * K_)
    C = 6 * We + 4 * Ht
    d["gemm_tc_ln_proj"] = gemm(C, We)
    d["gemm_tc_ln_eg_tri"] = gemm(4 * Ht, We)
    d["gemm_tc_lin_o"] = gemm(We, 2 * We, We)            # + residual read
    d["gemm_tc_ln_gelu"] = gemm(We, We, We)              # writes u and a
    d["gemm_tc_w2"] = gemm(We, We, We)
    d["gemm_tc_ln_eg"] = gemm(2 * Hn, We)
    d["gemm_tc_lin_o_e"] = gemm(We, Hn, We)
    d["gemm_tc_du"] = gemm(We, We, We)                   # + pre-activation read
    d["gemm_tc_dva"] = gemm(2 * We, We)
    d["gemm_tc_dhhat"] = gemm(Hn, We)
    d["triplet_fused_fwd"] = dict(bytes=(2 * R * We + 2 * R * We + 4 * R * Ht) * s + 4 * R,
                                  flops=R * We * 12 * We + 8 * B * N ** 3 * We)
    return d


DEVICE_KERNELS = {   # main kernels timed inside the library (tgt_kernel_timer_*) -> roofline model
    "tri_attn_fwd_tc": "triplet_attn_fwd", "tri_attn_bwd_tc": "triplet_attn_bwd",
    "tri_attn_fwd_tma": "triplet_attn_fwd", "tri_attn_fwd_mma": "triplet_attn_fwd",
    "tri_attn_bwd_tma": "triplet_attn_bwd", "tri_attn_bwd_mma": "triplet_attn_bwd",
    "tri_fused_fwd": "triplet_fused_fwd",
}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    return json.load(open(path)) if os.path.exists(path) else {}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from tgt_b200 import _C, ops
    from tgt_b200.harness.models import TGT_Multi, pretrain_loss
    from tgt_b200.harness.synthetic import add_scheme_fields
    from tgt_b200.harness.dist import env_rank_world, max_over_ranks, rank_batch, wrap_ddp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU path)")
    rank, local, world = env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _C.lib()                                                   # fail loudly if the CUDA library is missing

    cfg = model_cfg(args)
    torch.manual_seed(0)
    model = TGT_Multi(**cfg).to(dev).train()
    net = wrap_ddp(model, world, device_ids=[local], bucket_cap_mb=args.ddp_bucket_mb, bf16_grads=bool(args.ddp_bf16))
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, fused=True)
    B, N = args.batch, args.nodes
    micro = args.micro_batch or B
    assert B % micro == 0

    host = {k: v.pin_memory() for k, v in rank_batch(B, N, rank).items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    resident = {k: v.to(dev) for k, v in host.items()}

    def step(raw):
        total = None
        for mb in range(0, B, micro):
            part = {k: v[mb:mb + micro] for k, v in raw.items()} if micro != B else raw
            batch = add_scheme_fields(part, with_3d=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                gap, logits = net(batch)
                loss = pretrain_loss(gap.float(), logits.float() if args.fp32_logits else logits, batch,
                                     cfg["num_dist_bins"]) * (micro / B)
            if args.ddp_no_sync and world > 1:              # diagnostic: the same step without the gradient all-reduce
                with net.no_sync():
                    loss.backward()
            else:
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return total

    def step_e2e():
        raw = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        return float(step(raw).item())                          # D2H read of the loss (4 bytes) every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), world, dev)

    for _ in range(args.warmup):
        step(resident)
    torch.cuda.reset_peak_memory_stats()
    clocks = ClockSampler(local) if rank == 0 else None
    n0 = _C.launch_count()
    ms = timed_region(lambda: step(resident), args.steps)
    launches = _C.launch_count() - n0
    clk = clocks.stop() if clocks else None
    ms_e2e = timed_region(step_e2e, args.steps)
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    # per-kernel timing inside the real step (CUDA events on the launching stream)
    ops.KernelTimer.reset(True)
    _C.kernel_timer(True)
    barrier()
    ksteps = max(1, min(args.steps, 2))
    ops.LIBRARY_CALLS.clear()
    for _ in range(ksteps):
        step(resident)
    barrier()
    lib_calls_raw = dict(ops.LIBRARY_CALLS)
    ksum = ops.KernelTimer.summary()
    dsum = _C.kernel_timer_read()          # main kernels only, CUDA events recorded inside the library
    ops.KernelTimer.reset(False)
    _C.kernel_timer(False)

    if rank == 0:
        pk = peaks()
        km = kernel_models(micro, N, cfg["edge_width"], cfg["triplet_heads"], cfg["node_width"], cfg["num_heads"])
        step_ms = ms / args.steps
        kernels = {}
        for name, (n, tot) in ksum.items():
            per = tot / n
            ent = dict(launches_per_step=n / ksteps, ms_per_launch=per, share_of_step=(tot / ksteps) / step_ms)
            if name in km:
                gbs = km[name]["bytes"] / (per * 1e-3) / 1e9
                tfs = km[name]["flops"] / (per * 1e-3) / 1e12
                ent.update(alg_bytes=km[name]["bytes"], alg_flops=km[name]["flops"], hbm_gbs=gbs,
                           hbm_frac=gbs / pk["hbm"], tflops=tfs, tc_frac=tfs / pk["tc_sustained"])
            kernels[name] = ent
        # main kernels (helper kernels of the same API call excluded): these carry the roofline
        traffic = ncu_traffic()
        dev_kernels = {}
        for name, (n, tot) in dsum.items():
            per = tot / n
            ent = dict(launches_per_step=n / ksteps, ms_per_launch=per, share_of_step=(tot / ksteps) / step_ms)
            model = DEVICE_KERNELS.get(name)
            if model in km:
                gbs = km[model]["bytes"] / (per * 1e-3) / 1e9
                tfs = km[model]["flops"] / (per * 1e-3) / 1e12
                ent.update(alg_bytes=km[model]["bytes"], alg_flops=km[model]["flops"], hbm_gbs=gbs,
                           hbm_frac=gbs / pk["hbm"], tflops=tfs, tc_frac=tfs / pk["tc_sustained"],
                           traffic=traffic.get(name))
            dev_kernels[name] = ent
        cand = [k for k in dev_kernels if "alg_bytes" in dev_kernels[k]]
        top = max(cand, key=lambda k: dev_kernels[k]["share_of_step"]) if cand else None
        roofline = None
        if top:
            e = dev_kernels[top]
            roofline = dict(kernel=top, bound="hbm", achieved=e["hbm_gbs"], peak=pk["hbm"], unit="GB/s",
                            frac=e["hbm_frac"], traffic=e["traffic"], peak_source=pk["source"],
                            ms_per_launch=e["ms_per_launch"], share_of_step=e["share_of_step"],
                            tensor_tflops=e["tflops"], tensor_frac_of_sustained=e["tc_frac"],
                            note="HBM-bound O(N^3) core: algorithmic bytes / CUDA-event duration of the kernel alone, "
                                 "timed inside the 24-layer step")
        # the triplet MODULE (LayerNorm -> projection -> attention core -> lin_O; SURVEY 8d's fused boundary) against the
        # tensor roofline: CUDA events around the whole autograd Function, forward and backward, inside the real step.
        # F_mod = R*We*(16*We + 8*Ht) + 8*B*N^3*We forward, 2*F_mod backward (projection recompute is NOT counted).
        module = None
        if "triplet_module_fwd" in kernels and "triplet_module_bwd" in kernels:
            R_ = micro * N * N
            We_, Ht_ = cfg["edge_width"], cfg["triplet_heads"]
            f_mod = R_ * We_ * (16 * We_ + 8 * Ht_) + 8 * micro * N ** 3 * We_
            t_f, t_b = kernels["triplet_module_fwd"]["ms_per_launch"], kernels["triplet_module_bwd"]["ms_per_launch"]
            frac = lambda fl, ms_: fl / (ms_ * 1e-3) / 1e12 / pk["tc_sustained"]
            module = dict(fwd_ms=t_f, bwd_ms=t_b, alg_flops_fwd=f_mod, alg_flops_fwd_bwd=3 * f_mod,
                          module_tc_frac_fwd=frac(f_mod, t_f), module_tc_frac_fwd_bwd=frac(3 * f_mod, t_f + t_b),
                          parts_fwd={k: kernels[k]["ms_per_launch"] for k in
                                     ("gemm_tc_ln_proj", "triplet_attn_fwd", "gemm_tc_lin_o") if k in kernels},
                          parts_bwd={k: kernels[k]["ms_per_launch"] for k in kernels
                                     if k in ("gemm_tc_dva", "triplet_attn_bwd", "layernorm_bwd_W%d" % We_, "gemm_wgrad_proj",
                                              "gemm_tc_dy_proj", "lib:triplet.dy", "lib:triplet.dW",
                                              "lib:linear_residual_bwd.dW(bmm)")})
        if roofline and module:
            roofline["module_tc_frac_fwd"] = module["module_tc_frac_fwd"]
            roofline["module_tc_frac_fwd_bwd"] = module["module_tc_frac_fwd_bwd"]
            # one TGT-At layer is ~1.77 TF forward at B = 256, N = 64 => 3 x 1.77 TF x layers per step
            roofline["whole_step_tc_frac"] = (3 * 1.77e12 * cfg["model_height"] * (B / 256) / (step_ms * 1e-3) / 1e12
                                              / pk["tc_sustained"]) if N == 64 else None
        lib_calls = {k: v / ksteps for k, v in sorted(lib_calls_raw.items())}
        eager = None
        if not args.no_gpu_eager_baseline and world == 1:
            del net, opt, model, resident
            import gc
            gc.collect()
            eager = gpu_eager_baseline(args, dev)
            torch.cuda.empty_cache()
        base = None
        if not args.no_cpu_baseline:
            base, _ = cpu_baseline(args, args.cpu_batch or 4, 2, 1)       # bounded sample: 1 warm-up + 2 timed steps, ~18 s on 16 cores
        out = dict(metric=METRIC, value=world * B / (step_ms * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
                   warmup=args.warmup, ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                   dtype="bf16", data="synthetic",
                   config=dict(workload=f"TGT-At {args.layers}L (TGT_Multi, pretrain loss) training step = fwd + loss + "
                                        f"bwd + fused Adam, per-GPU batch {B} x N={N}, bf16 autocast, train mode "
                                        f"(source_dropout .3, drop_path .2, act_dropout .1)",
                               micro_batch=micro, parallelism=f"dp{world}",
                               ddp=dict(bucket_cap_mb=args.ddp_bucket_mb, bf16_grads=bool(args.ddp_bf16),
                                        no_sync_diagnostic=bool(args.ddp_no_sync)) if world > 1 else None,
                               inputs="synthetic PCQM-shaped batch (tgt_b200/harness/synthetic.py), random-init weights",
                               l2="inputs and activations per step (>10 GB) exceed the 126 MB L2; no explicit flush"),
                   e2e=dict(value=world * B / (ms_e2e / args.steps * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d_bytes,
                            d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps),
                   gpu_launches=int(launches), clocks=clk, roofline=roofline, device_kernels=dev_kernels,
                   triplet_module=module, library_calls_per_step=lib_calls, kernels=kernels, cpu_baseline=base,
                   gpu_eager_baseline=eager, peak_mem_gib=peak_mem)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def _reserve_stdout():
    """stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to fd 1) and stray prints are
    sent to stderr for the whole run, the result goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)
    return os.fdopen(saved, "w", buffering=1)


RESULT_OUT = None


def emit(obj):
    (RESULT_OUT or sys.__stdout__).write(json.dumps(obj) + "\n")
    (RESULT_OUT or sys.__stdout__).flush()


def main():
    global RESULT_OUT
    RESULT_OUT = _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (BASELINE: 256)")
    ap.add_argument("--micro-batch", type=int, default=0, help="gradient-accumulation micro-batch (0 = whole batch)")
    ap.add_argument("--nodes", type=int, default=64)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--cpu-batch", type=int, default=0)
    ap.add_argument("--cpu-warm", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--fp32-logits", action="store_true")
    ap.add_argument("--ddp-bucket-mb", type=int, default=25, help="DDP gradient bucket size (torch default 25)")
    ap.add_argument("--ddp-bf16", type=int, default=0, help="1 = all-reduce the gradients in bf16 (stock DDP comm hook)")
    ap.add_argument("--ddp-no-sync", action="store_true", help="diagnostic: skip the gradient all-reduce (INVALID as a result)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
