"""torch.autograd.Function wrappers over the C ABI (include/tgt_b200.h).

Design rules:
  * outputs are never saved for backward -- the caller (TGT_Layer) adds residuals to them;
  * nothing O(N^3) and no LayerNorm output is kept between forward and backward; the [R, 6We+4Ht]
    projection is kept only for the last k layers that HBM has room for (keep_projection), the other
    layers recompute it with the same LN-folded GEMM, so a 24-layer TGT-At step at B=256, N=64 fits
    in one B200's HBM;
  * GEMMs on the edge rows run on our tcgen05 kernels (csrc/gemm_tc.cu, csrc/gemm_wgrad.cu); what
    still goes to cuBLAS (node-side [B*N, 768] projections, fp32 parity path) is counted per call
    site in LIBRARY_CALLS and reported by bench.py;
  * no fallback: CPU tensors or a missing library raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function

from . import _C

Tensor = torch.Tensor


class KernelTimer:
    """Optional per-kernel timing with CUDA events on the launching (current) stream; used by bench.py to
    measure each of our kernels INSIDE the real step.  Disabled (zero overhead) by default."""
    enabled = False
    records = {}

    @classmethod
    def reset(cls, enabled: bool) -> None:
        cls.enabled = enabled
        cls.records = {}

    @classmethod
    def summary(cls):
        """name -> (launches, total_ms); call after torch.cuda.synchronize()."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in cls.records.items()}


class timed:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if KernelTimer.enabled:
            self.ev0 = torch.cuda.Event(enable_timing=True)
            self.ev0.record()
        return self

    def __exit__(self, *exc):
        if KernelTimer.enabled:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            KernelTimer.records.setdefault(self.name, []).append((self.ev0, ev1))
        return False


def compute_dtype(t: Tensor) -> torch.dtype:
    """dtype the kernels run in: the autocast dtype when autocast is on, else the tensor's."""
    if t.is_cuda and torch.is_autocast_enabled("cuda"):
        return torch.get_autocast_dtype("cuda")
    return t.dtype


def _require_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("tgt_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _guarded(fn):
    """Run a Function.forward / backward with the device of its first CUDA tensor argument current: every launch takes
    `torch.cuda.current_stream()` of the CURRENT device, so a model on cuda:1 in a process whose current device is cuda:0
    would otherwise launch on the wrong device.  No-op (one attribute check) when the device is already current."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args):
        dev = None
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None:
            for a in getattr(ctx, "saved_tensors", ()) if fn.__name__ == "backward" else ():
                if isinstance(a, torch.Tensor) and a.is_cuda:
                    dev = a.device
                    break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)
    return wrapper


def _f32c(t: Tensor) -> Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------
# library (cuBLAS via torch) GEMMs: every call site is named, counted and timed -- there is no silent dispatch.
# bench.py prints LIBRARY_CALLS per step next to the launch count of our own kernels.
# ------------------------------------------------------------------------------------------
import collections

LIBRARY_CALLS = collections.Counter()


def lib_mm(site: str, a: Tensor, b: Tensor) -> Tensor:
    LIBRARY_CALLS[site] += 1
    with timed("lib:" + site):
        return torch.mm(a, b)


def lib_addmm(site: str, bias: Tensor, a: Tensor, b: Tensor, out: Optional[Tensor] = None) -> Tensor:
    LIBRARY_CALLS[site] += 1
    with timed("lib:" + site):
        return torch.addmm(bias, a, b, out=out) if out is not None else torch.addmm(bias, a, b)


def lib_linear(site: str, x: Tensor, lin) -> Tensor:
    """nn.Linear module call (node-side [B*N, Wn] projections: plain library GEMMs with no fusable neighbour)."""
    LIBRARY_CALLS[site] += 1
    with timed("lib:" + site):
        return lin(x)


# ------------------------------------------------------------------------------------------
# raw kernel wrappers (no autograd)
# ------------------------------------------------------------------------------------------
AUG = 8          # augmentation columns appended to LayerNorm outputs: [1, 0, 0, 0, 0, 0, 0, 0]


def layernorm_fwd(x2: Tensor, gamma: Tensor, beta: Tensor, out_dtype: torch.dtype, eps: float = 1e-5,
                  aug: bool = False):
    """y = LN(x).  With aug=True y is [rows, W+8] = [LN(x) | 1 | 0...]: `y[:, :W]` feeds the projection GEMM
    (strided A operand) and the weight-gradient GEMM dP^T @ y_aug yields the bias gradient as its column W, so no
    separate column-sum pass over the (up to 6.25x edge-sized) projection gradient is needed."""
    rows, W = x2.shape
    ldy = W + AUG if aug else W
    y = torch.empty((rows, ldy), dtype=out_dtype, device=x2.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x2.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x2.device)
    with timed(f"layernorm_fwd_W{W}"):
        _C.check(_C.lib().tgt_layernorm_fwd(_C.ptr(x2), _C.ptr(gamma), _C.ptr(beta), _C.ptr(y), _C.ptr(mean),
                                            _C.ptr(rstd), rows, W, ldy, eps, _C.dtype_code(x2.dtype),
                                            _C.dtype_code(out_dtype), _C.stream_ptr()), "layernorm_fwd")
    return y, mean, rstd


def ln_bwd_emits_y(x2: Tensor, cd: torch.dtype) -> bool:
    """The W = 256 LayerNorm-backward kernel can also write y = LN(x) (augmented) for the weight-gradient GEMM."""
    return x2.shape[1] == 256 and cd in (torch.bfloat16, torch.float16) and x2.dtype in (cd, torch.float32)


def layernorm_bwd(dy: Tensor, x2: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor,
                  dres: Optional[Tensor] = None, beta: Optional[Tensor] = None, colsum: Optional[tuple] = None,
                  aug: bool = True):
    """(dx, dgamma, dbeta[, y][, dres_colsum]).  By-products of the same pass (see ln_bwd_emits_y):
    `beta` given   -> also the augmented y = [LN(x) | 1 | 0...] in dy's dtype, so callers order their weight-gradient
                      GEMM after this call instead of running a LayerNorm recompute pass for it;
    `colsum` = (scale [B] or None,) -> also sum_r scale[graph(r)] * dres[r, :]: dres is the upstream gradient of the module
                      this LayerNorm opens, i.e. the DropPath-weighted bias gradient of the module's output projection."""
    rows, W = x2.shape
    dx = torch.empty_like(x2)
    dgb = torch.zeros((3 if colsum is not None else 2, W), dtype=torch.float32, device=x2.device)
    if beta is not None or colsum is not None:
        ldy = W + AUG if aug else W            # aug=False: plain LN(x) (the caller has the bias gradient from elsewhere)
        y = torch.empty((rows, ldy), dtype=dy.dtype, device=x2.device) if beta is not None else None
        sc = colsum[0] if colsum is not None else None
        rps = rows // sc.numel() if sc is not None else 1
        with timed(f"layernorm_bwd_W{W}"):
            _C.check(_C.lib().tgt_layernorm_bwd_y(_C.ptr(dy), _C.ptr(x2), _C.ptr(gamma), _C.ptr(beta), _C.ptr(mean),
                                                  _C.ptr(rstd), _C.ptr(dres), _C.ptr(dx), _C.ptr(dgb[0]),
                                                  _C.ptr(dgb[1]), _C.ptr(y), ldy, rows, W,
                                                  _C.dtype_code(x2.dtype), _C.dtype_code(dy.dtype), _C.ptr(sc), rps,
                                                  _C.ptr(dgb[2]) if colsum is not None else None, _C.stream_ptr()),
                     "layernorm_bwd_y")
        out = (dx, dgb[0], dgb[1])
        if beta is not None:
            out += (y,)
        if colsum is not None:
            out += (dgb[2],)
        return out
    with timed(f"layernorm_bwd_W{W}"):
        _C.check(_C.lib().tgt_layernorm_bwd(_C.ptr(dy), _C.ptr(x2), _C.ptr(gamma), _C.ptr(mean), _C.ptr(rstd),
                                            _C.ptr(dres), _C.ptr(dx), _C.ptr(dgb[0]), _C.ptr(dgb[1]), rows, W,
                                            _C.dtype_code(x2.dtype), _C.dtype_code(dy.dtype), _C.stream_ptr()),
                 "layernorm_bwd")
    return dx, dgb[0], dgb[1]


def _workspace(nbytes: int, device):
    """Scratch for one kernel call from PyTorch's caching allocator (the library never allocates)."""
    nbytes = int(nbytes)
    if nbytes == 0:
        return None, 0
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _dres_for(dalias: Optional[Tensor], x2: Tensor) -> Optional[Tensor]:
    """Residual gradient (w.r.t. the aliased input) in the layout / dtype the LayerNorm backward kernel adds."""
    if dalias is None:
        return None
    d = dalias.reshape(x2.shape)
    if d.dtype != x2.dtype:
        d = d.to(x2.dtype)
    return d.contiguous()


def _x_for_ln(e: Tensor, cdtype: torch.dtype) -> Tensor:
    """LN kernels take x in fp32 or in the compute dtype."""
    if e.dtype != torch.float32 and e.dtype != cdtype:
        e = e.to(cdtype)
    return e.contiguous()


# ------------------------------------------------------------------------------------------
# tcgen05 / TMA projection GEMM with fused epilogues (csrc/gemm_tc.cu)
# ------------------------------------------------------------------------------------------
def row_stats(x2: Tensor, eps: float = 1e-5):
    """per-row LayerNorm statistics (mean, rstd) of a 16-bit [rows, W] matrix."""
    rows, W = x2.shape
    mean = torch.empty(rows, dtype=torch.float32, device=x2.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x2.device)
    with timed(f"row_stats_W{W}"):
        _C.check(_C.lib().tgt_row_stats(_C.ptr(x2), _C.ptr(mean), _C.ptr(rstd), rows, W, x2.stride(0), eps,
                                        _C.dtype_code(x2.dtype), _C.stream_ptr()), "row_stats")
    return mean, rstd


def gemm_tc(a: Tensor, w: Tensor, *, bias: Optional[Tensor] = None, ln: Optional[tuple] = None,
            gelu: Optional[tuple] = None, res: Optional[Tensor] = None, row_scale: Optional[Tensor] = None,
            rows_per_scale: int = 0, gelu_bwd: Optional[tuple] = None, out: Optional[Tensor] = None,
            stats_out: Optional[tuple] = None, ln_bwd: Optional[tuple] = None, name: str = "gemm_tc",
            store_gp: bool = False, mul: Optional[Tensor] = None):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T) on the tcgen05 kernel.

    ln   = (row_mean, row_rstd, col_sum): LayerNorm folded into the epilogue (`w` must already be W*gamma and
           `bias` = b + W beta);
    gelu = (p_drop, seed): exact GELU + dropout; returns (out, u) with u the 16-bit pre-activation -- or, with
           store_gp=True, (out, gp) with gp = GELU'(u) * dropout mask, which makes the backward epilogue `mul=gp`;
    mul: out = value * mul (16-bit [M, N]);
    row_scale / rows_per_scale: value *= row_scale[row // rows_per_scale] (DropPath);
    res: out = res + value;
    gelu_bwd = (u, p_drop, seed): out = value * GELU'(u) * dropout mask (backward of `gelu`);
    stats_out = (mean, rstd): fp32 [M] vectors that receive the LayerNorm statistics of the output rows (N <= 256);
    ln_bwd = (x, mean, rstd, gamma[, dres]): the GEMM result is dy of a LayerNorm over its N columns and `out` becomes
           dx = LN'(dy) (+ dres): LayerNorm backward (data gradient) as the epilogue."""
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    g = _C.GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = a.stride(0), w.stride(0), out.stride(0)
    g.dtype = _C.dtype_code(a.dtype)
    flags = 0
    if ln is not None:
        flags |= _C.EPI_LN
        g.row_mean, g.row_rstd, g.col_sum = ln[0].data_ptr(), ln[1].data_ptr(), ln[2].data_ptr()
    if bias is not None:
        flags |= _C.EPI_BIAS
        g.bias = bias.data_ptr()
    u = None
    if gelu is not None:
        flags |= _C.EPI_GELU | _C.EPI_STORE_U
        u = torch.empty((M, N), dtype=a.dtype, device=a.device)
        g.U, g.ldu = u.data_ptr(), u.stride(0)
        g.p_drop = float(gelu[0])
        if isinstance(gelu[1], torch.Tensor):        # device-resident seed: read by the kernel (CUDA-graph replays)
            g.seed, g.seed_ptr = 0, gelu[1].data_ptr()
        else:
            g.seed = int(gelu[1])
        if store_gp:
            flags |= _C.EPI_STORE_GP
    if mul is not None:
        flags |= _C.EPI_MULRES
        g.res, g.ldres, g.res_dtype = mul.data_ptr(), mul.stride(0), _C.dtype_code(mul.dtype)
    if res is not None:
        flags |= _C.EPI_RES
        g.res, g.ldres, g.res_dtype = res.data_ptr(), res.stride(0), _C.dtype_code(res.dtype)
    if gelu_bwd is not None:
        flags |= _C.EPI_GELU_BWD
        ub = gelu_bwd[0]
        g.res, g.ldres, g.res_dtype = ub.data_ptr(), ub.stride(0), _C.dtype_code(ub.dtype)
        g.p_drop, g.seed = float(gelu_bwd[1]), int(gelu_bwd[2])
    if row_scale is not None:
        if res is None:
            flags |= _C.EPI_ROWSCALE
        g.row_scale, g.rows_per_scale = row_scale.data_ptr(), int(rows_per_scale)
    if ln_bwd is not None:
        flags |= _C.EPI_LN_BWD
        xb = ln_bwd[0]
        g.res, g.ldres, g.res_dtype = xb.data_ptr(), xb.stride(0), _C.dtype_code(xb.dtype)
        g.row_mean, g.row_rstd, g.col_sum = ln_bwd[1].data_ptr(), ln_bwd[2].data_ptr(), ln_bwd[3].data_ptr()
        if len(ln_bwd) > 4 and ln_bwd[4] is not None:
            flags |= _C.EPI_RES
            g.res2, g.ldres2 = ln_bwd[4].data_ptr(), ln_bwd[4].stride(0)
    part = None
    if stats_out is not None:
        flags |= _C.EPI_STATS
        g.stat_mean, g.stat_rstd, g.stat_eps = stats_out[0].data_ptr(), stats_out[1].data_ptr(), 1e-5
        n_sl = _C.lib().tgt_gemm_tc_slices(N, K, flags | (_C.EPI_RES if res is not None else 0))
        if n_sl > 1:            # rows span several column slices: per-slice partial sums + a finalize kernel (in the library)
            part = torch.empty((n_sl, M, 2), dtype=torch.float32, device=a.device)
            g.stat_partial = part.data_ptr()
    g.flags = flags
    with timed(name):
        _C.check(_C.lib().tgt_gemm_tc(g, _C.ptr(a), _C.ptr(w), _C.ptr(out), _C.stream_ptr()), "gemm_tc")
    return (out, u) if gelu is not None else out


def take_stats(x: Tensor, x2: Tensor) -> Optional[tuple]:
    """LayerNorm statistics attached to `x` by the kernel that produced it (see linear_residual), if they describe
    exactly the rows of x2 (no dtype conversion / copy happened in between)."""
    st = getattr(x, "_tgt_ln_stats", None)
    if st is None or x2.data_ptr() != x.data_ptr() or x2.dtype != x.dtype or st[0].numel() != x2.shape[0]:
        return None
    if getattr(x, "_tgt_ln_stats_version", None) != x._version:
        return None                    # x was modified in place after the producing kernel ran: the statistics are stale
    return st


def attach_stats(out: Tensor, mean: Optional[Tensor], rstd: Optional[Tensor]) -> Tensor:
    if mean is not None:
        out._tgt_ln_stats = (mean, rstd)
        out._tgt_ln_stats_version = out._version
    return out


def tc_gemm_ok(a: Tensor, N: int, K: int) -> bool:
    """shapes / dtypes the tcgen05 GEMM accepts (everything else takes the cuBLAS + elementwise-kernel route)."""
    return (a.is_cuda and a.dtype in (torch.bfloat16, torch.float16) and a.dim() == 2 and a.stride(1) == 1
            and K % 8 == 0 and N % 8 == 0 and 8 <= K <= 512 and a.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0
            and N <= 148 * 16)


def _ln_fold(W: Tensor, b: Tensor, gamma: Tensor, beta: Tensor, cd: torch.dtype):
    """LayerNorm folded into the following Linear:  LN(x) W^T + b = rstd*(x Wg^T - mean*colsum) + b'  with
    Wg = W*gamma (rounded to the GEMM dtype), colsum = rowsum(Wg), b' = b + W beta."""
    Wf = W.detach().float()
    Wg = (Wf * gamma).to(cd).contiguous()
    return Wg, (b.detach().float() + Wf @ beta).contiguous(), Wg.float().sum(1).contiguous()


def _scale_rows(x2: Tensor, scale: Optional[Tensor]) -> Tensor:
    """x2 * scale[b] with x2 viewed [B, -1] (DropPath backward); identity when scale is None."""
    if scale is None:
        return x2
    xc = x2.contiguous()
    out = torch.empty_like(xc)
    B = scale.numel()
    _C.check(_C.lib().tgt_scaled_residual(_C.ptr(xc), None, _C.ptr(scale), _C.ptr(out), B, xc.numel() // B,
                                          _C.dtype_code(xc.dtype), _C.dtype_code(xc.dtype), _C.stream_ptr()),
             "scaled_residual(bwd)")
    return out


def linear_residual(a2: Tensor, Wc: Tensor, bias: Tensor, res2: Optional[Tensor], scale: Optional[Tensor],
                    out2: Optional[Tensor] = None, want_stats: bool = False, name: str = "gemm_tc_out"):
    """res2 + scale[b] * (a2 @ Wc^T + bias)   (res2 / scale optional; scale is per graph, rows are graph-major).
    One tcgen05 GEMM with the DropPath + residual epilogue when the shape allows, else cuBLAS + our residual kernel.
    `out2`: optional preallocated [M, N] result (a 2-D view of the tensor the caller returns).
    Returns (out2, stats): with want_stats the same GEMM also emits the LayerNorm statistics (mean, rstd) of the rows
    it writes, so the next LayerNorm-folded GEMM needs no pass of its own (stats is None when not available)."""
    M, K = a2.shape
    N = Wc.shape[0]
    if out2 is None:
        out2 = torch.empty((M, N), dtype=a2.dtype, device=a2.device)
    if tc_gemm_ok(a2, N, K) and Wc.dtype == a2.dtype and (res2 is None or (
            res2.dtype in (a2.dtype, torch.float32) and res2.stride(1) == 1 and res2.stride(0) % 8 == 0)):
        rps = M // scale.numel() if scale is not None else 0
        stats = None
        # statistics of the output rows from the epilogue (rows that span several column slices, e.g. lin_O with K = 512,
        # go through per-slice partial sums and a finalize kernel: still no pass over the edge-sized tensor)
        if want_stats and res2 is not None and N <= 256:
            stats = (torch.empty(M, dtype=torch.float32, device=a2.device),
                     torch.empty(M, dtype=torch.float32, device=a2.device))
        gemm_tc(a2, Wc, bias=bias.detach().float().contiguous(), res=res2,
                row_scale=scale if res2 is not None else None, rows_per_scale=rps, out=out2, stats_out=stats,
                name=name)
        return out2, stats
    if res2 is None:
        return lib_addmm("linear_residual", bias.detach().to(a2.dtype), a2, Wc.t(), out=out2), None
    y = lib_addmm("linear_residual", bias.detach().to(a2.dtype), a2, Wc.t())
    B = scale.numel() if scale is not None else 1
    rc = res2.contiguous()
    _C.check(_C.lib().tgt_scaled_residual(_C.ptr(y), _C.ptr(rc), _C.ptr(scale), _C.ptr(out2), B, y.numel() // B,
                                          _C.dtype_code(y.dtype), _C.dtype_code(rc.dtype), _C.stream_ptr()),
             "scaled_residual")
    return out2, None


_BMM_OUT_F32 = {}


def _bmm_f32(a: Tensor, b: Tensor) -> Tensor:
    """fp32 result of a 16-bit batched GEMM (PyTorch >= 2.8 on CUDA: `out_dtype`), else bmm + cast."""
    key = (a.device.type, a.dtype)
    if key not in _BMM_OUT_F32:
        try:
            torch.bmm(a[:1, :8, :8].contiguous(), b[:1, :8, :8].contiguous(), out_dtype=torch.float32)
            _BMM_OUT_F32[key] = True
        except (TypeError, RuntimeError):
            _BMM_OUT_F32[key] = False
    LIBRARY_CALLS["linear_residual_bwd.dW(bmm)"] += 1
    with timed("lib:linear_residual_bwd.dW(bmm)"):
        if _BMM_OUT_F32[key] and a.dtype != torch.float32:
            return torch.bmm(a, b, out_dtype=torch.float32)
        return torch.bmm(a, b).float()


def linear_residual_bwd(do2: Tensor, a2: Tensor, Wc: Tensor, scale: Optional[Tensor], gelu_bwd: Optional[tuple] = None,
                        name: str = "gemm_tc_dx", want_db: bool = True):
    """Gradients of  out = res + scale[b] * (a2 Wc^T + bias)  given do2 = d out: (da2, dW, dbias).
    The DropPath scale never costs a pass over an edge-sized tensor: da2 = scale[b] * (do2 Wc) is the row-scale
    epilogue of the GEMM, dW = sum_b scale[b] * do2_b^T a2_b is a per-graph batched GEMM followed by a weighted sum.
    gelu_bwd = (u, p_drop, seed) additionally multiplies da2 by GELU'(u) * dropout mask in the same epilogue."""
    M, N = do2.shape
    K = a2.shape[1]
    gp = None
    if gelu_bwd is not None and isinstance(gelu_bwd[0], str):      # ("gp", GELU'(u) * mask) stored by the forward epilogue
        gp, gelu_bwd = gelu_bwd[1], None
    if tc_gemm_ok(do2, K, N) and Wc.dtype == do2.dtype and (scale is not None or gelu_bwd is not None or gp is not None):
        rps = M // scale.numel() if scale is not None else 0
        da = gemm_tc(do2, Wc.t().contiguous(), row_scale=scale, rows_per_scale=rps, gelu_bwd=gelu_bwd, mul=gp, name=name)
    else:
        da = _scale_rows(lib_mm("linear_residual_bwd.dx", do2, Wc), scale)
        if gp is not None:
            da = da * gp
        if gelu_bwd is not None:
            u, p_drop, seed = gelu_bwd
            du = torch.empty_like(da)
            _C.check(_C.lib().tgt_gelu_dropout_bwd(_C.ptr(u), _C.ptr(da), _C.ptr(du), u.numel(), p_drop, seed,
                                                   _C.dtype_code(da.dtype), _C.stream_ptr()), "gelu_dropout_bwd")
            da = du
    if scale is None:
        return da, lib_mm("linear_residual_bwd.dW", do2.t(), a2), (do2.sum(0, dtype=torch.float32) if want_db else None)
    B = scale.numel()
    if B * N * K * 4 > M * (N + K):
        # few rows per graph (node-side tensors): the per-graph [B, N, K] intermediate would dwarf the operands --
        # scale the (small) gradient explicitly instead
        dos = _scale_rows(do2, scale)
        return da, lib_mm("linear_residual_bwd.dW", dos.t(), a2), dos.sum(0, dtype=torch.float32)
    dob = do2.view(B, M // B, N)
    dWb = _bmm_f32(dob.transpose(1, 2), a2.view(B, M // B, K))
    dW = torch.mv(dWb.view(B, N * K).t(), scale).view(N, K)
    # want_db=False: the caller gets the bias gradient as a by-product of its LayerNorm-backward kernel (layernorm_bwd)
    db = torch.mv(dob.sum(1, dtype=torch.float32).t(), scale) if want_db else None
    return da, dW, db


def ln_linear(x2: Tensor, g: Tensor, bt: Tensor, W: Tensor, b: Tensor, Wc: Tensor, cd: torch.dtype,
              stats: Optional[tuple] = None, name: str = "gemm_tc_ln"):
    """(LN(x2) W^T + b, mean, rstd, fold).  16-bit x2: row statistics + ONE tcgen05 GEMM on the raw rows with the
    LayerNorm folded into its epilogue (the normalised tensor never exists), fold = (Wg, b', colsum) for an identical
    recompute in backward; otherwise LN kernel + cuBLAS and fold = (None, None, None)."""
    N, K = W.shape
    if x2.dtype == cd and tc_gemm_ok(x2, N, K):
        # statistics: handed over by the GEMM that produced x2 (EPI_STATS) or one read-only pass
        mean, rstd = stats if stats is not None else row_stats(x2)
        Wg, bp, cs = _ln_fold(W, b, g, bt, cd)
        return gemm_tc(x2, Wg, bias=bp, ln=(mean, rstd, cs), name=name), mean, rstd, (Wg, bp, cs)
    y, mean, rstd = layernorm_fwd(x2, g, bt, cd)
    return lib_addmm("ln_linear", b.detach().to(cd), y, Wc.t()), mean, rstd, (None, None, None)


def _with_stats(ctx, out: Tensor, stats: Optional[tuple]):
    """Function.forward return value of the fused-residual modes: (out, mean, rstd); the statistics (or two empty
    placeholders) are non-differentiable side outputs that the module layer attaches to `out` (attach_stats)."""
    if stats is None:
        mean = rstd = torch.empty(0, dtype=torch.float32, device=out.device)
        rstd = torch.empty(0, dtype=torch.float32, device=out.device)
    else:
        mean, rstd = stats
    ctx.mark_non_differentiable(mean, rstd)
    return out, mean, rstd


def ctx_fused(ctx) -> bool:
    f = getattr(ctx, "fuse_res", None)
    if f is None:
        f = ctx.meta[5]
    return bool(f)


def unpack_fused(res):
    """(out, mean, rstd) from a fused-residual Function -> out with the statistics attached (if any)."""
    out, mean, rstd = res
    return attach_stats(out, mean if mean.numel() else None, rstd)


import os as _os

# Measured at config 3 (profiles/r1_16): the LayerNorm-backward GEMM epilogue takes 0.76 ms against 0.16 (cuBLAS dy GEMM)
# + 0.43 ms (LayerNorm-backward kernel) -- its row-per-lane reads of x are the bottleneck -- so it is opt-in for now.
_LN_BWD_FUSED = _os.environ.get("TGT_LN_BWD_FUSED", "0") == "1"


_LN_COLSUM = _os.environ.get("TGT_LN_COLSUM", "1") == "1"      # A/B switch for the bias-gradient by-product


def ln_bwd_colsum_ok(x2: Tensor, cd: torch.dtype, dout_dtype: torch.dtype) -> bool:
    """ln_linear_bwd(..., colsum=...) can deliver the weighted column sum of dres (see layernorm_bwd)."""
    return _LN_COLSUM and ln_bwd_emits_y(x2, cd) and dout_dtype == cd and not _LN_BWD_FUSED


def ln_linear_bwd(dout2: Tensor, x2: Tensor, g: Tensor, bt: Tensor, Wc: Tensor, mean: Tensor, rstd: Tensor,
                  dres: Optional[Tensor], cd: torch.dtype, name: str = "gemm_tc_ln_bwd", fused: Optional[bool] = None,
                  colsum: Optional[tuple] = None):
    """Backward of  out = LN(x2) Wc^T + b  given dout2 = d out [M, C]:  (dx, dgamma, dbeta, dW, db).

    16-bit x2 of width <= 256: ONE tcgen05 GEMM computes dy = dout2 Wc and applies the LayerNorm backward in its
    epilogue (dx = LN'(dy) + dres; dy never reaches HBM); the parameter gradients come from the weight-gradient GEMM
    against the affine-free normalised rows:  G|db = dout2^T [xhat | 1],  dW = G*gamma + db (x) beta,
    dgamma = colsum(Wc * G),  dbeta = db Wc  (tiny [C, W] operations, no pass over an edge-sized tensor).
    Otherwise: LN recompute + cuBLAS + the LayerNorm-backward kernel."""
    M, W = x2.shape
    C = Wc.shape[0]
    # (the epilogue needs whole rows in one CTA: the [W, C] weight panel must fit in shared memory next to the ring)
    if fused is None:
        fused = _LN_BWD_FUSED
    if (fused and x2.dtype == cd and dout2.dtype == cd and tc_gemm_ok(dout2, W, C) and W % 64 == 0 and W <= 256
            and ((C + 63) // 64) * W * 128 <= 159000 and (dres is None or dres.dtype == cd)):
        ones = torch.ones(W, dtype=torch.float32, device=x2.device)
        xh, _, _ = layernorm_fwd(x2, ones, torch.zeros_like(ones), cd, aug=True)          # [xhat | 1 | 0...]
        Ga = lib_mm("ln_linear_bwd.dW", dout2.t(), xh).float()                                              # [C, W+8]
        G, db = Ga[:, :W], Ga[:, W]
        del xh
        Wf = Wc.float()
        dW = G * g + db[:, None] * bt
        dg = (Wf * G).sum(0)
        dbt = db @ Wf
        dx = gemm_tc(dout2, Wc.t().contiguous(), ln_bwd=(x2, mean, rstd, g, dres), name=name)
        return dx, dg, dbt, dW, db
    if ln_bwd_emits_y(x2, cd) and dout2.dtype == cd:
        dy = lib_mm("ln_linear_bwd.dy", dout2, Wc)
        res = layernorm_bwd(dy, x2, g, mean, rstd, dres, beta=bt, colsum=colsum)
        dx, dg, dbt, y = res[:4]
        del dy
        dWa = lib_mm("ln_linear_bwd.dW", dout2.t(), y)
        return (dx, dg, dbt, dWa[:, :W], dWa[:, W]) + ((res[4],) if colsum is not None else ())
    assert colsum is None, "colsum needs the y-emitting LayerNorm backward (check ln_bwd_colsum_ok first)"
    y, _, _ = layernorm_fwd(x2, g, bt, cd, aug=True)
    dWa = lib_mm("ln_linear_bwd.dW", dout2.t(), y)
    dW, db = dWa[:, :W], dWa[:, W]
    dy = lib_mm("ln_linear_bwd.dy", dout2, Wc)
    del y
    dx, dg, dbt = layernorm_bwd(dy, x2, g, mean, rstd, dres)
    return dx, dg, dbt, dW, db


# ------------------------------------------------------------------------------------------
# LayerNorm -> Linear with LN recompute in backward  (EGT lin_EG / lin_E)
# ------------------------------------------------------------------------------------------
class LNLinearFn(Function):
    @staticmethod
    @_guarded
    def forward(ctx, x, ln_w, ln_b, W, b, cdtype):
        _require_cuda(x, W)
        with torch.autocast("cuda", enabled=False):
            shape = x.shape
            x2 = _x_for_ln(x, cdtype).view(-1, shape[-1])
            g, bt = _f32c(ln_w), _f32c(ln_b)
            Wc = W.detach().to(cdtype)
            out2, mean, rstd, _ = ln_linear(x2, g, bt, W, b, Wc, cdtype, stats=take_stats(x, x2), name="gemm_tc_ln_eg")
            out = out2.view(*shape[:-1], W.shape[0])      # not modified in place by any caller (feeds the EGT core)
            ctx.save_for_backward(x2, g, bt, Wc, mean, rstd)
            ctx.cdtype = cdtype
            ctx.in_dtype = x.dtype
            ctx.pdt = (ln_w.dtype, W.dtype, b.dtype)
            ctx.set_materialize_grads(False)
        # `out` is a fresh base tensor (never a view): callers may add residuals in place.  `x` is returned as a
        # second output (an alias): when the caller uses THAT for its residual add, the residual gradient arrives
        # here as `dalias` and is folded into the LayerNorm backward kernel (dx = LN'(dy) + dres) instead of
        # costing autograd an extra read-read-write accumulation pass over an edge-sized tensor.
        return out, x

    @staticmethod
    @_guarded
    def backward(ctx, dout, dalias=None):
        x2, g, bt, Wc, mean, rstd = ctx.saved_tensors
        cd = ctx.cdtype
        if dout is None:
            return (dalias, None, None, None, None, None)
        with torch.autocast("cuda", enabled=False):
            do = dout.reshape(-1, dout.shape[-1]).to(cd).contiguous()
            dx, dg, dbt, dW, db = ln_linear_bwd(do, x2, g, bt, Wc, mean, rstd, _dres_for(dalias, x2), cd,
                                                name="gemm_tc_ln_bwd_eg")
        return (dx.view(*dout.shape[:-1], x2.shape[-1]).to(ctx.in_dtype), dg.to(ctx.pdt[0]), dbt.to(ctx.pdt[0]),
                dW.to(ctx.pdt[1]), db.to(ctx.pdt[2]), None)


# ------------------------------------------------------------------------------------------
# triplet attention module: LN -> [QKV_in|QKV_out|EG_in|EG_out] GEMM -> core -> lin_O
# ------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------
# memory-for-time: keep the [R, 6We+4Ht] projection of the LAST k layers for backward when HBM allows
# ------------------------------------------------------------------------------------------
_LAYER_HINT = None          # (index, total) of the layer application being executed (set by TGT_Encoder)
_KEEP_STATE = {}            # per device: live bytes after layers 0 / 1, decided k


def set_layer_hint(index, total) -> None:
    global _LAYER_HINT
    _LAYER_HINT = None if index is None else (int(index), int(total))


def keep_projection(nbytes: int, device, needs_grad: bool) -> bool:
    """Should this triplet call keep its projection (nbytes) for backward instead of recomputing it (one GEMM, ~1.1 ms
    at config 3)?  Decided per step from the allocator's live-byte counter: the growth between layer 0 and layer 1 is
    extrapolated to the end of the forward; whatever stays below 88 % of the device memory after a 10 % reserve for the
    heads, the loss and the backward transients is handed to the LAST k layers (their backward runs first, so the extra
    buffers are the first to be freed).  TGT_KEEP_PROJ_LAYERS overrides k (0 disables).  Without a layer hint: never."""
    if _LAYER_HINT is None or not needs_grad:
        return False
    i, L = _LAYER_HINT
    st = _KEEP_STATE.setdefault(str(device), {})
    live = torch.cuda.memory_allocated(device)
    if i == 0:
        st["live0"], st["k"] = live, 0
        return False
    if i == 1 and "live0" in st:
        env = _os.environ.get("TGT_KEEP_PROJ_LAYERS")
        if env is not None:
            st["k"] = max(0, min(int(env), L - 2))
        else:
            # what this process can still get: free device memory plus the allocator's cached-but-unused blocks
            # (other processes on the GPU and the CUDA context are already excluded by mem_get_info)
            free, _ = torch.cuda.mem_get_info(device)
            total = free + torch.cuda.memory_reserved(device)
            predicted_end = live + (live - st["live0"]) * (L - 1)
            budget = 0.88 * total - predicted_end - 0.03 * total      # r1_21: k = 10 at config 3 (peak 151 GiB of 179)
            st["k"] = max(0, min(int(budget // (nbytes + (64 << 20))), L - 2))
        return False
    return i >= L - st.get("k", 0)


def _used_cols(wide, narrow, H: int, d: int) -> int:
    """One past the last projection column the core kernels write a gradient for: `wide` = offsets of the head-major
    q/k/v blocks (H*d columns each), `narrow` = offsets of the bias / gate blocks (H columns each, -1 = absent)."""
    hi = 0
    for offs, width in ((wide, H * d), (narrow, H)):
        for pair in offs:
            for o in tuple(pair):
                if o >= 0:
                    hi = max(hi, o + width)
    return hi


def _alloc_dproj(proj: Tensor, used: int) -> Tensor:
    """Gradient buffer of the projection.  The kernels write exactly the q/k/v/bias/gate columns; when the projection
    was padded to a multiple of 8 columns (layers/triplet.py `pad`) the pad columns must be ZERO, not uninitialised:
    they meet the zero pad rows of the weight in dy = dproj @ Wc, and NaN * 0 = NaN."""
    dproj = torch.empty_like(proj)
    if used < proj.shape[1]:
        dproj[:, used:].zero_()
    return dproj



_PANEL_IDX = {}


def _panel_rows(H: int, d: int, W: int, device) -> Tensor:
    """Row gather that regroups the kernel-order projection weights ([Q_in|K_in|V_in|Q_out|K_out|V_out], head-major)
    per PAIR of heads in the order the fused kernel's accumulator columns use:
    Qin h0,h1 | Qout h0,h1 | Kout h0,h1 | Vout h0,h1 | Kin h0,h1 | Vin h0,h1   (d rows each)."""
    key = (H, d, W, str(device))
    if key not in _PANEL_IDX:
        dd = torch.arange(d)
        rows = []
        for hp in range(H // 2):
            for base in (0, 3 * W, 4 * W, 5 * W, W, 2 * W):          # Qin, Qout, Kout, Vout, Kin, Vin
                for h in (2 * hp, 2 * hp + 1):
                    rows.append(base + h * d + dd)
        _PANEL_IDX[key] = torch.cat(rows).to(device)
    return _PANEL_IDX[key]


def _fused_triplet_fwd(x2, mean, rstd, fold, m3, va, stats, B, N, H, d, W, off_e, off_g, cdtype):
    Wg, bp, cs = fold
    idx = _panel_rows(H, d, W, x2.device)
    wf = Wg.index_select(0, idx).contiguous()
    wcs, wbs = cs.index_select(0, idx).contiguous(), bp.index_select(0, idx).contiguous()
    neg = Wg.shape[0] - 6 * W                      # E|G columns (padded to a multiple of 8); 0 for AxialAttention
    if neg > 0:
        proj_eg = gemm_tc(x2, Wg[6 * W:], bias=bp[6 * W:].contiguous(), ln=(mean, rstd, cs[6 * W:].contiguous()),
                          name="gemm_tc_ln_eg_tri")
    else:
        proj_eg = x2
    rel = lambda o: tuple(v - 6 * W if v >= 0 else -1 for v in o)
    desc = _C.TripletAttnDesc(B, N, H, d, max(neg, 8), (0, 0), (0, 0), (0, 0), rel(off_e), rel(off_g),
                              float(d) ** -0.5, _C.dtype_code(cdtype))
    ws, wsb = _workspace(_C.lib().tgt_triplet_attn_workspace_bytes(desc, 0), x2.device)
    with timed("triplet_fused_fwd"):
        _C.check(_C.lib().tgt_triplet_attn_fused_fwd(desc, _C.ptr(x2), x2.stride(0), W, _C.ptr(mean), _C.ptr(rstd),
                                                     _C.ptr(wf), _C.ptr(wcs), _C.ptr(wbs), _C.ptr(proj_eg), _C.ptr(m3),
                                                     _C.ptr(va), _C.ptr(stats), _C.ptr(ws), wsb, _C.stream_ptr()),
                 "triplet_attn_fused_fwd")
    return ws


_TRI_BWD_BIAS = _os.environ.get("TGT_TRI_BWD_BIAS", "1") == "1"   # A/B switch: projection bias gradient from the backward kernel


class TripletAttentionFn(Function):
    """e:[B,N,N,W]; Wcat:[C,W] rows in kernel order (head-major q/k/v blocks, then bias/gate blocks);
    Wo:[W,2W] with columns in kernel order (dir, h, dd).  `layout` = (H, d, off_q, off_k, off_v, off_e, off_g).

    fuse_res=False: returns (module(e), alias of e)  -- the caller adds the residual (reference layers.py:285).
    fuse_res=True : returns e + res_scale[b] * module(e) (DropPath + residual fused into the lin_O GEMM epilogue)."""

    @staticmethod
    @_guarded
    def forward(ctx, e, mask, ln_w, ln_b, Wcat, bcat, Wo, bo, layout, cdtype, res_scale=None, fuse_res=False):
        _require_cuda(e, mask, Wcat)
        H, d, off_q, off_k, off_v, off_e, off_g = layout
        B, N, _, W = e.shape
        R = B * N * N
        with timed("triplet_module_fwd"), torch.autocast("cuda", enabled=False):
            x2 = _x_for_ln(e, cdtype).view(R, W)
            g, bt = _f32c(ln_w), _f32c(ln_b)
            Wc, bc = Wcat.detach().to(cdtype).contiguous(), bcat.detach().to(cdtype).contiguous()
            Woc = Wo.detach().to(cdtype).contiguous()
            m3 = _f32c(mask).view(B, N, N)
            desc = _C.TripletAttnDesc(B, N, H, d, Wcat.shape[0], off_q, off_k, off_v, off_e, off_g,
                                      float(d) ** -0.5, _C.dtype_code(cdtype))
            va = torch.empty((R, 2 * H * d), dtype=cdtype, device=e.device)
            stats = torch.empty((B, 2, H, N, N, 2), dtype=torch.float32, device=e.device)
            kept_proj = None
            if (x2.dtype == cdtype and tc_gemm_ok(x2, Wcat.shape[0], W) and tuple(off_q) == (0, 3 * W)
                    and tuple(off_k) == (W, 4 * W) and tuple(off_v) == (2 * W, 5 * W)
                    and _C.lib().tgt_triplet_attn_fused_supported(desc, W)):
                # fused forward: the q/k/v projection is produced and consumed on chip (csrc/triplet_fused.cu)
                st_in = take_stats(e, x2)
                mean, rstd = st_in if st_in is not None else row_stats(x2)
                fold = _ln_fold(Wcat, bcat, g, bt, cdtype)
                ws = _fused_triplet_fwd(x2, mean, rstd, fold, m3, va, stats, B, N, H, d, W, off_e, off_g, cdtype)
            else:
                proj, mean, rstd, fold = ln_linear(x2, g, bt, Wcat, bcat, Wc, cdtype, stats=take_stats(e, x2),
                                                   name="gemm_tc_ln_proj")
                ws, wsb = _workspace(_C.lib().tgt_triplet_attn_workspace_bytes(desc, 0), e.device)
                with timed("triplet_attn_fwd"):
                    _C.check(_C.lib().tgt_triplet_attn_fwd(desc, _C.ptr(proj), _C.ptr(m3), _C.ptr(va), _C.ptr(stats),
                                                           _C.ptr(ws), wsb, _C.stream_ptr()), "triplet_attn_fwd")
                if fold[0] is not None and keep_projection(proj.numel() * proj.element_size(), e.device,
                                                           any(ctx.needs_input_grad)):
                    kept_proj = proj
                del proj
            # the [B,2,H,64,64] bias / gate tiles in `ws` (0.2 GB at config 3) are kept for the backward, which then
            # skips its prep pass; only the tensor-core families write them (policy fixed between forward and backward)
            tiles = ws if (ws is not None and _C.kernel_policy() != 1) else None
            ctx.tile_policy = _C.kernel_policy()
            sc = _f32c(res_scale).view(B) if (fuse_res and res_scale is not None) else None
            out = torch.empty((B, N, N, W), dtype=cdtype, device=e.device)
            _, ostats = linear_residual(va, Woc, bo, x2 if fuse_res else None, sc, out2=out.view(R, W),
                                        want_stats=fuse_res, name="gemm_tc_lin_o")
            ctx.save_for_backward(x2, m3, g, bt, Wc, bc, Woc, mean, rstd, stats, va, sc, *fold, tiles, kept_proj)
            ctx.desc = desc
            ctx.cdtype = cdtype
            ctx.in_dtype = e.dtype
            ctx.fuse_res = fuse_res
            ctx.pdt = (ln_w.dtype, Wcat.dtype, bcat.dtype, Wo.dtype, bo.dtype)
            ctx.set_materialize_grads(False)
        if fuse_res:
            return _with_stats(ctx, out, ostats)
        return out, e                   # (result, alias of the input for the caller's residual add: see LNLinearFn)

    @staticmethod
    @_guarded
    def backward(ctx, dout, dalias=None, _unused=None):
        if ctx_fused(ctx):
            dalias = None               # slots 2 and 3 are the (non-differentiable) statistics
        x2, m3, g, bt, Wc, bc, Woc, mean, rstd, stats, va, sc, Wg, bp, cs, tiles, kept_proj = ctx.saved_tensors
        if dout is None:
            return (dalias,) + (None,) * 11
        cd, desc = ctx.cdtype, ctx.desc
        B, N, W = desc.B, desc.N, x2.shape[1]
        with timed("triplet_module_bwd"), torch.autocast("cuda", enabled=False):
            do = dout.reshape(-1, W).to(cd).contiguous()
            if ctx.fuse_res:
                dalias = dout                               # residual branch: d(e) += dout, folded into the LN backward
            late_y = (kept_proj is not None or Wg is not None) and ln_bwd_emits_y(x2, cd)
            # fused residual: dout is also the residual gradient the LayerNorm backward adds, so that kernel delivers the
            # lin_O bias gradient (its DropPath-weighted column sum) as a by-product
            cs_db = late_y and ctx.fuse_res and _LN_COLSUM
            dva, dWo, dbo = linear_residual_bwd(do, va, Woc, sc, name="gemm_tc_dva", want_db=not cs_db)
            del do
            y = None if late_y else layernorm_fwd(x2, g, bt, cd, aug=True)[0]
            if kept_proj is not None:   # one of the last layers: HBM had room to keep the projection (keep_projection)
                proj = kept_proj
            elif Wg is not None:        # bit-identical recompute of the forward projection (same kernel, same inputs)
                proj = gemm_tc(x2, Wg, bias=bp, ln=(mean, rstd, cs), name="gemm_tc_ln_proj")
            else:
                proj = lib_addmm("triplet.proj", bc, y[:, :W], Wc.t())
            dproj = _alloc_dproj(proj, _used_cols((desc.off_q, desc.off_k, desc.off_v), (desc.off_e, desc.off_g),
                                                  desc.H, desc.d))
            ws, wsb = _workspace(_C.lib().tgt_triplet_attn_workspace_bytes(desc, 1), x2.device)
            if tiles is not None and ctx.tile_policy != _C.kernel_policy():
                tiles = None                                # kernel family changed since the forward: recompute
            # the pipelined backward kernel also sums the columns of dproj (= bias gradient of the projection): the
            # weight-gradient GEMM then runs on the plain 256-column LN(x) instead of the augmented 264-column one
            dbias = None
            if late_y and _TRI_BWD_BIAS and _C.lib().tgt_triplet_attn_bwd_bias_supported(desc):
                dbias = torch.zeros(proj.shape[1], dtype=torch.float32, device=x2.device)
            with timed("triplet_attn_bwd"):
                _C.check(_C.lib().tgt_triplet_attn_bwd_bias(desc, _C.ptr(proj), _C.ptr(m3), _C.ptr(va), _C.ptr(dva),
                                                            _C.ptr(stats), _C.ptr(dproj), _C.ptr(ws), wsb,
                                                            _C.ptr(tiles), _C.ptr(dbias), _C.stream_ptr()),
                         "triplet_attn_bwd")
            del ws
            del proj, dva
            if late_y:                  # LN(x) comes out of the LayerNorm-backward kernel: no recompute pass
                dy = lib_mm("triplet.dy", dproj, Wc)
                res = layernorm_bwd(dy, x2, g, mean, rstd, _dres_for(dalias, x2), beta=bt,
                                    colsum=(sc,) if cs_db else None, aug=dbias is None)
                dx, dg, dbt, y = res[:4]
                if cs_db:
                    dbo = res[4]
                del dy
                dWa = lib_mm("triplet.dW", dproj.t(), y)              # [C, W (+8)]: weight gradient (| bias gradient in column W)
                dWc, dbc = (dWa, dbias) if dbias is not None else (dWa[:, :W], dWa[:, W])
                del y, dproj
            else:
                dWa = lib_mm("triplet.dW", dproj.t(), y)
                dWc, dbc = dWa[:, :W], dWa[:, W]
                del y
                dy = lib_mm("triplet.dy", dproj, Wc)
                del dproj
                dx, dg, dbt = layernorm_bwd(dy, x2, g, mean, rstd, _dres_for(dalias, x2))
        p = ctx.pdt
        return (dx.view(B, N, N, W).to(ctx.in_dtype), None, dg.to(p[0]), dbt.to(p[0]), dWc.to(p[1]), dbc.to(p[2]),
                dWo.to(p[3]), dbo.to(p[4]), None, None, None, None)


class TripletAggregateFn(Function):
    """`layout` = (H, d, off_v, off_e, off_g, mask_dir)."""

    @staticmethod
    @_guarded
    def forward(ctx, e, mask, ln_w, ln_b, Wcat, bcat, Wo, bo, layout, cdtype, res_scale=None, fuse_res=False):
        _require_cuda(e, mask, Wcat)
        H, d, off_v, off_e, off_g, mask_dir = layout
        B, N, _, W = e.shape
        R = B * N * N
        with torch.autocast("cuda", enabled=False):
            x2 = _x_for_ln(e, cdtype).view(R, W)
            g, bt = _f32c(ln_w), _f32c(ln_b)
            Wc, bc = Wcat.detach().to(cdtype).contiguous(), bcat.detach().to(cdtype).contiguous()
            Woc = Wo.detach().to(cdtype).contiguous()
            m3 = _f32c(mask).view(B, N, N)
            proj, mean, rstd, fold = ln_linear(x2, g, bt, Wcat, bcat, Wc, cdtype, stats=take_stats(e, x2),
                                               name="gemm_tc_ln_proj")
            desc = _C.TripletAggrDesc(B, N, H, d, proj.shape[1], off_v, off_e, off_g, mask_dir,
                                      _C.dtype_code(cdtype))
            va = torch.empty((R, 2 * H * d), dtype=cdtype, device=e.device)
            aw = torch.empty((B, 2, H, N, N), dtype=torch.float32, device=e.device)
            with timed("triplet_aggr_fwd"):
                _C.check(_C.lib().tgt_triplet_aggr_fwd(desc, _C.ptr(proj), _C.ptr(m3), _C.ptr(va), _C.ptr(aw),
                                                       _C.stream_ptr()), "triplet_aggr_fwd")
            del proj
            sc = _f32c(res_scale).view(B) if (fuse_res and res_scale is not None) else None
            out = torch.empty((B, N, N, W), dtype=cdtype, device=e.device)
            _, ostats = linear_residual(va, Woc, bo, x2 if fuse_res else None, sc, out2=out.view(R, W),
                                        want_stats=fuse_res, name="gemm_tc_lin_o")
            ctx.save_for_backward(x2, m3, g, bt, Wc, bc, Woc, mean, rstd, aw, va, sc, *fold)
            ctx.fuse_res = fuse_res
            ctx.desc = desc
            ctx.cdtype = cdtype
            ctx.in_dtype = e.dtype
            ctx.pdt = (ln_w.dtype, Wcat.dtype, bcat.dtype, Wo.dtype, bo.dtype)
            ctx.set_materialize_grads(False)
        if fuse_res:
            return _with_stats(ctx, out, ostats)
        return out, e                   # (result, alias of the input for the caller's residual add: see LNLinearFn)

    @staticmethod
    @_guarded
    def backward(ctx, dout, dalias=None, _unused=None):
        if ctx_fused(ctx):
            dalias = None               # slots 2 and 3 are the (non-differentiable) statistics
        x2, m3, g, bt, Wc, bc, Woc, mean, rstd, aw, va, sc, Wg, bp, cs = ctx.saved_tensors
        if dout is None:
            return (dalias,) + (None,) * 11
        cd, desc = ctx.cdtype, ctx.desc
        B, N, W = desc.B, desc.N, x2.shape[1]
        with torch.autocast("cuda", enabled=False):
            do = dout.reshape(-1, W).to(cd).contiguous()
            if ctx.fuse_res:
                dalias = dout                               # residual branch: d(e) += dout, folded into the LN backward
            dva, dWo, dbo = linear_residual_bwd(do, va, Woc, sc, name="gemm_tc_dva")
            del do
            late_y = Wg is not None and ln_bwd_emits_y(x2, cd)
            y = None if late_y else layernorm_fwd(x2, g, bt, cd, aug=True)[0]
            if Wg is not None:          # bit-identical recompute of the forward projection (same kernel, same inputs)
                proj = gemm_tc(x2, Wg, bias=bp, ln=(mean, rstd, cs), name="gemm_tc_ln_proj")
            else:
                proj = lib_addmm("triplet.proj", bc, y[:, :W], Wc.t())
            dproj = _alloc_dproj(proj, _used_cols((desc.off_v,), (desc.off_e, desc.off_g), desc.H, desc.d))
            daw = torch.empty_like(aw)
            with timed("triplet_aggr_bwd"):
                _C.check(_C.lib().tgt_triplet_aggr_bwd(desc, _C.ptr(proj), _C.ptr(m3), _C.ptr(dva), _C.ptr(aw),
                                                       _C.ptr(daw), _C.ptr(dproj), _C.stream_ptr()),
                         "triplet_aggr_bwd")
            del proj, dva, daw
            if late_y:                  # LN(x) comes out of the LayerNorm-backward kernel: no recompute pass
                dy = lib_mm("triplet.dy", dproj, Wc)
                dx, dg, dbt, y = layernorm_bwd(dy, x2, g, mean, rstd, _dres_for(dalias, x2), beta=bt)
                del dy
                dWa = lib_mm("triplet.dW", dproj.t(), y)              # [C, W+8]: weight gradient | bias gradient (column W)
                dWc, dbc = dWa[:, :W], dWa[:, W]
                del y, dproj
            else:
                dWa = lib_mm("triplet.dW", dproj.t(), y)
                dWc, dbc = dWa[:, :W], dWa[:, W]
                del y
                dy = lib_mm("triplet.dy", dproj, Wc)
                del dproj
                dx, dg, dbt = layernorm_bwd(dy, x2, g, mean, rstd, _dres_for(dalias, x2))
        p = ctx.pdt
        return (dx.view(B, N, N, W).to(ctx.in_dtype), None, dg.to(p[0]), dbt.to(p[0]), dWc.to(p[1]), dbc.to(p[2]),
                dWo.to(p[3]), dbo.to(p[4]), None, None, None, None)


# ------------------------------------------------------------------------------------------
# TriangularUpdate core (triplet.py:154-170) and sigmoid(gate)*lin (triplet.py:129-131)
# ------------------------------------------------------------------------------------------
class TriangularCoreFn(Function):
    """proj [B,N,N,8H] = [lin_V | lin_E] outputs, mask -> Va [B,N,N,2H].  The gated O(N^2 H) operands live only in
    shared memory; backward recomputes them from `proj` (saved: it is the module's smallest tensor)."""

    @staticmethod
    @_guarded
    def forward(ctx, proj, mask, H, cdtype):
        _require_cuda(proj, mask)
        B, N = proj.shape[0], proj.shape[1]
        with torch.autocast("cuda", enabled=False):
            pc = proj.detach().to(cdtype).contiguous()
            m3 = _f32c(mask).view(B, N, N)
            desc = _C.TriangularDesc(B, N, H, pc.shape[-1], 0, 4 * H, _C.dtype_code(cdtype))
            va = torch.empty((B, N, N, 2 * H), dtype=cdtype, device=proj.device)
            with timed("triangular_fwd"):
                _C.check(_C.lib().tgt_triangular_fwd(desc, _C.ptr(pc), _C.ptr(m3), _C.ptr(va), _C.stream_ptr()),
                         "triangular_fwd")
        ctx.save_for_backward(pc, m3)
        ctx.desc = desc
        ctx.in_dtype = proj.dtype
        return va

    @staticmethod
    @_guarded
    def backward(ctx, dva):
        pc, m3 = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            dv = dva.to(pc.dtype).contiguous()
            dproj = torch.empty_like(pc)
            with timed("triangular_bwd"):
                _C.check(_C.lib().tgt_triangular_bwd(ctx.desc, _C.ptr(pc), _C.ptr(m3), _C.ptr(dv), _C.ptr(dproj),
                                                     _C.stream_ptr()), "triangular_bwd")
        return dproj.to(ctx.in_dtype), None, None, None


class SigLinFn(Function):
    """x [..., 2W] = (gates | lins) -> sigmoid(gates) * lins [..., W]; backward recomputes the sigmoid from x."""

    @staticmethod
    @_guarded
    def forward(ctx, x):
        _require_cuda(x)
        xc = x.detach().contiguous()
        W = xc.shape[-1] // 2
        rows = xc.numel() // (2 * W)
        y = torch.empty((*xc.shape[:-1], W), dtype=xc.dtype, device=xc.device)
        _C.check(_C.lib().tgt_siglin_fwd(_C.ptr(xc), _C.ptr(y), rows, W, _C.dtype_code(xc.dtype), _C.stream_ptr()),
                 "siglin_fwd")
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    @_guarded
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        W = xc.shape[-1] // 2
        dc = dy.to(xc.dtype).contiguous()
        dx = torch.empty_like(xc)
        _C.check(_C.lib().tgt_siglin_bwd(_C.ptr(xc), _C.ptr(dc), _C.ptr(dx), xc.numel() // (2 * W), W,
                                         _C.dtype_code(xc.dtype), _C.stream_ptr()), "siglin_bwd")
        return dx


# ------------------------------------------------------------------------------------------
# EGT node/edge attention core
# ------------------------------------------------------------------------------------------
class EGTCoreFn(Function):
    """qkv:[B,N,3*Wn] (or [B,N,2*Wn] when attend=False), eg:[B,N,N,2H] (or [..,H]); returns (hhat, vatt)."""

    @staticmethod
    @_guarded
    def forward(ctx, qkv, eg, mask, src_mask, H, attend, scale_degree, cdtype):
        _require_cuda(qkv, eg, mask)
        B, N = qkv.shape[0], qkv.shape[1]
        Wn = qkv.shape[2] // (3 if attend else 2)
        d = Wn // H
        with torch.autocast("cuda", enabled=False):
            q2 = qkv.detach().to(cdtype).contiguous()
            eg2 = eg.detach().to(cdtype).contiguous()
            m3 = _f32c(mask).view(B, N, N)
            src = _f32c(src_mask).view(B, N) if src_mask is not None else None
            desc = _C.EgtDesc(B, N, H, d, q2.shape[2], eg2.shape[3], float(d) ** -0.5, int(attend),
                              int(scale_degree), _C.dtype_code(cdtype))
            hhat = torch.empty((B, N, N, H), dtype=cdtype, device=qkv.device)
            vatt = torch.empty((B, N, Wn), dtype=cdtype, device=qkv.device) if attend else None
            stats = torch.empty((B, N, H, 3), dtype=torch.float32, device=qkv.device) if attend else None
            with timed("egt_attn_fwd"):
                _C.check(_C.lib().tgt_egt_attn_fwd(desc, _C.ptr(q2), _C.ptr(eg2), _C.ptr(m3), _C.ptr(src),
                                                   _C.ptr(hhat), _C.ptr(vatt), _C.ptr(stats), _C.stream_ptr()),
                         "egt_attn_fwd")
            ctx.save_for_backward(q2, eg2, m3, src, stats, vatt)
            ctx.desc = desc
            ctx.dts = (qkv.dtype, eg.dtype)
            ctx.set_materialize_grads(False)
        if attend:
            return hhat, vatt
        return hhat

    @staticmethod
    @_guarded
    def backward(ctx, dhhat, dvatt=None):
        q2, eg2, m3, src, stats, vatt = ctx.saved_tensors
        desc = ctx.desc
        cd = q2.dtype
        with torch.autocast("cuda", enabled=False):
            if desc.attend and dvatt is None:
                dvatt = torch.zeros((desc.B, desc.N, desc.H * desc.d), dtype=cd, device=q2.device)
            dh = dhhat.to(cd).contiguous() if dhhat is not None else None
            dv = dvatt.to(cd).contiguous() if dvatt is not None else None
            dqkv = torch.empty_like(q2)
            deg = torch.empty_like(eg2)
            ws, wsb = _workspace(_C.lib().tgt_egt_attn_workspace_bytes(desc), q2.device)
            with timed("egt_attn_bwd"):
                _C.check(_C.lib().tgt_egt_attn_bwd(desc, _C.ptr(q2), _C.ptr(eg2), _C.ptr(m3), _C.ptr(src),
                                                   _C.ptr(stats), _C.ptr(vatt), _C.ptr(dh), _C.ptr(dv), _C.ptr(dqkv),
                                                   _C.ptr(deg), _C.ptr(ws), wsb, _C.stream_ptr()), "egt_attn_bwd")
            del ws
        return dqkv.to(ctx.dts[0]), deg.to(ctx.dts[1]), None, None, None, None, None, None


# ------------------------------------------------------------------------------------------
# FFN: LN -> W1 -> gelu -> dropout -> W2   (saves the input, LN stats and the pre-activation only)
# ------------------------------------------------------------------------------------------
_STORE_GP = _os.environ.get("TGT_FFN_STORE_GP", "1") == "1"     # A/B switch: GELU' * mask stored by the forward epilogue


class FFNGeluFn(Function):
    """fuse_res=False: (FFN(x), alias of x); fuse_res=True: x + res_scale[b] * FFN(x) in one pass (the DropPath +
    residual add is the epilogue of the W2 GEMM; LN, bias, GELU and dropout are the epilogue of the W1 GEMM)."""

    @staticmethod
    @_guarded
    def forward(ctx, x, ln_w, ln_b, W1, b1, W2, b2, p_drop, seed, cdtype, res_scale=None, fuse_res=False):
        _require_cuda(x, W1)
        with torch.autocast("cuda", enabled=False):
            shape = x.shape
            x2 = _x_for_ln(x, cdtype).view(-1, shape[-1])
            g, bt = _f32c(ln_w), _f32c(ln_b)
            W1c, W2c = W1.detach().to(cdtype).contiguous(), W2.detach().to(cdtype).contiguous()
            inner, K = W1.shape
            if x2.dtype == cdtype and tc_gemm_ok(x2, inner, K):
                st = take_stats(x, x2)
                mean, rstd = st if st is not None else row_stats(x2)
                W1g, b1p, cs = _ln_fold(W1, b1, g, bt, cdtype)
                # the epilogue stores GELU'(u) * dropout mask in place of u: backward is then a multiply, not a re-evaluation
                a, u = gemm_tc(x2, W1g, bias=b1p, ln=(mean, rstd, cs),
                               gelu=(float(p_drop), seed if isinstance(seed, torch.Tensor) else int(seed)),
                               name="gemm_tc_ln_gelu", store_gp=_STORE_GP)
                u_is_gp = _STORE_GP
            else:
                u_is_gp = False
                y, mean, rstd = layernorm_fwd(x2, g, bt, cdtype)
                u = lib_addmm("ffn.W1", b1.detach().to(cdtype), y, W1c.t())
                del y
                a = torch.empty_like(u)
                if isinstance(seed, torch.Tensor):      # CUDA-graph capture: the kernel reads the seed at replay time
                    _C.check(_C.lib().tgt_gelu_dropout_fwd_dseed(_C.ptr(u), _C.ptr(a), u.numel(), float(p_drop),
                                                                 _C.ptr(seed), _C.dtype_code(cdtype), _C.stream_ptr()),
                             "gelu_dropout_fwd_dseed")
                else:
                    _C.check(_C.lib().tgt_gelu_dropout_fwd(_C.ptr(u), _C.ptr(a), u.numel(), float(p_drop), int(seed),
                                                           _C.dtype_code(cdtype), _C.stream_ptr()), "gelu_dropout_fwd")
            sc = None
            if fuse_res and res_scale is not None:
                sc = _f32c(res_scale).view(-1)
            out = torch.empty((*shape[:-1], W2.shape[0]), dtype=cdtype, device=x.device)
            _, ostats = linear_residual(a, W2c, b2, x2 if fuse_res else None, sc, out2=out.view(-1, W2.shape[0]),
                                        want_stats=fuse_res, name="gemm_tc_w2")
            ctx.save_for_backward(x2, g, bt, W1c, W2c, mean, rstd, u, a, sc)
            ctx.meta = (float(p_drop), 0 if isinstance(seed, torch.Tensor) else int(seed), cdtype, x.dtype,
                        (ln_w.dtype, W1.dtype, b1.dtype, W2.dtype, b2.dtype), fuse_res)
            ctx.u_is_gp = u_is_gp
            ctx.set_materialize_grads(False)
        if fuse_res:
            return _with_stats(ctx, out, ostats)
        return out, x                   # (result, alias of the input for the caller's residual add: see LNLinearFn)

    @staticmethod
    @_guarded
    def backward(ctx, dout, dalias=None, _unused=None):
        if ctx_fused(ctx):
            dalias = None               # slots 2 and 3 are the (non-differentiable) statistics
        x2, g, bt, W1c, W2c, mean, rstd, u, a, sc = ctx.saved_tensors
        p_drop, seed, cd, in_dtype, p, fuse_res = ctx.meta
        if dout is None:
            return (dalias,) + (None,) * 11
        with torch.autocast("cuda", enabled=False):
            do = dout.reshape(-1, dout.shape[-1]).to(cd).contiguous()
            if fuse_res:
                dalias = dout
            # du = scale[b] * (do W2) * GELU'(u) * dropout mask in ONE GEMM; `a` was saved by the forward epilogue
            # fused residual: W2's bias gradient is a by-product of the LayerNorm-backward kernel (its residual input is dout)
            cs_db = bool(fuse_res) and ln_bwd_colsum_ok(x2, cd, do.dtype)
            du, dW2, db2 = linear_residual_bwd(do, a, W2c, sc, name="gemm_tc_du", want_db=not cs_db,
                                               gelu_bwd=("gp", u) if ctx.u_is_gp else (u, p_drop, seed))
            res = ln_linear_bwd(du, x2, g, bt, W1c, mean, rstd, _dres_for(dalias, x2), cd, name="gemm_tc_ln_bwd_ffn",
                                colsum=(sc,) if cs_db else None)
            dx, dg, dbt, dW1, db1 = res[:5]
            if cs_db:
                db2 = res[5]
        return (dx.view(*dout.shape[:-1], x2.shape[-1]).to(in_dtype), dg.to(p[0]), dbt.to(p[0]), dW1.to(p[1]),
                db1.to(p[2]), dW2.to(p[3]), db2.to(p[4]), None, None, None, None, None)


class LinearResidualFn(Function):
    """res + scale[b] * (a W^T + bias): the edge-channel output projection of EGT_Attention / EdgeUpdate
    (lin_O_e, reference layers.py:82,126) with the caller's DropPath + residual add (layers.py:278-279) fused in."""

    @staticmethod
    @_guarded
    def forward(ctx, a, W, bias, res, scale, cdtype):
        _require_cuda(a, W, res)
        with torch.autocast("cuda", enabled=False):
            shape = res.shape
            a2 = a.detach().to(cdtype).reshape(-1, a.shape[-1]).contiguous()
            Wc = W.detach().to(cdtype).contiguous()
            r2 = res.detach().reshape(-1, shape[-1])
            if r2.dtype not in (cdtype, torch.float32):
                r2 = r2.to(cdtype)
            sc = _f32c(scale).view(-1) if scale is not None else None
            out = torch.empty(shape, dtype=cdtype, device=a.device)
            _, ostats = linear_residual(a2, Wc, bias, r2.contiguous(), sc, out2=out.view(-1, shape[-1]),
                                        want_stats=True, name="gemm_tc_lin_o_e")
            ctx.save_for_backward(a2, Wc, sc)
            ctx.dts = (a.dtype, W.dtype, bias.dtype, res.dtype, a.shape)
        return _with_stats(ctx, out, ostats)

    @staticmethod
    @_guarded
    def backward(ctx, dout, _m=None, _r=None):
        a2, Wc, sc = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            do = dout.reshape(-1, dout.shape[-1]).to(a2.dtype).contiguous()
            da, dW, db = linear_residual_bwd(do, a2, Wc, sc, name="gemm_tc_dhhat")
        d = ctx.dts
        return (da.view(d[4]).to(d[0]), dW.to(d[1]), db.to(d[2]), dout.to(d[3]), None, None)


# ------------------------------------------------------------------------------------------
# out = res + scale[b] * x      (DropPath + residual add, layers.py:163-177 / 269-290)
# ------------------------------------------------------------------------------------------
class ScaledResidualFn(Function):
    @staticmethod
    @_guarded
    def forward(ctx, x, res, scale):
        _require_cuda(x, res)
        B = x.shape[0]
        xc = x.contiguous()
        rc = res.contiguous()
        out = torch.empty_like(xc)
        sc = _f32c(scale).view(B) if scale is not None else None
        _C.check(_C.lib().tgt_scaled_residual(_C.ptr(xc), _C.ptr(rc), _C.ptr(sc), _C.ptr(out), B,
                                              xc.numel() // B, _C.dtype_code(xc.dtype), _C.dtype_code(rc.dtype),
                                              _C.stream_ptr()), "scaled_residual")
        ctx.scale = sc
        ctx.res_dtype = res.dtype
        return out

    @staticmethod
    @_guarded
    def backward(ctx, dout):
        dres = dout if dout.dtype == ctx.res_dtype else dout.to(ctx.res_dtype)
        if ctx.scale is None:
            return dout, dres, None
        dc = dout.contiguous()
        dx = torch.empty_like(dc)
        B = dc.shape[0]
        _C.check(_C.lib().tgt_scaled_residual(_C.ptr(dc), None, _C.ptr(ctx.scale), _C.ptr(dx), B, dc.numel() // B,
                                              _C.dtype_code(dc.dtype), _C.dtype_code(dc.dtype), _C.stream_ptr()),
                 "scaled_residual(bwd)")
        return dx, dres, None


def scaled_residual(x: Tensor, res: Tensor, scale: Optional[Tensor]) -> Tensor:
    return ScaledResidualFn.apply(x, res, scale)



# ------------------------------------------------------------------------------------------
# Gaussian basis of the 3-D distance embedding (models/pcqm/layers.py:25-48): one kernel each way
# ------------------------------------------------------------------------------------------
class GaussianBasisFn(Function):
    """x:[...] f32 (scaled distances), mu, sd:[K] f32  ->  [..., K] in `out_dtype`; backward recomputes the basis."""

    @staticmethod
    @_guarded
    def forward(ctx, x, mu, sd, out_dtype):
        _require_cuda(x, mu, sd)
        xc, muc, sdc = _f32c(x), _f32c(mu), _f32c(sd)
        K = muc.numel()
        out = torch.empty((*x.shape, K), dtype=out_dtype, device=x.device)
        _C.check(_C.lib().tgt_gaussian_basis_fwd(_C.ptr(xc), _C.ptr(muc), _C.ptr(sdc), _C.ptr(out), xc.numel(), K,
                                                 _C.dtype_code(out_dtype), _C.stream_ptr()), "gaussian_basis_fwd")
        ctx.save_for_backward(xc, muc, sdc)
        ctx.dts = (x.dtype, mu.dtype, sd.dtype)
        return out

    @staticmethod
    @_guarded
    def backward(ctx, dout):
        xc, muc, sdc = ctx.saved_tensors
        K = muc.numel()
        do = dout.contiguous()
        dx = torch.empty_like(xc)
        dms = torch.zeros((2, K), dtype=torch.float32, device=xc.device)
        _C.check(_C.lib().tgt_gaussian_basis_bwd(_C.ptr(xc), _C.ptr(muc), _C.ptr(sdc), _C.ptr(do), _C.ptr(dx),
                                                 _C.ptr(dms[0]), _C.ptr(dms[1]), xc.numel(), K, _C.dtype_code(do.dtype),
                                                 _C.stream_ptr()), "gaussian_basis_bwd")
        d = ctx.dts
        return dx.to(d[0]), dms[0].to(d[1]), dms[1].to(d[2]), None


# ------------------------------------------------------------------------------------------
# cross-entropy over distance bins (training_schemes/pcqm/commons.py:26-48): one read of the logits forward,
# d logits straight from the logits backward
# ------------------------------------------------------------------------------------------
class XentRowsFn(Function):
    """logits [R, nbins] (16-bit or fp32), target [R] int64  ->  xent [R] fp32 = F.cross_entropy(..., reduction='none')."""

    @staticmethod
    @_guarded
    def forward(ctx, logits, target):
        _require_cuda(logits, target)
        R, nb = logits.shape
        tg = target.contiguous()
        xent = torch.empty(R, dtype=torch.float32, device=logits.device)
        lse = torch.empty_like(xent)
        with timed("xent_rows_fwd"):
            _C.check(_C.lib().tgt_xent_rows_fwd(_C.ptr(logits), logits.stride(0), _C.ptr(tg), _C.ptr(xent), _C.ptr(lse), R, nb,
                                                _C.dtype_code(logits.dtype), _C.stream_ptr()), "xent_rows_fwd")
        ctx.save_for_backward(logits, tg, lse)
        return xent

    @staticmethod
    @_guarded
    def backward(ctx, gx):
        logits, tg, lse = ctx.saved_tensors
        R, nb = logits.shape
        dl = torch.empty((R, nb), dtype=logits.dtype, device=logits.device)
        with timed("xent_rows_bwd"):
            _C.check(_C.lib().tgt_xent_rows_bwd(_C.ptr(logits), logits.stride(0), _C.ptr(tg), _C.ptr(lse), _C.ptr(_f32c(gx)),
                                                _C.ptr(dl), R, nb, _C.dtype_code(logits.dtype), _C.stream_ptr()), "xent_rows_bwd")
        return dl, None


def xent_rows_ok(logits: Tensor, target: Tensor) -> bool:
    return (logits.is_cuda and logits.dim() == 2 and target.dim() == 1 and target.dtype == torch.int64
            and logits.dtype in (torch.float32, torch.bfloat16, torch.float16) and logits.shape[1] % 8 == 0
            and logits.shape[1] <= 1024 and logits.stride(1) == 1 and logits.stride(0) % 8 == 0 and logits.data_ptr() % 16 == 0)


def cross_entropy_rows(logits: Tensor, target: Tensor) -> Tensor:
    """Drop-in for `F.cross_entropy(logits, target, reduction='none')` on the [R, num_bins] distance-bin logits
    (DiscreteDistLoss, commons.py:38): fp32 row losses from logits of any supported dtype, no fp32 copy of the logits,
    no saved log-probabilities.  Shapes / dtypes the kernels do not take (and CPU tensors) go to torch."""
    if xent_rows_ok(logits, target):
        return XentRowsFn.apply(logits, target)
    return torch.nn.functional.cross_entropy(logits.float(), target, reduction='none')


# ------------------------------------------------------------------------------------------
# two-stage inference: logits -> symmetrised argmax bins -> distances, one kernel
# (dist_pred/scheme.py:186-194 + commons.py:72-82)
# ------------------------------------------------------------------------------------------
def bins_decode(logits: Tensor, range_bins: float = 8.0, shift_half: bool = True, zero_diag: bool = True,
                want_bins: bool = True):
    """logits [B,N,N,num_bins] (16-bit or fp32) -> (bins int16 [B,N,N] or None, dist_input fp32 [B,N,N])."""
    _require_cuda(logits)
    B, N, N2, nb = logits.shape
    assert N == N2
    lg = logits.detach().contiguous()
    bins = torch.empty((B, N, N), dtype=torch.int16, device=lg.device) if want_bins else None
    dist = torch.empty((B, N, N), dtype=torch.float32, device=lg.device)
    with timed("bins_decode"):
        _C.check(_C.lib().tgt_bins_decode(_C.ptr(lg), _C.ptr(bins), _C.ptr(dist), B, N, nb, range_bins / (nb - 1),
                                          int(shift_half), int(zero_diag), _C.dtype_code(lg.dtype), _C.stream_ptr()),
                 "bins_decode")
    return bins, dist
