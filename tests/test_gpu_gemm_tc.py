"""tcgen05 / TMA GEMM (tgt_gemm_tc) and row statistics against plain PyTorch fp32 references of the same ops.
Tolerances: the kernel accumulates bf16 products in fp32 and rounds once to bf16, so outputs must agree with the
fp32 reference rounded to bf16 within 2 bf16 ulps of the row scale (rel 1e-2 element-wise on O(1) data)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from tgt_b200 import ops
    return ops


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def _close(got, ref, tol):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item() + 1e-6
    assert err <= tol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (4096, 1600, 256), (1000, 128, 256), (5000, 256, 64),
                                   (777, 256, 512), (300, 64, 256), (130, 16, 16), (257, 1000, 136),
                                   (70000, 256, 256)])
def test_gemm_plain_and_bias(M, N, K, dtype):
    ops = _ops()
    a, w = _rand((M, K), dtype, 1), _rand((N, K), dtype, 2, K ** -0.5)
    bias = _rand((N,), torch.float32, 3)
    ref = a.float() @ w.float().t()
    _close(ops.gemm_tc(a, w), ref, 6e-3)
    _close(ops.gemm_tc(a, w, bias=bias), ref + bias, 6e-3)


def test_gemm_strided_operands():
    ops = _ops()
    big = _rand((3000, 304), torch.bfloat16, 4)
    a = big[:, 8:264]                                   # pitch 304 elements = 608 B, 16-byte aligned rows
    w = _rand((256, 256), torch.bfloat16, 5, 1 / 16)
    out = torch.zeros((3000, 512), dtype=torch.bfloat16, device="cuda")
    ops.gemm_tc(a, w, out=out[:, 256:])
    _close(out[:, 256:], a.float() @ w.float().t(), 6e-3)
    assert out[:, :256].abs().max().item() == 0


@pytest.mark.parametrize("M,N", [(1000, 72), (333, 200), (4097, 1608)])
def test_gemm_bulk_store_epilogue_clips_to_the_output_window(M, N):
    """The TMEM-domain epilogues leave through TMA bulk stores of 32 x 32 tiles (and vector stores for a slice's ragged
    last column group): rows >= M and the columns on either side of the [M, N] window must stay untouched."""
    ops = _ops()
    K = 256
    a, w = _rand((M, K), torch.bfloat16, 21), _rand((N, K), torch.bfloat16, 22, K ** -0.5)
    bias = _rand((N,), torch.float32, 23)
    big = torch.full((M + 40, N + 24), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm_tc(a, w, bias=bias, out=big[:M, 8:8 + N])
    _close(big[:M, 8:8 + N], a.float() @ w.float().t() + bias, 6e-3)
    assert (big[:, :8] == 7).all() and (big[:, 8 + N:] == 7).all() and (big[M:] == 7).all()


@pytest.mark.parametrize("W,N", [(256, 1600), (256, 128), (64, 256)])
def test_gemm_layernorm_fold(W, N):
    ops = _ops()
    M = 5000
    x = (_rand((M, W), torch.float32, 6) * 1.7 + 0.4).to(torch.bfloat16)
    gamma, beta = 1 + 0.1 * _rand((W,), torch.float32, 7), 0.1 * _rand((W,), torch.float32, 8)
    Wt, b = _rand((N, W), torch.float32, 9, W ** -0.5), _rand((N,), torch.float32, 10)
    ref = torch.nn.functional.layer_norm(x.float(), (W,), gamma, beta) @ Wt.t() + b
    mean, rstd = ops.row_stats(x)
    xf = x.float()
    assert torch.allclose(mean, xf.mean(1), atol=1e-5)
    assert torch.allclose(rstd, (xf.var(1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)
    wg = (Wt * gamma).to(torch.bfloat16)
    out = ops.gemm_tc(x, wg, bias=b + Wt @ beta, ln=(mean, rstd, wg.float().sum(1)))
    _close(out, ref, 1e-2)


@pytest.mark.parametrize("res_dtype", [torch.bfloat16, torch.float32])
def test_gemm_residual_droppath(res_dtype):
    ops = _ops()
    Bb, rows_per = 5, 900
    M, N, K = Bb * rows_per, 256, 512
    a, w = _rand((M, K), torch.bfloat16, 11), _rand((N, K), torch.bfloat16, 12, K ** -0.5)
    bias = _rand((N,), torch.float32, 13)
    res = _rand((M, N), res_dtype, 14)
    scale = torch.tensor([1.25, 0.0, 1.25, 1.25, 0.0], device="cuda")
    lin = (a.float() @ w.float().t() + bias).to(torch.bfloat16).float()
    ref = res.float() + scale.repeat_interleave(rows_per)[:, None] * lin
    _close(ops.gemm_tc(a, w, bias=bias, res=res, row_scale=scale, rows_per_scale=rows_per), ref, 8e-3)
    _close(ops.gemm_tc(a, w, bias=bias, res=res), res.float() + lin, 8e-3)


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_gemm_gelu_dropout_matches_elementwise_kernel(p_drop):
    """the fused epilogue must reproduce tgt_gelu_dropout_fwd (same hash, same element index) on its own U."""
    from tgt_b200 import _C
    ops = _ops()
    M, N, K = 3000, 256, 256
    a, w = _rand((M, K), torch.bfloat16, 15), _rand((N, K), torch.bfloat16, 16, K ** -0.5)
    bias = _rand((N,), torch.float32, 17)
    out, u = ops.gemm_tc(a, w, bias=bias, gelu=(p_drop, 1234567))
    _close(u, a.float() @ w.float().t() + bias, 6e-3)
    want = torch.empty_like(u)
    _C.check(_C.lib().tgt_gelu_dropout_fwd(_C.ptr(u), _C.ptr(want), u.numel(), p_drop, 1234567,
                                           _C.dtype_code(u.dtype), _C.stream_ptr()), "gelu_dropout_fwd")
    assert (out.float() - want.float()).abs().max().item() <= 1e-2 * want.float().abs().max().item()
    assert ((out == 0) == (want == 0)).float().mean().item() > 0.9999


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_gemm_gelu_backward_epilogue(p_drop):
    """value * GELU'(u) * mask with a DropPath row scale == tgt_gelu_dropout_bwd on the scaled GEMM result."""
    from tgt_b200 import _C
    ops = _ops()
    Bb, rows_per = 6, 500
    M, N, K = Bb * rows_per, 256, 256
    do, w = _rand((M, K), torch.bfloat16, 21), _rand((N, K), torch.bfloat16, 22, K ** -0.5)
    u = _rand((M, N), torch.bfloat16, 23)
    scale = torch.tensor([1.25, 0.0, 1.25, 1.25, 0.0, 1.25], device="cuda")
    got = ops.gemm_tc(do, w, row_scale=scale, rows_per_scale=rows_per, gelu_bwd=(u, p_drop, 777))
    da = ((do.float() @ w.float().t()) * scale.repeat_interleave(rows_per)[:, None]).to(torch.bfloat16)
    want = torch.empty_like(da)
    _C.check(_C.lib().tgt_gelu_dropout_bwd(_C.ptr(u), _C.ptr(da), _C.ptr(want), u.numel(), p_drop, 777,
                                           _C.dtype_code(u.dtype), _C.stream_ptr()), "gelu_dropout_bwd")
    _close(got, want, 1.2e-2)
    assert ((got == 0) == (want == 0)).float().mean().item() > 0.999
    plain = ops.gemm_tc(do, w, row_scale=scale, rows_per_scale=rows_per)
    _close(plain, da, 8e-3)


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_gemm_gelu_stored_derivative_and_multiply_epilogue(p_drop):
    """STORE_GP: same activation as the plain GELU epilogue, U holds GELU'(u) * mask; MULRES (+ROWSCALE) multiplies by it:
    together they equal the GELU_BWD epilogue on the stored pre-activation."""
    ops = _ops()
    Bb, rows_per = 6, 500
    M, N, K = Bb * rows_per, 256, 256
    a, w = _rand((M, K), torch.bfloat16, 15), _rand((N, K), torch.bfloat16, 16, K ** -0.5)
    bias = _rand((N,), torch.float32, 17)
    out0, u = ops.gemm_tc(a, w, bias=bias, gelu=(p_drop, 4242))
    out1, gp = ops.gemm_tc(a, w, bias=bias, gelu=(p_drop, 4242), store_gp=True)
    assert torch.equal(out0, out1)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf).sum().backward()
    keep = (out0 != 0) | (u == 0) if p_drop > 0 else torch.ones_like(u, dtype=torch.bool)
    scale_keep = 1.0 / (1.0 - p_drop) if p_drop > 0 else 1.0
    _close(gp.float()[keep], (uf.grad * scale_keep)[keep], 6e-3)
    if p_drop > 0:
        assert float((gp[~keep] != 0).float().sum()) == 0.0
    do, w2 = _rand((M, N), torch.bfloat16, 21), _rand((K, N), torch.bfloat16, 22, N ** -0.5)
    scale = torch.tensor([1.25, 0.0, 1.25, 1.25, 0.0, 1.25], device="cuda")
    got = ops.gemm_tc(do, w2, row_scale=scale, rows_per_scale=rows_per, mul=gp)
    want = ops.gemm_tc(do, w2, row_scale=scale, rows_per_scale=rows_per, gelu_bwd=(u, p_drop, 4242))
    _close(got, want, 1.2e-2)
    assert ((got == 0) == (want == 0)).float().mean().item() > 0.999
    got2 = ops.gemm_tc(do, w2, mul=gp)                 # without the row scale
    _close(got2, (do.float() @ w2.float().t()) * gp.float(), 8e-3)


def test_linear_residual_bwd_matches_autograd():
    """DropPath-scaled linear backward (row-scale epilogue + per-graph batched weight gradient) vs autograd."""
    ops = _ops()
    Bb, rows_per, N, K = 4, 256, 64, 128
    M = Bb * rows_per
    a = _rand((M, K), torch.bfloat16, 31).float().requires_grad_(True)
    W = _rand((N, K), torch.bfloat16, 32, K ** -0.5).float().requires_grad_(True)
    b = torch.zeros(N, device="cuda", requires_grad=True)
    scale = torch.tensor([1.25, 0.0, 1.25, 1.25], device="cuda")
    do = _rand((M, N), torch.bfloat16, 33)
    out = (a @ W.t() + b) * scale.repeat_interleave(rows_per)[:, None]
    out.backward(do.float())
    da, dW, db = ops.linear_residual_bwd(do, a.detach().bfloat16(), W.detach().bfloat16(), scale)
    _close(da, a.grad, 8e-3)
    _close(dW, W.grad, 8e-3)
    _close(db, b.grad, 8e-3)


@pytest.mark.parametrize("N,K", [(256, 256), (256, 64), (64, 128), (256, 512), (192, 448)])   # K > 256: two column slices
def test_gemm_output_row_stats(N, K):
    """EPI_STATS: mean / rstd of the rows the residual GEMM writes == LayerNorm statistics of its (rounded) output."""
    ops = _ops()
    M = 5003
    a, w = _rand((M, K), torch.bfloat16, 41), _rand((N, K), torch.bfloat16, 42, K ** -0.5)
    bias, res = _rand((N,), torch.float32, 43), (_rand((M, N), torch.float32, 44) * 2 + 0.7).to(torch.bfloat16)
    mean = torch.empty(M, device="cuda")
    rstd = torch.empty(M, device="cuda")
    out = ops.gemm_tc(a, w, bias=bias, res=res, stats_out=(mean, rstd))
    of = out.float()
    assert torch.allclose(mean, of.mean(1), atol=2e-5, rtol=1e-5)
    assert torch.allclose(rstd, (of.var(1, unbiased=False) + 1e-5).rsqrt(), rtol=2e-4)
    m2, r2 = ops.row_stats(out)
    assert torch.allclose(mean, m2, atol=2e-5) and torch.allclose(rstd, r2, rtol=2e-4)


def test_stats_handoff_between_modules():
    """the statistics attached by one fused module are consumed by the next one and change nothing numerically."""
    from tgt_b200 import layers as L
    from tgt_b200.harness.synthetic import make_edge_inputs
    torch.manual_seed(0)
    ffn1, ffn2 = L.FFN(64).cuda().eval(), L.FFN(64).cuda().eval()
    e, _ = make_edge_inputs(3, 12, 64, [12, 7, 5], seed=3)
    e = e.cuda().bfloat16()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y1 = ffn1.forward_residual(e, None)
        assert getattr(y1, "_tgt_ln_stats", None) is not None
        y2 = ffn2.forward_residual(y1, None)
        y1b = y1.clone()                                  # same values, no attached statistics
        y2b = ffn2.forward_residual(y1b, None)
    assert (y2.float() - y2b.float()).abs().max().item() <= 2e-2 * y2b.float().abs().max().item()


@pytest.mark.parametrize("W,C,with_res", [(256, 256, True), (256, 128, False), (64, 256, True), (128, 64, True)])
def test_gemm_layernorm_backward_epilogue(W, C, with_res):
    """dx of LayerNorm->Linear from ONE GEMM (LN backward in the epilogue) and the parameter gradients from the
    affine-free weight-gradient GEMM, against autograd on the same bf16-rounded operands."""
    ops = _ops()
    M = 4099
    x = (_rand((M, W), torch.float32, 51) * 1.3 + 0.2).to(torch.bfloat16)
    gamma = (1 + 0.2 * _rand((W,), torch.float32, 52)).requires_grad_(True)
    beta = (0.1 * _rand((W,), torch.float32, 53)).requires_grad_(True)
    Wt = _rand((C, W), torch.bfloat16, 54, W ** -0.5).float().requires_grad_(True)
    b = torch.zeros(C, device="cuda", requires_grad=True)
    dout = _rand((M, C), torch.bfloat16, 55)
    dres = _rand((M, W), torch.bfloat16, 56) if with_res else None
    xf = x.float().requires_grad_(True)
    out = torch.nn.functional.layer_norm(xf, (W,), gamma, beta) @ Wt.t() + b
    out.backward(dout.float())
    want_dx = xf.grad + (dres.float() if with_res else 0)
    mean, rstd = ops.row_stats(x)
    dx, dg, dbt, dW, db = ops.ln_linear_bwd(dout, x, gamma.detach(), beta.detach(), Wt.detach().bfloat16(), mean, rstd,
                                            dres, torch.bfloat16, fused=True)
    _close(dx, want_dx, 1.5e-2)
    _close(dg, gamma.grad, 1e-2)
    _close(dbt, beta.grad, 1e-2)
    _close(dW, Wt.grad, 1e-2)
    _close(db, b.grad, 1e-2)
