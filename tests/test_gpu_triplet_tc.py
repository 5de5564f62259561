"""The tcgen05 / TMEM triplet-attention core (csrc/triplet_tc.cu, kernel policy 4) through the C ABI.

  * second half of the bf16 parity protocol: the core alone, on 16-bit-rounded inputs, with fp32
    outputs (tgt_triplet_attn_fwd_f32out), relative L2 <= 1e-3 against the fp64 oracle evaluated on the SAME rounded
    inputs (oracle.triplet_attention_core, pinned to the reference in tests/test_oracle_pin.py) -- ragged N, padded
    graphs, fully-masked query rows, ungated / bias-free variants;
  * the 16-bit-output kernel inside the module (policy 4) against the mma.sync family (policy 5) and the fp64 oracle,
    with gradients through the backward that consumes the statistics the tcgen05 forward wrote;
  * the core backward on tcgen05 against autograd of the same oracle.
"""
import pytest
import torch

from conftest import rel_err
from oracle import tgt_oracle as O
from tgt_b200 import _C, ops
from tgt_b200 import layers as L
from tgt_b200.harness.synthetic import make_edge_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"
W, H, D_ = 256, 16, 16


def _layout(gated=True, biased=True):
    off_q, off_k, off_v = (0, 3 * W), (W, 4 * W), (2 * W, 5 * W)
    if not biased:
        return off_q, off_k, off_v, (-1, -1), (-1, -1), 6 * W
    if gated:
        return off_q, off_k, off_v, (6 * W, 6 * W + 2 * H), (6 * W + H, 6 * W + 3 * H), 6 * W + 4 * H
    return off_q, off_k, off_v, (6 * W, 6 * W + H), (-1, -1), 6 * W + 2 * H


def _proj_and_mask(B, N, nn_, C, dtype, seed, std=0.6):
    """std 0.6 is the scale of the reference's own projections at initialisation (LayerNorm output through a default-init
    nn.Linear(256, .): variance 256 / (3 * 256) = 1/3); std 1.3 makes the softmax peaky (a few keys carry the row)."""
    g = torch.Generator().manual_seed(seed)
    proj = (torch.randn(B, N, N, C, generator=g) * std).to(dtype)
    _, mask = make_edge_inputs(B, N, 8, nn_, seed=seed)
    return proj, mask[..., 0].contiguous()


def _desc(B, N, C, lay, dtype):
    off_q, off_k, off_v, off_e, off_g, _ = lay
    return _C.TripletAttnDesc(B, N, H, D_, C, off_q, off_k, off_v, off_e, off_g, float(D_) ** -0.5, _C.dtype_code(dtype))


@pytest.mark.parametrize("dtype,std", [(torch.bfloat16, 0.6), (torch.float16, 0.6), (torch.bfloat16, 1.3), (torch.float16, 1.3)])
@pytest.mark.parametrize("N,nn_", [(64, [64, 37, 50]), (48, [48, 25, 1]), (33, [33, 20, 9]), (1, [1, 1, 1])])
@pytest.mark.parametrize("variant", ["gated", "ungated", "axial"])
def test_core_fp32_out_on_rounded_inputs_within_1e3(dtype, std, N, nn_, variant):
    """fp32 outputs on 16-bit-rounded inputs vs the fp64 oracle on the same inputs.  The only rounding
    inside the core is that of the attention weights A = P * gate to the tensor-core operand type before A.V -- the
    reference's autocast path does the same (triplet.py:227).  For fp16 (2^-12) the north-star 1e-3 holds with a wide
    margin at any input scale.  For bf16 that single rounding is 2^-9 / sqrt(3) = 1.13e-3 relative per weight and, V
    having no preferred sign, it does not average out over the keys: NO bf16-operand implementation can be below
    ~1.1e-3 here.  So the binding criterion is the ATTAINABLE one -- the kernel's error may exceed that of the fp64 oracle
    with exactly that one rounding applied by at most 15 % (the gate also passes through an fp16 tile) -- next to the
    absolute ceilings 1e-3 (fp16) and 2.5e-3 (bf16)."""
    tol = 1e-3 if dtype == torch.float16 else 2.5e-3
    lay = _layout(gated=variant == "gated", biased=variant != "axial")
    C = lay[5]
    B = len(nn_)
    proj, mask = _proj_and_mask(B, N, nn_, C, dtype, seed=3 + N, std=std)
    ref = O.triplet_attention_core(proj.double(), mask.double(), H, D_, *lay[:5])
    desc = _desc(B, N, C, lay, dtype)
    pd, md = proj.to(DEV).view(B * N * N, C), mask.to(DEV)
    va = torch.full((B * N * N, 2 * W), float("nan"), dtype=torch.float32, device=DEV)
    stats = torch.empty((B, 2, H, N, N, 2), dtype=torch.float32, device=DEV)
    nb = _C.lib().tgt_triplet_attn_workspace_bytes(desc, 0)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    _C.check(_C.lib().tgt_triplet_attn_fwd_f32out(desc, _C.ptr(pd), _C.ptr(md), _C.ptr(va), _C.ptr(stats), _C.ptr(ws), nb,
                                                 _C.stream_ptr()), "triplet_attn_fwd_f32out")
    torch.cuda.synchronize()
    got = va.view(B, N, N, 2 * W).double().cpu()
    assert bool(torch.isfinite(got).all())
    err = rel_err(got, ref)
    eh = [rel_err(got.view(B, N, N, 2, H, D_)[..., hh::2, :], ref.view(B, N, N, 2, H, D_)[..., hh::2, :]) for hh in (0, 1)]
    print(f"tcgen05 core fp32-out {variant} {dtype} std={std} N={N}: rel-L2 {err:.3e} (even heads {eh[0]:.3e}, odd heads "
          f"{eh[1]:.3e}), max abs {float((got - ref).abs().max()):.3e}")
    attainable = rel_err(O.triplet_attention_core(proj.double(), mask.double(), H, D_, *lay[:5], round_a=dtype), ref)
    print(f"    attainable with A rounded to {dtype}: {attainable:.3e}")
    assert err <= tol, err
    assert err <= 1.15 * attainable + 5e-5, (err, attainable)
    # log-sum-exp statistics (log2 domain) against the oracle's logits, real rows only
    off_q, off_k, off_v, off_e, off_g, _ = lay
    P = proj.double()
    Q = P[..., off_q[0]:off_q[0] + W].reshape(B, N, N, H, D_)
    K = P[..., off_k[0]:off_k[0] + W].reshape(B, N, N, H, D_)
    S = D_ ** -0.5 * torch.einsum("bijhd,bjkhd->bijkh", Q, K) + mask.double()[:, :, None, :, None]
    if off_e[0] >= 0:
        S = S + P[..., off_e[0]:off_e[0] + H].unsqueeze(2)
    lse2 = torch.logsumexp(S, dim=3) / torch.log(torch.tensor(2.0, dtype=torch.float64))       # [B,i,j,h]
    st = stats.view(-1)[: B * 2 * H * N * N].view(B, 2, H, N, N).double().cpu()                # [b,dir,h,j,i]
    real = (mask[:, :, 0] == 0)                                                                 # atom i is real
    for b in range(B):
        idx = real[b].nonzero().flatten()
        a = st[b, 0][:, :, idx].permute(2, 1, 0)                                                # [i,j,h]
        assert torch.allclose(a, lse2[b, idx], atol=2e-3, rtol=1e-4)


@pytest.mark.parametrize("kind", ["attention", "attention_ungated", "axial_attention"])
@pytest.mark.parametrize("N,nn_", [(64, [64, 37]), (40, [40, 17]), (5, [5, 2])])
def test_module_policy4_vs_oracle_and_mma_family(kind, N, nn_):
    torch.manual_seed(3)
    mod = L.get_triplet_layer(kind)(W, H)
    e, mask = make_edge_inputs(2, N, W, nn_, seed=5)
    e = e.bfloat16()
    p = {k: v.double().requires_grad_(True) for k, v in mod.state_dict().items()}
    ed = e.double().requires_grad_(True)
    ref = O.TRIPLET_FNS[kind](p, ed, mask.double(), H)
    dout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    ref.backward(dout)
    mod = mod.to(DEV)
    res = {}
    for policy in (4, 5):
        _C.set_kernel_policy(policy)
        _C.kernel_timer(True)
        try:
            mod.zero_grad(set_to_none=True)
            eg = e.to(DEV).requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = mod(eg, mask.to(DEV))
            out.backward(dout.to(DEV).to(out.dtype))
            torch.cuda.synchronize()
            names = set(_C.kernel_timer_read())
            res[policy] = (out.float().cpu(), eg.grad.float().cpu(), {k: v.grad.cpu() for k, v in mod.named_parameters()})
        finally:
            _C.set_kernel_policy(0)
            _C.kernel_timer(False)
        assert ("tri_attn_fwd_tc" in names) == (policy == 4), names
    for policy in (4, 5):
        assert rel_err(res[policy][0], ref) < 1e-2 and rel_err(res[policy][1], ed.grad) < 2e-2
    # same arithmetic up to the summation order of the fp32 row sums and one 16-bit rounding of P
    assert rel_err(res[4][0], res[5][0]) < 2e-3, rel_err(res[4][0], res[5][0])
    assert rel_err(res[4][1], res[5][1]) < 5e-3, rel_err(res[4][1], res[5][1])
    e4 = rel_err(res[4][0], ref)
    e5 = rel_err(res[5][0], ref)
    print(f"policy 4 (tcgen05) out err {e4:.3e}, policy 5 (mma.sync) {e5:.3e}")
    assert e4 <= 1.05 * e5 + 1e-5


def _run_core(policy, proj, mask, dva, lay, dtype):
    """forward + backward of the core through the C ABI under a kernel policy -> (va, dproj) on the CPU."""
    B, N = proj.shape[0], proj.shape[1]
    C = proj.shape[-1]
    R = B * N * N
    desc = _desc(B, N, C, lay, dtype)
    pd, md, dv = proj.to(DEV).view(R, C), mask.to(DEV), dva.to(DEV).view(R, 2 * W)
    va = torch.empty((R, 2 * W), dtype=dtype, device=DEV)
    dproj = torch.full((R, C), float("nan"), dtype=dtype, device=DEV)
    stats = torch.empty((B, 2, H, N, N, 2), dtype=torch.float32, device=DEV)
    _C.set_kernel_policy(policy)
    _C.kernel_timer(True)
    try:
        lib = _C.lib()
        nf, nb = lib.tgt_triplet_attn_workspace_bytes(desc, 0), lib.tgt_triplet_attn_workspace_bytes(desc, 1)
        wf = torch.empty(nf, dtype=torch.uint8, device=DEV)
        wb = torch.empty(nb, dtype=torch.uint8, device=DEV)
        _C.check(lib.tgt_triplet_attn_fwd(desc, _C.ptr(pd), _C.ptr(md), _C.ptr(va), _C.ptr(stats), _C.ptr(wf), nf,
                                          _C.stream_ptr()), "fwd")
        _C.check(lib.tgt_triplet_attn_bwd_tiles(desc, _C.ptr(pd), _C.ptr(md), _C.ptr(va), _C.ptr(dv), _C.ptr(stats),
                                                _C.ptr(dproj), _C.ptr(wb), nb, _C.ptr(wf), _C.stream_ptr()), "bwd")
        torch.cuda.synchronize()
        names = set(_C.kernel_timer_read())
    finally:
        _C.set_kernel_policy(0)
        _C.kernel_timer(False)
    return va.view(B, N, N, 2 * W).float().cpu(), dproj.view(B, N, N, C).float().cpu(), names


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("N,nn_", [(64, [64, 37, 50]), (48, [48, 25, 1]), (33, [33, 20, 9]), (1, [1, 1, 1])])
@pytest.mark.parametrize("variant", ["gated", "ungated", "axial"])
def test_core_backward_tcgen05_vs_autograd_of_oracle(dtype, N, nn_, variant):
    """d(proj) of the tcgen05 backward (policy 4) against autograd of the fp64 core oracle on the same 16-bit inputs, per
    column block (q, k, v of both directions, bias, gate), next to the mma.sync backward (policy 5) on identical inputs.
    Only entries of real atom pairs are compared for q / k / v (padded rows of e receive no gradient in the model: the
    heads mask them), the bias / gate blocks everywhere."""
    lay = _layout(gated=variant == "gated", biased=variant != "axial")
    off_q, off_k, off_v, off_e, off_g, C = lay
    B = len(nn_)
    proj, mask = _proj_and_mask(B, N, nn_, C, dtype, seed=11 + N, std=0.6)
    g = torch.Generator().manual_seed(5)
    real = (mask == 0).float()                                           # [B,N,N] real pair (i, j)
    dva = (torch.randn(B, N, N, 2 * W, generator=g) * real.unsqueeze(-1)).to(dtype)
    pr = proj.double().requires_grad_(True)
    ref = O.triplet_attention_core(pr, mask.double(), H, D_, off_q, off_k, off_v, off_e, off_g)
    ref.backward(dva.double())
    want = pr.grad
    res = {pol: _run_core(pol, proj, mask, dva, lay, dtype) for pol in (4, 5)}
    assert "tri_attn_bwd_tc" in res[4][2] and "tri_attn_fwd_tc" in res[4][2], res[4][2]
    assert "tri_attn_bwd_tc" not in res[5][2]
    blocks = {"q_in": (off_q[0], W), "k_in": (off_k[0], W), "v_in": (off_v[0], W), "q_out": (off_q[1], W),
              "k_out": (off_k[1], W), "v_out": (off_v[1], W)}
    for dirn in (0, 1):
        if off_e[dirn] >= 0:
            blocks[f"e_{dirn}"] = (off_e[dirn], H)
        if off_g[dirn] >= 0:
            blocks[f"g_{dirn}"] = (off_g[dirn], H)
    m = real.unsqueeze(-1).double()
    for name, (o, wd) in blocks.items():
        w_ = want[..., o:o + wd] * m
        if float(w_.norm()) == 0.0:
            continue
        errs = {pol: rel_err(res[pol][1][..., o:o + wd].double() * m, w_) for pol in (4, 5)}
        print(f"bwd {variant} {dtype} N={N} {name}: tcgen05 {errs[4]:.3e}  mma.sync {errs[5]:.3e}")
        assert bool(torch.isfinite(res[4][1][..., o:o + wd]).all()), name
        tol = 2.5e-2 if name.startswith(("e_", "g_")) else 1e-2
        if name.startswith("k_") and errs[5] > tol:
            continue        # d(bias of K) is analytically ~0 (softmax shift invariance): both families only hold round-off
        assert errs[4] < tol, (name, errs)
        assert errs[4] <= 1.3 * errs[5] + 2e-4, (name, errs)
