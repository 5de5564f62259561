"""GPU parity tests: tgt_b200 CUDA path (through the C ABI) vs the golden vectors of the real reference and vs
the CPU oracle.  Tolerances (stated per BASELINE.json north_star):
  fp32 : max |ours - ref| / max |ref| <= 1e-5          (TF32 is off: torch default for matmul)
  bf16 : the binding criterion is  err(ours vs fp64 golden) <= 1.0 x err(reference algorithm under CUDA bf16 autocast
         vs fp64 golden)  on the same inputs, with NO additive slack, for module outputs and input gradients
         (relative L2; rounding a tensor to bf16 alone costs ~1.1e-3, so an element-wise 1e-3 is unreachable for ANY
         bf16 implementation, the reference's included -- SURVEY.md section 7).  BF16_TOL = 1e-2 is only a sanity
         ceiling.  The other half of SURVEY section 7's protocol -- the attention core with fp32 outputs on bf16-rounded
         inputs within 1e-3 of the fp64 oracle -- is tests/test_gpu_triplet_tc.py; whole layers / encoders at the
         shipped geometry are tests/test_gpu_shipped_geometry.py.
"""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, load_golden, max_rel, rel_err, assert_grads_close
from oracle import tgt_oracle as O
from tgt_b200 import TGT_Encoder, Graph, _C, ops
from tgt_b200 import layers as L
from tgt_b200.harness import models as HM
from tgt_b200.harness.synthetic import make_edge_inputs, make_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"
FP32_TOL = 1e-5
BF16_TOL = 1e-2

TRIPLET_FIX = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "triplet_*.pt")))


def _module_from_fixture(fx):
    mod = L.get_triplet_layer(fx["kind"])(fx["edge_width"], fx["num_heads"])
    mod.load_state_dict(fx["state"], strict=True)
    return mod.to(DEV)


def _run_triplet(mod, fx, autocast=None):
    e = fx["e"].to(DEV).requires_grad_(True)
    mask = fx["mask"].to(DEV)
    mod.zero_grad(set_to_none=True)
    if autocast is not None:
        with torch.autocast("cuda", dtype=autocast):
            out = mod(e, mask)
    else:
        out = mod(e, mask)
    out.backward(fx["dout"].to(DEV).to(out.dtype))
    return out, e.grad, {k: p.grad for k, p in mod.named_parameters()}


@pytest.mark.parametrize("name", TRIPLET_FIX)
def test_triplet_fp32_vs_golden(name):
    fx = load_golden(name)
    mod = _module_from_fixture(fx)
    out, de, grads = _run_triplet(mod, fx)
    assert out.dtype == torch.float32
    assert max_rel(out.cpu(), fx["out"]) < FP32_TOL
    assert max_rel(de.cpu(), fx["de"]) < FP32_TOL
    assert_grads_close(grads, fx["grads"], 2 * FP32_TOL, "max")


def _oracle_autocast_error(fx, dtype):
    """Error of the reference algorithm (oracle restatement) run on CUDA under autocast, vs the golden."""
    p = {k: v.to(DEV).requires_grad_(True) for k, v in fx["state"].items()}
    e = fx["e"].to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=dtype):
        out = O.TRIPLET_FNS[fx["kind"]](p, e, fx["mask"].to(DEV), fx["num_heads"])
    out.backward(fx["dout"].to(DEV).to(out.dtype))
    return rel_err(out.float().cpu(), fx["out"]), rel_err(e.grad.cpu(), fx["de"])


@pytest.mark.parametrize("policy", [0, 1, 4, 5])     # 0 = default (tcgen05 fwd + mma.sync bwd), 1 = generic SIMT, 4 = tcgen05 fwd + bwd, 5 = TMA-staged mma.sync
@pytest.mark.parametrize("name", TRIPLET_FIX)
def test_triplet_bf16_vs_golden(name, policy):
    fx = load_golden(name)
    mod = _module_from_fixture(fx)
    _C.set_kernel_policy(policy)
    try:
        out, de, grads = _run_triplet(mod, fx, autocast=torch.bfloat16)
    finally:
        _C.set_kernel_policy(0)
    assert out.dtype == torch.bfloat16
    ref_eo, ref_ed = _oracle_autocast_error(fx, torch.bfloat16)
    eo, ed = rel_err(out.float().cpu(), fx["out"]), rel_err(de.cpu(), fx["de"])
    print(f"bf16 {name} policy {policy}: out {eo:.3e} (reference autocast {ref_eo:.3e}), de {ed:.3e} ({ref_ed:.3e})")
    assert eo < BF16_TOL and eo <= ref_eo, (eo, ref_eo)
    assert ed < 2 * BF16_TOL and ed <= ref_ed, (ed, ref_ed)
    assert_grads_close(grads, fx["grads"], 3 * BF16_TOL, "l2")


@pytest.mark.parametrize("name", ["triplet_attention_a.pt", "triplet_aggregate_a.pt"])
def test_triplet_fp16_vs_golden(name):
    fx = load_golden(name)
    mod = _module_from_fixture(fx)
    out, de, grads = _run_triplet(mod, fx, autocast=torch.float16)
    assert out.dtype == torch.float16
    assert rel_err(out.float().cpu(), fx["out"]) < 3e-3
    assert rel_err(de.cpu(), fx["de"]) < 6e-3


@pytest.mark.parametrize("name", ["egt.pt", "egt_noscale_noedge.pt", "edge_update.pt"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_egt_vs_golden(name, mode):
    fx = load_golden(name)
    if fx["kind"] == "edge_update":
        mod = L.EdgeUpdate(fx["node_width"], fx["edge_width"], fx["num_heads"])
    else:
        mod = L.EGT_Attention(fx["node_width"], fx["edge_width"], fx["num_heads"], **fx["kwargs"])
    mod.load_state_dict(fx["state"], strict=True)
    mod = mod.to(DEV).eval()
    h = fx["h"].to(DEV).requires_grad_(True)
    e = fx["e"].to(DEV).requires_grad_(True)
    mask = fx["mask"].to(DEV)
    if mode == "bf16":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ho, eo = mod(h, e, mask)
    else:
        ho, eo = mod(h, e, mask)
    loss = (ho.float() * fx["dh_out"].to(DEV)).sum()
    if fx["de"] is not None or eo is e:
        loss = loss + (eo.float() * fx["de_out"].to(DEV)).sum()
    loss.backward()
    if mode == "fp32":
        chk = lambda a, b, k="": (max_rel(a.cpu(), b) < 2 * FP32_TOL) or pytest.fail(f"{k}: {max_rel(a.cpu(), b)}")
    else:
        chk = lambda a, b, k="": (rel_err(a.float().cpu(), b) < 2 * BF16_TOL) or pytest.fail(f"{k}: {rel_err(a.float().cpu(), b)}")
    chk(ho, fx["h_out"], "h_out")
    chk(eo, fx["e_out"], "e_out")
    if fx["dh"] is not None:
        chk(h.grad, fx["dh"], "dh")
    if fx["de"] is not None:
        chk(e.grad, fx["de"], "de")
    assert_grads_close({k: v.grad for k, v in mod.named_parameters() if k in fx["grads"]}, fx["grads"],
                       2 * FP32_TOL if mode == "fp32" else 2 * BF16_TOL, "max" if mode == "fp32" else "l2")


@pytest.mark.parametrize("name,cls", [("model_multi_at.pt", HM.TGT_Multi), ("model_gap_agx2.pt", HM.TGT_Gap),
                                      ("model_dist_at.pt", HM.TGT_Distance)])
def test_whole_model_fp32_vs_golden(name, cls):
    """Reference task models (lib/models/pcqm) on top of our encoder: eval-mode forward equals the reference's."""
    fx = load_golden(name)
    model = cls(**fx["cfg"], **fx["extra"])
    model.load_state_dict(fx["state"], strict=True)
    model = model.to(DEV).eval()
    batch = {k: v.to(DEV) for k, v in fx["batch"].items()}
    with torch.no_grad():
        out = model(batch)
    outs = out if isinstance(out, tuple) else (out,)
    for o, g in zip(outs, fx["outs"]):
        assert max_rel(o.cpu(), g) < 5 * FP32_TOL


def test_config1_triplet_attention_vs_oracle():
    """BASELINE.json configs[0]: TripletAttention(512, 8), B=4, N=16 (d=64), seed 0; fwd + all grads, fp32."""
    torch.manual_seed(0)
    mod = L.TripletAttention(512, 8)
    e, mask = make_edge_inputs(4, 16, 512, [16, 12, 9, 5], seed=0)
    p = {k: v.double().requires_grad_(True) for k, v in mod.state_dict().items()}
    ed = e.double().requires_grad_(True)
    ref = O.triplet_attention(p, ed, mask.double(), 8)
    dout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    ref.backward(dout)
    mod = mod.to(DEV)
    eg = e.to(DEV).requires_grad_(True)
    out = mod(eg, mask.to(DEV))
    out.backward(dout.float().to(DEV))
    assert max_rel(out.cpu(), ref) < FP32_TOL
    assert max_rel(eg.grad.cpu(), ed.grad) < FP32_TOL
    assert_grads_close({k: v.grad for k, v in mod.named_parameters()}, {k: v.grad for k, v in p.items()},
                       2 * FP32_TOL, "max")
    import json
    kat = json.load(open(os.path.join(GOLDEN, "config1_kat.json")))       # checksum produced by the real reference
    assert abs(float(out.double().sum()) - kat["sum"]) < 1e-4 * kat["abs_sum"]
    for (b, i, j, c), v in zip([(0, 0, 0, 0), (1, 3, 7, 100), (2, 8, 8, 511), (3, 4, 2, 17)], kat["samples"]):
        assert abs(float(out[b, i, j, c]) - v) < 1e-5 * max(1.0, abs(v))


@pytest.mark.parametrize("kind,N,nn_", [("attention", 64, [64, 37]), ("attention", 48, [48, 25]),
                                        ("attention", 33, [33, 1]), ("aggregate", 32, [32, 17]),
                                        ("attention", 1, [1, 1]), ("aggregate", 64, [64, 40])])
def test_shipped_width_bf16_vs_oracle(kind, N, nn_):
    """Shipped head geometry (We=256, Ht=16, d=16) at ragged N up to the 64 maximum, bf16 tensor-core paths (tcgen05 and
    mma.sync families) vs the fp64 oracle, and the tensor-core kernels vs the generic kernels (policy 1) on identical inputs."""
    torch.manual_seed(3)
    mod = L.get_triplet_layer(kind)(256, 16)
    e, mask = make_edge_inputs(2, N, 256, nn_, seed=5)
    p = {k: v.double().requires_grad_(True) for k, v in mod.state_dict().items()}
    ed = e.double().requires_grad_(True)
    ref = O.TRIPLET_FNS[kind](p, ed, mask.double(), 16)
    dout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    ref.backward(dout)
    mod = mod.to(DEV)
    res = {}
    for policy in (0, 1, 4, 5):
        _C.set_kernel_policy(policy)
        try:
            mod.zero_grad(set_to_none=True)
            eg = e.to(DEV).requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = mod(eg, mask.to(DEV))
            out.backward(dout.to(DEV).to(out.dtype))
            res[policy] = (out.float().cpu(), eg.grad.cpu(), {k: v.grad.cpu() for k, v in mod.named_parameters()})
        finally:
            _C.set_kernel_policy(0)
        assert rel_err(res[policy][0], ref) < BF16_TOL
        assert rel_err(res[policy][1], ed.grad) < 2 * BF16_TOL
        assert_grads_close(res[policy][2], {k: v.grad for k, v in p.items()}, 3 * BF16_TOL, "l2")
    assert rel_err(res[0][0], res[1][0]) < BF16_TOL
    # the tcgen05 and the mma.sync families run the same arithmetic; only the order of the fp32 row sums and the delta
    # of the backward (dO . O vs sum dP P) differ, i.e. a few 16-bit roundings flip
    assert rel_err(res[4][0], res[5][0]) < 2e-3 and rel_err(res[4][1], res[5][1]) < 5e-3


def test_outputs_survive_inplace_residual_add():
    """TGT_Layer adds residuals in place to our outputs (layers.py:270-290): nothing saved may alias them."""
    torch.manual_seed(0)
    mod = L.TripletAttention(32, 2).to(DEV)
    e, mask = make_edge_inputs(2, 6, 32, [6, 3], seed=1)
    e = e.to(DEV).requires_grad_(True)
    out = mod(e, mask.to(DEV))
    out.add_(e.detach())
    out.mul_(2.0)
    out.sum().backward()
    assert torch.isfinite(e.grad).all()


def test_unsupported_inputs_raise():
    mod = L.TripletAttention(32, 2).to(DEV)
    e, mask = make_edge_inputs(1, 65, 32, [65], seed=1)
    with pytest.raises(RuntimeError, match="N=65"):
        mod(e.to(DEV), mask.to(DEV))
    mod2 = L.TripletAttention(32, 2, attention_dropout=0.1).to(DEV).train()
    e, mask = make_edge_inputs(1, 4, 32, [4], seed=1)
    with pytest.raises(NotImplementedError):
        mod2(e.to(DEV), mask.to(DEV))
    mod2.eval()
    mod2(e.to(DEV), mask.to(DEV))          # dropout inactive in eval: fine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.TriangularUpdate(32, 2)(e, mask)             # CPU tensors: refuse, never fall back


def test_padding_leak_of_aggregate_is_reproduced():
    """TripletAggregate's outward softmax is unmasked (triplet.py:63-64): perturbing PADDED entries of e changes
    real outputs in the reference; ours must reproduce that, not 'fix' it."""
    torch.manual_seed(0)
    mod = L.TripletAggregate(32, 2).to(DEV)
    e, mask = make_edge_inputs(1, 8, 32, [5], seed=1)
    e2 = e.clone()
    e2[:, 5:, :, :] += torch.randn(e2[:, 5:, :, :].shape, generator=torch.Generator().manual_seed(7))   # (a constant shift would be removed by the LayerNorm)
    o1 = mod(e.to(DEV), mask.to(DEV))
    o2 = mod(e2.to(DEV), mask.to(DEV))
    assert (o1[:, :5, :5] - o2[:, :5, :5]).abs().max() > 1e-3
    moda = L.TripletAttention(32, 2).to(DEV)
    o1 = moda(e.to(DEV), mask.to(DEV))
    o2 = moda(e2.to(DEV), mask.to(DEV))
    assert (o1[:, :5, :5] - o2[:, :5, :5]).abs().max() < 1e-5


def test_rowwise_kernels():
    torch.manual_seed(0)
    for W, xdt, ydt in [(256, torch.float32, torch.float32), (256, torch.bfloat16, torch.bfloat16),
                        (768, torch.float32, torch.bfloat16), (768, torch.bfloat16, torch.bfloat16),
                        (36, torch.float32, torch.float32)]:
        x = torch.randn(1000, W, device=DEV).to(xdt)
        g = torch.rand(W, device=DEV) + 0.5
        b = torch.randn(W, device=DEV)
        y, mean, rstd = ops.layernorm_fwd(x, g, b, ydt)
        ref = torch.nn.functional.layer_norm(x.float(), (W,), g, b)
        tol = 1e-5 if ydt == torch.float32 else 1e-2
        assert max_rel(y.float(), ref) < tol
        dy = torch.randn(1000, W, device=DEV).to(ydt)
        xr = x.float().requires_grad_(True)
        gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xr, (W,), gr, br).backward(dy.float())
        dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd)
        assert max_rel(dx.float(), xr.grad) < (1e-5 if xdt == torch.float32 else 2e-2)
        assert max_rel(dg, gr.grad) < 1e-4 and max_rel(db, br.grad) < 1e-4
        if xdt == ydt:                                       # residual gradient folded in (node rows: ln_bwd_wide at W = 768)
            dres = torch.randn(1000, W, device=DEV).to(xdt)
            dx2, dg2, db2 = ops.layernorm_bwd(dy, x, g, mean, rstd, dres)
            assert max_rel(dx2.float(), xr.grad + dres.float()) < (1e-5 if xdt == torch.float32 else 2e-2)
            assert max_rel(dg2, gr.grad) < 1e-4 and max_rel(db2, br.grad) < 1e-4
    # LayerNorm backward that also emits the augmented y = [LN(x) | 1 | 0..] (tgt_layernorm_bwd_y, W = 256, 16-bit dy):
    # same dx / dgamma / dbeta as the plain call, y equal to the forward kernel's output for the same statistics
    for xdt in (torch.bfloat16, torch.float32):
        x = torch.randn(777, 256, device=DEV).to(xdt)
        g = torch.rand(256, device=DEV) + 0.5
        b = torch.randn(256, device=DEV)
        dy = torch.randn(777, 256, device=DEV).bfloat16()
        dres = torch.randn(777, 256, device=DEV).to(xdt)
        assert ops.ln_bwd_emits_y(x, torch.bfloat16)
        yf, mean, rstd = ops.layernorm_fwd(x, g, b, torch.bfloat16, aug=True)
        dx0, dg0, db0 = ops.layernorm_bwd(dy, x, g, mean, rstd, dres)
        dx1, dg1, db1, y1 = ops.layernorm_bwd(dy, x, g, mean, rstd, dres, beta=b)
        assert torch.equal(dx0, dx1) and max_rel(dg1, dg0) < 1e-5 and max_rel(db1, db0) < 1e-5
        assert y1.shape == (777, 264) and torch.equal(y1, yf)
        # ... and the DropPath-weighted column sum of the residual gradient (= bias gradient of the module's output Linear)
        sc = torch.tensor([0., 1.25, 2.5, 0., 1., 1., 0.5], device=DEV)          # 7 graphs x 111 rows
        dx2, dg2, db2, y2, cs = ops.layernorm_bwd(dy, x, g, mean, rstd, dres, beta=b, colsum=(sc,))
        assert torch.equal(dx2, dx0) and torch.equal(y2, yf)
        ref_cs = (dres.float().view(7, 111, 256) * sc.view(7, 1, 1)).sum((0, 1))
        assert max_rel(cs, ref_cs) < 1e-5
        cs1 = ops.layernorm_bwd(dy, x, g, mean, rstd, dres, colsum=(None,))[3]
        assert max_rel(cs1, dres.float().sum(0)) < 1e-5
    assert not ops.ln_bwd_emits_y(torch.empty(4, 768), torch.bfloat16)
    assert not ops.ln_bwd_emits_y(torch.empty(4, 256), torch.float32)
    # gelu + dropout: p=0 equals F.gelu; p>0 keeps ~(1-p), scales by 1/(1-p), fwd/bwd masks agree
    u = torch.randn(4096, 256, device=DEV)
    y = torch.empty_like(u)
    _C.check(_C.lib().tgt_gelu_dropout_fwd(_C.ptr(u), _C.ptr(y), u.numel(), 0.0, 0, 0, _C.stream_ptr()), "gelu")
    assert max_rel(y, torch.nn.functional.gelu(u)) < 1e-6
    _C.check(_C.lib().tgt_gelu_dropout_fwd(_C.ptr(u), _C.ptr(y), u.numel(), 0.25, 1234, 0, _C.stream_ptr()), "gelu")
    keep = (y != 0) | (u == 0)
    assert abs(float(keep.float().mean()) - 0.75) < 0.01
    assert max_rel(y[keep], torch.nn.functional.gelu(u)[keep] / 0.75) < 1e-6
    du = torch.empty_like(u)
    ones = torch.ones_like(u)
    _C.check(_C.lib().tgt_gelu_dropout_bwd(_C.ptr(u), _C.ptr(ones), _C.ptr(du), u.numel(), 0.25, 1234, 0,
                                           _C.stream_ptr()), "gelu_bwd")
    assert bool(((du != 0) == (y != 0)).float().mean() > 0.999)
    ur = u.clone().requires_grad_(True)
    torch.nn.functional.gelu(ur).sum().backward()
    assert max_rel(du[keep], ur.grad[keep] / 0.75) < 1e-5
    # scaled residual
    x = torch.randn(4, 8, 8, 32, device=DEV).bfloat16()
    r = torch.randn(4, 8, 8, 32, device=DEV)
    s = torch.tensor([0., 1.25, 1.25, 0.], device=DEV)
    out = ops.scaled_residual(x, r, s)
    assert out.dtype == torch.bfloat16
    assert max_rel(out.float(), r + s.view(4, 1, 1, 1) * x.float()) < 1e-2


def test_layer_train_mode_statistics_and_grads():
    """Train mode with all dropouts on: runs, finite, and drop_path=1-eps style sanity (p=0 equals eval)."""
    torch.manual_seed(0)
    kw = dict(node_width=48, edge_width=32, num_heads=4, triplet_heads=2, triplet_type="attention")
    layer = L.TGT_Layer(**kw, source_dropout=0.3, drop_path=0.2, node_act_dropout=0.1, edge_act_dropout=0.1).to(DEV)
    e, mask = make_edge_inputs(4, 10, 32, [10, 7, 5, 2], seed=1)
    h = torch.randn(4, 10, 48, device=DEV)
    g = Graph(h=h.requires_grad_(True), e=e.to(DEV).requires_grad_(True), mask=mask.to(DEV), node_mask=None)
    layer.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = layer(g)
    (out.h.float().sum() + out.e.float().sum()).backward()
    assert out.e.dtype == torch.bfloat16 and out.h.dtype == torch.bfloat16      # SURVEY 5.8 residual-stream dtype
    assert torch.isfinite(g.h.grad).all() and torch.isfinite(g.e.grad).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in layer.parameters())
    assert "node_mask" in out


def test_encoder_fp32_train_vs_oracle_grads():
    """3-layer encoder (attention + a last layer without edge branch), fp32, dropouts off but train mode:
    output and parameter gradients vs the fp64 oracle."""
    torch.manual_seed(4)
    cfg = dict(node_width=48, edge_width=32, num_heads=4, triplet_heads=2, triplet_type="attention")
    enc = TGT_Encoder(model_height=3, node_ended=True, edge_ended=False, **cfg)
    e, mask = make_edge_inputs(2, 9, 32, [9, 5], seed=3)
    h = torch.randn(2, 9, 48)
    p = {k: v.double().requires_grad_(True) for k, v in enc.state_dict().items()}
    ocfg = dict(cfg)
    ocfg.pop("node_width"), ocfg.pop("edge_width")
    ho, eo = O.encoder(p, h.double(), e.double(), mask.double(), model_height=3, node_ended=True,
                       edge_ended=False, **ocfg)
    (ho.sum() + 0.5 * eo.sum()).backward()
    enc = enc.to(DEV).train()
    g = enc(Graph(h=h.to(DEV), e=e.to(DEV), mask=mask.to(DEV)))
    (g.h.sum() + 0.5 * g.e.sum()).backward()
    assert max_rel(g.h.cpu(), ho) < 5 * FP32_TOL and max_rel(g.e.cpu(), eo) < 5 * FP32_TOL
    assert_grads_close({k: v.grad for k, v in enc.named_parameters()}, {k: v.grad for k, v in p.items()},
                       1e-4, "max")


@pytest.mark.parametrize("kind", ["attention", "attention_ungated", "axial_attention"])
@pytest.mark.parametrize("N,nn_", [(64, [64, 37, 50]), (33, [33, 20, 9]), (8, [8, 5, 1])])
def test_fused_forward_matches_unfused(kind, N, nn_):
    """bf16 edge input at the shipped geometry: the fused projection+attention kernel (policy 3) must reproduce the
    un-fused tensor-core path (policy 0: LN-folded GEMM -> [R, 6We] projection -> TMA attention kernel) and both must
    sit within bf16 tolerance of the fp64 oracle; gradients flow through the shared backward."""
    from tgt_b200 import ops
    torch.manual_seed(11)
    mod = L.get_triplet_layer(kind)(256, 16)
    e, mask = make_edge_inputs(3, N, 256, nn_, seed=7)
    e16 = e.bfloat16()
    p = {k: v.double().requires_grad_(True) for k, v in mod.state_dict().items()}
    ed = e16.double().requires_grad_(True)
    ref = O.TRIPLET_FNS[kind](p, ed, mask.double(), 16)
    dout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
    ref.backward(dout)
    mod = mod.to(DEV)
    res = {}
    for policy in (3, 0):
        _C.set_kernel_policy(policy)
        ops.KernelTimer.reset(True)
        try:
            mod.zero_grad(set_to_none=True)
            eg = e16.to(DEV).requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = mod(eg, mask.to(DEV))
            out.backward(dout.to(DEV).to(out.dtype))
            torch.cuda.synchronize()
            names = set(ops.KernelTimer.summary())
            res[policy] = (out.float().cpu(), eg.grad.float().cpu())
        finally:
            _C.set_kernel_policy(0)
            ops.KernelTimer.reset(False)
        assert ("triplet_fused_fwd" in names) == (policy == 3), names
        assert rel_err(res[policy][0], ref) < BF16_TOL
        assert rel_err(res[policy][1], ed.grad) < 2 * BF16_TOL
    assert rel_err(res[3][0], res[0][0]) < 2e-3, rel_err(res[3][0], res[0][0])
    assert rel_err(res[3][1], res[0][1]) < 2e-2


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gaussian_basis_kernel_vs_torch(out_dtype):
    """tgt_gaussian_basis_fwd/bwd against the reference's element-wise formula (models/pcqm/layers.py:25-48)."""
    from tgt_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(3, 17, 17, generator=g) * 6).to(DEV).requires_grad_(True)
    mu = (torch.rand(128, generator=g) * 3).to(DEV).requires_grad_(True)
    sd = (torch.rand(128, generator=g) * 3 + 0.01).to(DEV).requires_grad_(True)
    ref = torch.exp(-0.5 * ((x.unsqueeze(-1) - mu) / sd) ** 2) / ((2 * 3.14159) ** 0.5 * sd)
    dout = torch.randn(ref.shape, generator=g).to(DEV)
    ref.backward(dout)
    want = (x.grad.clone(), mu.grad.clone(), sd.grad.clone())
    x.grad = mu.grad = sd.grad = None
    out = ops.GaussianBasisFn.apply(x, mu, sd, out_dtype)
    out.backward(dout.to(out_dtype))
    tol = 1e-5 if out_dtype == torch.float32 else 1e-2
    assert rel_err(out.float().cpu(), ref.detach().cpu()) < tol
    for got, w in zip((x.grad, mu.grad, sd.grad), want):
        assert rel_err(got.cpu(), w.cpu()) < (1e-4 if out_dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bins_decode_kernel_vs_golden_and_oracle(dtype):
    """tgt_bins_decode (softmax + p + p^T + argmax + bins2dist in one pass) vs the reference-made fixture (fp32 logits:
    bit-exact bins and distances) and vs the oracle on the 16-bit-rounded logits.  A bin may only differ where the two
    best symmetrised probabilities are closer than fp32 softmax rounding (1e-6 relative)."""
    fx = load_golden("bins_decode.pt")
    lg = fx["logits"].to(dtype)
    B, S, N = lg.shape[:3]
    for s in range(S):
        bins, dist = ops.bins_decode(lg[:, s].to(DEV), range_bins=fx["range_bins"])
        ref_bins = O.bins_from_logits(lg[:, s].float())
        ref_dist = O.bins2dist(ref_bins, fx["num_bins"], fx["range_bins"])
        if dtype == torch.float32:
            assert torch.equal(ref_bins.to(torch.int16), fx["bins"][:, s])
        p = torch.softmax(lg[:, s].float(), -1)
        p = p + p.transpose(-2, -3)
        top2 = p.topk(2, dim=-1).values
        clear = (top2[..., 0] - top2[..., 1]) > 1e-6 * top2[..., 0]
        same = bins.cpu().long() == ref_bins
        assert bool((same | ~clear).all()) and float(same.float().mean()) > 0.999
        assert torch.equal(dist.cpu()[same], ref_dist[same])
        assert torch.equal(bins, bins.transpose(-1, -2)) and torch.equal(dist, dist.transpose(-1, -2))


def test_bins_decode_edge_cases():
    """ragged sizes, num_bins not a multiple of 32, exact ties (lowest bin wins like torch.argmax), limits."""
    g = torch.Generator().manual_seed(5)
    for B, N, nb in [(1, 1, 8), (2, 7, 100), (1, 64, 512), (3, 5, 33)]:
        lg = torch.randn(B, N, N, nb, generator=g)
        bins, dist = ops.bins_decode(lg.to(DEV), range_bins=8.0)
        rb = O.bins_from_logits(lg)
        assert float((bins.cpu().long() == rb).float().mean()) > 0.999
    lg = torch.zeros(1, 4, 4, 64)                                  # all bins tie -> bin 0 everywhere
    lg[0, 1, 2, 10] = lg[0, 1, 2, 20] = 5.0                       # two-way tie of the pair (1, 2)
    lg[0, 2, 1, 10] = lg[0, 2, 1, 20] = 5.0
    bins, dist = ops.bins_decode(lg.to(DEV), range_bins=8.0)
    assert int(bins[0, 0, 3]) == 0 and int(bins[0, 1, 2]) == 10 and int(bins[0, 2, 1]) == 10
    assert float(dist[0, 0, 0]) == 0.0 and abs(float(dist[0, 1, 2]) - 2 * 10.5 * 8.0 / 63) < 1e-6
    with pytest.raises(RuntimeError, match="num_bins"):
        ops.bins_decode(torch.zeros(1, 2, 2, 1024, device=DEV))


def test_two_stage_inference_matches_oracle_pipeline():
    """Config 5 glue: TGT_Distance -> bins -> distances -> TGT_Gap (harness.inference) vs the same pipeline with the oracle's
    bins_from_logits / bins2dist on OUR logits (eval mode, fp32, one sample: deterministic)."""
    from tgt_b200.harness import inference as INF
    torch.manual_seed(0)
    cfg = dict(node_width=64, edge_width=32, num_heads=4, triplet_heads=2, triplet_type="attention")
    dm = HM.TGT_Distance(model_height=2, num_dist_bins=64, embed_3d_type="none", **cfg).to(DEV).eval()
    gm = HM.TGT_Gap(model_height=2, **cfg).to(DEV).eval()
    batch = {k: v.to(DEV) for k, v in make_batch(3, 9, seed=2).items()}
    bins, d = INF.predict_dist_inputs(dm, batch, samples=1, amp_dtype=None, want_bins=True)
    from tgt_b200.harness.synthetic import add_scheme_fields
    with torch.no_grad():
        logits = dm(add_scheme_fields(batch, with_3d=False))
    rb = O.bins_from_logits(logits.cpu())
    assert float((bins[:, 0].cpu().long() == rb).float().mean()) > 0.99
    assert torch.allclose(d[:, 0].cpu(), O.bins2dist(bins[:, 0].cpu().long(), 64), atol=1e-6)
    pred = INF.predict_gap(gm, batch, d, samples=1, amp_dtype=None)
    b2 = dict(add_scheme_fields(batch, with_3d=False))
    b2["dist_input"] = d[:, 0]
    with torch.no_grad():
        ref = gm(b2)
    assert torch.allclose(pred, ref.float(), rtol=1e-5, atol=1e-5)
    assert pred.shape == (3,) and bool(torch.isfinite(pred).all())


def test_full_size_batch_invariance_and_padding_independence():
    """BASELINE config-3 size (B = 256, N = 64, Wn = 768, We = 256, Hn = 64, Ht = 16, bf16 autocast) through two
    size-independent properties of the reference path: (1) a graph's outputs and input gradients do not depend on which
    other graphs share the batch -- the full batch must agree with the same graphs run four at a time (only cuBLAS's
    node-side tile choice may differ: bf16-level tolerance, edge kernels never mix graphs); (2) for TripletAttention the
    values on PADDED atoms of the input never reach real-pair outputs (unlike TripletAggregate, see the leak test)."""
    torch.manual_seed(0)
    B, N = 256, 64
    layer = L.TGT_Layer(768, 256, 64, triplet_heads=16, triplet_type="attention").to(DEV).eval()
    nn_ = [N] + [N // 2 + (i * 7) % (N // 2 + 1) for i in range(1, B)]
    e, mask = make_edge_inputs(B, N, 256, nn_, seed=3)
    gen = torch.Generator().manual_seed(9)
    h = torch.randn(B, N, 768, generator=gen)
    wh, we = torch.randn(B, N, 768, generator=gen), torch.randn(B, N, N, 256, generator=gen)

    def run(sl, e_src=e):
        hh = h[sl].to(DEV).requires_grad_(True)
        ee = e_src[sl].to(DEV).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            g = layer(Graph(h=hh, e=ee, mask=mask[sl].to(DEV)))
        ((g.h.float() * wh[sl].to(DEV)).sum() + (g.e.float() * we[sl].to(DEV)).sum()).backward()
        return g.h.float().detach().cpu(), g.e.float().detach().cpu(), hh.grad.cpu(), ee.grad.cpu()

    full = run(slice(0, B))
    assert all(bool(torch.isfinite(t).all()) for t in full)
    for b0 in (0, 100, 252):
        sub = run(slice(b0, b0 + 4))
        for name, f, s_ in zip(("h", "e", "dh", "de"), full, sub):
            assert rel_err(f[b0:b0 + 4], s_) < 5e-3, (name, b0, rel_err(f[b0:b0 + 4], s_))
    # (2) perturb padded entries of e (graph 1 has nn_[1] < N real atoms): real-pair outputs must not move
    n1 = nn_[1]
    e2 = e.clone()
    e2[1, n1:, :, :] += torch.randn(e2[1, n1:, :, :].shape, generator=gen)
    e2[1, :, n1:, :] += torch.randn(e2[1, :, n1:, :].shape, generator=gen)
    a = run(slice(0, 4))
    b_ = run(slice(0, 4), e_src=e2)
    assert torch.equal(a[1][1, :n1, :n1], b_[1][1, :n1, :n1]) and torch.equal(a[0][1, :n1], b_[0][1, :n1])


def test_full_size_properties_config2_and_config5():
    """Size-independent properties at the other BASELINE sizes.  Config 5 (B = 512, N = 48, 256 bins): the decoded bins /
    distances are symmetric with a zero diagonal, invariant to a per-pair logit shift (softmax) and equivariant to
    transposing the atom pair.  Config 2 (TripletAggregate, B = 256, N = 32, We = 256, Ht = 16, bf16): batch invariance of
    outputs and input gradients (bit-exact: no library GEMM on this path depends on the batch size per row)."""
    gen = torch.Generator(device=DEV).manual_seed(4)
    lg = (torch.randn(512, 48, 48, 256, device=DEV, generator=gen) * 2).bfloat16()
    bins, dist = ops.bins_decode(lg, range_bins=8.0)
    assert torch.equal(bins, bins.transpose(1, 2)) and torch.equal(dist, dist.transpose(1, 2))
    assert float(dist.diagonal(dim1=1, dim2=2).abs().max()) == 0.0 and int(bins.min()) >= 0 and int(bins.max()) <= 255
    off = torch.arange(48, device=DEV).float()
    assert torch.equal(dist[:, off.long(), off.long()], torch.zeros(512, 48, device=DEV))
    b2, d2 = ops.bins_decode(lg.transpose(1, 2).contiguous(), range_bins=8.0)
    assert torch.equal(b2, bins) and torch.equal(d2, dist)
    shift = torch.randn(512, 48, 48, 1, device=DEV, generator=gen)
    b3, _ = ops.bins_decode(lg.float() + shift, range_bins=8.0)
    assert float((b3 == ops.bins_decode(lg.float(), range_bins=8.0)[0]).float().mean()) > 0.9999
    del lg, shift

    torch.manual_seed(0)
    mod = L.TripletAggregate(256, 16).to(DEV)
    B, N = 256, 32
    nn_ = [N] + [N // 2 + (i * 5) % (N // 2 + 1) for i in range(1, B)]
    e, mask = make_edge_inputs(B, N, 256, nn_, seed=6)
    w = torch.randn(B, N, N, 256, generator=torch.Generator().manual_seed(2))

    def run(sl):
        ee = e[sl].to(DEV).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = mod(ee, mask[sl].to(DEV))
        (out.float() * w[sl].to(DEV)).sum().backward()
        return out.detach().cpu(), ee.grad.cpu()

    full = run(slice(0, B))
    for b0 in (0, 77, 252):
        sub = run(slice(b0, b0 + 4))
        assert torch.equal(full[0][b0:b0 + 4], sub[0]) and rel_err(full[1][b0:b0 + 4], sub[1]) < 1e-3


def test_cuda_graph_replay_of_mc_dropout_forward():
    """harness.inference.GraphedForward: a train-mode (MC-dropout) forward of TGT_Gap captured once as a CUDA graph.  With the
    dropouts off the replay reproduces the eager forward bit for bit; with them on, every replay draws fresh masks (device-
    resident seed of the fused GELU + dropout epilogue, torch's graph-safe generator for source dropout / DropPath) and the
    sample mean approaches the eager sample mean."""
    from tgt_b200.harness import inference as INF
    from tgt_b200.harness.synthetic import add_scheme_fields
    torch.manual_seed(0)
    cfg = dict(model_height=2, node_width=64, edge_width=32, num_heads=4, triplet_heads=2, triplet_type="aggregate")
    batch = add_scheme_fields({k: v.to(DEV) for k, v in make_batch(4, 12, seed=2).items()}, with_3d=True)
    det = HM.TGT_Gap(**cfg).to(DEV).train()                    # every dropout defaults to 0
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        want = det(batch).float()
    gf = INF.GraphedForward(det, batch)
    assert torch.equal(gf().float(), want) and torch.equal(gf().float(), want)
    mc = HM.TGT_Gap(**cfg, source_dropout=0.3, drop_path=0.2, node_act_dropout=0.1, edge_act_dropout=0.1).to(DEV).train()
    mc.load_state_dict(det.state_dict())
    gf = INF.GraphedForward(mc, batch)
    a, b_ = gf().float().clone(), gf().float().clone()
    assert bool(torch.isfinite(a).all()) and not torch.equal(a, b_)
    mean_g = torch.stack([gf().float().clone() for _ in range(64)]).mean(0)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        mean_e = torch.stack([mc(batch).float() for _ in range(64)]).mean(0)
    assert float((mean_g - mean_e).abs().max()) < 0.35 * float(mean_e.abs().max() + 1.0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("nbins,pitch", [(512, 512), (128, 128), (264, 272), (1024, 1024), (8, 16)])
def test_cross_entropy_rows_matches_torch(dtype, nbins, pitch):
    """DiscreteDistLoss's F.cross_entropy(..., reduction='none') (commons.py:38) and its gradient, incl. masked rows
    (zero upstream gradient: written as zeros without being read) and a strided logits view."""
    torch.manual_seed(3)
    R = 3001
    big = (torch.randn(R, pitch, device=DEV) * 3).to(dtype)
    logits = big[:, :nbins].detach().requires_grad_(True)
    target = torch.randint(0, nbins, (R,), device=DEV)
    w = torch.rand(R, device=DEV) * (torch.rand(R, device=DEV) > 0.4)
    assert ops.xent_rows_ok(logits, target)
    x = ops.cross_entropy_rows(logits, target)
    ref_in = big[:, :nbins].detach().float().requires_grad_(True)
    xr = torch.nn.functional.cross_entropy(ref_in, target, reduction='none')
    assert x.dtype == torch.float32 and max_rel(x, xr) < 2e-6
    (x * w).sum().backward()
    (xr * w).sum().backward()
    tol = 1e-5 if dtype == torch.float32 else (1e-2 if dtype == torch.bfloat16 else 2e-3)
    assert max_rel(logits.grad.float(), ref_in.grad) < tol
    assert (logits.grad[w == 0] == 0).all()
