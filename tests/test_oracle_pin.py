"""Pin the CPU oracle: (a) against golden vectors produced by the real reference (always), (b) against
the real reference imported live (build container only).  fp64 throughout: agreement must be ~1e-12
live and fp32-storage-limited (1e-6) against the stored fixtures."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, load_golden, max_rel
from oracle import tgt_oracle as O

TRIPLET_FIX = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "triplet_*.pt")))


def _dbl(d):
    return {k: v.double() for k, v in d.items()}


@pytest.mark.parametrize("name", TRIPLET_FIX)
def test_triplet_oracle_vs_golden(name):
    fx = load_golden(name)
    p = {k: v.double().requires_grad_(True) for k, v in fx["state"].items()}
    e = fx["e"].double().requires_grad_(True)
    out = O.TRIPLET_FNS[fx["kind"]](p, e, fx["mask"].double(), fx["num_heads"])
    assert max_rel(out, fx["out"]) < 2e-6
    out.backward(fx["dout"].double())
    assert max_rel(e.grad, fx["de"]) < 2e-6
    for k, g in fx["grads"].items():
        assert max_rel(p[k].grad, g) < 2e-6, k


@pytest.mark.parametrize("name", ["egt.pt", "egt_noscale_noedge.pt", "edge_update.pt"])
def test_egt_oracle_vs_golden(name):
    fx = load_golden(name)
    p = {k: v.double().requires_grad_(True) for k, v in fx["state"].items()}
    h = fx["h"].double().requires_grad_(True)
    e = fx["e"].double().requires_grad_(True)
    if fx["kind"] == "edge_update":
        ho, eo = O.edge_update(p, h, e, fx["mask"].double(), fx["num_heads"])
    else:
        ho, eo = O.egt_attention(p, h, e, fx["mask"].double(), fx["num_heads"], **fx["kwargs"])
    assert max_rel(ho, fx["h_out"]) < 2e-6
    assert max_rel(eo, fx["e_out"]) < 2e-6
    loss = (ho * fx["dh_out"].double()).sum()
    if eo.requires_grad and eo is not e:
        loss = loss + (eo * fx["de_out"].double()).sum()
    elif eo is e:
        loss = loss + (eo * fx["de_out"].double()).sum()
    loss.backward()
    if fx["dh"] is not None:
        assert max_rel(h.grad, fx["dh"]) < 2e-6
    if fx["de"] is not None:
        assert max_rel(e.grad, fx["de"]) < 2e-6
    for k, g in fx["grads"].items():
        assert max_rel(p[k].grad, g) < 2e-6, k


@pytest.mark.parametrize("name,fn", [("model_multi_at.pt", O.tgt_multi), ("model_gap_agx2.pt", O.tgt_gap),
                                     ("model_dist_at.pt", O.tgt_distance)])
def test_model_oracle_vs_golden(name, fn):
    fx = load_golden(name)
    p = _dbl(fx["state"])
    batch = {k: (v.double() if v.is_floating_point() else v) for k, v in fx["batch"].items()}
    cfg = dict(fx["cfg"])
    cfg.pop("node_width"), cfg.pop("edge_width")
    out = fn(p, batch, **cfg)
    outs = out if isinstance(out, tuple) else (out,)
    for o, g in zip(outs, fx["outs"]):
        assert max_rel(o, g) < 2e-6


def test_gated_core_backward_closed_form():
    """SURVEY.md appendix A backward formulas == autograd (second oracle for the CUDA backward)."""
    torch.manual_seed(0)
    N, d = 9, 5
    Q, K, V = (torch.randn(N, d, dtype=torch.float64, requires_grad=True) for _ in range(3))
    E, G = (torch.randn(N, N, dtype=torch.float64, requires_grad=True) for _ in range(2))
    M = torch.zeros(N, N, dtype=torch.float64)
    M[:, 7:] = torch.finfo(torch.float32).min
    dO = torch.randn(N, d, dtype=torch.float64)
    sc = d ** -0.5
    A = torch.softmax(sc * Q @ K.T + E + M, -1) * torch.sigmoid(G + M)
    (A @ V).backward(dO)
    dQ, dK, dV, dE, dG = O.gated_core_backward(Q.detach(), K.detach(), V.detach(), E.detach(), G.detach(), M, dO, sc)
    for a, b in [(dQ, Q.grad), (dK, K.grad), (dV, V.grad), (dE, E.grad), (dG, G.grad)]:
        assert (a - b).abs().max() < 1e-12


# ------------------------------------------------------------------ live reference (build container)
@pytest.mark.parametrize("kind", list(O.TRIPLET_FNS))
def test_triplet_oracle_vs_live_reference(reference_lib, kind):
    from lib.tgt.layers.triplet import get_triplet_layer
    from tgt_b200.harness.synthetic import make_edge_inputs
    torch.manual_seed(1)
    W, H = 64, 8
    mod = get_triplet_layer(kind)(W, H).double()
    e, mask = make_edge_inputs(2, 10, W, [10, 6], seed=9)
    e, mask = e.double(), mask.double()
    ref = mod(e, mask)
    out = O.TRIPLET_FNS[kind](_dbl(mod.state_dict()), e, mask, H)
    assert (out - ref).abs().max() < 1e-11


def test_layer_and_encoder_oracle_vs_live_reference(reference_lib):
    from lib.tgt import TGT_Encoder, Graph
    from tgt_b200.harness.synthetic import make_edge_inputs
    for node_ended, edge_ended, ttype in [(True, True, "attention"), (True, False, "aggregate"),
                                          (False, True, "attention")]:
        torch.manual_seed(2)
        cfg = dict(node_width=48, edge_width=32, num_heads=4, triplet_heads=2, triplet_type=ttype)
        enc = TGT_Encoder(model_height=3, layer_multiplier=2, node_ended=node_ended, edge_ended=edge_ended,
                          **cfg).double().eval()
        e, mask = make_edge_inputs(2, 9, 32, [9, 5], seed=3)
        h = torch.randn(2, 9, 48, dtype=torch.float64)
        g = enc(Graph(h=h.clone(), e=e.double().clone(), mask=mask.double()))
        cfg.pop("node_width"), cfg.pop("edge_width")
        ho, eo = O.encoder(_dbl(enc.state_dict()), h, e.double(), mask.double(), model_height=3,
                           layer_multiplier=2, node_ended=node_ended, edge_ended=edge_ended, **cfg)
        assert (ho - g.h).abs().max() < 1e-10
        assert (eo - g.e).abs().max() < 1e-10


def test_config1_known_answer(reference_lib):
    """Config 1 of BASELINE.json: TripletAttention(512, 8), B=4 N=16, seed 0 -- oracle vs stored checksums."""
    import json
    from lib.tgt.layers.triplet import TripletAttention
    from tgt_b200.harness.synthetic import make_edge_inputs
    kat = json.load(open(os.path.join(GOLDEN, "config1_kat.json")))
    torch.manual_seed(0)
    mod = TripletAttention(512, 8)
    e, mask = make_edge_inputs(4, 16, 512, [16, 12, 9, 5], seed=0)
    out = O.triplet_attention(_dbl(mod.state_dict()), e.double(), mask.double(), 8)
    assert abs(float(out.sum()) - kat["sum"]) < 1e-8 * max(1.0, abs(kat["abs_sum"]))
    assert abs(float(out.abs().sum()) - kat["abs_sum"]) < 1e-8 * kat["abs_sum"]


def test_bins_decode_oracle_vs_golden():
    """bins_from_logits / bins2dist against the fixture made by the reference's own predict_bins + BinsProcessor."""
    fx = load_golden("bins_decode.pt")
    S = fx["logits"].shape[1]
    bins = torch.stack([O.bins_from_logits(fx["logits"][:, s]) for s in range(S)], 1)
    assert torch.equal(bins.to(torch.int16), fx["bins"])
    d = O.bins2dist(bins, fx["num_bins"], fx["range_bins"])
    assert torch.equal(d, fx["dist"])
    assert torch.equal(d, d.transpose(-1, -2)) and float(d.diagonal(dim1=-2, dim2=-1).abs().max()) == 0.0


@pytest.mark.parametrize("kind", ["attention", "attention_ungated", "axial_attention"])
def test_core_oracle_matches_module_oracle(kind):
    """O.triplet_attention_core (the projection-level restatement the core-kernel tests use) composed with the
    LayerNorm / projection / lin_O of the module equals the module-level oracle, which is pinned to the reference
    above.  Also pins the host-side channel bookkeeping (head-major gather, Va -> lin_O column order)."""
    from tgt_b200 import layers as L
    from tgt_b200.layers.triplet import _head_major_perm, _out_perm
    torch.manual_seed(0)
    W, H, B, N = 32, 2, 2, 7
    d = W // H
    mod = L.get_triplet_layer(kind)(W, H).double()
    p = {k: v.detach() for k, v in mod.state_dict().items()}
    e = torch.randn(B, N, N, W, dtype=torch.float64)
    m = (torch.arange(N)[None, :] < torch.tensor([7, 4])[:, None]).double()
    mask = ((1 - m[:, :, None] * m[:, None, :]) * torch.finfo(torch.float32).min).unsqueeze(-1)
    ref = O.TRIPLET_FNS[kind](p, e, mask, H)
    x = torch.nn.functional.layer_norm(e, (W,), p["tri_ln_e.weight"], p["tri_ln_e.bias"])
    hm = _head_major_perm(W, H, "cpu")
    qp = torch.cat([hm, hm + W, hm + 2 * W])
    ws, bs = [p["lin_QKV_in.weight"][qp], p["lin_QKV_out.weight"][qp]], [p["lin_QKV_in.bias"][qp], p["lin_QKV_out.bias"][qp]]
    off_e, off_g, col = [-1, -1], [-1, -1], 6 * W
    for dirn, name in enumerate(mod._bias_names):
        if name is None:
            continue
        ws.append(p[name + ".weight"]); bs.append(p[name + ".bias"])
        off_e[dirn] = col
        if mod._gated:
            off_g[dirn] = col + H
        col += p[name + ".weight"].shape[0]
    proj = x @ torch.cat(ws).t() + torch.cat(bs)
    va = O.triplet_attention_core(proj, mask[..., 0], H, d, (0, 3 * W), (W, 4 * W), (2 * W, 5 * W), off_e, off_g)
    out = va @ p["lin_O.weight"][:, _out_perm(W, H, "cpu")].t() + p["lin_O.bias"]
    assert float((out - ref).abs().max()) < 1e-12
