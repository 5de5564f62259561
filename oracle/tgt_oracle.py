"""CPU oracle for the TGT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict driven) PyTorch restatement of the arithmetic of the
reference's ``lib/tgt`` package, written from the index-level specification in
SURVEY.md Appendix A.  It materialises the O(N^3 H) tensors exactly like the
reference does, so it doubles as the "reference CPU path" that ``bench.py``
times (``cpu_baseline.kind == "port"``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` leg may import this module.  ``tgt_b200`` (the product)
never does: it fails loudly when its CUDA library is missing.

Parity pinning: ``tests/test_oracle_pin.py`` checks every function here
 (a) against the real reference imported from ``/root/reference`` when that
     tree is present (this container), and
 (b) against the golden vectors in ``tests/golden/*.pt`` which were produced by
     the real reference through ``oracle/make_golden.py`` (committed).
The reference ships no golden vectors / tests of its own.

Reference citations (relative to /root/reference):
  triplet_attention    lib/tgt/layers/triplet.py:205-250
  triplet_aggregate    lib/tgt/layers/triplet.py:45-73
  triplet_*ungated/axial/triangular  lib/tgt/layers/triplet.py:77-387
  egt_attention        lib/tgt/layers/layers.py:46-84 (+ 8-12)
  edge_update          lib/tgt/layers/layers.py:110-130
  ffn                  lib/tgt/layers/layers.py:155-160
  tgt_layer            lib/tgt/layers/layers.py:262-294
  encoder              lib/tgt/encoder.py:52-90
  embed_input / heads  lib/models/pcqm/layers.py:62-83, multitask.py:53-68,
                       gap_predictor.py:46-60, distance_predictor.py:46-54
  losses               lib/training_schemes/pcqm/commons.py:6-48,
                       pretrain/scheme.py:78-88
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def sub(p: Params, prefix: str) -> Params:
    """View of a state dict restricted to `prefix.` with the prefix removed."""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix + ".")}


def _lin(p: Params, name: str, x: Tensor) -> Tensor:
    return F.linear(x, p[name + ".weight"], p.get(name + ".bias"))


def _ln(p: Params, name: str, x: Tensor) -> Tensor:
    w = p[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, p[name + ".bias"], 1e-5)


def _heads(t: Tensor, H: int) -> Tensor:
    """[..., d*H] -> [..., d, H]   (channel c = dd*H + hh; triplet.py:213, layers.py:62)."""
    return t.reshape(*t.shape[:-1], t.shape[-1] // H, H)


# --------------------------------------------------------------------------
# triplet interaction modules  (forward(e, mask) -> e_delta)
# --------------------------------------------------------------------------
def _pair_attention(x: Tensor, mask: Tensor, p: Params, H: int, qkv: str,
                    bias: Optional[str], gated: bool, outward: bool) -> Tensor:
    """One direction of (gated / ungated / axial) triplet attention.

    inward  (triplet.py:210-227):  S[b,i,j,k,h] = <Q[b,i,j],K[b,j,k]>/sqrt(d) + E[b,i,k,h]
                                   softmax over k, mask[b,i,k], Va = sum_k A V[b,j,k]
    outward (triplet.py:229-246):  S[b,k,i,j,h] = <Q[b,i,j],K[b,k,j]>/sqrt(d) + E[b,k,i,h]
                                   softmax over k (dim 1), mask[b,k,i], Va = sum_k A V[b,k,j]
    """
    W = x.shape[-1]
    d = W // H
    Q, K, V = (_heads(t, H) for t in _lin(p, qkv, x).chunk(3, dim=-1))
    Q = Q * d ** -0.5
    if outward:
        S = torch.einsum("bijdh,bkjdh->bkijh", Q, K)
        m = mask.unsqueeze(3)            # [B,k,i,1,1]
        ax, un = 1, 3
    else:
        S = torch.einsum("bijdh,bjkdh->bijkh", Q, K)
        m = mask.unsqueeze(2)            # [B,i,1,k,1]
        ax, un = 3, 2
    G = None
    if bias is not None:
        EG = _lin(p, bias, x).unsqueeze(un)
        if gated:
            E, G = EG.chunk(2, dim=-1)
        else:
            E = EG
        S = S + E
    A = torch.softmax(S + m, dim=ax)
    if G is not None:
        A = A * torch.sigmoid(G + m)
    if outward:
        return torch.einsum("bkijh,bkjdh->bijdh", A, V)
    return torch.einsum("bijkh,bjkdh->bijdh", A, V)


def _tri_attention_family(p: Params, e: Tensor, mask: Tensor, H: int,
                          bias_in, bias_out, gated: bool) -> Tensor:
    B, N, _, W = e.shape
    x = _ln(p, "tri_ln_e", e)
    va_in = _pair_attention(x, mask, p, H, "lin_QKV_in", bias_in, gated, False)
    va_out = _pair_attention(x, mask, p, H, "lin_QKV_out", bias_out, gated, True)
    va = torch.cat([va_in, va_out], dim=-1).reshape(B, N, N, 2 * W)   # c' = dd*2H + dir*H + hh
    return _lin(p, "lin_O", va)


def triplet_attention(p: Params, e: Tensor, mask: Tensor, num_heads: int) -> Tensor:
    """TripletAttention.forward -- triplet.py:205-250."""
    return _tri_attention_family(p, e, mask, num_heads, "lin_EG_in", "lin_EG_out", True)


def triplet_attention_ungated(p: Params, e: Tensor, mask: Tensor, num_heads: int) -> Tensor:
    """TripletAttentionUngated.forward -- triplet.py:276-322."""
    return _tri_attention_family(p, e, mask, num_heads, "lin_E_in", "lin_E_out", False)


def axial_attention(p: Params, e: Tensor, mask: Tensor, num_heads: int) -> Tensor:
    """AxialAttention.forward -- triplet.py:345-387."""
    return _tri_attention_family(p, e, mask, num_heads, None, None, False)


def triplet_aggregate(p: Params, e: Tensor, mask: Tensor, num_heads: int) -> Tensor:
    """TripletAggregate.forward -- triplet.py:45-73.

    NOTE the outward weights carry NO mask term (triplet.py:63-64); parity must
    hold on the whole padded tensor.
    """
    B, N, _, W = e.shape
    H = num_heads
    x = _ln(p, "tri_ln_e", e)
    V_in, V_out = (_heads(t, H) for t in _lin(p, "lin_V", x).chunk(2, dim=-1))
    E_in, G_in, E_out, G_out = _lin(p, "lin_EG", x).chunk(4, dim=-1)
    A_in = torch.softmax(E_in + mask, dim=2) * torch.sigmoid(G_in + mask)
    A_out = torch.softmax(E_out, dim=1) * torch.sigmoid(G_out)
    va_in = torch.einsum("bikh,bjkdh->bijdh", A_in, V_in)
    va_out = torch.einsum("bkih,bkjdh->bijdh", A_out, V_out)
    va = torch.cat([va_in, va_out], dim=-1).reshape(B, N, N, 2 * W)
    return _lin(p, "lin_O", va)


def triplet_aggregate_ungated(p: Params, e: Tensor, mask: Tensor, num_heads: int) -> Tensor:
    """TripletAggregateUngated.forward -- triplet.py:99-127 (both directions masked)."""
    B, N, _, W = e.shape
    H = num_heads
    x = _ln(p, "tri_ln_e", e)
    V_in, V_out = (_heads(t, H) for t in _lin(p, "lin_V", x).chunk(2, dim=-1))
    E_in, E_out = _lin(p, "lin_E", x).chunk(2, dim=-1)
    A_in = torch.softmax(E_in + mask, dim=2)
    A_out = torch.softmax(E_out + mask, dim=1)
    va_in = torch.einsum("bikh,bjkdh->bijdh", A_in, V_in)
    va_out = torch.einsum("bkih,bkjdh->bijdh", A_out, V_out)
    va = torch.cat([va_in, va_out], dim=-1).reshape(B, N, N, 2 * W)
    return _lin(p, "lin_O", va)


def triangular_update(p: Params, e: Tensor, mask: Tensor, num_heads: int) -> Tensor:
    """TriangularUpdate.forward -- triplet.py:150-176 (sigmoid(gate)*lin on every branch)."""
    x = _ln(p, "tri_ln_e", e)
    vg_i, vl_i, vg_o, vl_o = _lin(p, "lin_V", x).chunk(4, dim=-1)
    eg_i, el_i, eg_o, el_o = _lin(p, "lin_E", x).chunk(4, dim=-1)
    V_in = torch.sigmoid(vg_i + mask) * vl_i
    V_out = torch.sigmoid(vg_o + mask) * vl_o
    E_in = torch.sigmoid(eg_i + mask) * el_i
    E_out = torch.sigmoid(eg_o + mask) * el_o
    va_in = torch.einsum("bikh,bjkh->bijh", E_in, V_in)
    va_out = torch.einsum("bkih,bkjh->bijh", E_out, V_out)
    og, ol = _lin(p, "lin_O", torch.cat([va_in, va_out], dim=-1)).chunk(2, dim=-1)
    return torch.sigmoid(og) * ol


TRIPLET_FNS = {
    "attention": triplet_attention,
    "attention_ungated": triplet_attention_ungated,
    "axial_attention": axial_attention,
    "aggregate": triplet_aggregate,
    "aggregate_ungated": triplet_aggregate_ungated,
    "tiangular_update": triangular_update,      # (sic) the typo is the reference's API, triplet.py:16
}


# --------------------------------------------------------------------------
# EGT node/edge attention
# --------------------------------------------------------------------------
def egt_attention(p: Params, h: Tensor, e: Tensor, mask: Tensor, num_heads: int,
                  scale_degree: bool = True, edge_update: bool = True,
                  source_mask: Optional[Tensor] = None):
    """EGT_Attention.forward -- layers.py:46-84.

    `source_mask` ([B,1,N,1], additive) stands for the random column mask the
    reference draws at layers.py:55-59; the caller owns the RNG.
    """
    B, N, Wn = h.shape
    H = num_heads
    d = Wn // H
    Q, K, V = (_heads(t, H) for t in _lin(p, "lin_QKV", _ln(p, "mha_ln_h", h)).chunk(3, dim=-1))
    E, G = _lin(p, "lin_EG", _ln(p, "mha_ln_e", e)).chunk(2, dim=-1)
    if source_mask is not None:
        mask = mask + source_mask
    gates = torch.sigmoid(G + mask)
    H_hat = torch.einsum("bldh,bmdh->blmh", Q * d ** -0.5, K) + E
    A = torch.softmax(H_hat + mask, dim=2) * gates
    V_att = torch.einsum("blmh,bmkh->blkh", A, V)
    if scale_degree:
        V_att = V_att * torch.log(1 + gates.sum(dim=2, keepdim=True))     # layers.py:8-12
    h_out = _lin(p, "lin_O_h", V_att.reshape(B, N, Wn))
    e_out = _lin(p, "lin_O_e", H_hat) if edge_update else e
    return h_out, e_out


def edge_update(p: Params, h: Tensor, e: Tensor, mask: Tensor, num_heads: int):
    """EdgeUpdate.forward -- layers.py:110-130 (returns h unchanged)."""
    B, N, Wn = h.shape
    H = num_heads
    d = Wn // H
    Q, K = (_heads(t, H) for t in _lin(p, "lin_QK", _ln(p, "mha_ln_h", h)).chunk(2, dim=-1))
    E = _lin(p, "lin_E", _ln(p, "mha_ln_e", e))
    H_hat = torch.einsum("bldh,bmdh->blmh", Q * d ** -0.5, K) + E
    return h, _lin(p, "lin_O_e", H_hat)


# --------------------------------------------------------------------------
# FFN, layer, encoder
# --------------------------------------------------------------------------
def _act(name: str, x: Tensor) -> Tensor:
    """activations.py:4-25."""
    if name in ("geglu", "glu", "swiglu"):
        g, v = x.chunk(2, dim=-1)
        if name == "geglu":
            return v * F.gelu(g)
        if name == "glu":
            return v * torch.sigmoid(g)
        return v * torch.sigmoid(g) * g
    return getattr(F, name)(x)


def ffn(p: Params, x: Tensor, activation: str = "gelu",
        drop_mask: Optional[Tensor] = None) -> Tensor:
    """FFN.forward -- layers.py:155-160.  `drop_mask` (already scaled by 1/keep)
    stands for nn.Dropout's random mask."""
    y = _act(activation, _lin(p, "lin_W1", _ln(p, "ffn_ln", x)))
    if drop_mask is not None:
        y = y * drop_mask
    return _lin(p, "lin_W2", y)


def tgt_layer(p: Params, h: Tensor, e: Tensor, mask: Tensor, *, num_heads: int,
              triplet_heads: int = 0, triplet_type: str = "aggregate",
              activation: str = "gelu", scale_degree: bool = True,
              node_update: bool = True, edge_update_: bool = True):
    """TGT_Layer.forward in eval mode -- layers.py:262-294 (drop_path/dropout are
    identities when not training)."""
    if node_update:
        dh, de = egt_attention(sub(p, "update"), h, e, mask, num_heads,
                               scale_degree=scale_degree, edge_update=edge_update_)
    elif edge_update_:
        dh, de = edge_update(sub(p, "update"), h, e, mask, num_heads)
    else:
        raise ValueError("At least one of node_update and edge_update must be True")
    if node_update:
        h = _residual(dh, h)
        h = _residual(ffn(sub(p, "node_ffn"), h, activation), h)
    if edge_update_:
        e = _residual(de, e)
        if triplet_heads > 0:
            e = _residual(TRIPLET_FNS[triplet_type](sub(p, "tria"), e, mask, triplet_heads), e)
        e = _residual(ffn(sub(p, "edge_ffn"), e, activation), e)
    return h, e


def _residual(x: Tensor, res: Tensor) -> Tensor:
    """`x.add_(res)` of layers.py:270-290: the sum is formed in the promoted type and stored in x's dtype.  In a
    uniform-precision run this is x + res; under autocast x is the 16-bit Linear output and res the (initially fp32)
    stream, so the residual stream is 16-bit from the first add onwards -- kept here so that the
    oracle run under CUDA autocast is the reference's arithmetic, not a more accurate one."""
    return (x + res).to(x.dtype)


def encoder(p: Params, h: Tensor, e: Tensor, mask: Tensor, *, model_height: int,
            layer_multiplier: int = 1, node_ended: bool = True, edge_ended: bool = True,
            egt_simple: bool = False, **layer_cfg):
    """TGT_Encoder.forward in eval mode -- encoder.py:52-90."""
    for i in range(model_height):
        last = i == model_height - 1
        nu = not (last and not node_ended)
        eu = False if egt_simple else not (last and not edge_ended)
        lp = sub(p, f"TGT_layers.{i}")
        for _ in range(layer_multiplier):
            h, e = tgt_layer(lp, h, e, mask, node_update=nu, edge_update_=eu, **layer_cfg)
    return h, e


# --------------------------------------------------------------------------
# task models (input embedding + heads) and losses -- needed only so that the
# CPU baseline can time the same whole-model step as bench.py's GPU arm.
# --------------------------------------------------------------------------
NODE_FEATURES_OFFSET = 128      # lib/models/pcqm/consts.py:1-7
NUM_NODE_FEATURES = 9
EDGE_FEATURES_OFFSET = 8
NUM_EDGE_FEATURES = 3
HL_MEAN = 5.6894608
HL_STD = 1.1621397


def coords2dist(x: Tensor) -> Tensor:
    """commons.py:6-8."""
    return torch.norm(x.unsqueeze(-2) - x.unsqueeze(-3), dim=-1)


def embed_input(p: Params, batch: Dict[str, Tensor], upto_hop: int = 32):
    """EmbedInput.forward with the gaussian 3-D embedding -- models/pcqm/layers.py:62-83, 112-157."""
    nf = batch["node_features"].long()
    h = F.embedding(nf, p["nodef_embed.weight"], padding_idx=0).sum(dim=2)
    dm = batch["distance_matrix"].long().clamp(max=upto_hop + 1)
    fm = batch["feature_matrix"].long()
    e = F.embedding(dm, p["dist_embed.weight"]) \
        + F.embedding(fm, p["featm_embed.weight"], padding_idx=0).sum(dim=-2)
    if "m3d_embed.gbf.means.weight" in p:
        N = nf.shape[1]
        ti = nf[:, :, 0]
        tj = ti + NODE_FEATURES_OFFSET
        types = torch.stack([ti.unsqueeze(2).expand(-1, -1, N),
                             tj.unsqueeze(1).expand(-1, N, -1)], dim=-1)
        mul = F.embedding(types, p["m3d_embed.gbf.mul.weight"], padding_idx=0).sum(dim=-2)
        bia = F.embedding(types, p["m3d_embed.gbf.bias.weight"], padding_idx=0).sum(dim=-2)
        xk = mul * batch["dist_input"].unsqueeze(-1) + bia
        mean = p["m3d_embed.gbf.means.weight"].float().view(-1)
        std = p["m3d_embed.gbf.stds.weight"].float().view(-1).abs() + 1e-2
        a = (2 * 3.14159) ** 0.5
        gb = torch.exp(-0.5 * (((xk.float() - mean) / std) ** 2)) / (a * std)
        gb = gb.type_as(p["m3d_embed.gbf.means.weight"])
        y = F.gelu(_lin(p, "m3d_embed.gbf_proj.layer1", gb))
        e = e + _lin(p, "m3d_embed.gbf_proj.layer2", y)
    em = batch["edge_mask"].unsqueeze(-1).to(e.dtype)
    mask = (1 - em) * torch.finfo(e.dtype).min
    return h, e, mask


def _pool_head(p: Params, h: Tensor, node_mask: Tensor) -> Tensor:
    h = _ln(p, "final_ln_node", h)
    m = node_mask.float().unsqueeze(-1)
    h = (h * m).sum(dim=1) / (m.sum(dim=1) + 1e-9)
    return _lin(p, "pred", h).squeeze(-1)


def tgt_multi(p: Params, batch: Dict[str, Tensor], *, model_height: int, upto_hop: int = 32, **cfg):
    """TGT_Multi.forward (eval mode) -- multitask.py:53-68."""
    h, e, mask = embed_input(sub(p, "input_embed"), batch, upto_hop)
    h, e = encoder(sub(p, "encoder"), h, e, mask, model_height=model_height,
                   node_ended=True, edge_ended=True, **cfg)
    gap = _pool_head(p, h, batch["node_mask"])
    logits = _lin(p, "dist_pred", _ln(p, "final_ln_edge", e))
    return gap, logits


def tgt_gap(p: Params, batch: Dict[str, Tensor], *, model_height: int, upto_hop: int = 32, **cfg):
    """TGT_Gap.forward (eval mode) -- gap_predictor.py:46-60."""
    h, e, mask = embed_input(sub(p, "input_embed"), batch, upto_hop)
    h, e = encoder(sub(p, "encoder"), h, e, mask, model_height=model_height,
                   node_ended=True, edge_ended=False, **cfg)
    return _pool_head(p, h, batch["node_mask"])


def tgt_distance(p: Params, batch: Dict[str, Tensor], *, model_height: int, upto_hop: int = 32, **cfg):
    """TGT_Distance.forward (eval mode) -- distance_predictor.py:46-54."""
    h, e, mask = embed_input(sub(p, "input_embed"), batch, upto_hop)
    h, e = encoder(sub(p, "encoder"), h, e, mask, model_height=model_height,
                   node_ended=False, edge_ended=True, **cfg)
    return _lin(p, "dist_pred", _ln(p, "final_ln_edge", e))


def discrete_dist_loss(logits: Tensor, dist_targ: Tensor, edge_mask: Tensor,
                       num_bins: int, range_bins: float) -> Tensor:
    """DiscreteDistLoss.__call__(reduce=True) -- commons.py:19-48."""
    t = (dist_targ * ((num_bins - 1) / range_bins)).long().clamp(0, num_bins - 1)
    xent = F.cross_entropy(logits.reshape(-1, num_bins), t.reshape(-1), reduction="none")
    B = logits.shape[0]
    xent = xent.view(B, -1)
    m = edge_mask.to(xent.dtype).view(B, -1)
    return (xent * m).sum() / (m.sum() + 1e-9)


def pretrain_loss(gap: Tensor, logits: Tensor, batch: Dict[str, Tensor], num_bins: int,
                  range_bins: float = 8.0, dist_loss_weight: float = 0.1) -> Tensor:
    """pretrain/scheme.py:78-88."""
    prim = F.l1_loss(gap, batch["target"].to(gap.dtype))
    dl = discrete_dist_loss(logits, coords2dist(batch["dft_coords"]), batch["edge_mask"],
                            num_bins, range_bins)
    return prim + dist_loss_weight * dl


# --------------------------------------------------------------------------
# two-stage inference glue (BASELINE config 5)
# --------------------------------------------------------------------------
def bins_from_logits(logits: Tensor) -> Tensor:
    """One sample of predict_bins -- lib/training_schemes/pcqm/dist_pred/scheme.py:190-193:
    softmax over bins, symmetrise over the atom pair (p + p^T), argmax."""
    p = torch.softmax(logits.float(), dim=-1)
    p = p + p.transpose(-2, -3)
    return p.argmax(dim=-1)


def bins2dist(bins: Tensor, num_bins: int, range_bins: float = 8.0, shift_half: bool = True,
              zero_diag: bool = True) -> Tensor:
    """BinsProcessor.bins2dist -- lib/training_schemes/pcqm/commons.py:68-82 (bin_size = range / (num_bins - 1))."""
    d = bins.float()
    if shift_half:
        d = d + 0.5
    d = d * (range_bins / (num_bins - 1))
    d = d + d.transpose(-2, -1)
    if zero_diag:
        d = d * (1 - torch.eye(d.size(-1), dtype=d.dtype, device=d.device))
    return d


# --------------------------------------------------------------------------
# the attention CORE alone, on a given projection buffer (what the CUDA core kernels see)
# --------------------------------------------------------------------------
def triplet_attention_core(proj: Tensor, mask: Tensor, H: int, d: int, off_q, off_k, off_v, off_e, off_g,
                           scale: Optional[float] = None, round_a: Optional[torch.dtype] = None) -> Tensor:
    """triplet.py:213-227 (inward) and :232-246 (outward) on an explicit projection `proj` [B,N,N,C] whose column blocks
    are laid out the way the kernels want them (include/tgt_b200.h: HEAD-major q/k/v blocks of H*d columns at
    off_q/off_k/off_v[dir], H-wide bias / gate blocks at off_e/off_g[dir], -1 = absent).  mask: [B,N,N] additive.
    Returns Va [B,N,N,2*H*d] with channel = dir*H*d + h*d + dd.  Differentiable: autograd of this function is the
    oracle of the core backward kernels.
    round_a: round the attention weights A to this 16-bit dtype before the A.V contraction -- the ONE rounding a 16-bit
    tensor-core implementation cannot avoid (the reference's own autocast path casts A to 16 bit for this einsum,
    triplet.py:227 / SURVEY App. A "Autocast dtype flow"); used as the attainable-accuracy yardstick in the tests."""
    B, N = proj.shape[0], proj.shape[1]
    scale = d ** -0.5 if scale is None else scale
    outs = []
    for dirn in range(2):
        blk = lambda o: proj[..., o[dirn]:o[dirn] + H * d].reshape(B, N, N, H, d)
        Q, K, V = blk(off_q), blk(off_k), blk(off_v)
        if dirn == 0:      # S[b,i,j,k,h] = <Q[b,i,j], K[b,j,k]> + E[b,i,k,h] + M[b,i,k]
            S = scale * torch.einsum("bijhd,bjkhd->bijkh", Q, K)
            tile = lambda t: t.unsqueeze(2)                          # [B,i,k,.] -> [B,i,1,k,.]
        else:              # S[b,i,j,k,h] = <Q[b,i,j], K[b,k,j]> + E[b,k,i,h] + M[b,k,i]
            S = scale * torch.einsum("bijhd,bkjhd->bijkh", Q, K)
            tile = lambda t: t.transpose(1, 2).unsqueeze(2)          # [B,k,i,.] -> [B,i,1,k,.]
        M = tile(mask.unsqueeze(-1))
        if off_e[dirn] >= 0:
            S = S + tile(proj[..., off_e[dirn]:off_e[dirn] + H])
        A = torch.softmax(S + M, dim=3)
        if off_g[dirn] >= 0:
            A = A * torch.sigmoid(tile(proj[..., off_g[dirn]:off_g[dirn] + H]) + M)
        if round_a is not None:
            A = A.to(round_a).to(proj.dtype)
        if dirn == 0:
            Va = torch.einsum("bijkh,bjkhd->bijhd", A, V)
        else:
            Va = torch.einsum("bijkh,bkjhd->bijhd", A, V)
        outs.append(Va.reshape(B, N, N, H * d))
    return torch.cat(outs, dim=-1)


# --------------------------------------------------------------------------
# closed-form backward of the gated attention core (second oracle for the CUDA
# backward; SURVEY.md Appendix A, verified against autograd in the tests)
# --------------------------------------------------------------------------
def gated_core_backward(Qt, Kt, Vt, E, G, M, dO, scale):
    """Single (b, j, h) problem.  Qt,Kt,Vt: [N,d]; E,G,M: [N,N] indexed (query i, key k);
    returns dQ, dK, dV, dE_contrib, dG_contrib (the last two are summed over j by the caller)."""
    S = scale * Qt @ Kt.T + E + M
    P = torch.softmax(S, dim=-1)
    g = torch.sigmoid(G + M)
    A = P * g
    dA = dO @ Vt.T
    dV = A.T @ dO
    dG = g * (1 - g) * dA * P
    dP = dA * g
    delta = (dP * P).sum(-1, keepdim=True)
    dS = P * (dP - delta)
    return scale * dS @ Kt, scale * dS.T @ Qt, dV, dS, dG
