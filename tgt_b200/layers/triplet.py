"""Triplet interaction modules -- the API surface of the reference's lib/tgt/layers/triplet.py.

Same constructor signatures, attribute names, parameter creation order (so seeded initialisation
matches) and state-dict keys / shapes as the reference; ``forward(e, mask) -> e`` runs our CUDA
kernels (ops.TripletAttentionFn / ops.TripletAggregateFn) instead of the O(N^3 H) einsum chain.

Channel bookkeeping: the reference views projections as [..., d, H] (channel c = dd*H + hh,
triplet.py:213) and concatenates the two directions on the head axis (c' = dd*2H + dir*H + hh,
triplet.py:248).  The kernels want head-major blocks (c = hh*d + dd), so the weight rows/columns are
gathered into kernel order on every call with differentiable index ops -- autograd scatters the
weight gradients back into the reference layout and the state dict never changes.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


def get_triplet_layer(layer_type):
    """Reference factory, triplet.py:6-20 (the 'tiangular_update' spelling is the reference's)."""
    table = {
        "aggregate": TripletAggregate,
        "aggregate_ungated": TripletAggregateUngated,
        "attention": TripletAttention,
        "attention_ungated": TripletAttentionUngated,
        "tiangular_update": TriangularUpdate,
        "axial_attention": AxialAttention,
    }
    if layer_type not in table:
        raise ValueError(f"Invalid layer_type: {layer_type}")
    return table[layer_type]


def _head_major_perm(width: int, heads: int, device) -> torch.Tensor:
    """perm[h*d + dd] = dd*H + h : gathers reference (d-major) channels into head-major order."""
    d = width // heads
    return torch.arange(width, device=device).view(d, heads).t().reshape(-1)


def _out_perm(width: int, heads: int, device) -> torch.Tensor:
    """Kernel Va channel (dir*W + h*d + dd)  ->  reference lin_O input column dd*2H + dir*H + h."""
    d = width // heads
    dirs = torch.arange(2, device=device).view(2, 1, 1)
    hh = torch.arange(heads, device=device).view(1, heads, 1)
    dd = torch.arange(d, device=device).view(1, 1, d)
    return (dd * 2 * heads + dirs * heads + hh).reshape(-1)


class _TripletBase(nn.Module):
    def __init__(self, edge_width, num_heads, attention_dropout=0):
        super().__init__()
        self.edge_width = edge_width
        self.num_heads = num_heads
        self.attention_dropout = attention_dropout

    def _common_checks(self, e, mask):
        if self.attention_dropout > 0 and self.training:
            # every shipped config has triplet_dropout = 0 (tgt_training.py:35); dropping entries of the
            # O(N^3) attention tensor would require materialising it -- refuse rather than fall back.
            raise NotImplementedError("tgt_b200: attention_dropout > 0 is not supported by the fused triplet kernels")
        if e.dim() != 4 or e.shape[1] != e.shape[2] or e.shape[3] != self.edge_width:
            raise ValueError(f"expected e of shape [B,N,N,{self.edge_width}], got {tuple(e.shape)}")


class _AttentionFamily(_TripletBase):
    """TripletAttention / TripletAttentionUngated / AxialAttention share one kernel."""
    _bias_names = (None, None)
    _gated = False

    def _build(self):
        W, H = self.edge_width, self.num_heads
        assert not (W % H), "edge_width must be divisible by num_heads"
        self._dot_dim = W // H
        self._scale_factor = self._dot_dim ** -0.5

    def forward(self, e, mask):
        return self.forward_alias(e, mask)[0]

    def forward_alias(self, e, mask):
        """(module(e), alias of e for the caller's residual add -- see ops.LNLinearFn)."""
        return self._run(e, mask, None, False)

    def forward_residual(self, e, mask, scale):
        """e + scale[b] * module(e): DropPath + residual add (reference layers.py:284-285) fused into the kernels."""
        return self._run(e, mask, scale, True)

    def _run(self, e, mask, scale, fuse_res):
        self._common_checks(e, mask)
        W, H, d = self.edge_width, self.num_heads, self._dot_dim
        dev = e.device
        hm = _head_major_perm(W, H, dev)
        qkv_perm = torch.cat([hm, hm + W, hm + 2 * W])
        w_parts = [self.lin_QKV_in.weight.index_select(0, qkv_perm), self.lin_QKV_out.weight.index_select(0, qkv_perm)]
        b_parts = [self.lin_QKV_in.bias.index_select(0, qkv_perm), self.lin_QKV_out.bias.index_select(0, qkv_perm)]
        off_q, off_k, off_v = (0, 3 * W), (W, 4 * W), (2 * W, 5 * W)
        off_e, off_g = [-1, -1], [-1, -1]
        col = 6 * W
        for dirn, name in enumerate(self._bias_names):
            if name is None:
                continue
            lin = getattr(self, name)
            w_parts.append(lin.weight)
            b_parts.append(lin.bias)
            off_e[dirn] = col
            if self._gated:
                off_g[dirn] = col + H
            col += lin.weight.shape[0]
        pad = (-col) % 8                                  # keep rows 16-byte aligned for the vector loads
        if pad:
            w_parts.append(self.lin_O.weight.new_zeros(pad, W))
            b_parts.append(self.lin_O.weight.new_zeros(pad))
        Wcat = torch.cat(w_parts, 0)
        bcat = torch.cat(b_parts, 0)
        Wo = self.lin_O.weight.index_select(1, _out_perm(W, H, dev))
        layout = (H, d, off_q, off_k, off_v, tuple(off_e), tuple(off_g))
        res = ops.TripletAttentionFn.apply(e, mask, self.tri_ln_e.weight, self.tri_ln_e.bias, Wcat, bcat, Wo,
                                           self.lin_O.bias, layout, ops.compute_dtype(e), scale, fuse_res)
        return ops.unpack_fused(res) if fuse_res else res


class TripletAttention(_AttentionFamily):
    """Reference: triplet.py:179-250 (TGT-At)."""
    _bias_names = ("lin_EG_in", "lin_EG_out")
    _gated = True

    def __init__(self, edge_width, num_heads, attention_dropout=0):
        super().__init__(edge_width, num_heads, attention_dropout)
        self._build()
        W, H = self.edge_width, self.num_heads
        self.tri_ln_e = nn.LayerNorm(W)
        self.lin_QKV_in = nn.Linear(W, W * 3)
        self.lin_EG_in = nn.Linear(W, H * 2)
        self.lin_QKV_out = nn.Linear(W, W * 3)
        self.lin_EG_out = nn.Linear(W, H * 2)
        self.lin_O = nn.Linear(W * 2, W)


class TripletAttentionUngated(_AttentionFamily):
    """Reference: triplet.py:253-322."""
    _bias_names = ("lin_E_in", "lin_E_out")
    _gated = False

    def __init__(self, edge_width, num_heads, attention_dropout=0):
        super().__init__(edge_width, num_heads, attention_dropout)
        self._build()
        W, H = self.edge_width, self.num_heads
        self.tri_ln_e = nn.LayerNorm(W)
        self.lin_QKV_in = nn.Linear(W, W * 3)
        self.lin_E_in = nn.Linear(W, H)
        self.lin_QKV_out = nn.Linear(W, W * 3)
        self.lin_E_out = nn.Linear(W, H)
        self.lin_O = nn.Linear(W * 2, W)


class AxialAttention(_AttentionFamily):
    """Reference: triplet.py:325-387."""
    _bias_names = (None, None)
    _gated = False

    def __init__(self, edge_width, num_heads, attention_dropout=0):
        super().__init__(edge_width, num_heads, attention_dropout)
        self._build()
        W = self.edge_width
        self.tri_ln_e = nn.LayerNorm(W)
        self.lin_QKV_in = nn.Linear(W, W * 3)
        self.lin_QKV_out = nn.Linear(W, W * 3)
        self.lin_O = nn.Linear(W * 2, W)


class _AggregateFamily(_TripletBase):
    _gated = True

    def _build(self):
        W, H = self.edge_width, self.num_heads
        assert not (W % H), "edge_width must be divisible by num_heads"
        self._dot_dim = W // H
        self._scale_factor = self._dot_dim ** -0.5

    def forward(self, e, mask):
        return self.forward_alias(e, mask)[0]

    def forward_alias(self, e, mask):
        return self._run(e, mask, None, False)

    def forward_residual(self, e, mask, scale):
        """e + scale[b] * module(e) (see _AttentionFamily.forward_residual)."""
        return self._run(e, mask, scale, True)

    def _run(self, e, mask, scale, fuse_res):
        self._common_checks(e, mask)
        W, H, d = self.edge_width, self.num_heads, self._dot_dim
        dev = e.device
        hm = _head_major_perm(W, H, dev)
        v_perm = torch.cat([hm, hm + W])
        lin_b = self.lin_EG if self._gated else self.lin_E
        w_parts = [self.lin_V.weight.index_select(0, v_perm), lin_b.weight]
        b_parts = [self.lin_V.bias.index_select(0, v_perm), lin_b.bias]
        col = 2 * W
        if self._gated:                       # chunk(4): E_in, G_in, E_out, G_out  (triplet.py:51)
            off_e, off_g = (col, col + 2 * H), (col + H, col + 3 * H)
            mask_dir = (1, 0)                 # outward direction is NOT masked (triplet.py:63-64)
            col += 4 * H
        else:                                 # chunk(2): E_in, E_out (triplet.py:105)
            off_e, off_g = (col, col + H), (-1, -1)
            mask_dir = (1, 1)
            col += 2 * H
        pad = (-col) % 8
        if pad:
            w_parts.append(self.lin_O.weight.new_zeros(pad, W))
            b_parts.append(self.lin_O.weight.new_zeros(pad))
        Wcat = torch.cat(w_parts, 0)
        bcat = torch.cat(b_parts, 0)
        Wo = self.lin_O.weight.index_select(1, _out_perm(W, H, dev))
        layout = (H, d, (0, W), off_e, off_g, mask_dir)
        res = ops.TripletAggregateFn.apply(e, mask, self.tri_ln_e.weight, self.tri_ln_e.bias, Wcat, bcat, Wo,
                                           self.lin_O.bias, layout, ops.compute_dtype(e), scale, fuse_res)
        return ops.unpack_fused(res) if fuse_res else res


class TripletAggregate(_AggregateFamily):
    """Reference: triplet.py:22-73 (TGT-Agx2)."""
    _gated = True

    def __init__(self, edge_width, num_heads, attention_dropout=0):
        super().__init__(edge_width, num_heads, attention_dropout)
        self._build()
        W, H = self.edge_width, self.num_heads
        self.tri_ln_e = nn.LayerNorm(W)
        self.lin_V = nn.Linear(W, W * 2)
        self.lin_EG = nn.Linear(W, H * 4)
        self.lin_O = nn.Linear(W * 2, W)


class TripletAggregateUngated(_AggregateFamily):
    """Reference: triplet.py:77-127."""
    _gated = False

    def __init__(self, edge_width, num_heads, attention_dropout=0):
        super().__init__(edge_width, num_heads, attention_dropout)
        self._build()
        W, H = self.edge_width, self.num_heads
        self.tri_ln_e = nn.LayerNorm(W)
        self.lin_V = nn.Linear(W, W * 2)
        self.lin_E = nn.Linear(W, H * 2)
        self.lin_O = nn.Linear(W * 2, W)


class TriangularUpdate(_TripletBase):
    """Reference: triplet.py:134-176.  LN -> [lin_V | lin_E] (one LayerNorm-folded GEMM) -> gated triangular
    contraction kernel (ops.TriangularCoreFn) -> lin_O -> sigmoid(gate)*lin kernel (ops.SigLinFn)."""

    def __init__(self, edge_width, num_heads, attention_dropout=0):
        super().__init__(edge_width, num_heads, attention_dropout)
        W, H = self.edge_width, self.num_heads
        self.tri_ln_e = nn.LayerNorm(W)
        self.lin_V = nn.Linear(W, H * 4)
        self.lin_E = nn.Linear(W, H * 4)
        self.lin_O = nn.Linear(H * 2, W * 2)

    def forward(self, e, mask):
        return self.forward_alias(e, mask)[0]

    def forward_alias(self, e, mask):
        """(module(e), alias of e for the caller's residual add -- see ops.LNLinearFn)."""
        self._common_checks(e, mask)       # the reference never applies attention_dropout in this variant either
        cd = ops.compute_dtype(e)
        Wcat = torch.cat([self.lin_V.weight, self.lin_E.weight], 0)
        bcat = torch.cat([self.lin_V.bias, self.lin_E.bias], 0)
        proj, alias = ops.LNLinearFn.apply(e, self.tri_ln_e.weight, self.tri_ln_e.bias, Wcat, bcat, cd)
        va = ops.TriangularCoreFn.apply(proj, mask, self.num_heads, cd)
        lo = nn.functional.linear(va, self.lin_O.weight.to(cd), self.lin_O.bias.to(cd))     # K = 2H: plain library GEMM
        return ops.SigLinFn.apply(lo), alias

    def forward_residual(self, e, mask, scale):
        """e + scale[b] * module(e) (see _AttentionFamily.forward_residual)."""
        out, alias = self.forward_alias(e, mask)
        return ops.scaled_residual(out, alias, scale)
