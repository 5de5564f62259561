"""EGT attention, edge update, FFN, DropPath and the TGT layer -- API surface of the reference's
lib/tgt/layers/layers.py with the hot ops running on tgt_b200's CUDA kernels.

State-dict keys / shapes, constructor signatures, sub-module names (update, node_ffn, tria, edge_ffn,
drop_path) and parameter creation order are the reference's (layers.py:15-260).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .activations import get_activation
from .triplet import get_triplet_layer


class EGT_Attention(nn.Module):
    """Reference: layers.py:15-84."""

    def __init__(self, node_width, edge_width, num_heads, source_dropout=0, scale_degree=True, edge_update=True):
        super().__init__()
        self.node_width = node_width
        self.edge_width = edge_width
        self.num_heads = num_heads
        self.source_dropout = source_dropout
        self.scale_degree = scale_degree
        self.edge_update = edge_update

        assert not (self.node_width % self.num_heads), "node_width must be divisible by num_heads"
        self._dot_dim = self.node_width // self.num_heads
        self._scale_factor = self._dot_dim ** -0.5

        self.mha_ln_h = nn.LayerNorm(self.node_width)
        self.mha_ln_e = nn.LayerNorm(self.edge_width)
        self.lin_QKV = nn.Linear(self.node_width, self.node_width * 3)
        self.lin_EG = nn.Linear(self.edge_width, self.num_heads * 2)
        self.lin_O_h = nn.Linear(self.node_width, self.node_width)
        if self.edge_update:
            self.lin_O_e = nn.Linear(self.num_heads, self.edge_width)

    def forward(self, h, e, mask):
        h, e, _ = self.forward_alias(h, e, mask)
        return h, e

    def forward_alias(self, h, e, mask):
        """forward() plus an alias of the input `e` to be used for the caller's residual add (its gradient is
        then folded into the LayerNorm-backward kernel, see ops.LNLinearFn)."""
        h, hhat, e_alias = self.forward_parts(h, e, mask)
        if self.edge_update:
            e = ops.lib_linear("egt.lin_O_e(unfused)", hhat, self.lin_O_e)
        return h, e, e_alias

    def edge_residual(self, hhat, e_alias, scale):
        """e + scale[b] * lin_O_e(H_hat): the edge output projection (layers.py:81-82) with the caller's DropPath +
        residual add (layers.py:278-279) as the epilogue of one GEMM."""
        return ops.unpack_fused(ops.LinearResidualFn.apply(hhat, self.lin_O_e.weight, self.lin_O_e.bias, e_alias,
                                                           scale, ops.compute_dtype(e_alias)))

    def forward_parts(self, h, e, mask):
        """(h_out, H_hat, alias of e): everything except the edge output projection."""
        cd = ops.compute_dtype(e)
        # node side: small [B*N, Wn] GEMMs -- plain library calls behind our LayerNorm kernels
        qkv = ops.LNLinearFn.apply(h, self.mha_ln_h.weight, self.mha_ln_h.bias, self.lin_QKV.weight,
                                   self.lin_QKV.bias, cd)[0]
        # edge side: LN + projection to 2H channels (LN output recomputed in backward)
        eg, e_alias = ops.LNLinearFn.apply(e, self.mha_ln_e.weight, self.mha_ln_e.bias, self.lin_EG.weight,
                                           self.lin_EG.bias, cd)
        src = None
        if self.source_dropout > 0 and self.training:       # layers.py:55-59; RNG stays in PyTorch
            src = torch.empty((h.shape[0], h.shape[1]), dtype=torch.float32, device=h.device) \
                       .bernoulli_(self.source_dropout) * torch.finfo(torch.float32).min
        hhat, vatt = ops.EGTCoreFn.apply(qkv, eg, mask, src, self.num_heads, True, self.scale_degree, cd)
        h = ops.lib_linear("egt.lin_O_h", vatt, self.lin_O_h)
        return h, hhat, e_alias


class EdgeUpdate(nn.Module):
    """Reference: layers.py:87-130."""

    def __init__(self, node_width, edge_width, num_heads):
        super().__init__()
        self.node_width = node_width
        self.edge_width = edge_width
        self.num_heads = num_heads

        assert not (self.node_width % self.num_heads), "node_width must be divisible by num_heads"
        self._dot_dim = self.node_width // self.num_heads
        self._scale_factor = self._dot_dim ** -0.5

        self.mha_ln_h = nn.LayerNorm(self.node_width)
        self.mha_ln_e = nn.LayerNorm(self.edge_width)
        self.lin_QK = nn.Linear(self.node_width, self.node_width * 2)
        self.lin_E = nn.Linear(self.edge_width, self.num_heads)
        self.lin_O_e = nn.Linear(self.num_heads, self.edge_width)

    def forward(self, h, e, mask):
        h, e, _ = self.forward_alias(h, e, mask)
        return h, e

    def forward_alias(self, h, e, mask):
        h, hhat, e_alias = self.forward_parts(h, e, mask)
        return h, ops.lib_linear("edge_update.lin_O_e(unfused)", hhat, self.lin_O_e), e_alias

    edge_residual = EGT_Attention.edge_residual

    def forward_parts(self, h, e, mask):
        cd = ops.compute_dtype(e)
        qk = ops.LNLinearFn.apply(h, self.mha_ln_h.weight, self.mha_ln_h.bias, self.lin_QK.weight,
                                  self.lin_QK.bias, cd)[0]
        eb, e_alias = ops.LNLinearFn.apply(e, self.mha_ln_e.weight, self.mha_ln_e.bias, self.lin_E.weight,
                                           self.lin_E.bias, cd)
        hhat = ops.EGTCoreFn.apply(qk, eb, mask, None, self.num_heads, False, False, cd)
        return h, hhat, e_alias


class FFN(nn.Module):
    """Reference: layers.py:134-160."""

    def __init__(self, width, multiplier=1., act_dropout=0., activation='gelu'):
        super().__init__()
        self.width = width
        self.multiplier = multiplier
        self.act_dropout = act_dropout
        self.activation = activation

        self.ffn_fn, self.act_mul = get_activation(activation)
        inner_dim = round(self.width * self.multiplier)

        self.ffn_ln = nn.LayerNorm(self.width)
        self.lin_W1 = nn.Linear(self.width, inner_dim * self.act_mul)
        self.lin_W2 = nn.Linear(inner_dim, self.width)
        self.dropout = nn.Dropout(self.act_dropout)

    def forward(self, x):
        return self.forward_alias(x)[0]

    def forward_alias(self, x):
        """(FFN(x), alias of x for the caller's residual add -- see ops.LNLinearFn)."""
        return self._run(x, None, False)

    def forward_residual(self, x, scale):
        """x + scale[b] * FFN(x): DropPath + residual add (reference layers.py:272-273, 289-290) fused in."""
        return self._run(x, scale, True)

    def _run(self, x, scale, fuse_res):
        if self.activation == 'gelu':
            p = self.act_dropout if self.training else 0.
            if p > 0 and x.is_cuda and torch.cuda.is_current_stream_capturing():
                # CUDA-graph capture (MC-dropout inference): the seed lives on the device and comes from torch's
                # graph-safe generator, so every replay of the graph draws a fresh dropout mask
                seed = torch.randint(0, 2 ** 62, (1,), device=x.device, dtype=torch.int64)
            else:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0      # CPU generator: no sync
            res = ops.FFNGeluFn.apply(x, self.ffn_ln.weight, self.ffn_ln.bias, self.lin_W1.weight,
                                      self.lin_W1.bias, self.lin_W2.weight, self.lin_W2.bias, p, seed,
                                      ops.compute_dtype(x), scale, fuse_res)
            return ops.unpack_fused(res) if fuse_res else res
        # gated activations (geglu/glu/swiglu) are not used by any shipped config: plain library ops on the
        # LayerNorm / GEMMs (still CUDA-only: the residual kernels around this module refuse CPU tensors)
        ops._require_cuda(x)
        x_ln = self.ffn_ln(x)
        y = self.ffn_fn(self.lin_W1(x_ln))
        y = self.dropout(y)
        if fuse_res:
            return ops.scaled_residual(self.lin_W2(y), x, scale)
        return self.lin_W2(y), x


class DropPath(nn.Module):
    """Reference: layers.py:163-177 (per-sample stochastic depth)."""

    def __init__(self, drop_path=0.):
        super().__init__()
        self.drop_path = drop_path
        self._keep_prob = 1 - self.drop_path

    def sample_scale(self, x):
        """[B] fp32 vector mask/keep_prob, or None when inactive."""
        if self.drop_path > 0 and self.training:
            m = torch.empty(x.size(0), dtype=torch.float32, device=x.device).bernoulli_(self._keep_prob)
            return m / self._keep_prob
        return None

    def forward(self, x):
        s = self.sample_scale(x)
        if s is not None:
            x = x * s.view(-1, *([1] * (x.ndim - 1))).to(x.dtype)
        return x

    def __repr__(self):
        return f'{self.__class__.__name__}(drop_path={self.drop_path})'


class TGT_Layer(nn.Module):
    """Reference: layers.py:180-302."""

    def __init__(self, node_width, edge_width, num_heads, activation='gelu', scale_degree=True,
                 node_update=True, edge_update=True, triplet_heads=0, triplet_type='aggregate',
                 triplet_dropout=0, node_ffn_multiplier=1., edge_ffn_multiplier=1., source_dropout=0,
                 drop_path=0, node_act_dropout=0, edge_act_dropout=0):
        super().__init__()
        self.node_width = node_width
        self.edge_width = edge_width
        self.num_heads = num_heads
        self.activation = activation
        self.node_ffn_multiplier = node_ffn_multiplier
        self.edge_ffn_multiplier = edge_ffn_multiplier
        self.node_act_dropout = node_act_dropout
        self.edge_act_dropout = edge_act_dropout
        self.source_dropout = source_dropout
        self.drop_path = drop_path
        self.scale_degree = scale_degree
        self.node_update = node_update
        self.edge_update = edge_update
        self.triplet_heads = triplet_heads
        self.triplet_type = triplet_type
        self.triplet_dropout = triplet_dropout

        self._triplet_update = self.triplet_heads > 0

        if self.node_update:
            self.update = EGT_Attention(node_width=self.node_width, edge_width=self.edge_width,
                                        num_heads=self.num_heads, source_dropout=self.source_dropout,
                                        scale_degree=self.scale_degree, edge_update=self.edge_update)
        elif self.edge_update:
            self.update = EdgeUpdate(node_width=self.node_width, edge_width=self.edge_width,
                                     num_heads=self.num_heads)
        else:
            raise ValueError('At least one of node_update and edge_update must be True')

        if self.node_update:
            self.node_ffn = FFN(width=self.node_width, multiplier=self.node_ffn_multiplier,
                                act_dropout=self.node_act_dropout, activation=self.activation)
        if self.edge_update:
            if self._triplet_update:
                TripletLayer = get_triplet_layer(self.triplet_type)
                self.tria = TripletLayer(edge_width=self.edge_width, num_heads=self.triplet_heads,
                                         attention_dropout=self.triplet_dropout)
            self.edge_ffn = FFN(width=self.edge_width, multiplier=self.edge_ffn_multiplier,
                                act_dropout=self.edge_act_dropout, activation=self.activation)

        self.drop_path = DropPath(self.drop_path)

    def _residual(self, x, res):
        """drop_path(x) then add the residual (layers.py:269-290), as one fused kernel."""
        return ops.scaled_residual(x, res, self.drop_path.sample_scale(x))

    def forward(self, g):
        h, e, mask = g.h, g.e, g.mask

        # every sub-module also hands back an alias of its edge/node input; adding the residual to THAT routes the
        # residual gradient into the sub-module's LayerNorm-backward kernel (ops.LNLinearFn) -- same values as the
        # reference's `x.add_(x_r)` (layers.py:269-290)
        h_r1 = h
        h, hhat, e_r1 = self.update.forward_parts(h, e, mask)

        if self.node_update:
            h = self._residual(h, h_r1)
            h = self.node_ffn.forward_residual(h, self.drop_path.sample_scale(h))

        if self.edge_update:
            # each edge branch is ONE fused pass: x + drop_path(f(x)) comes out of the last GEMM's epilogue
            e = self.update.edge_residual(hhat, e_r1, self.drop_path.sample_scale(e))
            if self._triplet_update:
                e = self.tria.forward_residual(e, mask, self.drop_path.sample_scale(e))
            e = self.edge_ffn.forward_residual(e, self.drop_path.sample_scale(e))

        g = g.copy()
        g.h, g.e = h, e
        return g

    def __repr__(self):
        rep = super().__repr__()
        return rep + f' (activation: {self.activation}, source_dropout: {self.source_dropout})'
