"""Task models for the bench harness: input embedding + TGT encoder + heads.

Functionally equivalent to (and state-dict compatible with) the reference's lib/models/pcqm
{layers.py:11-173, multitask.py:10-68, gap_predictor.py:10-60, distance_predictor.py:9-54}; written as
plain PyTorch because these O(N^2) pieces are outside the hot path ("reused unchanged").
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import TGT_Encoder, Graph, ops
from .synthetic import (NODE_FEATURES_OFFSET, NUM_NODE_FEATURES, EDGE_FEATURES_OFFSET, NUM_EDGE_FEATURES)

HL_MEAN = 5.6894608


class _GaussianBasis(nn.Module):
    def __init__(self, K, edge_types):
        super().__init__()
        self.K = K
        self.means = nn.Embedding(1, K)
        self.stds = nn.Embedding(1, K)
        self.mul = nn.Embedding(edge_types, 1, padding_idx=0)
        self.bias = nn.Embedding(edge_types, 1, padding_idx=0)
        nn.init.uniform_(self.means.weight, 0, 3)
        nn.init.uniform_(self.stds.weight, 0, 3)
        nn.init.constant_(self.bias.weight, 0)
        nn.init.constant_(self.mul.weight, 1)

    def forward(self, dist, types):
        if isinstance(types, tuple):
            # (row types [B,N], column types [B,N]) of the pair-type tensor types[b,i,j] = (row[b,i], col[b,j]): the two
            # lookups per pair become two lookups per NODE plus a broadcast add -- same values, and the embedding
            # backward sees B*N indices instead of sorting 2*B*N^2 of them
            row, col = types
            scale = (self.mul(row).unsqueeze(2) + self.mul(col).unsqueeze(1))
            shift = (self.bias(row).unsqueeze(2) + self.bias(col).unsqueeze(1))
        else:
            scale = self.mul(types).sum(dim=-2)
            shift = self.bias(types).sum(dim=-2)
        mu = self.means.weight.float().view(-1)
        sd = self.stds.weight.float().view(-1).abs() + 1e-2
        if dist.is_cuda and self.K % 4 == 0 and self.K <= 128:
            # one kernel each way instead of ~6 element-wise passes over [B,N,N,K]; under autocast the basis is written
            # directly in the dtype the following Linear would cast it to
            x = (scale * dist.unsqueeze(-1) + shift).squeeze(-1).float()
            odt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else self.means.weight.dtype
            return ops.GaussianBasisFn.apply(x, mu, sd, odt)
        x = (scale * dist.unsqueeze(-1) + shift).expand(-1, -1, -1, self.K).float()
        norm = (2 * 3.14159) ** 0.5
        return (torch.exp(-0.5 * ((x - mu) / sd) ** 2) / (norm * sd)).type_as(self.means.weight)


class _TwoLayer(nn.Module):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.layer1 = nn.Linear(n_in, n_in)
        self.layer2 = nn.Linear(n_in, n_out)

    def forward(self, x):
        return self.layer2(F.gelu(self.layer1(x)))


class _Gaussian3D(nn.Module):
    def __init__(self, width, edge_types, K):
        super().__init__()
        self.gbf = _GaussianBasis(K, edge_types)
        self.gbf_proj = _TwoLayer(K, width)

    def forward(self, dist, types):
        return self.gbf_proj(self.gbf(dist, types if isinstance(types, tuple) else types.long()))


class _EdgeEmbedFn(torch.autograd.Function):
    """dist_embed(hops) + featm_embed(feature_matrix).sum(-2)  (reference models/pcqm/layers.py:69-72) as ONE
    multi-hot GEMM.  Same values as the gather+sum; the point is the backward: the stock embedding backward sorts the
    3*B*N^2 indices (~8 ms per step at B=256, N=64), here it is a [V, R] x [R, W] GEMM."""

    @staticmethod
    def forward(ctx, hops, fm, w_dist, w_featm):
        shape = hops.shape
        R = hops.numel()
        Vd, Vf = w_dist.shape[0], w_featm.shape[0]
        idx = torch.cat([hops.reshape(R, 1), fm.reshape(R, -1) + Vd], 1)
        M = torch.zeros((R, Vd + Vf), dtype=w_dist.dtype, device=hops.device)
        M.scatter_add_(1, idx, torch.ones_like(idx, dtype=M.dtype))
        ctx.save_for_backward(M)
        ctx.Vd = Vd
        with torch.autocast(hops.device.type, enabled=False):          # embeddings stay fp32 under autocast
            return (M @ torch.cat([w_dist, w_featm], 0)).view(*shape, -1)

    @staticmethod
    def backward(ctx, dout):
        (M,) = ctx.saved_tensors
        with torch.autocast(dout.device.type, enabled=False):
            dW = M.t() @ dout.reshape(M.shape[0], -1).to(M.dtype)
        dwf = dW[ctx.Vd:].clone()
        dwf[0].zero_()                                   # padding_idx = 0 of featm_embed
        return None, None, dW[:ctx.Vd], dwf


class EmbedInput(nn.Module):
    def __init__(self, node_width, edge_width, upto_hop=32, embed_3d_type='gaussian', num_3d_kernels=128):
        super().__init__()
        if embed_3d_type not in ('gaussian', 'none'):
            raise ValueError('harness supports embed_3d_type gaussian|none')
        self.upto_hop = upto_hop
        self.nodef_embed = nn.Embedding(NUM_NODE_FEATURES * NODE_FEATURES_OFFSET + 1, node_width, padding_idx=0)
        self.dist_embed = nn.Embedding(upto_hop + 2, edge_width)
        self.featm_embed = nn.Embedding(NUM_EDGE_FEATURES * EDGE_FEATURES_OFFSET + 1, edge_width, padding_idx=0)
        self.uses_3d = embed_3d_type == 'gaussian'
        if self.uses_3d:
            self.m3d_embed = _Gaussian3D(edge_width, 2 * NODE_FEATURES_OFFSET + 1, num_3d_kernels)

    def forward(self, inputs):
        g = Graph(inputs)
        nf = g.node_features.long()
        h = self.nodef_embed(nf).sum(dim=2)
        hops = g.distance_matrix.long().clamp(max=self.upto_hop + 1)
        e = _EdgeEmbedFn.apply(hops, g.feature_matrix.long(), self.dist_embed.weight, self.featm_embed.weight)
        if self.uses_3d:
            a = nf[:, :, 0]
            # reference layers.py:74-77 stacks [a_i, a_j + offset] per pair; the factored form is passed instead
            e = e + self.m3d_embed(g.dist_input, (a, a + NODE_FEATURES_OFFSET))
        em = g.edge_mask.unsqueeze(-1).to(e.dtype)
        g.h, g.e, g.mask = h, e, (1 - em) * torch.finfo(e.dtype).min
        return g


class _TaskModel(nn.Module):
    node_ended, edge_ended = True, True

    def __init__(self, model_height, layer_multiplier=1, upto_hop=32, embed_3d_type='gaussian',
                 num_3d_kernels=128, num_dist_bins=None, **layer_configs):
        super().__init__()
        nw, ew = layer_configs['node_width'], layer_configs['edge_width']
        self.encoder = TGT_Encoder(model_height=model_height, layer_multiplier=layer_multiplier,
                                   node_ended=self.node_ended, edge_ended=self.edge_ended, egt_simple=False,
                                   **layer_configs)
        self.input_embed = EmbedInput(nw, ew, upto_hop, embed_3d_type, num_3d_kernels)
        if self.node_ended:
            self.final_ln_node = nn.LayerNorm(nw)
            self.pred = nn.Linear(nw, 1)
            nn.init.constant_(self.pred.bias, HL_MEAN)
        if self.edge_ended:
            self.final_ln_edge = nn.LayerNorm(ew)
            self.dist_pred = nn.Linear(ew, num_dist_bins)

    def _gap(self, g):
        h = self.final_ln_node(g.h)
        m = g.node_mask.float().unsqueeze(-1)
        return self.pred((h * m).sum(dim=1) / (m.sum(dim=1) + 1e-9)).squeeze(-1)

    def _bins(self, g):
        if g.e.is_cuda:
            # final LayerNorm folded into the distance-bin GEMM (statistics come with g.e from the last edge FFN)
            return ops.LNLinearFn.apply(g.e, self.final_ln_edge.weight, self.final_ln_edge.bias, self.dist_pred.weight,
                                        self.dist_pred.bias, ops.compute_dtype(g.e))[0]
        return self.dist_pred(self.final_ln_edge(g.e))


class TGT_Multi(_TaskModel):
    node_ended, edge_ended = True, True

    def __init__(self, model_height, num_dist_bins=128, **kw):
        super().__init__(model_height, num_dist_bins=num_dist_bins, **kw)

    def forward(self, inputs):
        g = self.encoder(self.input_embed(inputs))
        return self._gap(g), self._bins(g)


class TGT_Gap(_TaskModel):
    node_ended, edge_ended = True, False

    def forward(self, inputs):
        return self._gap(self.encoder(self.input_embed(inputs)))


class TGT_Distance(_TaskModel):
    node_ended, edge_ended = False, True

    def __init__(self, model_height, num_dist_bins=128, **kw):
        super().__init__(model_height, num_dist_bins=num_dist_bins, **kw)

    def forward(self, inputs):
        return self._bins(self.encoder(self.input_embed(inputs)))


def coords2dist(x):
    return torch.norm(x.unsqueeze(-2) - x.unsqueeze(-3), dim=-1)


def pretrain_loss(gap, logits, batch, num_bins, range_bins=8.0, dist_loss_weight=0.1):
    """L1(gap) + w * masked cross-entropy over distance bins (pretrain/scheme.py:78-88, commons.py:19-48)."""
    prim = F.l1_loss(gap, batch['target'].to(gap.dtype))
    tgt = (coords2dist(batch['dft_coords']) * ((num_bins - 1) / range_bins)).long().clamp(0, num_bins - 1)
    xent = ops.cross_entropy_rows(logits.reshape(-1, num_bins), tgt.reshape(-1))      # fp32 rows; one pass each way
    bsz = logits.size(0)
    m = batch['edge_mask'].to(xent.dtype).view(bsz, -1)
    dist = (xent.view(bsz, -1) * m).sum() / (m.sum() + 1e-9)
    return prim + dist_loss_weight * dist


TGT_AT_CONFIG = dict(       # configs/pcqm/tgt_at_200m/pretrain/tgt_at_tp.yaml
    model_height=24, node_width=768, edge_width=256, num_heads=64, triplet_heads=16, triplet_type='attention',
    source_dropout=0.3, drop_path=0.2, node_act_dropout=0.1, edge_act_dropout=0.1, activation='gelu',
    scale_degree=True, node_ffn_multiplier=1.0, edge_ffn_multiplier=1.0, upto_hop=32, num_dist_bins=512)
TGT_AGX2_CONFIG = dict(     # configs/pcqm/tgt_agx2_100m/gap_pred/tgt_agx2_tp_rdkit.yaml
    model_height=12, layer_multiplier=2, node_width=768, edge_width=256, num_heads=64, triplet_heads=16,
    triplet_type='aggregate', source_dropout=0.3, drop_path=0.2, node_act_dropout=0.1, edge_act_dropout=0.1,
    activation='gelu', scale_degree=True, node_ffn_multiplier=1.0, edge_ffn_multiplier=1.0, upto_hop=32)
