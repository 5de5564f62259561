"""Synthetic PCQM4Mv2-shaped batches: same keys / dtypes / value ranges as the reference's
`padded_collate` output plus the scheme-side extras (edge_mask, dist_input)."""
from __future__ import annotations

from typing import Dict

import torch

NODE_FEATURES_OFFSET, NUM_NODE_FEATURES = 128, 9      # lib/models/pcqm/consts.py:1-4
EDGE_FEATURES_OFFSET, NUM_EDGE_FEATURES = 8, 3


def make_batch(B: int, N: int, seed: int = 1, device="cpu", with_3d: bool = True) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    num_nodes = torch.randint(max(N // 2, 1), N + 1, (B,), generator=g)
    num_nodes[0] = N                                   # fix the padded shape, exercise padding elsewhere
    ar = torch.arange(N)
    node_mask = (ar[None, :] < num_nodes[:, None])
    edge_mask = node_mask[:, :, None] & node_mask[:, None, :]

    raw = torch.randint(0, 100, (B, N, NUM_NODE_FEATURES), generator=g)
    nf = raw + 1 + NODE_FEATURES_OFFSET * torch.arange(NUM_NODE_FEATURES)
    nf = nf * node_mask[:, :, None]

    # shortest-path-like hop matrix: |i-j| along a chain with a few shortcuts, 0 on diag / padding
    dm = (ar[None, :, None] - ar[None, None, :]).abs().expand(B, N, N).clone()
    dm = torch.minimum(dm, torch.randint(1, 40, (B, N, N), generator=g))
    dm = torch.minimum(dm, dm.transpose(1, 2))
    dm = dm * edge_mask
    dm.diagonal(dim1=1, dim2=2).zero_()

    bond = (dm == 1)
    fraw = torch.randint(0, 6, (B, N, N, NUM_EDGE_FEATURES), generator=g)
    fm = (fraw + 1 + EDGE_FEATURES_OFFSET * torch.arange(NUM_EDGE_FEATURES)) * bond[..., None]

    coords = torch.randn(B, N, 3, generator=g) * 3.0 * node_mask[:, :, None]
    target = torch.randn(B, generator=g, dtype=torch.float64) * 1.1621397 + 5.6894608

    batch = dict(
        num_nodes=num_nodes.long(),
        node_mask=node_mask.to(torch.uint8),
        node_features=nf.to(torch.int16),
        distance_matrix=dm.to(torch.int16),
        feature_matrix=fm.to(torch.int16),
        dft_coords=coords.float(),
        target=target,
    )
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return add_scheme_fields(batch, with_3d)


def add_scheme_fields(batch: Dict[str, torch.Tensor], with_3d: bool = True) -> Dict[str, torch.Tensor]:
    """edge_mask = m (x) m and dist_input = |x_i - x_j| (pretrain/scheme.py:62-74; coordinate noise omitted:
    it is RNG-only and off the hot path)."""
    nm = batch["node_mask"]
    batch = dict(batch)
    batch["edge_mask"] = nm.unsqueeze(-1) * nm.unsqueeze(-2)
    if with_3d:
        x = batch["dft_coords"]
        batch["dist_input"] = torch.norm(x.unsqueeze(-2) - x.unsqueeze(-3), dim=-1)
    return batch


def make_edge_inputs(B: int, N: int, W: int, num_nodes=None, seed: int = 0, dtype=torch.float32):
    """Config-1 style module inputs: e ~ N(0,1) [B,N,N,W] and the additive mask [B,N,N,1]
    (mask built exactly like lib/models/pcqm/layers.py:78-80)."""
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(B, N, N, W, generator=g, dtype=torch.float32).to(dtype)
    if num_nodes is None:
        num_nodes = [max(1, N - 3 * i) for i in range(B)]
    nn_ = torch.tensor(num_nodes)
    m = (torch.arange(N)[None, :] < nn_[:, None]).float()
    em = (m[:, :, None] * m[:, None, :]).unsqueeze(-1)
    mask = (1 - em) * torch.finfo(torch.float32).min
    return e, mask
