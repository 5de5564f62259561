"""Two-stage inference (BASELINE config 5): distance predictor -> distance bins -> gap predictor.

Mirrors the reference prediction schemes with the parquet round trip removed: `predict_bins`
(lib/training_schemes/pcqm/dist_pred/scheme.py:181-205: S stochastic forward passes, softmax, p + p^T, argmax),
`BinsProcessor.bins2dist` (commons.py:72-82) and the gap predictor's sample loop (gap_pred/scheme.py:78-107: sample s
uses distance input s mod S, predictions are averaged, :116).  Both models run in TRAIN mode under no_grad like the
reference (`predict_in_train`, tgt_training.py:42): dropout is the Monte-Carlo sampler."""
from __future__ import annotations

from typing import Dict

import torch

from .. import ops
from .synthetic import add_scheme_fields


class GraphedForward:
    """One stochastic forward pass of a task model captured as a CUDA graph and replayed per Monte-Carlo sample.

    The MC-dropout loops of the reference (dist_pred/scheme.py:181-205, gap_pred/scheme.py:78-107) run 50-100 forwards of
    ONE fixed-shape batch: a 12-24 layer TGT forward is ~2.5 k kernel launches, and at N = 32 (BASELINE config 2) the
    launches, not the kernels, bound the batch.  Capture is possible because nothing on our path synchronises: RNG comes
    from torch's graph-safe generator (source dropout, DropPath) and, for the fused GELU + dropout GEMM epilogue, from a
    device-resident seed drawn inside the graph (layers.FFN._run), so every replay samples fresh masks.
    `inputs`: dict of static input tensors (cloned); `call(**updates)` copies the given tensors into them, replays and
    returns the static output(s) -- consume or clone them before the next call."""

    def __init__(self, model, inputs: Dict[str, torch.Tensor], amp_dtype=torch.bfloat16, warmup: int = 2):
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inputs.items()}
        self.model, self.amp = model, amp_dtype

        def run():
            with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
                return model(self.static)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = run()

    def __call__(self, **updates):
        for k, v in updates.items():
            self.static[k].copy_(v)
        self.graph.replay()
        return self.out


@torch.no_grad()
def predict_dist_inputs(dist_model, batch: Dict[str, torch.Tensor], samples: int, range_bins: float = 8.0,
                        amp_dtype=torch.bfloat16, want_bins: bool = False):
    """S stochastic passes of TGT_Distance -> (bins [B,S,N,N] int16 or None, dist_input [B,S,N,N] fp32)."""
    b = add_scheme_fields(batch, with_3d=False)
    dists, bins = [], []
    for _ in range(samples):
        with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
            logits = dist_model(b)
        bn, d = ops.bins_decode(logits, range_bins=range_bins, want_bins=want_bins)
        dists.append(d)
        bins.append(bn)
    return (torch.stack(bins, 1) if want_bins else None), torch.stack(dists, 1)


@torch.no_grad()
def predict_gap(gap_model, batch: Dict[str, torch.Tensor], dist_inputs: torch.Tensor, samples: int,
                amp_dtype=torch.bfloat16, graph: bool = False) -> torch.Tensor:
    """S stochastic passes of TGT_Gap on the decoded distances; returns the per-molecule mean prediction [B].
    graph=True replays one captured forward per sample (GraphedForward) instead of launching it eagerly."""
    b = dict(add_scheme_fields(batch, with_3d=False))
    S = dist_inputs.shape[1]
    preds = []
    if graph:
        b["dist_input"] = dist_inputs[:, 0].contiguous()
        gf = GraphedForward(gap_model, b, amp_dtype)
        for s in range(samples):
            preds.append(gf(dist_input=dist_inputs[:, s % S]).float().clone())
        return torch.stack(preds, -1).mean(-1)
    for s in range(samples):
        b["dist_input"] = dist_inputs[:, s % S]
        with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
            preds.append(gap_model(b).float())
    return torch.stack(preds, -1).mean(-1)


def two_stage_predict(dist_model, gap_model, batch, samples: int, range_bins: float = 8.0, amp_dtype=torch.bfloat16):
    _, d = predict_dist_inputs(dist_model, batch, samples, range_bins, amp_dtype)
    return predict_gap(gap_model, batch, d, samples, amp_dtype)
