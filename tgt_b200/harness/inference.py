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
                amp_dtype=torch.bfloat16) -> torch.Tensor:
    """S stochastic passes of TGT_Gap on the decoded distances; returns the per-molecule mean prediction [B]."""
    b = dict(add_scheme_fields(batch, with_3d=False))
    S = dist_inputs.shape[1]
    preds = []
    for s in range(samples):
        b["dist_input"] = dist_inputs[:, s % S]
        with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
            preds.append(gap_model(b).float())
    return torch.stack(preds, -1).mean(-1)


def two_stage_predict(dist_model, gap_model, batch, samples: int, range_bins: float = 8.0, amp_dtype=torch.bfloat16):
    _, d = predict_dist_inputs(dist_model, batch, samples, range_bins, amp_dtype)
    return predict_gap(gap_model, batch, d, samples, amp_dtype)
