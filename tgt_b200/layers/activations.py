"""Activation table -- API of the reference's lib/tgt/layers/activations.py:4-25.

`get_activation(name)` returns (callable, width multiplier): gated variants (geglu / glu / swiglu) split
their input in two halves and therefore need a 2x wide first FFN projection.
"""
import torch
import torch.nn.functional as F


def _split(x: torch.Tensor):
    return x.chunk(2, dim=-1)


def geglu(x: torch.Tensor) -> torch.Tensor:
    gate, val = _split(x)
    return val * F.gelu(gate)


def glu(x: torch.Tensor) -> torch.Tensor:
    gate, val = _split(x)
    return val * torch.sigmoid(gate)


def swiglu(x: torch.Tensor) -> torch.Tensor:
    gate, val = _split(x)
    return val * torch.sigmoid(gate) * gate


glu_dict = {'geglu': geglu, 'glu': glu, 'swiglu': swiglu}


def get_activation(activation):
    if activation in glu_dict:
        return glu_dict[activation], 2
    return getattr(F, activation), 1
