"""Graph container and layer stack -- API surface of the reference's lib/tgt/encoder.py:7-90."""
from __future__ import annotations

from torch import nn

from . import ops
from .layers import TGT_Layer


class Graph(dict):
    """Attribute-style dict carrying at least h, e, mask (encoder.py:7-21); extra keys pass through."""

    def __dir__(self):
        return list(super().__dir__()) + list(self.keys())

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError('No such attribute: ' + key)

    def __setattr__(self, key, value):
        self[key] = value

    def copy(self):
        return self.__class__(self)


class TGT_Encoder(nn.Module):
    """Stack of TGT layers.  Constructor / attribute / state-dict compatible with encoder.py:24-90:
    per-layer values via `IndivConfig` lists, linear stochastic-depth schedule p*i/(L-1), the last layer
    may drop its node or edge branch, and each layer is applied `layer_multiplier` times (shared weights)."""

    class IndivConfig(list):
        pass

    def __init__(self, model_height=4, layer_multiplier=1, node_ended=True, edge_ended=True,
                 egt_simple=False, **layer_configs):
        super().__init__()
        self.model_height = model_height
        self.layer_multiplier = layer_multiplier
        self.node_ended = node_ended
        self.edge_ended = edge_ended
        self.egt_simple = egt_simple
        self.layer_configs = layer_configs
        for name, value in layer_configs.items():
            setattr(self, name, value)

        assert (self.node_ended or self.edge_ended), 'At least one of node_ended and edge_ended must be True'

        self.TGT_layers = nn.ModuleList(TGT_Layer(**self.get_layer_kwargs(i)) for i in range(self.model_height))

    def get_layer_kwargs(self, i):
        is_last = i == self.model_height - 1
        kwargs = {}
        for name, value in self.layer_configs.items():
            if isinstance(value, self.IndivConfig):
                kwargs[name] = value[i]
            elif name == 'drop_path':
                kwargs[name] = value * i / (self.model_height - 1)
            else:
                kwargs[name] = value
        kwargs['node_update'] = not (is_last and not self.node_ended)
        kwargs['edge_update'] = (not self.egt_simple) and not (is_last and not self.edge_ended)
        return kwargs

    def apply_layer(self, layer_idx, graph):
        layer = self.TGT_layers[layer_idx]
        total = self.model_height * self.layer_multiplier
        for rep in range(self.layer_multiplier):
            # position of this layer application in the stack: lets the triplet op decide how many of the LAST layers
            # may keep their projection for backward instead of recomputing it (ops.keep_projection)
            ops.set_layer_hint(layer_idx * self.layer_multiplier + rep, total)
            try:
                graph = layer(graph)
            finally:
                ops.set_layer_hint(None, None)
        return graph

    def forward(self, inputs):
        g = Graph(inputs)
        for i in range(self.model_height):
            g = self.apply_layer(i, g)
        return g
