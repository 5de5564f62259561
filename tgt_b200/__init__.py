"""tgt_b200 -- B200-native (sm_100a) implementation of the TGT triplet / EGT attention hot path.

Drop-in for the reference's `lib.tgt` package: `from tgt_b200 import TGT_Encoder, Graph`
(reference: lib/tgt/__init__.py:1) and `from tgt_b200.layers import TGT_Layer` (lib/tgt/layers/__init__.py:1).
See INTEGRATION.md for how the reference tree is pointed at this package.
"""
from .encoder import TGT_Encoder, Graph
from . import layers

__all__ = ["TGT_Encoder", "Graph", "layers"]
__version__ = "0.1.0"
