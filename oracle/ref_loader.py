"""Import the staged reference (oracle/_ref, see build_ref.py) -- test / baseline infrastructure only.

Two ways of using it:
  * `reference()`                 : the real `lib.tgt` + `lib.models.pcqm` + `lib.training_schemes.pcqm.commons`,
                                    exactly as shipped (CPU baseline arm of bench.py, GPU-eager comparator, ground truth
                                    of the drop-in test);
  * `reference_models_over(pkg)`  : the real `lib.models.pcqm.*` modules imported with `lib.tgt` aliased to `pkg`
                                    (`sys.modules["lib.tgt"] = tgt_b200`) -- the reference's
                                    own model classes running on top of our kernels, with nothing else changed.
Nothing under tgt_b200/ may import this module.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_MODEL_MODULES = ("lib.models.pcqm.consts", "lib.models.pcqm.layers", "lib.models.pcqm.gap_predictor",
                  "lib.models.pcqm.distance_predictor", "lib.models.pcqm.multitask")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "lib", "tgt", "encoder.py"))


def _purge(prefixes):
    saved = {}
    for name in list(sys.modules):
        if any(name == p or name.startswith(p + ".") for p in prefixes):
            saved[name] = sys.modules.pop(name)
    return saved


def _import_all(extra=()):
    ns = types.SimpleNamespace()
    for name in _MODEL_MODULES + tuple(extra):
        setattr(ns, name.rsplit(".", 1)[1], importlib.import_module(name))
    ns.TGT_Multi = ns.multitask.TGT_Multi
    ns.TGT_Gap = ns.gap_predictor.TGT_Gap
    ns.TGT_Distance = ns.distance_predictor.TGT_Distance
    return ns


@contextlib.contextmanager
def _on_path():
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/build_ref.py` where /root/reference exists")
    sys.dont_write_bytecode = True
    saved = _purge(("lib",))
    sys.path.insert(0, REF)
    try:
        yield
    finally:
        sys.path.remove(REF)
        _purge(("lib",))
        sys.modules.update(saved)


@contextlib.contextmanager
def reference():
    """The unmodified reference: namespace with TGT_Multi / TGT_Gap / TGT_Distance, `tgt` (= lib.tgt), `triplet`,
    `tgt_layers` (= lib.tgt.layers.layers) and `commons`."""
    with _on_path():
        ns = _import_all(("lib.training_schemes.pcqm.commons",))
        ns.tgt = importlib.import_module("lib.tgt")
        ns.triplet = importlib.import_module("lib.tgt.layers.triplet")
        ns.tgt_layers = importlib.import_module("lib.tgt.layers.layers")
        yield ns


@contextlib.contextmanager
def reference_models_over(pkg):
    """The reference's lib.models.pcqm with `lib.tgt` replaced by `pkg` (which must expose TGT_Encoder, Graph and a
    `layers` sub-package) -- the one-line integration of INTEGRATION.md, done through sys.modules."""
    with _on_path():
        importlib.import_module("lib")
        sys.modules["lib.tgt"] = pkg
        sys.modules["lib.tgt.layers"] = pkg.layers
        ns = _import_all(("lib.training_schemes.pcqm.commons",))
        assert ns.multitask.TGT_Encoder is pkg.TGT_Encoder, "alias did not take"
        yield ns
