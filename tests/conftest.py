import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("TGT_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def reference_lib():
    """The real reference package (only in the build container; tests using it skip elsewhere)."""
    if not os.path.isdir(os.path.join(REFERENCE, "lib", "tgt")):
        pytest.skip("reference tree not present")
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import lib.tgt                      # noqa: F401
    import lib.tgt.layers.triplet      # noqa: F401
    import lib.tgt.layers.layers       # noqa: F401
    return sys.modules["lib"]


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-6))


def max_rel(a, b, floor=1e-6):
    """max |a-b| normalised by max |b| (the 'max-rel' figure the fp32 tolerance is stated on)."""
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + floor))


def assert_grads_close(ours, ref, tol, metric="max"):
    """Compare dicts of parameter gradients.  Gradients that are analytically zero in the reference (a bias that
    feeds a softmax is shift-invariant, so its gradient is round-off only) are checked on the scale of the largest
    gradient of the set instead of on their own (meaningless) magnitude."""
    scale = max(float(g.detach().double().abs().max()) for g in ref.values())
    for k, g in ref.items():
        a, b = ours[k].detach().double().cpu(), g.detach().double().cpu()
        if float(b.abs().max()) < 1e-6 * scale:
            assert float((a - b).abs().max()) < tol * scale, f"{k}: zero-grad noise {float((a - b).abs().max())}"
            continue
        err = max_rel(a, b) if metric == "max" else rel_err(a, b)
        assert err < tol, f"{k}: {err}"
