"""Stage the UNMODIFIED reference under oracle/_ref so that it travels to the GPU box.

The reference (shamim-hussain/tgt) is a pure-Python code base with no setup.py / pyproject.toml, so there is nothing
to pip-install or compile: "building" it means placing the modules of the hot path and of the models / loss that sit on
top of it where `import lib...` finds them.  This script copies, byte for byte and only into the git-ignored
`oracle/_ref/`, the files listed in FILES from `$TGT_REFERENCE` (default /root/reference):

    lib/__init__.py, lib/tgt/**            the hot path itself (EGT attention, triplet modules, encoder)
    lib/models/pcqm/**                     TGT_Multi / TGT_Gap / TGT_Distance + EmbedInput (the callers of the path)
    lib/training_schemes/pcqm/commons.py   DiscreteDistLoss, coords2dist, BinsProcessor.bins2dist (pretrain loss,
                                           two-stage inference glue)

`oracle/_ref/` is listed in .gitignore (it never enters the history) and NOT in .gpurunignore (it ships with the
snapshot like our own .so).  Nothing under tgt_b200/ imports it: it is test / baseline infrastructure --
`bench.py --impl reference` (the CPU arm), `bench.py`'s `gpu_eager_baseline` leg, and the drop-in test that runs the
real `lib.models.pcqm` on top of tgt_b200 (tests/test_gpu_dropin_reference_models.py).

    python oracle/build_ref.py            # called by __graft_entry__.build() when the reference tree is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "lib/__init__.py",
    "lib/tgt/__init__.py",
    "lib/tgt/encoder.py",
    "lib/tgt/layers/__init__.py",
    "lib/tgt/layers/layers.py",
    "lib/tgt/layers/triplet.py",
    "lib/tgt/layers/activations.py",
    "lib/models/pcqm/consts.py",
    "lib/models/pcqm/layers.py",
    "lib/models/pcqm/gap_predictor.py",
    "lib/models/pcqm/distance_predictor.py",
    "lib/models/pcqm/multitask.py",
    "lib/training_schemes/pcqm/commons.py",
]


def reference_root() -> str:
    return os.environ.get("TGT_REFERENCE", "/root/reference")


def build(verbose: bool = True) -> str | None:
    """Copy FILES into oracle/_ref; returns the destination, or None when the reference tree is absent (GPU box:
    the staged copy from the build container is used as is)."""
    src_root = reference_root()
    if not os.path.isdir(os.path.join(src_root, "lib", "tgt")):
        return DEST if os.path.isdir(os.path.join(DEST, "lib", "tgt")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    sub = os.path.join(src_root, ".SUBMODULES.json")
    commit = None
    if os.path.exists(sub):
        try:
            commit = json.load(open(sub)).get("commit")
        except (ValueError, AttributeError):
            commit = None
    json.dump(dict(source=src_root, commit=commit, files=manifest), open(os.path.join(DEST, "MANIFEST.json"), "w"),
              indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files under {DEST}")
    return DEST


if __name__ == "__main__":
    out = build()
    if out is None:
        print("reference tree not found and nothing staged", file=sys.stderr)
        sys.exit(1)
