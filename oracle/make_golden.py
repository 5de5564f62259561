"""Generate tests/golden/*.pt from the REAL reference (imported from /root/reference, never copied).

Run in the build container:   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
Each fixture holds: module kind + constructor kwargs, the reference module's state_dict, the inputs,
the reference outputs and the reference gradients (w.r.t. inputs and every parameter) for a fixed
upstream gradient, all in fp32 computed by the reference in fp64 (so they are the ground truth to
~1e-15, rounded once to fp32 for storage).  Shapes are small so the files stay < 1 MB each.
Also writes tests/golden/state_dict_abi.json: key -> shape for the shipped model sizes (checkpoint ABI).
"""
import json
import os
import sys

import torch

REF = os.environ.get("TGT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lib.tgt.layers import triplet as rtri            # noqa: E402
from lib.tgt.layers import layers as rlay             # noqa: E402
from lib.tgt import TGT_Encoder                        # noqa: E402
from lib.models.pcqm.multitask import TGT_Multi       # noqa: E402
from lib.models.pcqm.gap_predictor import TGT_Gap     # noqa: E402
from lib.models.pcqm.distance_predictor import TGT_Distance  # noqa: E402
from tgt_b200.harness.synthetic import make_edge_inputs, make_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def f32(t):
    return t.detach().to(torch.float32).clone()


def run_edge_module(mod, e, mask, seed):
    mod = mod.double().eval()
    e = e.double().requires_grad_(True)
    out = mod(e, mask.double())
    g = torch.Generator().manual_seed(seed)
    dout = torch.randn(out.shape, generator=g, dtype=torch.float64)
    out.backward(dout)
    return dict(state={k: f32(v) for k, v in mod.state_dict().items()},
                e=f32(e), mask=f32(mask), out=f32(out), dout=f32(dout), de=f32(e.grad),
                grads={k: f32(p.grad) for k, p in mod.named_parameters()})


def main():
    torch.manual_seed(0)
    # ---- triplet variants: B=3, N=12 (ragged 12/9/5), W=64, H=4 (d=16)  and a d=8, N=7 odd case
    for kind in ["attention", "aggregate", "attention_ungated", "aggregate_ungated", "axial_attention",
                 "tiangular_update"]:
        for tag, (B, N, W, H, nn_) in {"a": (3, 12, 64, 4, [12, 9, 5]), "b": (2, 7, 32, 4, [7, 4])}.items():
            torch.manual_seed((sum(map(ord, kind + tag))))
            mod = rtri.get_triplet_layer(kind)(W, H)
            # non-trivial LN affine so gamma/beta gradients are exercised
            with torch.no_grad():
                mod.tri_ln_e.weight.uniform_(0.5, 1.5)
                mod.tri_ln_e.bias.uniform_(-0.2, 0.2)
            e, mask = make_edge_inputs(B, N, W, nn_, seed=3)
            fx = run_edge_module(mod, e, mask, seed=4)
            fx.update(kind=kind, edge_width=W, num_heads=H)
            torch.save(fx, os.path.join(OUT, f"triplet_{kind}_{tag}.pt"))

    # ---- EGT_Attention / EdgeUpdate: B=3, N=12, Wn=48, We=32, H=4 (d=12 like the shipped model)
    for kind, cls, kw in [("egt", rlay.EGT_Attention, dict(scale_degree=True, edge_update=True)),
                          ("egt_noscale_noedge", rlay.EGT_Attention, dict(scale_degree=False, edge_update=False)),
                          ("edge_update", rlay.EdgeUpdate, {})]:
        torch.manual_seed(7)
        B, N, Wn, We, H = 3, 12, 48, 32, 4
        mod = cls(Wn, We, H, **kw).double().eval()
        with torch.no_grad():
            mod.mha_ln_e.weight.uniform_(0.5, 1.5)
            mod.mha_ln_e.bias.uniform_(-0.2, 0.2)
            mod.mha_ln_h.weight.uniform_(0.5, 1.5)
        e, mask = make_edge_inputs(B, N, We, [12, 9, 5], seed=5)
        h = torch.randn(B, N, Wn, generator=torch.Generator().manual_seed(6))
        hd, ed = h.double().requires_grad_(True), e.double().requires_grad_(True)
        ho, eo = mod(hd, ed, mask.double())
        g = torch.Generator().manual_seed(8)
        dh = torch.randn(ho.shape, generator=g, dtype=torch.float64)
        de = torch.randn(eo.shape, generator=g, dtype=torch.float64)
        loss = (ho * dh).sum() + ((eo * de).sum() if eo.requires_grad else 0.)
        loss.backward()
        fx = dict(kind=kind, node_width=Wn, edge_width=We, num_heads=H, kwargs=kw,
                  state={k: f32(v) for k, v in mod.state_dict().items()},
                  h=f32(h), e=f32(e), mask=f32(mask), h_out=f32(ho), e_out=f32(eo), dh_out=f32(dh), de_out=f32(de),
                  dh=f32(hd.grad) if hd.grad is not None else None,
                  de=f32(ed.grad) if ed.grad is not None else None,
                  grads={k: f32(p.grad) for k, p in mod.named_parameters() if p.grad is not None})
        torch.save(fx, os.path.join(OUT, f"{kind}.pt"))

    # ---- whole encoder, eval mode: TGT-At-like and TGT-Agx2-like miniatures + known answers for models
    for tag, cfg, cls, extra in [
        ("multi_at", dict(model_height=3, node_width=48, edge_width=32, num_heads=4, triplet_heads=2,
                          triplet_type="attention"), TGT_Multi, dict(num_dist_bins=16)),
        ("gap_agx2", dict(model_height=2, layer_multiplier=2, node_width=48, edge_width=32, num_heads=4,
                          triplet_heads=2, triplet_type="aggregate"), TGT_Gap, {}),
        ("dist_at", dict(model_height=2, node_width=48, edge_width=32, num_heads=4, triplet_heads=2,
                         triplet_type="attention"), TGT_Distance, dict(num_dist_bins=16)),
    ]:
        torch.manual_seed(11)
        model = cls(**cfg, **extra).double().eval()
        batch = make_batch(3, 10, seed=2)
        bd = {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}
        out = model(bd)
        outs = [f32(o) for o in (out if isinstance(out, tuple) else (out,))]
        fx = dict(kind=tag, cfg=cfg, extra=extra, state={k: f32(v) for k, v in model.state_dict().items()},
                  batch=batch, outs=outs)
        torch.save(fx, os.path.join(OUT, f"model_{tag}.pt"))

    # ---- checkpoint ABI of the shipped sizes (keys + shapes only)
    abi = {}
    for tag, cls, cfg in [
        ("TGT_Multi_At", TGT_Multi, dict(model_height=24, node_width=768, edge_width=256, num_heads=64,
                                          triplet_heads=16, triplet_type="attention", num_dist_bins=512)),
        ("TGT_Gap_Agx2", TGT_Gap, dict(model_height=12, layer_multiplier=2, node_width=768, edge_width=256,
                                        num_heads=64, triplet_heads=16, triplet_type="aggregate")),
        ("TGT_Distance_At", TGT_Distance, dict(model_height=24, node_width=768, edge_width=256, num_heads=64,
                                                triplet_heads=16, triplet_type="attention", num_dist_bins=256)),
    ]:
        with torch.device("meta"):
            m = cls(**cfg)
        abi[tag] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_abi.json"), "w") as f:
        json.dump(abi, f, indent=0, sort_keys=True)

    # ---- config-1 known answers (weights regenerated from the seed, only checksums stored)
    torch.manual_seed(0)
    mod = rtri.TripletAttention(512, 8)
    e, mask = make_edge_inputs(4, 16, 512, [16, 12, 9, 5], seed=0)
    out = mod.double()(e.double(), mask.double())
    kat = dict(sum=float(out.sum()), abs_sum=float(out.abs().sum()),
               samples=[float(out[b, i, j, c]) for b, i, j, c in [(0, 0, 0, 0), (1, 3, 7, 100), (2, 8, 8, 511), (3, 4, 2, 17)]],
               wsum=float(sum(p.double().sum() for p in mod.parameters())))
    with open(os.path.join(OUT, "config1_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("golden fixtures written to", OUT)
    for n in sorted(os.listdir(OUT)):
        print(f"  {n:40s} {os.path.getsize(os.path.join(OUT, n)) / 1024:8.1f} KB")


def bins_fixture():
    """tests/golden/bins_decode.pt: the reference's own `predict_bins` (dist_pred/scheme.py:181-205, called unbound on a
    stub that replays fixed logits) and `BinsProcessor.bins2dist` (commons.py:72-82) on random logits."""
    import tempfile
    from types import SimpleNamespace
    from lib.training_schemes.pcqm.dist_pred.scheme import SCHEME as DistScheme
    from lib.training_schemes.pcqm.commons import BinsProcessor

    g = torch.Generator().manual_seed(11)
    B, N, nb, S = 3, 10, 256, 2
    logits = [torch.randn(B, N, N, nb, generator=g) * 3.0 for _ in range(S)]
    it = iter(logits)
    stub = SimpleNamespace(nb_draw_samples=S, model=lambda batch: next(it))
    bins = DistScheme.predict_bins(stub, None)                      # [B, S, N, N] int64
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "meta.json"), "w") as f:
            json.dump(dict(num_samples=S, num_bins=nb, range_bins=8), f)
        dist = BinsProcessor(d).bins2dist(bins)
    torch.save(dict(logits=torch.stack(logits, 1), bins=bins.to(torch.int16), dist=f32(dist), num_bins=nb, range_bins=8.0),
               os.path.join(OUT, "bins_decode.pt"))
    print("bins_decode.pt written", tuple(bins.shape), float(dist.max()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bins":
        bins_fixture()
    else:
        main()
        bins_fixture()
