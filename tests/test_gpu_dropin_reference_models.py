"""The drop-in claim itself: the reference's OWN model classes (`lib.models.pcqm.{multitask,gap_predictor,
distance_predictor}`, staged unmodified under oracle/_ref by oracle/build_ref.py) imported with `lib.tgt` aliased to
tgt_b200 -- constructed with the reference's kwargs, loaded with the reference's state dict
(strict), run on the GPU -- against the same classes on the real `lib.tgt` (CPU, fp64).

reference call sites exercised: gap_predictor.py:5,50  distance_predictor.py:5,50  multitask.py:5,55
(`from lib.tgt import TGT_Encoder` / `self.encoder(g)`), models/pcqm/layers.py:9 (`Graph`).
"""
import pytest
import torch

from conftest import max_rel, rel_err
from oracle import ref_loader as R
from tgt_b200.harness.synthetic import make_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"

SMALL = dict(model_height=2, node_width=64, edge_width=32, num_heads=4, triplet_heads=2, activation="gelu",
             scale_degree=True, node_ffn_multiplier=1.0, edge_ffn_multiplier=1.0, upto_hop=32)


def _need_ref():
    if not R.available():
        pytest.skip("oracle/_ref not staged (run python oracle/build_ref.py in the build container)")


def _cases():
    return [("TGT_Multi", dict(triplet_type="attention", num_dist_bins=64)),
            ("TGT_Gap", dict(triplet_type="aggregate", layer_multiplier=2)),
            ("TGT_Distance", dict(triplet_type="attention", num_dist_bins=32, embed_3d_type="none"))]


def _to(batch, device, fdtype=None):
    out = {}
    for k, v in batch.items():
        v = v.to(device)
        if fdtype is not None and v.is_floating_point():
            v = v.to(fdtype)
        out[k] = v
    return out


@pytest.mark.parametrize("cls,extra", _cases())
def test_reference_models_on_tgt_b200_fp32(cls, extra):
    _need_ref()
    import tgt_b200
    kw = dict(SMALL, **extra)
    batch = make_batch(3, 11, seed=5)
    torch.manual_seed(0)
    with R.reference() as ref:
        m_ref = getattr(ref, cls)(**kw).double().eval()
        state = {k: v.detach().clone() for k, v in m_ref.state_dict().items()}
        with torch.no_grad():
            want = m_ref(_to(batch, "cpu", torch.float64))
    want = want if isinstance(want, tuple) else (want,)
    with R.reference_models_over(tgt_b200) as ours:
        m = getattr(ours, cls)(**kw)
        assert type(m.encoder) is tgt_b200.TGT_Encoder
        m.load_state_dict({k: v.float() for k, v in state.items()}, strict=True)       # the reference's checkpoint ABI
        m = m.to(DEV).eval()
        with torch.no_grad():
            got = m(_to(batch, DEV))
    got = got if isinstance(got, tuple) else (got,)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.dtype == torch.float32 and max_rel(g.cpu(), w) < 5e-5, max_rel(g.cpu(), w)


def test_reference_multitask_training_step_on_tgt_b200_bf16():
    """TGT-At widths (Wn 768, We 256, Hn 64, Ht 16), 2 layers: the reference's TGT_Multi + the reference's own loss
    (F.l1_loss + DiscreteDistLoss, pretrain/scheme.py:78-88) under bf16 autocast in train mode on our kernels --
    loss and parameter gradients against the real reference in fp64, judged by the error the real reference makes
    under CUDA bf16 autocast on the same batch."""
    _need_ref()
    import tgt_b200
    kw = dict(model_height=2, node_width=768, edge_width=256, num_heads=64, triplet_heads=16, triplet_type="attention",
              num_dist_bins=64, upto_hop=32)
    batch = make_batch(2, 24, seed=3)

    def loss_of(ns, model, b, acc):
        gap, logits = model(b)
        dist_targ = ns.commons.coords2dist(b["dft_coords"])
        return (torch.nn.functional.l1_loss(gap.to(acc), b["target"].to(acc))
                + 0.1 * ns.commons.DiscreteDistLoss(64, 8)(logits.to(acc), dist_targ, b["edge_mask"]))

    torch.manual_seed(1)
    with R.reference() as ref:
        m64 = ref.TGT_Multi(**kw).double().train()           # every dropout defaults to 0: train mode is deterministic
        state = {k: v.detach().clone() for k, v in m64.state_dict().items()}
        l64 = loss_of(ref, m64, _to(batch, "cpu", torch.float64), torch.float64)
        l64.backward()
        g64 = {k: p.grad.clone() for k, p in m64.named_parameters()}
        # the real reference under CUDA bf16 autocast (the comparator of the bf16 criterion)
        mac = ref.TGT_Multi(**kw)
        mac.load_state_dict({k: v.float() for k, v in state.items()})
        mac = mac.to(DEV).train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lac = loss_of(ref, mac, _to(batch, DEV), torch.float32)
        lac.backward()
        gac = {k: p.grad.double().cpu() for k, p in mac.named_parameters()}
    with R.reference_models_over(tgt_b200) as ours:
        m = ours.TGT_Multi(**kw)
        m.load_state_dict({k: v.float() for k, v in state.items()}, strict=True)
        m = m.to(DEV).train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lo = loss_of(ours, m, _to(batch, DEV), torch.float32)
        lo.backward()
        go = {k: p.grad.double().cpu() for k, p in m.named_parameters()}

    def gerr(g):
        num = sum(float((g[k] - g64[k]).pow(2).sum()) for k in g64)
        return (num / sum(float(v.pow(2).sum()) for v in g64.values())) ** 0.5
    e_loss_o, e_loss_r = abs(float(lo) - float(l64)) / abs(float(l64)), abs(float(lac) - float(l64)) / abs(float(l64))
    print(f"loss {float(lo):.6f} (fp64 {float(l64):.6f}, reference autocast {float(lac):.6f}); "
          f"grad rel-L2 ours {gerr(go):.3e} reference autocast {gerr(gac):.3e}")
    assert set(go) == set(g64)
    assert e_loss_o <= max(e_loss_r, 2e-3), (e_loss_o, e_loss_r)
    assert gerr(go) <= 1.0 * gerr(gac), (gerr(go), gerr(gac))
