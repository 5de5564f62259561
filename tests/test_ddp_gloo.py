"""World-size-2 data-parallel host logic on CPU/gloo: per-rank batches differ, DDP averages the
gradients exactly like a single process seeing both batches, and the bench timing reduction is a max over ranks.
The encoder has model_height=0 here (the CUDA kernels have no CPU path by design), so only the plumbing runs."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(model_height=0, node_width=32, edge_width=16, num_heads=4, triplet_heads=2, triplet_type="attention",
           upto_hop=32, num_dist_bins=8, num_3d_kernels=8)


def _loss(model, batch):
    from tgt_b200.harness.models import pretrain_loss
    from tgt_b200.harness.synthetic import add_scheme_fields
    batch = add_scheme_fields(batch, with_3d=True)
    gap, logits = model(batch)
    return pretrain_loss(gap, logits, batch, CFG["num_dist_bins"])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tgt_b200.harness.dist import env_rank_world, max_over_ranks, rank_batch, wrap_ddp
        from tgt_b200.harness.models import TGT_Multi
        assert env_rank_world() == (rank, rank, world)
        torch.manual_seed(0)
        model = TGT_Multi(**CFG)
        net = wrap_ddp(model, world)
        batch = rank_batch(3, 6, rank)
        _loss(net, batch).backward()
        grads = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        t = max_over_ranks(10.0 + rank, world, "cpu")
        torch.save(dict(grads=grads, t=t, nf=batch["node_features"]), os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_ddp_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in (0, 1))
    assert not torch.equal(r0["nf"], r1["nf"])                 # ranks own different graphs
    assert torch.allclose(r0["grads"], r1["grads"])            # all-reduced gradients agree
    assert r0["t"] == r1["t"] == 11.0                          # max over ranks
    # single-process reference: average of the two per-rank gradients
    from tgt_b200.harness.dist import rank_batch
    from tgt_b200.harness.models import TGT_Multi
    acc = None
    for rank in (0, 1):
        torch.manual_seed(0)
        model = TGT_Multi(**CFG)
        _loss(model, rank_batch(3, 6, rank)).backward()
        g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        acc = g if acc is None else acc + g
    assert torch.allclose(r0["grads"], acc / 2, atol=1e-6)
