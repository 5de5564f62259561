"""CPU tests of the host-side algebra and bookkeeping that feed the tcgen05 kernels (no GPU, no kernel calls):
LayerNorm folding, the parameter gradients recovered from the affine-free weight-gradient GEMM, the per-head-pair weight
panel order of the fused forward, the DropPath weight-gradient identity, the ctypes mirror of tgt_gemm_desc and the
keep-projection policy."""
import ctypes as C
import re

import pytest
import torch

from tgt_b200 import _C, ops


def test_ln_fold_identity():
    """LN(x) W^T + b == rstd * (x Wg^T - mean * colsum) + b'  with Wg = W*gamma, colsum = rowsum(Wg), b' = b + W beta."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(50, 64, generator=g, dtype=torch.float64) * 1.5 + 0.3
    W, b = torch.randn(24, 64, generator=g, dtype=torch.float64), torch.randn(24, generator=g, dtype=torch.float64)
    gamma, beta = 1 + 0.2 * torch.randn(64, generator=g, dtype=torch.float64), torch.randn(64, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.layer_norm(x, (64,), gamma, beta) @ W.t() + b
    # the helper works in fp32 (its outputs feed the fp32 epilogue vectors); evaluate the identity in fp64 on them
    Wg, bp, cs = ops._ln_fold(W.float(), b.float(), gamma.float(), beta.float(), torch.float32)
    mean = x.mean(1, keepdim=True)
    rstd = (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    got = rstd * (x @ Wg.double().t() - mean * cs.double()) + bp.double()
    assert torch.allclose(got, ref, atol=1e-4)


def test_ln_backward_parameter_gradients_from_affine_free_gemm():
    """G|db = dout^T [xhat | 1]  =>  dW = G*gamma + db (x) beta, dgamma = colsum(W*G), dbeta = db W  (ops.ln_linear_bwd)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(40, 32, generator=g, dtype=torch.float64, requires_grad=True)
    W = torch.randn(12, 32, generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(12, dtype=torch.float64, requires_grad=True)
    gamma = (1 + 0.3 * torch.randn(32, generator=g, dtype=torch.float64)).requires_grad_(True)
    beta = torch.randn(32, generator=g, dtype=torch.float64).requires_grad_(True)
    dout = torch.randn(40, 12, generator=g, dtype=torch.float64)
    (torch.nn.functional.layer_norm(x, (32,), gamma, beta) @ W.t() + b).backward(dout)
    xd = x.detach()
    xhat = (xd - xd.mean(1, keepdim=True)) * (xd.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    Ga = dout.t() @ torch.cat([xhat, torch.ones(40, 1, dtype=torch.float64)], 1)
    G, db = Ga[:, :32], Ga[:, 32]
    Wd, gd, bd = W.detach(), gamma.detach(), beta.detach()
    assert torch.allclose(G * gd + db[:, None] * bd, W.grad, atol=1e-10)
    assert torch.allclose((Wd * G).sum(0), gamma.grad, atol=1e-10)
    assert torch.allclose(db @ Wd, beta.grad, atol=1e-10)
    assert torch.allclose(db, b.grad, atol=1e-10)


def test_droppath_weight_gradient_identity():
    """dW of out = s_b * (a W^T): sum_b s_b * do_b^T a_b  (the per-graph batched GEMM + weighted sum of linear_residual_bwd)."""
    g = torch.Generator().manual_seed(2)
    B, rows, N, K = 5, 7, 6, 4
    a = torch.randn(B * rows, K, generator=g, dtype=torch.float64)
    W = torch.randn(N, K, generator=g, dtype=torch.float64, requires_grad=True)
    s = torch.tensor([1.25, 0.0, 1.25, 0.0, 1.25], dtype=torch.float64)
    do = torch.randn(B * rows, N, generator=g, dtype=torch.float64)
    ((a @ W.t()) * s.repeat_interleave(rows)[:, None]).backward(do)
    dWb = torch.bmm(do.view(B, rows, N).transpose(1, 2), a.view(B, rows, K))
    assert torch.allclose(torch.mv(dWb.view(B, N * K).t(), s).view(N, K), W.grad, atol=1e-12)


def test_fused_panel_row_order():
    """weight rows regrouped per pair of heads: Qin h0,h1 | Qout | Kout | Vout | Kin | Vin, d rows each (triplet_fused.cu)."""
    H, d, W = 4, 16, 64
    idx = ops._panel_rows(H, d, W, "cpu")
    assert idx.numel() == 6 * W and sorted(idx.tolist()) == list(range(6 * W))
    blocks = idx.view(H // 2, 12, d)
    bases = [0, 0, 3 * W, 3 * W, 4 * W, 4 * W, 5 * W, 5 * W, W, W, 2 * W, 2 * W]        # Qin, Qout, Kout, Vout, Kin, Vin
    for hp in range(H // 2):
        for slot in range(12):
            h = 2 * hp + (slot & 1)
            want = bases[slot] + h * d + torch.arange(d)
            assert torch.equal(blocks[hp, slot], want), (hp, slot)


def test_gemm_desc_mirrors_header():
    """field order / count of the ctypes GemmDesc == tgt_gemm_desc in include/tgt_b200.h, flag values too."""
    import os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "tgt_b200.h")).read()
    body = re.search(r"typedef struct \{([^}]*)\} tgt_gemm_desc;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.sub(r"[\*\s]", " ", part).split()[-1])
    assert names == [f[0] for f in _C.GemmDesc._fields_]
    for flag in ("LN", "BIAS", "GELU", "RES", "STORE_U", "ROWSCALE", "GELU_BWD", "STATS", "LN_BWD", "STORE_GP", "MULRES"):
        val = int(re.search(rf"#define TGT_EPI_{flag}\s+(\d+)", hdr).group(1))
        assert getattr(_C, f"EPI_{flag}") == val
    assert C.sizeof(_C.GemmDesc) % 8 == 0


def test_keep_projection_policy(monkeypatch):
    """k of the LAST layers keep their projection, decided from the live-byte growth between layer 0 and layer 1."""
    live = {"v": 0}
    monkeypatch.setattr(torch.cuda, "memory_allocated", lambda dev=None: live["v"])

    # what the process can still get = free device memory + the allocator's cached blocks (100 GiB here, of which
    # some other tenant of the GPU already holds the rest: it never enters the budget)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda dev=None: (60 << 30, 140 << 30))
    monkeypatch.setattr(torch.cuda, "memory_reserved", lambda dev=None: 40 << 30)
    monkeypatch.delenv("TGT_KEEP_PROJ_LAYERS", raising=False)
    ops._KEEP_STATE.clear()
    L, per_layer, proj = 10, 4 << 30, 3 << 30
    kept = []
    for i in range(L):
        live["v"] = (5 << 30) + i * per_layer
        ops.set_layer_hint(i, L)
        kept.append(ops.keep_projection(proj, "cuda:0", True))
    ops.set_layer_hint(None, None)
    # predicted end = 9 + 4*9 = 45 GiB; budget = 0.78*100 - 45 = 33 GiB -> k = 33 // (3 GiB + 64 MiB) = 10 -> capped at L-2
    assert kept == [False, False] + [True] * 8
    assert ops.keep_projection(proj, "cuda:0", True) is False            # no layer hint
    ops.set_layer_hint(5, L)
    assert ops.keep_projection(proj, "cuda:0", False) is False           # no gradient needed
    ops.set_layer_hint(None, None)
    monkeypatch.setenv("TGT_KEEP_PROJ_LAYERS", "0")
    ops._KEEP_STATE.clear()
    out = []
    for i in range(L):
        ops.set_layer_hint(i, L)
        out.append(ops.keep_projection(proj, "cuda:0", True))
    ops.set_layer_hint(None, None)
    assert not any(out)


def test_other_descriptors_mirror_header():
    """ctypes mirrors of the attention / aggregate / triangular / EGT descriptors: same field names in the same order."""
    import os
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "tgt_b200.h")).read()
    for cname, cls in (("tgt_triplet_attn_desc", _C.TripletAttnDesc), ("tgt_triplet_aggr_desc", _C.TripletAggrDesc),
                       ("tgt_triangular_desc", _C.TriangularDesc), ("tgt_egt_desc", _C.EgtDesc)):
        body = re.search(r"typedef struct \{([^}]*)\} " + cname + ";", hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[\d+\]", "", re.sub(r"[\*\s]", " ", part)).split()[-1])
        assert names == [f[0] for f in cls._fields_], cname


def test_two_stage_oracle_helpers():
    """bins_from_logits / bins2dist (oracle): symmetry, zero diagonal, the (bin + 0.5) * range / (bins - 1) * 2 scale."""
    from oracle import tgt_oracle as O
    g = torch.Generator().manual_seed(0)
    lg = torch.randn(2, 6, 6, 16, generator=g)
    b = O.bins_from_logits(lg)
    assert torch.equal(b, b.transpose(1, 2)) and b.dtype == torch.int64
    d = O.bins2dist(b, 16, 8.0)
    assert torch.equal(d, d.transpose(1, 2)) and float(d.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    i, j = 0, 3
    assert abs(float(d[0, i, j]) - 2 * (float(b[0, i, j]) + 0.5) * 8.0 / 15) < 1e-6
    d2 = O.bins2dist(b, 16, 8.0, shift_half=False, zero_diag=False)
    assert abs(float(d2[0, 2, 2]) - 2 * float(b[0, 2, 2]) * 8.0 / 15) < 1e-6


def test_cross_entropy_rows_falls_back_to_torch_off_gpu():
    import torch
    from tgt_b200 import ops
    logits, target = torch.randn(7, 16), torch.randint(0, 16, (7,))
    assert not ops.xent_rows_ok(logits, target)
    assert torch.allclose(ops.cross_entropy_rows(logits, target), torch.nn.functional.cross_entropy(logits, target, reduction='none'))
