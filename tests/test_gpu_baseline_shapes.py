"""Parity at BASELINE.json's own graph shapes (-m gpu): the citation2-shape (2.93 M nodes, ~64 M stored entries after
symmetrisation + diagonal, hub rows split into many work items) and collab-shape adjacencies are too big for a full
CPU product in a test, so the kernels are compared with the in-order C oracle on SAMPLED rows -- every split hub row
plus a few thousand random ones -- for the forward product, the row-subset plan of the last conv and the row-sparse
backward; and through size-independent identities on the whole output."""
import pytest
import torch

import bench
from oracle import cspmm
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _shape_graph(name):
    from plnlp_b200.graph import CSRGraph, structure_of
    from plnlp_b200.utils import gcn_normalization
    cfg = dict(bench.WORKLOADS[name])
    torch.manual_seed(0)
    data, _ = bench.build_workload(cfg, torch.device("cuda"), CSRGraph, gcn_normalization)
    return cfg, data.adj_t, structure_of(data.adj_t)


def _sub_csr(rowptr, col, val, rows):
    """CPU CSR of just ``rows`` (int64 tensors) for the oracle"""
    cnt = rowptr[rows + 1] - rowptr[rows]
    sub_ptr = torch.zeros(rows.numel() + 1, dtype=torch.int64)
    sub_ptr[1:] = torch.cumsum(cnt, 0)
    idx = torch.repeat_interleave(rowptr[rows] - sub_ptr[:-1], cnt) + torch.arange(int(sub_ptr[-1]))
    return sub_ptr, col[idx], None if val is None else val[idx]


def _sample_rows(st, n_random, seed):
    g = torch.Generator().manual_seed(seed)
    hubs = st.fwd.fix_row.cpu().long()
    rnd = torch.randint(0, st.n_rows, (n_random,), generator=g)
    return torch.unique(torch.cat([hubs, rnd])), hubs


@pytest.mark.parametrize("shape,F,reduce", [("citation2", 50, "sum"), ("citation2", 200, "sum"), ("collab", 256, "mean")])
def test_spmm_sampled_rows_at_baseline_shape(shape, F, reduce):
    from plnlp_b200 import _ops
    cfg, adj, st = _shape_graph(shape)
    N = st.n_rows
    if shape == "citation2":
        assert st.fwd.n_fix >= 50 and st.symmetric                       # hub rows are split; A_hat is symmetric
    rowptr, col, val = (None if t is None else t.cpu() for t in adj.csr())
    rows, hubs = _sample_rows(st, 3000, F)
    g = torch.Generator().manual_seed(F)
    x = torch.randn(N, F, generator=g)
    xg = x.cuda()
    mean = reduce == "mean"
    plan = st.fwd_noval if mean else st.fwd
    full = _ops.spmm_raw(plan, xg, use_val=not mean, div_rows=mean)
    sp, sc, sv = _sub_csr(rowptr, col, None if mean else val, rows)
    want32 = cspmm.spmm(sp, sc, sv, x, reduce)
    want64 = cspmm.spmm(sp, sc, sv, x, reduce, f64=True)
    got = full[rows.cuda()].cpu()
    is_hub = torch.isin(rows, hubs)
    assert torch.equal(got[~is_hub], want32[~is_hub])                    # unsplit rows: the bits of the in-order loop
    assert rel_err(got, want64) < TOL                                    # split hub rows: fixed-order combine
    # ---- the row-subset plan of the last conv: ~10 % of the rows incl. every hub, compact output, identical bits
    sel = torch.unique(torch.cat([hubs, torch.randint(0, N, (N // 10,), generator=g)])).cuda()
    sub = _ops.spmm_rows(adj, xg, sel, reduce)
    assert torch.equal(sub, full[sel])
    # ---- its backward: the compact gradient as a row-sparse operand of the transposed plan == the dense product of
    # the gradient scattered into a zero matrix; and, on sampled rows, the oracle
    gc = torch.randn(sel.numel(), F, generator=g).cuda()
    xr = xg.clone().requires_grad_(True)
    _ops.spmm_rows(adj, xr, sel, reduce).backward(gc)
    dense_g = torch.zeros(N, F, device="cuda")
    dense_g[sel] = gc
    plan_t = st.bwd_mean if mean else st.bwd
    ref = _ops.spmm_raw(plan_t, dense_g, use_val=True if mean else st.has_value, div_rows=False)
    assert torch.equal(xr.grad, ref)
    if st.symmetric and not mean:                                        # A^T = A: the oracle rows apply directly
        want = cspmm.spmm(sp, sc, sv, dense_g.cpu(), "sum", f64=True)
        assert rel_err(xr.grad[rows.cuda()].cpu(), want) < TOL
    # ---- whole-output identities: linearity and adjointness
    y = torch.randn(N, F, generator=g).cuda()
    ay = _ops.spmm_raw(plan, y, use_val=not mean, div_rows=mean)
    lin = _ops.spmm_raw(plan, 2.0 * xg - 3.0 * y, use_val=not mean, div_rows=mean)
    assert rel_err(lin, 2.0 * full - 3.0 * ay) < TOL
    aty = _ops.spmm_raw(plan_t, y, use_val=True if mean else st.has_value, div_rows=False)
    lhs, rhs = float((full.double() * y.double()).sum()), float((xg.double() * aty.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs))


def test_tall_skinny_gemms_at_citation2_shape():
    """the encoder's dense layers at M = 2 927 963 on the TMA-fed kernel: sampled rows against fp64, and bit-identical
    to the CTA-pair kernel"""
    from plnlp_b200 import _ops
    M = 2927963
    g = torch.Generator(device="cuda").manual_seed(3)
    for N, K, tb in ((200, 178, True), (50, 200, False)):
        A = torch.randn(M, (K + 3) // 4 * 4, device="cuda", generator=g)[:, :K]
        B = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g)
        bias = torch.randn(N, device="cuda", generator=g)
        _ops.GEMM_TMA = "auto"
        C = _ops.gemm_raw(A, B, transb=tb, bias=bias, act=_ops.ACT_RELU)
        rows = torch.randint(0, M, (4096,), device="cuda", generator=g)
        rows[:3] = torch.tensor([0, M - 1, M - 2], device="cuda")
        want = torch.relu(A[rows].double() @ (B.t() if tb else B).double() + bias.double())
        assert rel_err(C[rows], want) < TOL
        _ops.GEMM_TMA = "0"
        try:
            C2 = _ops.gemm_raw(A, B, transb=tb, bias=bias, act=_ops.ACT_RELU)
        finally:
            _ops.GEMM_TMA = "auto"
        assert torch.equal(C, C2)
