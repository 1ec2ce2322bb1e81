"""Property tests (hypothesis, -m gpu): the CUDA kernels, through the C ABI, on generated graphs -- empty rows, isolated
nodes, self loops, duplicate edges, a hub row far longer than the chunk, N not a multiple of 32, the feature widths
of the recipes -- against the in-order CPU oracle."""
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import cspmm, plnlp_ref, sparse
from tests.helpers import rel_err
from tests.test_properties_cpu import graphs

pytestmark = pytest.mark.gpu
WIDTHS = [1, 3, 50, 64, 128, 178, 200, 256, 512]


def _gpu(o):
    from plnlp_b200.graph import CSRGraph
    rowptr, col, val = o.csr()
    return CSRGraph(rowptr.cuda(), col.cuda(), None if val is None else val.cuda(), o.sparse_sizes())


@settings(max_examples=40, deadline=None)
@given(graphs(max_n=70, max_e=400), st.sampled_from(WIDTHS), st.sampled_from(["sum", "mean"]),
       st.sampled_from([32, 1024]))
def test_spmm_kernel_matches_in_order_loop(g, F, reduce, chunk):
    """rows that fit one work item carry the bits of the in-order CPU loop; split (hub) rows are within 1e-5 of fp64;
    empty rows are exactly 0; the transposed product (backward) is the adjoint"""
    from plnlp_b200 import _ops
    from plnlp_b200.graph import Structure
    n, ei, w = g
    o = sparse.to_sparse_tensor(ei, w if reduce == "sum" else None, n)
    rowptr, col, val = o.csr()
    st_ = Structure(_gpu(o), chunk=chunk)
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(F + n))
    got = _ops.spmm_raw(st_.fwd, x.cuda(), use_val=(reduce == "sum"), div_rows=(reduce == "mean")).cpu()
    want = cspmm.spmm(rowptr, col, val, x, reduce)
    split = torch.zeros(n, dtype=torch.bool)
    split[st_.fwd.fix_row.cpu().long()] = True
    assert torch.equal(got[~split], want[~split])
    if split.any():
        assert rel_err(got[split], cspmm.spmm(rowptr, col, val, x, reduce, f64=True)[split]) < 1e-5
    assert torch.all(got[(rowptr[1:] - rowptr[:-1]) == 0] == 0)
    y = torch.randn(n, F, generator=torch.Generator().manual_seed(F)).cuda()
    plan_t = st_.bwd_mean if reduce == "mean" else st_.bwd
    aty = _ops.spmm_raw(plan_t, y, use_val=True if reduce == "mean" else st_.has_value, div_rows=False)
    lhs, rhs = float((got.cuda().double() * y.double()).sum()), float((x.cuda().double() * aty.double()).sum())
    # <Ax, y> = <x, A^T y>: both sides are sums with cancellation, so the yardstick is the size of the summands
    # (hypothesis found n = 51, F = 1 with |sum| = 0.02 against terms of order 1: 8e-7 apart, i.e. fp32 rounding)
    scale = float((got.cuda().double().abs() * y.double().abs()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1e-3) or abs(lhs - rhs) <= 1e-6 * scale


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 300), st.sampled_from([1, 3]),
       st.sampled_from(["AUC", "HingeAUC", "WeightedHingeAUC", "WeightedAUC", "AdaAUC", "AdaHingeAUC", "CE", "LogRank", "InfoNCE"]))
def test_pair_loss_kernel_any_batch(B, k, name):
    """loss value and d loss / d score for ragged batch sizes and num_neg in {1, 3} (loss.py), against the oracle's
    autograd"""
    from plnlp_b200 import _ops
    g = torch.Generator().manual_seed(B * 7 + k)
    pos, neg = torch.randn(B, generator=g), torch.randn(B * k, generator=g)
    wgt = torch.rand(B, generator=g) + 0.1 if ("Weighted" in name or "Ada" in name) else None
    loss, dpos, dneg = _ops.pair_loss_raw(_ops.LOSS_KINDS[name], pos.cuda(), neg.cuda(), k, None if wgt is None else wgt.cuda())
    rl, rdp, rdn = plnlp_ref.pair_loss_autograd(name, pos.double(), neg.double(), k, None if wgt is None else wgt.double())
    assert rel_err(loss.cpu(), rl.reshape(1)) < 1e-5
    scale = max(float(rdp.abs().max()), float(rdn.abs().max()), 1e-30)
    assert float((dpos.cpu().double() - rdp.reshape(-1)).abs().max()) <= 1e-5 * scale
    assert float((dneg.cpu().double() - rdn.reshape(-1)).abs().max()) <= 1e-5 * scale


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 200), st.integers(1, 400), st.sampled_from([1, 12, 50, 200]))
def test_edge_gather_kernels_any_shape(N, P, H):
    """endpoint gather + Hadamard (bit-exact), dot scores, and the deterministic scatter-add backward for any P / N"""
    from plnlp_b200 import _ops
    g = torch.Generator().manual_seed(N * 1000 + P)
    h = torch.randn(N, H, generator=g)
    edges = torch.randint(0, N, (P, 2), generator=g)
    want = h[edges[:, 0]] * h[edges[:, 1]]
    assert torch.equal(_ops.gather_hadamard_raw(h.cuda(), edges.cuda()).cpu(), want)
    da = torch.randn(P, H, generator=g)
    gh = _ops.edge_scatter_raw(h.cuda(), edges.cuda(), da=da.cuda()).cpu()
    ref = torch.zeros(N, H, dtype=torch.float64)
    ref.index_add_(0, edges[:, 0], (da * h[edges[:, 1]]).double())
    ref.index_add_(0, edges[:, 1], (da * h[edges[:, 0]]).double())
    assert rel_err(gh, ref) < 1e-5


@settings(max_examples=40, deadline=None)
@given(graphs(max_n=60, max_e=300), st.booleans())
def test_graph_build_kernels_match_oracle(g, weighted):
    """csrc/graph_build.cu behind CSRGraph on CUDA tensors -- ToSparseTensor (stable sort, duplicates kept),
    to_symmetric (equal keys merged, values summed), set_diag (stored diagonal dropped, unit diagonal inserted),
    D^-1/2 A D^-1/2 -- against the oracle's restatement of torch_sparse: index arrays bit-exact, values to rounding"""
    from plnlp_b200.graph import CSRGraph
    from plnlp_b200.utils import gcn_normalization
    n, ei, w = g

    def same(a, b, exact_values):
        (ra, ca, va), (rb, cb, vb) = a.csr(), b.csr()
        assert torch.equal(ra.cpu(), rb) and torch.equal(ca.cpu(), cb)
        assert (va is None) == (vb is None)
        if va is not None:
            if exact_values:
                assert torch.equal(va.cpu(), vb)
            elif vb.numel():
                assert float((va.cpu() - vb).abs().max()) <= 1e-6 * max(float(vb.abs().max()), 1e-30)

    a = CSRGraph.from_edge_index(ei.cuda(), w.cuda() if weighted else None, n)
    b = sparse.to_sparse_tensor(ei, w if weighted else None, n)
    same(a, b, True)
    same(a.to_symmetric(), b.to_symmetric(), False)            # merged values: summation order may differ
    same(a.set_diag(), b.set_diag(), True)
    same(gcn_normalization(a.to_symmetric()), sparse.gcn_normalization(b.to_symmetric()), False)
    same(gcn_normalization(CSRGraph.from_edge_index(ei.cuda(), None, n)), sparse.gcn_normalization(sparse.to_sparse_tensor(ei, None, n)), False)
