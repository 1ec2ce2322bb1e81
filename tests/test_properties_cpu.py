"""Property tests (hypothesis, CPU, -m "not gpu") of the host logic around the kernels, on generated graphs with the
shapes SURVEY.md section 4 lists: empty rows, isolated nodes, self loops, duplicate edges, a hub row far longer than
the chunk, node counts that are not multiples of 32.  The kernels themselves see the same generator on the GPU in
tests/test_gpu_properties.py."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import sparse
from plnlp_b200 import parallel
from plnlp_b200.graph import CSRGraph, build_plan, build_subset_plan
from plnlp_b200.utils import gcn_normalization
from tests.helpers import rel_err
from tests.test_host_logic import _emulate, _same


@st.composite
def graphs(draw, max_n=40, max_e=160):
    """edge list [2, E] with duplicates and self loops allowed, optional hub row, trailing isolated nodes"""
    n = draw(st.integers(2, max_n))
    live = draw(st.integers(1, n))                        # nodes >= live stay isolated
    e = draw(st.integers(0, max_e))
    src = draw(st.lists(st.integers(0, live - 1), min_size=e, max_size=e))
    dst = draw(st.lists(st.integers(0, live - 1), min_size=e, max_size=e))
    if draw(st.booleans()) and live > 1:                  # a hub: every node points at node 0, twice
        src += list(range(live)) * 2
        dst += [0] * (2 * live)
    ei = torch.tensor([src, dst], dtype=torch.int64).reshape(2, -1)
    w = torch.tensor(draw(st.lists(st.floats(0.25, 4.0, width=32), min_size=ei.size(1), max_size=ei.size(1))),
                     dtype=torch.float32)
    return n, ei, w


@settings(max_examples=40, deadline=None)
@given(graphs(), st.booleans())
def test_csrgraph_preparation_matches_oracle(g, weighted):
    """ToSparseTensor / to_symmetric / set_diag / gcn_normalization: same index arrays as the oracle's restatement
    of torch_sparse (bit-exact), values within fp32 rounding"""
    n, ei, w = g
    a = CSRGraph.from_edge_index(ei, w if weighted else None, n)
    b = sparse.to_sparse_tensor(ei, w if weighted else None, n)
    _same(a, b)
    _same(a.to_symmetric(), b.to_symmetric())
    _same(a.set_diag(), b.set_diag())
    _same(gcn_normalization(a.to_symmetric()), sparse.gcn_normalization(b.to_symmetric()))


@settings(max_examples=40, deadline=None)
@given(graphs(), st.sampled_from([1, 2, 4, 32]), st.sampled_from(["sum", "mean"]), st.integers(1, 7))
def test_plan_walk_equals_oracle_spmm(g, chunk, reduce, F):
    """the work plan (rows cut into <= chunk items, hub rows combined in a fixed-order second pass), walked exactly
    as csrc/spmm.cu walks it, reproduces the oracle's SpMM for every chunk size; empty rows give 0"""
    n, ei, w = g
    o = sparse.to_sparse_tensor(ei, w if reduce == "sum" else None, n)
    rowptr, col, val = o.csr()
    plan = build_plan(rowptr, col, val, n, n, chunk)
    assert int(plan.item_ptr[-1]) == col.numel() and plan.n_items >= n
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(F))
    got = _emulate(plan, x, use_val=(reduce == "sum"), div=(reduce == "mean"))
    assert not torch.isnan(got).any()
    assert rel_err(got, sparse.matmul(o, x, reduce)) < 1e-5
    empty = (rowptr[1:] - rowptr[:-1]) == 0
    assert torch.all(got[empty] == 0)


@settings(max_examples=40, deadline=None)
@given(graphs(), st.sampled_from([2, 32]), st.data())
def test_subset_plan_and_row_sparse_operand(g, chunk, data):
    """a row-subset plan yields exactly the selected rows of the full product, and an x_index with -1 entries the
    product of the operand with those rows zeroed (the two tricks behind the last conv)"""
    n, ei, w = g
    o = sparse.to_sparse_tensor(ei, w, n)
    rowptr, col, val = o.csr()
    parent = build_plan(rowptr, col, val, n, n, chunk)
    rows = torch.tensor(sorted(data.draw(st.sets(st.integers(0, n - 1), min_size=1, max_size=n))))
    x = torch.randn(n, 3, generator=torch.Generator().manual_seed(n))
    full = _emulate(parent, x, True, False)
    sub = _emulate(build_subset_plan(parent, rowptr, rows), x, True, False)
    assert torch.equal(sub, full[rows])
    keep = torch.tensor(data.draw(st.lists(st.booleans(), min_size=n, max_size=n)))
    x_index = torch.where(keep, torch.cumsum(keep.long(), 0) - 1, torch.full((n,), -1)).to(torch.int32)
    xz = x.clone()
    xz[~keep] = 0
    assert torch.equal(_emulate(parent, x[keep], True, False, x_index=x_index), _emulate(parent, xz, True, False))


@settings(max_examples=40, deadline=None)
@given(graphs(), st.integers(1, 5))
def test_row_partition_tiles_the_matrix(g, ws):
    """row blocks of every rank, stacked, are the matrix (entries in their original order, padding rows empty); the
    column block view used by the restricted last conv is the transpose of the row block of A^T"""
    n, ei, w = g
    full = sparse.to_sparse_tensor(ei, w, n)
    rowptr, col, val = full.csr()
    blk = parallel.block_size(n, ws)
    cols, vals, cnt = [], [], []
    mk = lambda rp, c, v, sizes: sparse.SparseTensor(rowptr=rp, col=c, value=v, sparse_sizes=sizes, is_sorted=True)  # noqa: E731
    dense = full.to_dense()
    for r in range(ws):
        sadj = parallel.shard_graph(full, r, ws, mk)
        lrp, lcol, lval = sadj.local.csr()
        assert sadj.local.size(0) == blk and sadj.local.size(1) == blk * ws
        lo, hi = parallel.row_block(n, r, ws)
        assert torch.all(lrp[hi - lo:] == lrp[hi - lo])                      # padding rows are empty
        cols.append(lcol); vals.append(lval); cnt.append(lrp[1:hi - lo + 1] - lrp[:hi - lo])
        # column block A[:, lo:hi] == (rows lo:hi of A^T)^T
        t_dense = sadj.local_t.to_dense()[: hi - lo, :n]
        assert torch.allclose(t_dense.t(), dense[:, lo:hi])
    assert torch.equal(torch.cat(cols), col) and torch.equal(torch.cat(vals), val)
    assert torch.equal(torch.cat(cnt), rowptr[1:] - rowptr[:-1])


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 300), st.integers(1, 9))
def test_row_blocks_partition_the_index_range(n, ws):
    blocks = [parallel.row_block(n, r, ws) for r in range(ws)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert all(0 <= hi - lo <= parallel.block_size(n, ws) for lo, hi in blocks)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 200), st.sampled_from([1, 3]), st.integers(1, 64))
def test_batches_cover_every_training_edge_once(E, num_neg, batch):
    """model.py:147-171 batch loop over a shuffled epoch: every positive (and its num_neg negatives) is visited
    exactly once, the last batch may be ragged"""
    order = torch.randperm(E, generator=torch.Generator().manual_seed(E))
    perms = [order[i:i + batch] for i in range(0, E, batch)]
    assert sum(p.numel() for p in perms) == E and perms[-1].numel() == (E % batch or min(batch, E))
    neg = torch.arange(E * num_neg * 2).reshape(E, num_neg, 2)
    seen = torch.cat([neg[p].reshape(-1, 2) for p in perms])
    assert torch.equal(torch.sort(seen[:, 0])[0], torch.sort(neg.reshape(-1, 2)[:, 0])[0])
