"""CPU tests (-m "not gpu") of the multi-GPU plumbing with torch.distributed gloo, world_size 2:
row partition bookkeeping, the all-gather / reduce-scatter autograd pair, the partitioned SpMM
(with the oracle as the injected local operator -- the CUDA kernel is exercised under gpurun) and
the flat gradient all-reduce.  Parity target: the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sparse
from plnlp_b200 import parallel
from tests.helpers import rand_graph, rel_err


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_graph(rowptr, col, val, sizes):
    return sparse.SparseTensor(rowptr=rowptr, col=col, value=val, sparse_sizes=sizes, is_sorted=True)


def _local_op(adj, x, reduce, **kw):
    return sparse.matmul(adj, x, reduce)


def _worker(rank, world_size, port, N, F, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        torch.manual_seed(0)
        ei, w = rand_graph(N, 400, seed=3, weighted=True, hub=True)
        full = sparse.to_sparse_tensor(ei, w, N)
        x = torch.randn(N, F)
        gout = torch.randn(N, F)
        lo, hi = parallel.row_block(N, rank, world_size)
        blk = parallel.block_size(N, world_size)
        sadj = parallel.shard_graph(full, rank, world_size, _oracle_graph)
        assert sadj.local.size(0) == blk and sadj.local.size(1) == blk * world_size
        # index work: local rows are the global rows, bit for bit
        rp, col, val = full.csr()
        lrp, lcol, lval = sadj.local.csr()
        assert torch.equal(lcol, col[rp[lo]:rp[hi]]) and torch.equal(lval, val[rp[lo]:rp[hi]])
        assert torch.equal(lrp[: hi - lo + 1], rp[lo:hi + 1] - rp[lo])
        out = {}
        for reduce in ("sum", "mean"):
            xl = x[lo:hi].clone().requires_grad_(True)
            adj_r = sadj if reduce == "sum" else parallel.ShardedAdj(
                _oracle_graph(lrp, lcol, None, sadj.local.sparse_sizes()), N, rank, world_size)
            y = parallel.pspmm(adj_r, xl, reduce, local_op=_local_op)
            assert y.shape == (blk, F)
            g = torch.zeros(blk, F)
            g[: hi - lo] = gout[lo:hi]
            y.backward(g)
            out[reduce] = (y.detach()[: hi - lo], xl.grad.clone())
        # flat gradient all-reduce
        p1, p2 = torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))
        p1.grad = torch.full((3, 2), float(rank + 1))
        p2.grad = torch.arange(5.0) * (rank + 1)
        parallel.allreduce_grads([p1, p2])
        assert torch.equal(p1.grad, torch.full((3, 2), 3.0)) and torch.equal(p2.grad, torch.arange(5.0) * 3)
        ret[rank] = {k: (v[0], v[1], lo, hi) for k, v in out.items()}
    finally:
        dist.destroy_process_group()


def test_row_partition_bookkeeping():
    assert parallel.block_size(10, 4) == 3
    assert [parallel.row_block(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert parallel.row_block(2, 3, 4) == (2, 2)                 # empty tail block
    blocks = [parallel.row_block(2927963, r, 8) for r in range(8)]
    assert blocks[0][0] == 0 and blocks[-1][1] == 2927963
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))


def test_partitioned_spmm_matches_single_process_gloo_ws2():
    N, F, ws = 37, 6, 2                                          # odd N: last block is padded
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(ws, port, N, F, ret), nprocs=ws, join=True)
    ei, w = rand_graph(N, 400, seed=3, weighted=True, hub=True)
    torch.manual_seed(0)
    x = torch.randn(N, F)
    gout = torch.randn(N, F)
    for reduce in ("sum", "mean"):
        full = sparse.to_sparse_tensor(ei, w if reduce == "sum" else None, N)
        xr = x.clone().requires_grad_(True)
        y = sparse.matmul(full, xr, reduce)
        y.backward(gout)
        for r in range(ws):
            yl, gl, lo, hi = ret[r][reduce]
            assert rel_err(yl, y.detach()[lo:hi]) < 1e-6
            assert rel_err(gl, xr.grad[lo:hi]) < 1e-5             # summation order differs across ranks


def _fetch_worker(rank, world_size, port, N, F, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(5)
        H = torch.randn(N, F, generator=g)
        blk = parallel.block_size(N, world_size)
        lo, hi = parallel.row_block(N, rank, world_size)
        h_local = parallel.pad_rows(H[lo:hi].clone(), blk).requires_grad_(True)
        gr = torch.Generator().manual_seed(100 + rank)
        # this rank's distinct endpoint rows: rank 1 asks for nothing owned by rank 0's first rows, both ask for
        # rows of both owners, one id is requested by every rank
        ids = torch.unique(torch.cat([torch.randint(0, N, (9 + 4 * rank,), generator=gr), torch.tensor([N - 1])]))
        rows = parallel.fetch_rows(
            h_local, ids, gather_fn=lambda h, i: h[i],
            scatter_fn=lambda gg, i, b: torch.zeros(b, gg.size(1)).index_add_(0, i, gg))
        assert torch.equal(rows.detach(), H[ids])
        gout = torch.randn(ids.numel(), F, generator=gr)
        rows.backward(gout)
        ret[rank] = (ids, gout, h_local.grad.clone(), lo, hi)
    finally:
        dist.destroy_process_group()


def test_fetch_rows_exchange_gloo_ws3():
    """the compact endpoint-row exchange of the partitioned scoring step: every rank receives exactly H[ids],
    and the owners' gradient blocks add up to the single-process index_add of all ranks' gradient rows"""
    N, F, ws = 23, 5, 3
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_fetch_worker, args=(ws, port, N, F, ret), nprocs=ws, join=True)
    want = torch.zeros(N, F)
    for r in range(ws):
        ids, gout, _, _, _ = ret[r]
        want.index_add_(0, ids, gout)
    for r in range(ws):
        _, _, grad_local, lo, hi = ret[r]
        assert rel_err(grad_local[: hi - lo], want[lo:hi]) < 1e-6
        assert torch.all(grad_local[hi - lo:] == 0)               # padding rows receive nothing


def _restricted_worker(rank, world_size, port, N, F, symmetric, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        full = _restricted_graph(N, symmetric)
        g = torch.Generator().manual_seed(6)
        X = torch.randn(N, F, generator=g)
        lo, hi = parallel.row_block(N, rank, world_size)
        sadj = parallel.shard_graph(full, rank, world_size, _oracle_graph)
        assert (sadj.local_t is sadj.local) == symmetric          # a symmetric matrix shares its row block
        x_local = X[lo:hi].clone().requires_grad_(True)
        gr = torch.Generator().manual_seed(200 + rank)
        mine = torch.cat([torch.randint(0, N, (8,), generator=gr), torch.tensor([1, N - 1])])   # equal counts
        ids = parallel.union_ids(mine)
        # every rank multiplies the COLUMN block it owns the operand rows of; the partial rows are all-reduced
        rows = parallel.pspmm_rows(
            sadj, x_local, ids, "sum",
            local_op=lambda adj, x, r, reduce: sparse.matmul(adj.base.t(), x, reduce)[r])
        want_rows = sparse.matmul(full, X, "sum")[ids]
        assert rows.shape == (ids.numel(), F) and rel_err(rows.detach(), want_rows) < 1e-6
        # the sharded form (reduce-scatter; a row-wise map on 1 / R of the rows; all-gather) gives the same rows
        shard = parallel.pspmm_rows(sadj, x_local.detach(), ids, "sum", sharded=True,
                                    local_op=lambda adj, x, r, reduce: sparse.matmul(adj.base.t(), x, reduce)[r])
        back = parallel.gather_rows(shard * 2.0, tag="restricted rows")[: ids.numel()]
        assert rel_err(back, 2.0 * want_rows) < 1e-6
        # each rank's loss reads its own endpoints only
        pos = torch.searchsorted(ids, mine)
        gout = torch.randn(mine.numel(), F, generator=gr)
        (rows[pos] * gout).sum().backward()
        ret[rank] = (ids, mine, gout, x_local.grad.clone(), lo, hi)
    finally:
        dist.destroy_process_group()


def _restricted_graph(N, symmetric):
    ei, w = rand_graph(N, 500, seed=4, weighted=True, hub=True)
    if symmetric:       # the prepared citation2 graph: symmetrised, unit diagonal, D^-1/2 A D^-1/2 (bit-symmetric values)
        return sparse.gcn_normalization(sparse.to_sparse_tensor(ei, None, N).to_symmetric())
    return sparse.to_sparse_tensor(ei, w, N)


@pytest.mark.parametrize("symmetric", [True, False])
def test_restricted_rows_column_block_products_gloo_ws3(symmetric):
    """the row-partitioned form of 'the last conv computes only the rows the batches read' (parallel.union_ids ->
    pspmm_rows): every rank ends up with exactly (A @ X)[ids] for the union of all ranks' endpoints, and the
    gradient w.r.t. every rank's block of X equals the single-process gradient of the summed losses"""
    N, F, ws = 31, 4, 3
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_restricted_worker, args=(ws, port, N, F, symmetric, ret), nprocs=ws, join=True)
    full = _restricted_graph(N, symmetric)
    X = torch.randn(N, F, generator=torch.Generator().manual_seed(6)).requires_grad_(True)
    Y = sparse.matmul(full, X, "sum")
    total = sum((Y[ret[r][1]] * ret[r][2]).sum() for r in range(ws))
    total.backward()
    ids0 = ret[0][0]
    for r in range(ws):
        ids, _, _, grad_local, lo, hi = ret[r]
        assert torch.equal(ids, ids0)                              # the same sorted union on every rank
        assert rel_err(grad_local, X.grad[lo:hi]) < 1e-5
