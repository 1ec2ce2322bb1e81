"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL / NVLink).

The reference has no multi-device code; the design here follows SURVEY.md
section 8e and is checked against the single-device result:

* small graphs (ddi / collab shape): the encoder is replicated, EDGE BATCHES are data parallel and the
  gradients are all-reduced (``allreduce_grads``).
* citation2-shape: the full-graph encoder is ROW PARTITIONED.  Rank r owns the contiguous block of
  ``blk = ceil(N / R)`` nodes ``[r*blk, (r+1)*blk)``: those rows of the embedding table / features /
  activations and those rows of the adjacency (all columns).  Per layer and direction there is exactly
  one collective: ``all_gather`` of the SpMM operand in forward, ``reduce_scatter`` of the transposed
  product in backward (``pspmm``).  Scoring gathers ``h`` once per step (``gather_rows``) and each rank
  scores its share of the edge batch; dense-weight gradients are all-reduced, embedding rows need no
  collective (owner computes).

Everything here is device agnostic torch.distributed code so the bookkeeping is testable on CPU with
the gloo backend (tests/test_parallel_cpu.py); the local SpMM is the CUDA kernel unless a test injects
another local operator.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

# how the scoring step of the row-partitioned model gets its endpoint embeddings: "rows" = only the distinct
# endpoint rows of the batch, by all_to_all (FetchRows); "allgather" = the whole matrix
EXCHANGE = os.environ.get("PLNLP_EXCHANGE", "rows")
# row-partitioned run: exchange the row requests BEFORE the last conv and let every owner compute only the rows
# that were requested (DESIGN.md 4a item 3 for the partitioned encoder).  Bookkeeping verified on CPU (gloo,
# tests/test_parallel_cpu.py); OFF by default until it has been through the 2-rank NCCL parity test on GPUs.
RESTRICT_LAST = os.environ.get("PLNLP_PARTITIONED_RESTRICT", "0") == "1"


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def block_size(n, world_size):
    return (n + world_size - 1) // world_size


def row_block(n, rank, world_size):
    """[lo, hi) of the rows rank owns (the last blocks may be short or empty)."""
    blk = block_size(n, world_size)
    lo = min(rank * blk, n)
    return lo, min(lo + blk, n)


def allreduce_grads(params, group=None):
    """sum the gradients of ``params`` over ranks with ONE flat all-reduce (weak-scaling edge batches:
    the loss is a sum over pairs, so the summed gradient is the gradient of the global batch)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    from . import profiling
    with profiling.span("nccl all_reduce (grads)", flat.numel() * 4, 0):
        dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def _reduce_scatter_rows(full, blk, group=None):
    """sum ``full`` [R*blk, F] over ranks and return this rank's [blk, F] block."""
    rank, ws = world()
    out = torch.empty(blk, full.size(1), dtype=full.dtype, device=full.device)
    if dist.get_backend(group) == "gloo":          # gloo has no reduce_scatter: test-only path
        dist.all_reduce(full, group=group)
        out.copy_(full[rank * blk:(rank + 1) * blk])
    else:
        from . import profiling
        # bytes RECEIVED per rank: (R-1)/R of the full matrix (SURVEY 8d all-gather model)
        with profiling.span("nccl reduce_scatter (rows)", (ws - 1) * out.numel() * 4, 0):
            dist.reduce_scatter_tensor(out, full.contiguous(), group=group)
    return out


def all_gather_rows(x_local, group=None):
    """[blk, F] row block -> [R*blk, F] (no autograd)"""
    _, ws = world()
    full = torch.empty(ws * x_local.size(0), x_local.size(1), dtype=x_local.dtype, device=x_local.device)
    from . import profiling
    with profiling.span("nccl all_gather (rows)", (ws - 1) * x_local.numel() * 4, 0):
        dist.all_gather_into_tensor(full, x_local.contiguous(), group=group)
    return full


class GatherRows(torch.autograd.Function):
    """x_local [blk, F] (row block of a row-partitioned matrix, zero padded to blk) -> x_full [R*blk, F].
    Backward: reduce-scatter of the incoming gradient."""

    @staticmethod
    def forward(ctx, x_local, group):
        ctx.group, ctx.blk = group, x_local.size(0)
        return all_gather_rows(x_local, group)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_rows(g.contiguous(), ctx.blk, ctx.group), None


def gather_rows(x_local, group=None):
    return GatherRows.apply(x_local, group)


class FetchRows(torch.autograd.Function):
    """rows = H[ids] for a ROW-PARTITIONED matrix H (rank r owns rows [r*blk, (r+1)*blk) as ``h_local``) and a
    sorted list ``ids`` of the distinct global rows THIS rank needs -- the endpoints of its edge batch
    (a batch touches at most 2P = 524 288 of citation2-shape's 2.9 M rows, so moving just those
    rows costs ~5x less NVLink traffic than all-gathering H, 2.34 GB).

    Forward: the id lists go to their owners (all_to_all), every owner gathers the requested rows of its block
    (``gather_fn``: the CUDA row-gather kernel), the rows travel back (all_to_all).  Backward: the gradient
    rows travel the same way in reverse and every owner sums them into its block in a fixed order
    (``scatter_fn``: the sorted segment-sum kernel) -> deterministic.  One host read per call (the split sizes
    of the variable-length exchange)."""

    @staticmethod
    def forward(ctx, h_local, ids, group, gather_fn, scatter_fn):
        rank, ws = world()
        blk = h_local.size(0)
        dev = h_local.device
        bounds = torch.arange(ws + 1, device=dev, dtype=ids.dtype) * blk
        cut = torch.searchsorted(ids, bounds)                     # ids are sorted: owner o holds ids[cut[o]:cut[o+1]]
        send_cnt = (cut[1:] - cut[:-1]).to(torch.int64)
        recv_cnt = torch.empty_like(send_cnt)
        dist.all_to_all_single(recv_cnt, send_cnt, group=group)
        both = torch.stack([send_cnt, recv_cnt]).tolist()         # the one host read
        send_split, recv_split = both[0], both[1]
        owner = torch.repeat_interleave(torch.arange(ws, device=dev, dtype=ids.dtype), send_cnt,
                                        output_size=ids.numel())
        want = torch.empty(sum(recv_split), dtype=ids.dtype, device=dev)
        dist.all_to_all_single(want, ids - owner * blk, recv_split, send_split, group=group)
        served = gather_fn(h_local, want)                         # [n_served, F]: rows other ranks asked me for
        rows = torch.empty(ids.numel(), h_local.size(1), dtype=h_local.dtype, device=dev)
        from . import profiling
        with profiling.span("nccl all_to_all (endpoint rows)", (ids.numel() - send_split[rank]) * h_local.size(1) * 4, 0):
            dist.all_to_all_single(rows, served, send_split, recv_split, group=group)
        ctx.group, ctx.blk, ctx.scatter_fn = group, blk, scatter_fn
        ctx.splits = (send_split, recv_split)
        ctx.save_for_backward(want)
        return rows

    @staticmethod
    def backward(ctx, g):
        (want,) = ctx.saved_tensors
        send_split, recv_split = ctx.splits
        back = torch.empty(want.numel(), g.size(1), dtype=g.dtype, device=g.device)
        rank, _ = world()
        from . import profiling
        with profiling.span("nccl all_to_all (endpoint row grads)", (g.size(0) - send_split[rank]) * g.size(1) * 4, 0):
            dist.all_to_all_single(back, g.contiguous(), recv_split, send_split, group=ctx.group)
        return ctx.scatter_fn(back, want, ctx.blk), None, None, None, None


def fetch_rows(h_local, ids, group=None, gather_fn=None, scatter_fn=None):
    """see ``FetchRows``; the default gather / scatter are the CUDA kernels"""
    if gather_fn is None or scatter_fn is None:
        from . import _ops
        gather_fn = gather_fn or _ops.gather_rows_idx_raw
        scatter_fn = scatter_fn or _ops.row_scatter_raw
    return FetchRows.apply(h_local, ids, group, gather_fn, scatter_fn)


class RowRequests:
    """result of ``exchange_row_requests``: ``want`` = the local row ids other ranks (and this one) asked THIS
    rank for, concatenated in requester order; the split sizes of the exchange in both directions"""

    def __init__(self, want, send_split, recv_split, n_ids):
        self.want, self.send_split, self.recv_split, self.n_ids = want, send_split, recv_split, n_ids


def exchange_row_requests(ids, blk, group=None):
    """phase 1 of the endpoint-row exchange on its own: every rank sends the sorted distinct global row ids it
    needs to their owners.  Knowing ``want`` BEFORE the last conv runs lets the owner compute only the requested
    rows (``pspmm_rows``) and then serve them (``serve_rows``).  One host read (the split sizes)."""
    _, ws = world()
    dev = ids.device
    bounds = torch.arange(ws + 1, device=dev, dtype=ids.dtype) * blk
    cut = torch.searchsorted(ids, bounds)
    send_cnt = (cut[1:] - cut[:-1]).to(torch.int64)
    recv_cnt = torch.empty_like(send_cnt)
    dist.all_to_all_single(recv_cnt, send_cnt, group=group)
    both = torch.stack([send_cnt, recv_cnt]).tolist()
    send_split, recv_split = both[0], both[1]
    owner = torch.repeat_interleave(torch.arange(ws, device=dev, dtype=ids.dtype), send_cnt, output_size=ids.numel())
    want = torch.empty(sum(recv_split), dtype=ids.dtype, device=dev)
    dist.all_to_all_single(want, ids - owner * blk, recv_split, send_split, group=group)
    return RowRequests(want, send_split, recv_split, ids.numel())


class ServeRows(torch.autograd.Function):
    """phase 2: ``rows[i] = src[pos[j]]`` for every request j this rank received, delivered to the requester.
    ``src`` is whatever table the owner holds the requested rows in -- its whole block (pos = want) or the compact
    output of a row-restricted last conv (pos = position of want in the sorted distinct requested rows).
    Backward: gradient rows return to the owner and are summed into ``src``'s rows in a fixed order."""

    @staticmethod
    def forward(ctx, src, pos, req, group, gather_fn, scatter_fn):
        rank, _ = world()
        served = gather_fn(src, pos)
        rows = torch.empty(req.n_ids, src.size(1), dtype=src.dtype, device=src.device)
        from . import profiling
        with profiling.span("nccl all_to_all (endpoint rows)", (req.n_ids - req.send_split[rank]) * src.size(1) * 4, 0):
            dist.all_to_all_single(rows, served, req.send_split, req.recv_split, group=group)
        ctx.group, ctx.n_src, ctx.scatter_fn, ctx.req = group, src.size(0), scatter_fn, req
        ctx.save_for_backward(pos)
        return rows

    @staticmethod
    def backward(ctx, g):
        (pos,) = ctx.saved_tensors
        req = ctx.req
        rank, _ = world()
        back = torch.empty(pos.numel(), g.size(1), dtype=g.dtype, device=g.device)
        from . import profiling
        with profiling.span("nccl all_to_all (endpoint row grads)", (g.size(0) - req.send_split[rank]) * g.size(1) * 4, 0):
            dist.all_to_all_single(back, g.contiguous(), req.recv_split, req.send_split, group=ctx.group)
        return ctx.scatter_fn(back, pos, ctx.n_src), None, None, None, None, None


def serve_rows(src, pos, req, group=None, gather_fn=None, scatter_fn=None):
    if gather_fn is None or scatter_fn is None:
        from . import _ops
        gather_fn = gather_fn or _ops.gather_rows_idx_raw
        scatter_fn = scatter_fn or _ops.row_scatter_raw
    return ServeRows.apply(src, pos, req, group, gather_fn, scatter_fn)


class ShardedAdj:
    """Rows ``[lo, hi)`` of an adjacency, columns in the padded global index space ``[0, R*blk)``."""

    def __init__(self, local_adj, n_global, rank, world_size, group=None):
        self.local = local_adj                  # CSRGraph-like, shape [blk, R*blk]
        self.n_global, self.rank, self.world_size, self.group = n_global, rank, world_size, group
        self.blk = block_size(n_global, world_size)

    def size(self, dim):
        return self.local.size(dim)


def shard_graph(adj, rank, world_size, graph_cls, group=None):
    """slice a full adjacency (``csr()`` / ``size()``) into this rank's ``ShardedAdj``.  Index work only;
    entries keep their order, so every local row is bit-identical to the corresponding global row."""
    rowptr, col, val = adj.csr()
    n = adj.size(0)
    blk = block_size(n, world_size)
    lo, hi = row_block(n, rank, world_size)
    e0, e1 = int(rowptr[lo]), int(rowptr[hi])
    lptr = torch.full((blk + 1,), e1 - e0, dtype=torch.int64, device=rowptr.device)
    lptr[: hi - lo + 1] = rowptr[lo:hi + 1] - e0
    local = graph_cls(lptr, col[e0:e1].clone(), None if val is None else val[e0:e1].clone(),
                      (blk, blk * world_size))
    return ShardedAdj(local, n, rank, world_size, group)


def pad_rows(x, blk):
    if x.size(0) == blk:
        return x
    pad = torch.zeros(blk - x.size(0), x.size(1), dtype=x.dtype, device=x.device)
    return torch.cat([x, pad], 0)


def pspmm_rows(sadj, x_local, rows_local, reduce="sum", local_op=None):
    """row-partitioned SpMM restricted to the output rows ``rows_local`` (local ids, sorted, distinct) of this
    rank's block, written compactly: (A[lo:hi, :] @ all_gather(x_local))[rows_local].  ``local_op(adj, x, rows,
    reduce)`` defaults to the CUDA row-subset SpMM (``_ops.spmm_rows``)."""
    if local_op is None:
        from . import _ops
        local_op = _ops.spmm_rows
    x_full = gather_rows(pad_rows(x_local, sadj.blk), sadj.group)
    return local_op(sadj.local, x_full, rows_local, reduce)


def pspmm(sadj, x_local, reduce="sum", local_op=None, **epilogue):
    """row-partitioned SpMM: out_local = A[lo:hi, :] @ all_gather(x_local).  ``local_op(adj, x, reduce,
    **epilogue)`` defaults to the CUDA SpMM (whose backward is the transposed kernel; the transposed
    product over ALL columns is then reduce-scattered by ``GatherRows.backward``)."""
    if local_op is None:
        from . import _ops
        local_op = _ops.spmm
    x_full = gather_rows(pad_rows(x_local, sadj.blk), sadj.group)
    return local_op(sadj.local, x_full, reduce, **epilogue)
