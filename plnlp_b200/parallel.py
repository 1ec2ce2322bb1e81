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

import torch
import torch.distributed as dist


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


class GatherRows(torch.autograd.Function):
    """x_local [blk, F] (row block of a row-partitioned matrix, zero padded to blk) -> x_full [R*blk, F].
    Backward: reduce-scatter of the incoming gradient."""

    @staticmethod
    def forward(ctx, x_local, group):
        _, ws = world()
        ctx.group, ctx.blk = group, x_local.size(0)
        full = torch.empty(ws * x_local.size(0), x_local.size(1), dtype=x_local.dtype, device=x_local.device)
        from . import profiling
        with profiling.span("nccl all_gather (rows)", (ws - 1) * x_local.numel() * 4, 0):
            dist.all_gather_into_tensor(full, x_local.contiguous(), group=group)
        return full

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_rows(g.contiguous(), ctx.blk, ctx.group), None


def gather_rows(x_local, group=None):
    return GatherRows.apply(x_local, group)


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


def pspmm(sadj, x_local, reduce="sum", local_op=None, **epilogue):
    """row-partitioned SpMM: out_local = A[lo:hi, :] @ all_gather(x_local).  ``local_op(adj, x, reduce,
    **epilogue)`` defaults to the CUDA SpMM (whose backward is the transposed kernel; the transposed
    product over ALL columns is then reduce-scattered by ``GatherRows.backward``)."""
    if local_op is None:
        from . import _ops
        local_op = _ops.spmm
    x_full = gather_rows(pad_rows(x_local, sadj.blk), sadj.group)
    return local_op(sadj.local, x_full, reduce, **epilogue)
