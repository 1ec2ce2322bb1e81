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
# row-partitioned run: the last conv computes only the rows some rank's edge batch reads (DESIGN.md 4a item 3 for
# the partitioned encoder, ``pspmm_rows``).  PLNLP_PARTITIONED_RESTRICT=0 switches it off.
RESTRICT_LAST = os.environ.get("PLNLP_PARTITIONED_RESTRICT", "1") != "0"
# how the partial rows of the restricted last conv are combined: "rs" = reduce-scatter, linear map on 1/R of the rows,
# all-gather of the result; "ar" = all-reduce, every rank maps all rows
RESTRICT_COMBINE = os.environ.get("PLNLP_RESTRICT_COMBINE", "rs")


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


def allreduce_grads(params, group=None, average=False):
    """sum (``average``: mean) the gradients of ``params`` over ranks with ONE flat all-reduce.  Data-parallel
    edge batches: for a loss that is a SUM over pairs the summed gradient is the gradient of the global batch;
    for a MEAN loss (CE / LogRank / InfoNCE) it is the average."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    from . import profiling
    with profiling.span("nccl all_reduce (grads)", flat.numel() * 4, 0):
        dist.all_reduce(flat, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def _reduce_scatter_rows(full, blk, group=None, async_op=False, name="nccl reduce_scatter (rows)"):
    """sum ``full`` [R*blk, F] over ranks and return this rank's [blk, F] block.  ``async_op``: -> (block, wait)
    where ``wait()`` must be called before the block is read; the collective then runs beside whatever the caller
    enqueues in between."""
    rank, ws = world()
    out = torch.empty(blk, full.size(1), dtype=full.dtype, device=full.device)
    work = None
    if dist.get_backend(group) == "gloo":          # gloo has no reduce_scatter: test-only path
        dist.all_reduce(full, group=group)
        out.copy_(full[rank * blk:(rank + 1) * blk])
    else:
        from . import profiling
        full = full.contiguous()
        # bytes RECEIVED per rank: (R-1)/R of the full matrix (SURVEY 8d all-gather model)
        if async_op and not profiling.enabled():
            work = dist.reduce_scatter_tensor(out, full, group=group, async_op=True)
        else:
            with profiling.span(name, (ws - 1) * out.numel() * 4, 0):
                dist.reduce_scatter_tensor(out, full, group=group)
    if async_op:
        return out, ((lambda: work.wait()) if work is not None else (lambda: None))
    return out


# ---------------------------------------------------------------------------
# all-gather of row blocks over NVLink peer memory
# ---------------------------------------------------------------------------
# "p2p" (default where it works): every rank publishes its block in a symmetric-memory buffer (torch.distributed.
# _symmetric_memory: the same allocation mapped into every rank of the box) and PULLS the R-1 other blocks with
# peer-to-peer copies over NVLink -- copy engines, no SMs, one device-side barrier before and one after.  Measured
# against NCCL's all_gather for the [N / R, 50] embedding blocks: see DESIGN.md section 6.  "nccl": always NCCL.
ALLGATHER = os.environ.get("PLNLP_ALLGATHER", "p2p")
BARRIER_TIMEOUT_MS = int(os.environ.get("PLNLP_P2P_TIMEOUT_MS", "20000"))   # a lost peer traps instead of hanging
_SYMM = {}          # (group name, tag) -> (symmetric-memory handle, capacity in bytes)
_SYMM_BROKEN = []   # first failure: remember and use NCCL from then on


def _symm_handle(group, tag, nbytes, device):
    import torch.distributed._symmetric_memory as sm
    pg = group if group is not None else dist.group.WORLD
    gname = pg.group_name
    key = (gname, tag)
    hit = _SYMM.get(key)
    if hit is None or hit[1] < nbytes:
        try:
            sm.enable_symm_mem_for_group(gname)
        except Exception:
            pass                                   # newer torch enables groups implicitly
        cap = max(int(nbytes * 1.25), 1 << 20)
        t = sm.empty(cap, dtype=torch.uint8, device=device)
        hit = _SYMM[key] = (sm.rendezvous(t, gname), cap, t)
    return hit[0]


def _p2p_all_gather(x_local, group, tag, full):
    """pull-based all-gather on the CURRENT stream; x_local [blk, F] contiguous, full [R*blk, F]"""
    rank, ws = world()
    hdl = _symm_handle(group, tag, x_local.numel() * x_local.element_size(), x_local.device)
    key = (id(hdl), tuple(x_local.shape), x_local.dtype)
    peers = _PEER_VIEWS.get(key)
    if peers is None:                              # tensor views of every rank's buffer, built once per shape
        if len(_PEER_VIEWS) > 32:
            _PEER_VIEWS.clear()
        peers = _PEER_VIEWS[key] = [hdl.get_buffer(r, x_local.shape, x_local.dtype) for r in range(ws)]
    blk = x_local.size(0)
    peers[rank].copy_(x_local)                     # publish (a device-to-device copy of one block)
    full[rank * blk:(rank + 1) * blk].copy_(x_local)
    hdl.barrier(channel=0, timeout_ms=BARRIER_TIMEOUT_MS)      # every rank's block is in place
    for step in range(1, ws):
        r = (rank - step) % ws                     # each rank starts at a different peer
        full[r * blk:(r + 1) * blk].copy_(peers[r])
    hdl.barrier(channel=0, timeout_ms=BARRIER_TIMEOUT_MS)      # nobody republishes before every peer has read
    return full


_PEER_VIEWS = {}


def all_gather_rows(x_local, group=None, name="nccl all_gather (rows)", tag="rows", async_op=False):
    """[blk, F] row block -> [R*blk, F] (no autograd).  ``async_op``: -> (full, wait) -- the gather runs on a side
    stream and ``wait()`` makes the current stream wait for it."""
    _, ws = world()
    x_local = x_local.contiguous()
    full = torch.empty(ws * x_local.size(0), x_local.size(1), dtype=x_local.dtype, device=x_local.device)
    from . import profiling
    use_p2p = (ALLGATHER == "p2p" and not _SYMM_BROKEN and x_local.is_cuda and ws > 1
               and dist.get_backend(group) == "nccl")
    nbytes = (ws - 1) * x_local.numel() * 4
    if use_p2p:
        try:
            if async_op and not profiling.enabled():
                cur = torch.cuda.current_stream()
                side = _side_stream(x_local.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    # (the SAME buffer / barrier as the synchronous form: ranks may differ in which form they take --
                    # bench.py instruments rank 0 only -- and must still meet at one barrier)
                    _p2p_all_gather(x_local, group, tag, full)
                    ev = torch.cuda.Event()
                    ev.record(side)
                x_local.record_stream(side)
                full.record_stream(side)
                return full, (lambda: torch.cuda.current_stream().wait_event(ev))
            with profiling.span(name.replace("nccl", "p2p"), nbytes, 0):
                _p2p_all_gather(x_local, group, tag, full)
            return (full, (lambda: None)) if async_op else full
        except Exception as ex:                    # no symmetric memory on this box / torch: NCCL from now on
            _SYMM_BROKEN.append(repr(ex))
    if async_op and not profiling.enabled():
        work = dist.all_gather_into_tensor(full, x_local, group=group, async_op=True)
        return full, (lambda: work.wait())
    with profiling.span(name, nbytes, 0):
        dist.all_gather_into_tensor(full, x_local, group=group)
    return (full, (lambda: None)) if async_op else full


_SIDE = {}


def _side_stream(device):
    s = _SIDE.get(device)
    if s is None:
        s = _SIDE[device] = torch.cuda.Stream(device=device)
    return s


class GatherRows(torch.autograd.Function):
    """x_local [blk, F] (row block of a row-partitioned matrix, zero padded to blk) -> x_full [R*blk, F].
    Backward: reduce-scatter of the incoming gradient."""

    @staticmethod
    def forward(ctx, x_local, group, tag):
        ctx.group, ctx.blk, ctx.tag = group, x_local.size(0), tag
        return all_gather_rows(x_local, group, name=f"nccl all_gather ({tag})")

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_rows(g.contiguous(), ctx.blk, ctx.group, name=f"nccl reduce_scatter ({ctx.tag} grads)"), None, None


def gather_rows(x_local, group=None, tag="rows"):
    return GatherRows.apply(x_local, group, tag)


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


class AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x (every rank gets y).  Used where every rank holds a PARTIAL value of a replicated
    quantity and then continues with its own share of the loss: the total loss is the sum of the ranks' losses,
    so d loss / d x on every rank is the sum of the ranks' gradients w.r.t. y -- the backward is an all-reduce
    too."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.contiguous()
        if y.data_ptr() == x.data_ptr():
            ctx.mark_dirty(x)
        from . import profiling
        _, ws = world()
        with profiling.span("nccl all_reduce (restricted rows)", 2 * (ws - 1) * y.numel() * 4 // max(ws, 1), 0):
            dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        from . import profiling
        _, ws = world()
        with profiling.span("nccl all_reduce (restricted row grads)", 2 * (ws - 1) * g.numel() * 4 // max(ws, 1), 0):
            dist.all_reduce(g, group=ctx.group)
        return g, None


def all_reduce_sum(x, group=None):
    return AllReduceSum.apply(x, group)


class ReduceScatterSum(torch.autograd.Function):
    """x [T, F] (a partial value of a replicated quantity on every rank) -> this rank's block of
    ``ceil(T / R)`` rows of the SUM over ranks (zero rows past T).  Backward: the blocks' gradients are all-gathered
    -- every rank's partial value feeds every block."""

    @staticmethod
    def forward(ctx, x, group):
        _, ws = world()
        T = x.size(0)
        tb = (T + ws - 1) // ws
        ctx.group, ctx.T = group, T
        if tb * ws != T:
            x = torch.cat([x, x.new_zeros(tb * ws - T, x.size(1))], 0)
        return _reduce_scatter_rows(x.contiguous(), tb, group, name="nccl reduce_scatter (restricted rows)")

    @staticmethod
    def backward(ctx, g):
        full = all_gather_rows(g.contiguous(), ctx.group, name="nccl all_gather (restricted row grads)")
        return full[: ctx.T], None


def reduce_scatter_sum(x, group=None):
    return ReduceScatterSum.apply(x, group)


def union_ids(ids_local, group=None):
    """sorted distinct union over ranks of every rank's id list (all ranks must pass the same count -- the
    data-parallel edge batches have equal sizes) -> the same tensor on every rank.  One all-gather of the raw ids;
    ``torch.unique`` is the one host read (the size of the union)."""
    _, ws = world()
    ids_local = ids_local.contiguous()
    if ws == 1:
        return torch.unique(ids_local)
    allv = torch.empty(ws * ids_local.numel(), dtype=ids_local.dtype, device=ids_local.device)
    if ALLGATHER == "p2p" and not _SYMM_BROKEN and ids_local.is_cuda and dist.get_backend(group) == "nccl":
        # over peer memory, not NCCL: this runs on the side stream that prepares batch i + 1 while batch i is in
        # flight, and an NCCL collective would queue behind every collective of batch i on NCCL's own stream
        try:
            _p2p_all_gather(ids_local.view(-1, 1), group, "ids", allv.view(-1, 1))
            return torch.unique(allv)
        except Exception as ex:
            _SYMM_BROKEN.append(repr(ex))
    dist.all_gather_into_tensor(allv, ids_local, group=group)
    return torch.unique(allv)


class ShardedAdj:
    """Rows ``[lo, hi)`` of an adjacency A, columns in the padded global index space ``[0, R*blk)`` (``local``),
    plus the same rows of A^T (``local_t``; the SAME object when A is symmetric, as every prepared OGB graph
    is).  ``cols()`` is the COLUMN block A[:, lo:hi] = local_t^T as a transposed view."""

    def __init__(self, local_adj, n_global, rank, world_size, group=None, local_t=None):
        self.local = local_adj                  # CSRGraph-like, shape [blk, R*blk]
        self.local_t = local_t if local_t is not None else local_adj
        self.symmetric = local_t is None        # A == A^T entry for entry (see shard_graph)
        self.n_global, self.rank, self.world_size, self.group = n_global, rank, world_size, group
        self.blk = block_size(n_global, world_size)
        self._cols = None

    def size(self, dim):
        return self.local.size(dim)

    def cols(self):
        if self._cols is None:
            from .graph import TransposedAdj
            self._cols = TransposedAdj(self.local_t)
        return self._cols


def _row_slice(rowptr, col, val, n, rank, world_size, graph_cls):
    blk = block_size(n, world_size)
    lo, hi = row_block(n, rank, world_size)
    e0, e1 = int(rowptr[lo]), int(rowptr[hi])
    lptr = torch.full((blk + 1,), e1 - e0, dtype=torch.int64, device=rowptr.device)
    lptr[: hi - lo + 1] = rowptr[lo:hi + 1] - e0
    return graph_cls(lptr, col[e0:e1].clone(), None if val is None else val[e0:e1].clone(),
                     (blk, blk * world_size))


def transpose_csr(rowptr, col, val, n):
    """CSR arrays of A^T for a square [n, n] CSR matrix (entries of a column keep their row order)"""
    deg = rowptr[1:] - rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(n, device=col.device), deg, output_size=col.numel())
    perm = torch.argsort(col * n + row, stable=True)
    t_rowptr = torch.zeros(n + 1, dtype=torch.int64, device=col.device)
    if col.numel():
        t_rowptr[1:] = torch.cumsum(torch.bincount(col, minlength=n), 0)
    return t_rowptr, row[perm], None if val is None else val[perm], perm


def shard_graph(adj, rank, world_size, graph_cls, group=None, symmetric=None):
    """slice a full adjacency (``csr()`` / ``size()``) into this rank's ``ShardedAdj``.  Index work only;
    entries keep their order, so every local row is bit-identical to the corresponding global row.
    ``symmetric`` (None = check): A == A^T, entry for entry -- then the row block doubles as the row block of A^T."""
    rowptr, col, val = adj.csr()
    n = adj.size(0)
    local = _row_slice(rowptr, col, val, n, rank, world_size, graph_cls)
    local_t = None
    if symmetric is not True and adj.size(0) == adj.size(1):
        t_rowptr, t_col, t_val, _ = transpose_csr(rowptr, col, val, n)
        if symmetric is None:
            symmetric = bool(torch.equal(t_rowptr, rowptr) and torch.equal(t_col, col)
                             and (val is None or torch.equal(t_val, val)))
        if not symmetric:
            local_t = _row_slice(t_rowptr, t_col, t_val, n, rank, world_size, graph_cls)
    return ShardedAdj(local, n, rank, world_size, group, local_t=local_t)


def pad_rows(x, blk):
    if x.size(0) == blk:
        return x
    pad = torch.zeros(blk - x.size(0), x.size(1), dtype=x.dtype, device=x.device)
    return torch.cat([x, pad], 0)


def pspmm_rows(sadj, x_local, rows, reduce="sum", local_op=None, sharded=False, premask=None):
    """(A @ X)[rows] on a row-partitioned X, for GLOBAL row ids ``rows`` (sorted, distinct, the SAME list on every
    rank -- ``union_ids``), returned in full, compactly, on every rank.

    The last conv's output is read only at the endpoint rows of the edge batches.  Gathering the operand to the
    rows (all-gather of X: 2.05 GB received per rank for citation2-shape's [N, 200] activations) would move ten
    times more than the result is worth, so the product is formed the other way round: rank r multiplies the
    COLUMN block it owns the operand rows of, ``A[rows, lo:hi] @ X[lo:hi]`` -- a row-subset SpMM on the transposed
    view of its block of A^T -- and the [len(rows), F] partial results are summed over ranks (all-reduce:
    ~0.23 GB).  Backward: all-reduce of the compact gradient, then ``A[rows, lo:hi]^T @ g`` with the compact
    gradient as a row-sparse operand of the rank's own row block.  ``local_op(adj, x, rows, reduce)`` defaults to
    the CUDA row-subset SpMM (``_ops.spmm_rows``).  ``sharded``: see below."""
    if reduce not in ("sum", "add"):
        raise NotImplementedError("row-restricted partitioned products are sums (GCNConv)")
    if local_op is None:
        from . import _ops
        local_op = _ops.spmm_rows
    kw = {} if premask is None else {"premask": premask}      # see layer.GCNConv.forward(premask_input=)
    partial = local_op(sadj.cols(), pad_rows(x_local, sadj.blk), rows, "sum", **kw)
    if sharded:
        # -> this rank's ceil(T / R) rows of the sum only: the caller runs its row-wise work (the conv's linear map)
        # on 1 / R of the rows and all-gathers the result (``gather_rows``) -- same bytes on the wire as the
        # all-reduce, the dense layer and both of its backward GEMMs R times smaller
        return reduce_scatter_sum(partial, sadj.group)
    return all_reduce_sum(partial, sadj.group)


def pspmm(sadj, x_local, reduce="sum", local_op=None, **epilogue):
    """row-partitioned SpMM: out_local = A[lo:hi, :] @ all_gather(x_local).  ``local_op(adj, x, reduce,
    **epilogue)`` defaults to the CUDA SpMM (whose backward is the transposed kernel; the transposed
    product over ALL columns is then reduce-scattered by ``GatherRows.backward``)."""
    if local_op is None:
        from . import _ops
        local_op = _ops.spmm
    x_full = gather_rows(pad_rows(x_local, sadj.blk), sadj.group)
    return local_op(sadj.local, x_full, reduce, **epilogue)
