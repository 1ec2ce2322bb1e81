"""Adjacency holder and the per-graph structure cache behind the SpMM kernels.

``CSRGraph`` is the adjacency object of this package.  It exposes the subset of the
``torch_sparse.SparseTensor`` surface the reference touches (``csr()``, ``coo()``,
``size()``, ``set_value()``, ``to_symmetric()``, ``set_diag()``, ``sum(dim=1)``, row / column
broadcast multiplication, ``t()``, ``to()``), built from torch index ops on whatever device
the tensors live on, so ``/root/reference/main.py:81-83,109-110,177-179`` style graph
preparation works on it unchanged.  A real ``torch_sparse.SparseTensor`` is accepted anywhere
a ``CSRGraph`` is (duck-typed through ``.csr()`` / ``.size()``).

``structure_of(adj)`` returns the cached device-side structure the kernels consume: int32
column indices, fp32 values, the SpMM work plan (rows cut into <= chunk slices so hub rows
of power-law graphs do not serialise one warp), and the same for the transposed matrix used by
the backward pass.  The adjacency is constant for a whole run (main.py:175-186), so this is
built once.
"""
from __future__ import annotations

import os

import torch

# allow the dense tensor-core path for small dense-ish graphs (see Structure.dense_ok)
DENSE_SPMM = os.environ.get("PLNLP_DENSE_SPMM", "1") != "0"


class CSRGraph:
    def __init__(self, rowptr, col, value=None, sparse_sizes=None, row=None):
        self._rowptr = rowptr.to(torch.int64)
        self._col = col.to(torch.int64)
        self._value = value
        if sparse_sizes is None:
            sparse_sizes = (rowptr.numel() - 1, rowptr.numel() - 1)
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        self._row = row

    # ---- construction -----------------------------------------------------
    @staticmethod
    def from_coo(row, col, value=None, sparse_sizes=None, is_sorted=False):
        """SparseTensor(row=, col=, value=) (main.py:124-126): sort by (row, col), keep duplicates."""
        if sparse_sizes is None:
            n = int(torch.max(torch.stack([row.max(), col.max()]))) + 1 if row.numel() else 0
            sparse_sizes = (n, n)
        M, N = int(sparse_sizes[0]), int(sparse_sizes[1])
        row, col = row.to(torch.int64), col.to(torch.int64)
        if row.is_cuda and not is_sorted and 0 < row.numel() < 2 ** 31:
            # hand-written path (csrc/graph_build.cu): one stable radix sort of (row * N + col, position) pairs, then
            # row pointers and column indices straight from the sorted keys
            keys, pos = _gb_make_keys(row, col, N, both=False, drop_diag=False)
            keys, perm = _gb_sort(keys, pos, M, N, has_dead=False)
            rowptr, col = _gb_keys_to_csr(keys, M, N)
            return CSRGraph(rowptr, col, None if value is None else value[perm], (M, N))
        if not is_sorted:
            perm = torch.argsort(row * N + col, stable=True)
            row, col = row[perm], col[perm]
            if value is not None:
                value = value[perm]
        rowptr = torch.zeros(M + 1, dtype=torch.int64, device=row.device)
        if row.numel():
            rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=M), 0)
        return CSRGraph(rowptr, col, value, (M, N), row=row)

    @staticmethod
    def from_edge_index(edge_index, edge_weight=None, num_nodes=None):
        """``T.ToSparseTensor()`` (main.py:81): adj_t[dst, src]."""
        N = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
        return CSRGraph.from_coo(edge_index[1], edge_index[0], edge_weight, (N, N))

    # ---- torch_sparse-like accessors ---------------------------------------
    def size(self, dim):
        return self._sizes[dim]

    def sizes(self):
        return list(self._sizes)

    def sparse_sizes(self):
        return self._sizes

    def nnz(self):
        return self._col.numel()

    @property
    def device(self):
        return self._col.device

    def has_value(self):
        return self._value is not None

    def csr(self):
        return self._rowptr, self._col, self._value

    def _rows(self):
        if self._row is None:
            cnt = self._rowptr[1:] - self._rowptr[:-1]
            self._row = torch.repeat_interleave(
                torch.arange(self._sizes[0], device=self._col.device), cnt, output_size=self._col.numel())
        return self._row

    def coo(self):
        return self._rows(), self._col, self._value

    def set_value(self, value, layout=None):
        return CSRGraph(self._rowptr, self._col, value, self._sizes, row=self._row)

    def to(self, device, *args, **kwargs):
        mv = lambda t: None if t is None else t.to(device)  # noqa: E731
        return CSRGraph(mv(self._rowptr), mv(self._col), mv(self._value), self._sizes, row=mv(self._row))

    def cuda(self):
        return self.to("cuda")

    def t(self):
        row, col, value = self.coo()
        return CSRGraph.from_coo(col, row, value, (self._sizes[1], self._sizes[0]))

    def to_symmetric(self):
        """main.py:110: union of (r,c) and (c,r), sorted, duplicates merged (values summed)."""
        N = max(self._sizes)
        row, col, value = self.coo()
        if col.is_cuda and 0 < col.numel() < 2 ** 30 and (value is None or value.dtype == torch.float32):
            keys, pos = _gb_make_keys(row, col, N, both=True, drop_diag=False)
            keys, pos = _gb_sort(keys, pos, N, N, has_dead=False)
            ukeys, uval = _gb_unique(keys, pos, value, col.numel())
            rowptr, ucol = _gb_keys_to_csr(ukeys, N, N)
            return CSRGraph(rowptr, ucol, uval, (N, N))
        key = torch.cat([row * N + col, col * N + row])
        if value is None:
            key = torch.unique(key)
            return CSRGraph.from_coo(torch.div(key, N, rounding_mode="floor"), key % N, None, (N, N), True)
        ukey, inv = torch.unique(key, return_inverse=True)
        v = torch.zeros(ukey.numel(), dtype=value.dtype, device=value.device)
        v.index_add_(0, inv, torch.cat([value, value]))
        return CSRGraph.from_coo(torch.div(ukey, N, rounding_mode="floor"), ukey % N, v, (N, N), True)

    def set_diag(self):
        """utils.py:84: drop the stored diagonal, add one unit entry per row."""
        M, N = self._sizes
        row, col, value = self.coo()
        if col.is_cuda and 0 < col.numel() < 2 ** 31 - max(M, N):
            # key space: stored diagonal entries get a dead key, one unit entry per row is appended, one sort
            n, nd = col.numel(), min(M, N)
            keys = torch.empty(n + nd, dtype=torch.int64, device=col.device)
            pos = torch.empty(n + nd, dtype=torch.int64, device=col.device)
            _gb_make_keys(row, col, N, both=False, drop_diag=True, out=(keys, pos))
            _gb_diag_keys(nd, N, keys[n:], pos[n:], n)
            n_dead = int((row == col).sum())
            keys, pos = _gb_sort(keys, pos, M, N, has_dead=n_dead > 0)
            keys, pos = keys[: n + nd - n_dead], pos[: n + nd - n_dead]
            rowptr, ncol = _gb_keys_to_csr(keys, M, N)
            nval = None
            if value is not None:
                nval = torch.cat([value, torch.ones(nd, dtype=value.dtype, device=col.device)])[pos]
            return CSRGraph(rowptr, ncol, nval, (M, N))
        keep = row != col
        d = torch.arange(min(M, N), dtype=torch.int64, device=col.device)
        nval = None
        if value is not None:
            nval = torch.cat([value[keep], torch.ones(d.numel(), dtype=value.dtype, device=col.device)])
        return CSRGraph.from_coo(torch.cat([row[keep], d]), torch.cat([col[keep], d]), nval, (M, N))

    def sum(self, dim=1):
        if dim != 1:
            raise NotImplementedError("only row sums are used by the reference (utils.py:85,93)")
        if self._value is None:
            return self._rowptr[1:] - self._rowptr[:-1]
        out = torch.zeros(self._sizes[0], dtype=self._value.dtype, device=self._value.device)
        return out.index_add_(0, self._rows(), self._value)

    def _scaled(self, dense):
        M, N = self._sizes
        if dense.dim() == 2 and dense.size(0) == M and dense.size(1) == 1:
            f = dense.reshape(-1)[self._rows()]
        elif dense.dim() == 2 and dense.size(0) == 1 and dense.size(1) == N:
            f = dense.reshape(-1)[self._col]
        else:
            raise NotImplementedError("only [M,1] / [1,N] broadcasts are used by the reference (utils.py:88,96)")
        return self.set_value(f if self._value is None else f.to(self._value.dtype) * self._value)

    def __mul__(self, dense):
        return self._scaled(dense)

    def __rmul__(self, dense):
        return self._scaled(dense)

    def to_dense(self):
        M, N = self._sizes
        v = self._value if self._value is not None else torch.ones(self.nnz(), device=self._col.device)
        out = torch.zeros(M, N, dtype=v.dtype, device=v.device)
        return out.index_put_((self._rows(), self._col), v, accumulate=True)


SparseTensor = CSRGraph  # name used by main.py-style code


# ---------------------------------------------------------------------------
# graph-construction kernels (csrc/graph_build.cu) -- CUDA tensors only
# ---------------------------------------------------------------------------
def _gb_make_keys(row, col, n_cols, both, drop_diag, out=None):
    from . import _lib
    lib = _lib.load()
    n = row.numel()
    row, col = row.contiguous(), col.contiguous()
    if out is None:
        m = 2 * n if both else n
        out = (torch.empty(m, dtype=torch.int64, device=row.device), torch.empty(m, dtype=torch.int64, device=row.device))
    keys, pos = out
    _lib.check(lib.plnlp_graph_make_keys(_lib.ptr(row), _lib.ptr(col), n, int(n_cols), int(both), int(drop_diag),
                                         _lib.ptr(keys), _lib.ptr(pos), _lib.stream()), "plnlp_graph_make_keys")
    return keys, pos


def _gb_diag_keys(n_diag, n_cols, keys, pos, pos0):
    from . import _lib
    lib = _lib.load()
    _lib.check(lib.plnlp_graph_diag_keys(int(n_diag), int(n_cols), _lib.ptr(keys), _lib.ptr(pos), int(pos0), _lib.stream()),
               "plnlp_graph_diag_keys")


def _gb_sort(keys, pos, n_rows, n_cols, has_dead):
    from . import _lib
    lib = _lib.load()
    n = keys.numel()
    ko, po = torch.empty_like(keys), torch.empty_like(pos)
    nbytes = lib.plnlp_graph_sort_workspace_bytes(n)
    ws = _lib.workspace.get("graph_sort", nbytes, keys.device)
    _lib.check(lib.plnlp_graph_sort_pairs(_lib.ptr(keys), _lib.ptr(pos), n, int(n_rows), int(n_cols), int(has_dead),
                                          _lib.ptr(ko), _lib.ptr(po), _lib.ptr(ws), nbytes, _lib.stream()),
               "plnlp_graph_sort_pairs")
    return ko, po


def _gb_unique(keys, pos, value, n_src):
    """distinct keys of a sorted key array (+ the values of equal keys summed in sorted order); one host read"""
    from . import _lib
    lib = _lib.load()
    n = keys.numel()
    uk = torch.empty_like(keys)
    cnt = torch.empty(1, dtype=torch.int64, device=keys.device)
    uv = None if value is None else torch.empty(n, dtype=torch.float32, device=keys.device)
    nbytes = lib.plnlp_graph_unique_workspace_bytes(n)
    ws = _lib.workspace.get("graph_unique", nbytes, keys.device)
    _lib.check(lib.plnlp_graph_unique(_lib.ptr(keys), n, _lib.ptr(uk), _lib.ptr(cnt), _lib.ptr(pos),
                                      _lib.ptr(None if value is None else value.contiguous()), int(n_src), _lib.ptr(uv),
                                      _lib.ptr(ws), nbytes, _lib.stream()), "plnlp_graph_unique")
    m = int(cnt.item())
    return uk[:m], None if uv is None else uv[:m]


def _gb_keys_to_csr(keys, n_rows, n_cols):
    from . import _lib
    lib = _lib.load()
    n = keys.numel()
    rowptr = torch.empty(n_rows + 1, dtype=torch.int64, device=keys.device)
    col = torch.empty(n, dtype=torch.int64, device=keys.device)
    _lib.check(lib.plnlp_graph_keys_to_csr(_lib.ptr(keys.contiguous()), n, int(n_rows), int(n_cols), _lib.ptr(rowptr),
                                           _lib.ptr(col), _lib.stream()), "plnlp_graph_keys_to_csr")
    return rowptr, col


def sym_normalize(adj):
    """D^-1/2 A D^-1/2 of a CUDA CSRGraph in one kernel pair (deg = row sums, inf -> 0): the arithmetic of
    utils.gcn_normalization after set_diag (plnlp/utils.py:85-88)"""
    from . import _lib
    lib = _lib.load()
    rowptr, col, val = adj.csr()
    n_rows = adj.size(0)
    dis = torch.empty(n_rows, dtype=torch.float32, device=col.device)
    out = torch.empty(col.numel(), dtype=torch.float32, device=col.device)
    vin = None if val is None else val.to(torch.float32).contiguous()
    _lib.check(lib.plnlp_graph_sym_normalize(_lib.ptr(rowptr.contiguous()), _lib.ptr(col.contiguous()), _lib.ptr(vin),
                                             n_rows, _lib.ptr(dis), _lib.ptr(out), _lib.stream()),
               "plnlp_graph_sym_normalize")
    return adj.set_value(out)


# ---------------------------------------------------------------------------
# SpMM work plan + structure cache
# ---------------------------------------------------------------------------
class SpmmPlan:
    """Device arrays consumed by ``plnlp_spmm_csr_f32`` for one CSR matrix."""

    __slots__ = ("n_rows", "n_cols", "nnz", "chunk", "col", "val", "item_ptr", "item_row", "item_slot",
                 "n_items", "item_end", "subset", "fix_ptr", "fix_row", "n_fix", "n_partial", "row_cnt")

    def alg_bytes(self, F, elem=4):
        """ALGORITHMIC bytes of one launch: gathered rows + indices +
        values + row pointers + output write."""
        return (self.nnz * F * elem + self.nnz * 4 + (self.nnz * 4 if self.val is not None else 0)
                + (self.n_rows + 1) * 8 + self.n_rows * F * elem)


def _pick_chunk(nnz, n_rows):
    # aim for >= ~19k warp-items (148 SMs x 32 resident warps x 4) while never splitting
    # rows shorter than 64 entries and never letting one warp walk more than 1024
    target = max(nnz // 19000, 1)
    chunk = ((target + 31) // 32) * 32
    return int(min(1024, max(64, chunk)))


def build_plan(rowptr, col, val, n_rows, n_cols, chunk=None):
    dev = col.device
    nnz = col.numel()
    if nnz >= 2 ** 31 - 1 or n_rows >= 2 ** 31 - 1:
        raise RuntimeError("plnlp_b200 SpMM plans use int32 offsets: nnz and rows must be < 2^31")
    if chunk is None:
        chunk = _pick_chunk(nnz, n_rows)
    deg = rowptr[1:] - rowptr[:-1]
    n_it = torch.clamp((deg + chunk - 1) // chunk, min=1)
    n_items = int(n_it.sum())
    rows = torch.arange(n_rows, device=dev)
    item_row = torch.repeat_interleave(rows, n_it, output_size=n_items)
    first = torch.cumsum(n_it, 0) - n_it
    k = torch.arange(n_items, device=dev) - first[item_row]
    item_beg = rowptr[:-1][item_row] + k * chunk
    item_ptr = torch.cat([item_beg, torch.tensor([nnz], device=dev, dtype=torch.int64)])
    multi_row = n_it > 1
    multi_item = multi_row[item_row]
    slot = torch.cumsum(multi_item.to(torch.int64), 0) - 1
    item_slot = torch.where(multi_item, slot, torch.full_like(slot, -1))
    fix_row = torch.nonzero(multi_row).reshape(-1)
    fix_cnt = n_it[fix_row]
    fix_ptr = torch.zeros(fix_row.numel() + 1, dtype=torch.int64, device=dev)
    if fix_row.numel():
        fix_ptr[1:] = torch.cumsum(fix_cnt, 0)
    p = SpmmPlan()
    p.n_rows, p.n_cols, p.nnz, p.chunk = int(n_rows), int(n_cols), int(nnz), int(chunk)
    p.col = col.to(torch.int32).contiguous()
    p.val = None if val is None else val.to(torch.float32).contiguous()
    p.item_ptr = item_ptr.to(torch.int32).contiguous()
    p.item_row = item_row.to(torch.int32).contiguous()
    p.item_slot = item_slot.to(torch.int32).contiguous()
    p.n_items = n_items
    p.item_end, p.subset = None, False
    p.fix_ptr = fix_ptr.to(torch.int32).contiguous()
    p.fix_row = fix_row.to(torch.int32).contiguous()
    p.n_fix = int(fix_row.numel())
    p.n_partial = int(fix_ptr[-1]) if fix_row.numel() else 0
    p.row_cnt = torch.clamp(deg, min=1).to(torch.float32).contiguous()  # mean divisor, max(row_nnz, 1)
    return p


def build_subset_plan(parent, rowptr, rows):
    """Plan for the rows ``rows`` (int64 device vector of distinct row ids) of the matrix behind ``parent``:
    output row t of the SpMM is matrix row rows[t].  Same chunking as ``build_plan`` (hub rows are cut into
    several items whose partial sums a fixed-order pass combines); items carry explicit ends because the
    selected rows are not contiguous.  Index arrays (col / val) are shared with the parent.  One host read
    (the item count)."""
    dev = rows.device
    T = rows.numel()
    chunk = parent.chunk
    if rows.is_cuda and T > 0:
        return _build_subset_plan_cuda(parent, rowptr, rows)
    start = rowptr[rows]
    length = rowptr[rows + 1] - start
    n_it = torch.clamp((length + chunk - 1) // chunk, min=1)
    n_items = int(n_it.sum())
    t_ids = torch.arange(T, device=dev)
    item_row = torch.repeat_interleave(t_ids, n_it, output_size=n_items)
    first = torch.cumsum(n_it, 0) - n_it
    k = torch.arange(n_items, device=dev) - first[item_row]
    item_beg = start[item_row] + k * chunk
    item_end = torch.minimum(item_beg + chunk, (start + length)[item_row])
    multi_row = n_it > 1
    multi_item = multi_row[item_row]
    slot = torch.cumsum(multi_item.to(torch.int64), 0) - 1
    item_slot = torch.where(multi_item, slot, torch.full_like(slot, -1))
    fix_row = torch.nonzero(multi_row).reshape(-1)
    fix_ptr = torch.zeros(fix_row.numel() + 1, dtype=torch.int64, device=dev)
    if fix_row.numel():
        fix_ptr[1:] = torch.cumsum(n_it[fix_row], 0)
    p = SpmmPlan()
    p.n_rows, p.n_cols, p.chunk = int(T), parent.n_cols, chunk
    p.col, p.val = parent.col, parent.val
    p.item_ptr = item_beg.to(torch.int32).contiguous()
    p.item_end = item_end.to(torch.int32).contiguous()
    p.item_row = item_row.to(torch.int32).contiguous()
    p.item_slot = item_slot.to(torch.int32).contiguous()
    p.n_items, p.subset = n_items, True
    p.fix_ptr = fix_ptr.to(torch.int32).contiguous()
    p.fix_row = fix_row.to(torch.int32).contiguous()
    p.n_fix = int(fix_row.numel())
    p.n_partial = int(multi_item.sum()) if p.n_fix else 0
    p.nnz = int(length.sum())
    p.row_cnt = torch.clamp(length, min=1).to(torch.float32).contiguous()
    return p


def plan_tensors(plan):
    """the device tensors a plan owns or references (for stream bookkeeping)"""
    return [t for t in (getattr(plan, s) for s in SpmmPlan.__slots__) if isinstance(t, torch.Tensor)]


def _build_subset_plan_cuda(parent, rowptr, rows):
    """``build_subset_plan`` on the device (csrc/graph_build.cu): two kernels around one stacked prefix sum, ONE host
    read (item count, split-row count, partial-slot count, stored entries)"""
    from . import _lib
    lib = _lib.load()
    dev, T, chunk = rows.device, rows.numel(), parent.chunk
    rows = rows.contiguous()
    cnt = torch.empty(4, T, dtype=torch.int64, device=dev)          # n_it, multi, n_slot, len
    _lib.check(lib.plnlp_subset_plan_count(_lib.ptr(rowptr), _lib.ptr(rows), T, chunk, _lib.ptr(cnt[0]), _lib.ptr(cnt[1]),
                                           _lib.ptr(cnt[2]), _lib.ptr(cnt[3]), _lib.stream()), "plnlp_subset_plan_count")
    cum = torch.cumsum(cnt, 1)                                       # inclusive, all four at once
    n_items, n_fix, n_partial, nnz = cum[:, -1].tolist()             # the one host read
    p = SpmmPlan()
    p.n_rows, p.n_cols, p.chunk = int(T), parent.n_cols, chunk
    p.col, p.val = parent.col, parent.val
    i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=dev)  # noqa: E731
    p.item_ptr, p.item_end, p.item_row, p.item_slot = i32(n_items), i32(n_items), i32(n_items), i32(n_items)
    p.fix_ptr, p.fix_row = torch.zeros(n_fix + 1, dtype=torch.int32, device=dev), i32(n_fix)
    p.row_cnt = torch.empty(T, dtype=torch.float32, device=dev)
    _lib.check(lib.plnlp_subset_plan_fill(_lib.ptr(rowptr), _lib.ptr(rows), T, chunk, _lib.ptr(cum[0]), _lib.ptr(cum[1]),
                                          _lib.ptr(cum[2]), _lib.ptr(cnt[0]), _lib.ptr(p.item_ptr), _lib.ptr(p.item_end),
                                          _lib.ptr(p.item_row), _lib.ptr(p.item_slot), _lib.ptr(p.fix_ptr),
                                          _lib.ptr(p.fix_row), _lib.ptr(p.row_cnt), _lib.stream()), "plnlp_subset_plan_fill")
    p.item_ptr, p.item_end = p.item_ptr[:n_items], p.item_end[:n_items]
    p.item_row, p.item_slot = p.item_row[:n_items], p.item_slot[:n_items]
    p.fix_row = p.fix_row[:n_fix]
    p.n_items, p.subset, p.n_fix, p.n_partial, p.nnz = int(n_items), True, int(n_fix), int(n_partial), int(nnz)
    return p


class Structure:
    """Everything the encoder kernels need about one adjacency, forward and transposed."""

    def __init__(self, adj, chunk=None):
        rowptr, col, val = adj.csr()
        if not col.is_cuda:
            raise RuntimeError("plnlp_b200 needs the adjacency on a CUDA device; there is no CPU path")
        M, N = adj.size(0), adj.size(1)
        rowptr, col = rowptr.to(torch.int64), col.to(torch.int64)
        self.n_rows, self.n_cols = M, N
        self.has_value = val is not None
        # forward plans: valued (GCN, 'sum') and value-less (SAGE drops values, 'mean')
        self.fwd = build_plan(rowptr, col, val, M, N, chunk)
        self.rowptr = rowptr                      # int64, for per-step row-subset plans (build_subset_plan)
        self.fwd_noval = self.fwd if val is None else _share_plan(self.fwd, None)
        # transposed structure (backward): sort entries by (col, row)
        deg = rowptr[1:] - rowptr[:-1]
        row = torch.repeat_interleave(torch.arange(M, device=col.device), deg, output_size=col.numel())
        perm = torch.argsort(col * M + row, stable=True)
        self.t_perm = perm                        # entry e of the transposed matrix is entry perm[e] of this one
        t_row, t_col = col[perm], row[perm]
        t_rowptr = torch.zeros(N + 1, dtype=torch.int64, device=col.device)
        if t_row.numel():
            t_rowptr[1:] = torch.cumsum(torch.bincount(t_row, minlength=N), 0)
        t_val = None if val is None else val.to(torch.float32)[perm]
        self.bwd = build_plan(t_rowptr, t_col, t_val, N, M, chunk)
        self.t_rowptr = t_rowptr                  # int64 row pointers of the transposed matrix (TransposedStructure)
        # mean backward: A^T D^-1 g  ->  transposed entries carry 1/max(deg_row,1) of their source row
        inv = 1.0 / torch.clamp(deg, min=1).to(torch.float32)
        self.bwd_mean = _share_plan(self.bwd, inv[t_col].contiguous())
        self.symmetric = bool(M == N and torch.equal(t_rowptr, rowptr) and torch.equal(t_col, col))
        if self.symmetric:  # share the index arrays, halve the footprint
            for p in (self.bwd, self.bwd_mean):
                p.col, p.item_ptr, p.item_row, p.item_slot = (self.fwd.col, self.fwd.item_ptr,
                                                              self.fwd.item_row, self.fwd.item_slot)
                p.fix_ptr, p.fix_row = self.fwd.fix_ptr, self.fwd.fix_row
        # Dense fast path.  A small graph whose adjacency is more than a few per cent dense (ddi-shape:
        # 4 267 nodes, 11.7 % dense, source matrix L2 resident) is cheaper to multiply as a DENSE matrix on
        # the tensor cores than to gather row by row: gather cost ~ nnz*F*4 B at L2 speed, dense cost
        # 2*M*N*F FLOP at 3xTF32 speed; measured break-even on B200 is ~5 % density.
        self.density = col.numel() / max(M * N, 1)
        self.dense_ok = DENSE_SPMM and max(M, N) <= 16384 and self.density >= 0.06
        self._dense = {}
        self._coo = (row, col, None if val is None else val.to(torch.float32), inv)

    def entry_pairs(self):
        """int64 [nnz, 2]: (row, n_rows + col) of every stored entry, in CSR order -- the pair list that makes the
        edge-dot kernels compute one score per entry from the stacked matrix [query; key] (layer.TransformerConv)"""
        if getattr(self, "_entry_pairs", None) is None:
            row, col = self._coo[0], self._coo[1]
            self._entry_pairs = torch.stack([row, col + self.n_rows], 1).contiguous()
        return self._entry_pairs

    def dense(self, mean):
        """[M, N] fp32 dense form: stored values (or 1) summed per cell ('sum'), or 1/max(deg,1) per stored
        entry ('mean', SAGEConv).  Built once, on first use."""
        if mean not in self._dense:
            row, col, val, inv = self._coo
            v = inv[row] if mean else (val if val is not None else torch.ones(col.numel(), device=col.device))
            d = torch.zeros(self.n_rows, self.n_cols, dtype=torch.float32, device=col.device)
            d.index_put_((row, col), v, accumulate=True)
            self._dense[mean] = d
        return self._dense[mean]


class TransposedStructure:
    """``Structure`` of A^T as a VIEW of the structure of A: forward and backward plans swapped, nothing copied.
    Used by the row-partitioned encoder: the COLUMN block A[:, block] a rank multiplies its own activation rows
    with is the transpose of the row block of A^T it already holds (``parallel.ShardedAdj.local_t``).
    'sum' products only (the mean divisor of the transposed matrix is not kept)."""

    def __init__(self, st):
        if st.symmetric:
            raise RuntimeError("a symmetric structure shares its index arrays: use it directly")
        self.base = st
        self.n_rows, self.n_cols = st.n_cols, st.n_rows
        self.has_value = st.has_value
        self.fwd, self.bwd = st.bwd, st.fwd
        self.fwd_noval = st.bwd if st.bwd.val is None else _share_plan(st.bwd, None)
        self.bwd_mean = None
        self.rowptr, self.t_rowptr = st.t_rowptr, st.rowptr
        self.symmetric, self.dense_ok, self.density = False, False, st.density


class TransposedAdj:
    """adjacency handle whose ``structure_of`` is the transposed view of ``base``'s (no data of its own)"""

    def __init__(self, base):
        self.base = base
        self._plnlp_structure = None

    def size(self, dim):
        return self.base.size(1 - dim)


def _share_plan(plan, val):
    q = SpmmPlan()
    for s in SpmmPlan.__slots__:
        setattr(q, s, getattr(plan, s))
    q.val = val
    return q


_CACHE = {}


def structure_of(adj):
    """Cached ``Structure`` of an adjacency object (CSRGraph or torch_sparse.SparseTensor)."""
    if isinstance(adj, TransposedAdj):
        if adj._plnlp_structure is None:
            adj._plnlp_structure = TransposedStructure(structure_of(adj.base))
        return adj._plnlp_structure
    key = id(adj)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is adj:
        return hit[1]
    st = Structure(adj)
    if len(_CACHE) > 16:
        _CACHE.clear()
    _CACHE[key] = (adj, st)
    return st
