"""Oracle restatement of the torch_sparse pieces the reference relies on.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  torch_sparse is not vendored
under /root/reference and not installable here; this file restates its
published behaviour for exactly the calls the reference makes:

* ``T.ToSparseTensor()(data)``            (/root/reference/main.py:81)
* ``adj_t.coo()``                          (main.py:82, 229)
* ``adj_t.to_symmetric()``                 (main.py:110)
* ``SparseTensor(row=, col=, value=)``     (main.py:124-126, 136-138)
* ``adj_t.set_diag()``, ``adj_t.sum(dim=1)``, ``dense[N,1] * adj_t * dense[1,N]``
                                           (plnlp/utils.py:83-89, 92-97)
* ``matmul(adj_t, x, reduce='sum'|'mean')`` inside SAGEConv / GCNConv
                                           (plnlp/layer.py:20,23)

Assumptions about upstream behaviour (recorded because parity of this layer is
unpinned by the reference):
 (1) rows of ``adj_t`` are destinations, columns are sources; entries are sorted by
     (row, col); duplicate entries are KEPT by the constructor.
 (2) ``to_symmetric`` = union of (r,c) and (c,r), sorted, duplicates merged with
     their values summed.
 (3) ``set_diag`` drops any existing diagonal and inserts one entry per row with
     value 1 (value-less tensors stay value-less).
 (4) ``sum(dim=1)`` = row sum of values, or the row entry count when value-less.
 (5) ``dense[N,1] * sparse`` scales rows, ``sparse * dense[1,N]`` scales columns; a
     value-less operand takes the broadcast dense factor as its values.
 (6) ``matmul(reduce='sum'|'mean')`` accumulates in CSR order in fp32; ``mean``
     divides by ``max(row_nnz, 1)`` so empty rows give 0.
"""
from __future__ import annotations

import warnings

import torch


class SparseTensor:
    """Minimal CSR holder with the torch_sparse.SparseTensor methods used by
    the reference.  Index tensors are int64 like upstream."""

    def __init__(self, row=None, col=None, value=None, sparse_sizes=None,
                 rowptr=None, is_sorted=False):
        if row is None and rowptr is None:
            raise ValueError("need row or rowptr")
        if sparse_sizes is None:
            m = int(row.max()) + 1 if row is not None and row.numel() else 0
            n = int(col.max()) + 1 if col.numel() else 0
            if rowptr is not None:
                m = rowptr.numel() - 1
            sparse_sizes = (m, n)
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        M, N = self._sizes
        if row is not None and not is_sorted:
            key = row.to(torch.int64) * N + col.to(torch.int64)
            perm = torch.argsort(key, stable=True)
            row, col = row[perm], col[perm]
            if value is not None:
                value = value[perm]
        if rowptr is None:
            counts = torch.bincount(row, minlength=M)
            rowptr = torch.zeros(M + 1, dtype=torch.int64, device=col.device)
            torch.cumsum(counts, 0, out=rowptr[1:])
        self._rowptr = rowptr.to(torch.int64)
        self._col = col.to(torch.int64)
        self._value = value
        self._row = row.to(torch.int64) if row is not None else None

    # -- accessors ---------------------------------------------------------
    def size(self, dim):
        return self._sizes[dim]

    def sizes(self):
        return list(self._sizes)

    def sparse_sizes(self):
        return self._sizes

    def nnz(self):
        return self._col.numel()

    def _rows(self):
        if self._row is None:
            M = self._sizes[0]
            self._row = torch.repeat_interleave(
                torch.arange(M, device=self._col.device), self._rowptr[1:] - self._rowptr[:-1])
        return self._row

    def csr(self):
        return self._rowptr, self._col, self._value

    def coo(self):
        return self._rows(), self._col, self._value

    def has_value(self):
        return self._value is not None

    def set_value(self, value, layout=None):
        return SparseTensor(row=self._row, col=self._col, value=value,
                            sparse_sizes=self._sizes, rowptr=self._rowptr, is_sorted=True)

    def to(self, *args, **kwargs):
        dev_like = torch.empty(0).to(*args, **kwargs)
        v = self._value
        if v is not None:
            v = v.to(device=dev_like.device)
            if dev_like.dtype.is_floating_point and dev_like.dtype != torch.float32:
                v = v.to(dev_like.dtype)
        return SparseTensor(row=None if self._row is None else self._row.to(dev_like.device),
                            col=self._col.to(dev_like.device), value=v,
                            sparse_sizes=self._sizes, rowptr=self._rowptr.to(dev_like.device),
                            is_sorted=True)

    # -- structure ops -----------------------------------------------------
    def t(self):
        row, col, value = self.coo()
        return SparseTensor(row=col, col=row, value=value,
                            sparse_sizes=(self._sizes[1], self._sizes[0]))

    def to_symmetric(self):
        N = max(self._sizes)
        row, col, value = self.coo()
        r2 = torch.cat([row, col])
        c2 = torch.cat([col, row])
        key = r2 * N + c2
        if value is None:
            key = torch.unique(key)  # sorted, duplicates merged
            return SparseTensor(row=key // N, col=key % N, value=None,
                                sparse_sizes=(N, N), is_sorted=True)
        v2 = torch.cat([value, value])
        ukey, inv = torch.unique(key, return_inverse=True)
        v = torch.zeros(ukey.numel(), dtype=value.dtype).index_add_(0, inv, v2)
        return SparseTensor(row=ukey // N, col=ukey % N, value=v,
                            sparse_sizes=(N, N), is_sorted=True)

    def set_diag(self):
        M, N = self._sizes
        row, col, value = self.coo()
        keep = row != col
        d = torch.arange(min(M, N), dtype=torch.int64)
        nrow = torch.cat([row[keep], d])
        ncol = torch.cat([col[keep], d])
        nval = None
        if value is not None:
            nval = torch.cat([value[keep], torch.ones(d.numel(), dtype=value.dtype)])
        return SparseTensor(row=nrow, col=ncol, value=nval, sparse_sizes=(M, N))

    def sum(self, dim=1):
        assert dim == 1
        M = self._sizes[0]
        if self._value is None:
            return self._rowptr[1:] - self._rowptr[:-1]
        return torch.zeros(M, dtype=self._value.dtype).index_add_(0, self._rows(), self._value)

    def _scale(self, dense, left):
        M, N = self._sizes
        if dense.dim() == 2 and dense.size(0) == M and dense.size(1) == 1:
            f = dense.reshape(-1)[self._rows()]
        elif dense.dim() == 2 and dense.size(0) == 1 and dense.size(1) == N:
            f = dense.reshape(-1)[self._col]
        else:
            raise ValueError("only [M,1] and [1,N] broadcasts are restated")
        v = f if self._value is None else f.to(self._value.dtype) * self._value
        return self.set_value(v)

    def __mul__(self, dense):
        return self._scale(dense, left=False)

    def __rmul__(self, dense):
        return self._scale(dense, left=True)

    def matmul(self, x, reduce="sum"):
        return matmul(self, x, reduce)

    def to_dense(self):
        M, N = self._sizes
        v = self._value if self._value is not None else torch.ones(self.nnz())
        out = torch.zeros(M, N, dtype=v.dtype)
        out.index_put_((self._rows(), self._col), v, accumulate=True)
        return out


def to_sparse_tensor(edge_index, edge_weight=None, num_nodes=None):
    """``T.ToSparseTensor()`` (main.py:81): adj_t[dst, src], sorted by (dst, src),
    value = edge_weight when present, duplicates kept."""
    N = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    return SparseTensor(row=edge_index[1], col=edge_index[0], value=edge_weight,
                        sparse_sizes=(N, N))


# ---------------------------------------------------------------------------
# matmul(adj_t, x, reduce) with the backward torch_sparse implements
# ---------------------------------------------------------------------------
def _torch_csr(rowptr, col, val, shape, dtype):
    v = val.to(dtype) if val is not None else torch.ones(col.numel(), dtype=dtype)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.sparse_csr_tensor(rowptr, col, v, size=shape, check_invariants=False)


class _SpMM(torch.autograd.Function):
    """out = A @ x (sum) or D^-1 A @ x (mean, D = max(row_nnz, 1)).
    Backward w.r.t. x only: A^T g (sum) / A^T D^-1 g (mean) -- the reference
    never differentiates w.r.t. adjacency values."""

    @staticmethod
    def forward(ctx, x, adj, reduce):
        rowptr, col, val = adj.csr()
        M, N = adj.sparse_sizes()
        A = _torch_csr(rowptr, col, val, (M, N), x.dtype)
        out = A @ x
        inv = None
        if reduce == "mean":
            cnt = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(x.dtype)
            inv = 1.0 / cnt
            out = out / cnt[:, None]
        ctx.adj, ctx.inv = adj, inv
        return out

    @staticmethod
    def backward(ctx, g):
        adj, inv = ctx.adj, ctx.inv
        if inv is not None:
            g = g * inv[:, None]
        at = adj.t()
        rowptr, col, val = at.csr()
        At = _torch_csr(rowptr, col, val, at.sparse_sizes(), g.dtype)
        return At @ g, None, None


def matmul(adj, x, reduce="sum"):
    if reduce in ("sum", "add"):
        return _SpMM.apply(x, adj, "sum")
    if reduce == "mean":
        return _SpMM.apply(x, adj, "mean")
    raise NotImplementedError(reduce)


def matmul_indexadd(adj, x, reduce="sum"):
    """Independent formulation (gather + index_add_, natively differentiable)
    used only to cross-check ``matmul`` on small graphs."""
    row, col, val = adj.coo()
    msg = x[col]
    if val is not None:
        msg = msg * val.to(x.dtype)[:, None]
    out = torch.zeros(adj.size(0), x.size(1), dtype=x.dtype).index_add(0, row, msg)
    if reduce == "mean":
        rowptr = adj.csr()[0]
        cnt = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(x.dtype)
        out = out / cnt[:, None]
    return out


def gcn_normalization(adj_t):
    """plnlp/utils.py:83-89 applied to the oracle SparseTensor."""
    adj_t = adj_t.set_diag()
    deg = adj_t.sum(dim=1).to(torch.float)
    deg_inv_sqrt = deg.pow(-0.5)
    deg_inv_sqrt[deg_inv_sqrt == float("inf")] = 0
    return deg_inv_sqrt.view(-1, 1) * adj_t * deg_inv_sqrt.view(1, -1)
