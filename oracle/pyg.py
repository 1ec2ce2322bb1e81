"""Oracle restatement of the torch_geometric 2.0.1 pieces the reference uses.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  torch_geometric is pinned by
/root/reference/README.md:19 (pyg 2.0.1) but is neither vendored nor installable
here, so its published behaviour is restated for exactly these call sites:

* ``SAGEConv(in, out)`` / ``GCNConv(in, out, normalize=False)``
                                  (/root/reference/plnlp/layer.py:4,30-45)
* ``negative_sampling(edge_index, num_nodes, num_neg_samples, method='sparse')``
  and ``add_self_loops``          (/root/reference/plnlp/negative_sample.py:3,8-10)

Assumptions (parity of this layer is unpinned by the reference):
 (7) SAGEConv drops adjacency values before aggregating (``set_value(None)``), uses
     mean aggregation, and returns ``lin_l(agg) + lin_r(x)`` with a bias only in
     ``lin_l``; ``normalize=False`` and ``root_weight=True`` are the defaults.
 (8) GCNConv(normalize=False) = bias-free linear, then valued SpMM (sum), then
     ``+ bias``.
 (9) PyG 2.0.1 ``negative_sampling(method='sparse')``: linear ids ``row*N+col``;
     oversampling factor ``alpha = |1/(1 - 1.1*E/N^2)|``; draws ``int(alpha*n)``
     DISTINCT ids with python ``random.sample``; removes ids that are existing
     edges (``numpy.isin``); truncates to ``n``; returns ``[2, <=n]``.
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch

from .sparse import matmul


def _glorot(w):
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    torch.nn.init.uniform_(w, -a, a)


class SAGEConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = torch.nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def forward(self, x, adj_t):
        agg = matmul(adj_t.set_value(None), x, reduce="mean")
        return self.lin_l(agg) + self.lin_r(x)


class GCNConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, normalize=False):
        super().__init__()
        assert not normalize, "the reference only uses normalize=False (layer.py:45)"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        _glorot(self.lin.weight)
        torch.nn.init.zeros_(self.bias)

    def forward(self, x, adj_t):
        x = self.lin(x)
        out = matmul(adj_t, x, reduce="sum")
        return out + self.bias


class _Unavailable(torch.nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("out of scope (SURVEY.md section 2.1 rows 2)")


class GraphConv(torch.nn.Module):
    """torch_geometric 2.0.1 nn.GraphConv(in, out, aggr='add') restated [from memory -- parity unpinned, like
    the rest of this layer]: lin_rel(sum_j w_ij x_j) + lin_root(x_i); with a SparseTensor ``adj_t`` the
    aggregation is ``matmul(adj_t, x, reduce='add')``, which uses the stored values as edge weights; bias in
    lin_rel only.  Used by the reference's WSAGE encoder (plnlp/layer.py:48-54)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel = torch.nn.Linear(in_channels, out_channels, bias=True)
        self.lin_root = torch.nn.Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x, adj_t):
        return self.lin_rel(matmul(adj_t, x, reduce="sum")) + self.lin_root(x)


class TransformerConv(torch.nn.Module):
    """torch_geometric 2.0.1 nn.TransformerConv(in, out) with its defaults (heads=1, concat=True, beta=False,
    dropout=0, edge_dim=None, bias=True, root_weight=True) restated [from memory -- parity unpinned]:
    alpha_ij = softmax_j(<lin_query(x_i), lin_key(x_j)> / sqrt(out)) over the stored entries of row i of adj_t
    (row = target), out_i = sum_j alpha_ij lin_value(x_j) + lin_skip(x_i); all four linears carry a bias.
    Used by the reference's Transformer encoder (plnlp/layer.py:57-63) on a value-less adjacency."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_key = torch.nn.Linear(in_channels, out_channels)
        self.lin_query = torch.nn.Linear(in_channels, out_channels)
        self.lin_value = torch.nn.Linear(in_channels, out_channels)
        self.lin_skip = torch.nn.Linear(in_channels, out_channels)

    def reset_parameters(self):
        for lin in (self.lin_key, self.lin_query, self.lin_value, self.lin_skip):
            lin.reset_parameters()

    def forward(self, x, adj_t):
        row, col, _ = adj_t.coo()
        q, k, v = self.lin_query(x), self.lin_key(x), self.lin_value(x)
        score = (q[row] * k[col]).sum(-1) / (self.out_channels ** 0.5)
        N = x.size(0)
        mx = torch.full((N,), float("-inf"), dtype=score.dtype).scatter_reduce(0, row, score, "amax", include_self=True)
        e = torch.exp(score - mx[row])
        z = torch.zeros(N, dtype=score.dtype).index_add_(0, row, e)
        alpha = e / z[row]
        out = torch.zeros(N, self.out_channels, dtype=x.dtype).index_add_(0, row, alpha.unsqueeze(-1) * v[col])
        return out + self.lin_skip(x)


def add_self_loops(edge_index, edge_weight=None, num_nodes=None):
    N = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    loop = torch.arange(N, dtype=edge_index.dtype)
    return torch.cat([edge_index, torch.stack([loop, loop])], dim=1), edge_weight


def negative_sampling(edge_index, num_nodes=None, num_neg_samples=None, method="sparse"):
    N = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
    n = int(num_neg_samples) if num_neg_samples is not None else edge_index.size(1)
    size = N * N
    n = min(n, size - edge_index.size(1))
    row, col = edge_index
    idx = row * N + col
    alpha = abs(1 / (1 - 1.1 * (edge_index.size(1) / size)))
    k = min(int(alpha * n), size)
    perm = torch.tensor(random.sample(range(size), k), dtype=torch.int64)
    mask = torch.from_numpy(np.isin(perm.numpy(), idx.numpy())).to(torch.bool)
    perm = perm[~mask][:n]
    return torch.stack([perm // N, perm % N], dim=0)
