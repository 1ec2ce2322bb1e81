"""Random-walk augmentation on the GPU.

``random_walk`` keeps torch_cluster's call shape (``random_walk(row, col, start, walk_length)``,
/root/reference/main.py:242); ``random_walk_pairs`` produces what main.py:241-253 assigns to
``split_edge['train']['edge']`` / ``['weight']`` for one epoch: the (start, visited) pairs of every
walk, j-major, weights 1/(j+1), self pairs removed.
"""
from __future__ import annotations

import torch

from . import _lib, _ops, profiling
from ._lib import check, ptr, stream


def _rowptr_of(row, num_nodes):
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=row.device)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=num_nodes), 0)
    return rowptr


def random_walk(row, col, start, walk_length, num_nodes=None, rand=None, rowptr=None):
    """-> int64 [len(start), walk_length + 1].  ``row`` must be sorted (it is the row array of
    ``adj_t.coo()``, main.py:229).  ``rand`` ([n_walks, walk_length] uniforms) replaces the Philox
    stream -- used by the parity tests."""
    if not (col.is_cuda and start.is_cuda):
        raise RuntimeError("plnlp_b200.augment runs on the GPU; there is no CPU path")
    lib = _lib.load()
    if rowptr is None:
        # torch_cluster sizes the graph from row, col AND start (main.py --walk_start_type=node passes
        # arange(num_nodes): trailing isolated nodes are legal start points that simply stay where they are)
        n = int(num_nodes) if num_nodes is not None else \
            int(torch.max(torch.stack([row.max(), col.max(), start.max()]))) + 1
        rowptr = _rowptr_of(row, n)
    rowptr, col = rowptr.to(torch.int64).contiguous(), col.to(torch.int64).contiguous()
    start = start.to(torch.int64).contiguous()
    W = start.numel()
    if W and (int(start.max()) >= rowptr.numel() - 1 or int(start.min()) < 0):
        raise RuntimeError(f"random_walk: start node out of range for a graph of {rowptr.numel() - 1} nodes")
    walk = torch.empty(W, walk_length + 1, dtype=torch.int64, device=col.device)
    if rand is not None:
        rand = rand.to(torch.float32).contiguous()
    with profiling.span("random_walk", W * (walk_length + 1) * 8, 0):
        check(lib.plnlp_random_walk(ptr(rowptr), ptr(col), ptr(start), W, int(walk_length), ptr(rand),
                                    _ops.new_seed() if rand is None else 0, ptr(walk), stream()),
              "plnlp_random_walk")
    return walk


def walk_pairs(walk):
    """main.py:243-253 -> (edges int64 [M, 2], weights float32 [M]) with self pairs removed."""
    lib = _lib.load()
    W, L = walk.size(0), walk.size(1) - 1
    pairs = torch.empty(W * L, 2, dtype=torch.int64, device=walk.device)
    weight = torch.empty(W * L, dtype=torch.float32, device=walk.device)
    keep = torch.empty(W * L, dtype=torch.uint8, device=walk.device)
    with profiling.span("walk_pairs", W * L * 29, 0):
        check(lib.plnlp_walk_pairs(ptr(walk.contiguous()), W, L, ptr(pairs), ptr(weight), ptr(keep), stream()),
              "plnlp_walk_pairs")
    sel = keep.to(torch.bool)
    return pairs[sel], weight[sel]


def random_walk_pairs(adj_t, start, walk_length, rand=None):
    """one epoch of augmented training pairs from an adjacency (``.csr()``), as main.py:228-253 does
    with ``rw_row, rw_col, _ = adj_t.coo()``."""
    rowptr, col, _ = adj_t.csr()
    walk = random_walk(None, col, start, walk_length, rowptr=rowptr, rand=rand)
    return walk_pairs(walk)
