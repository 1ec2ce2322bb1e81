"""Oracle restatement of torch_cluster.random_walk + the pair assembly of the reference.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  torch_cluster is an un-pinned, un-vendored dependency
(/root/reference/main.py:9,242).  Assumed upstream behaviour (parity of this layer is unpinned by the
reference): uniforms ``rand[n_walks, L]`` are drawn first; step l of walk n moves from ``cur`` to
``col[rowptr[cur] + int(rand[n,l] * deg(cur))]`` and stays put when ``deg(cur) == 0``; the offset is clamped
to ``deg - 1`` (fp32 ``rand*deg`` can round up to ``deg``).  The pair / weight assembly follows
main.py:243-253 literally.
"""
from __future__ import annotations

import torch


def random_walk(rowptr, col, start, walk_length, rand):
    W = start.numel()
    walk = torch.empty(W, walk_length + 1, dtype=torch.int64)
    cur = start.clone()
    walk[:, 0] = cur
    for l in range(walk_length):
        b, e = rowptr[cur], rowptr[cur + 1]
        deg = e - b
        off = (rand[:, l].to(torch.float32) * deg.to(torch.float32)).to(torch.int64)
        off = torch.minimum(off, (deg - 1).clamp(min=0))
        nxt = col[(b + off).clamp(max=max(col.numel() - 1, 0))] if col.numel() else cur
        cur = torch.where(deg > 0, nxt, cur)
        walk[:, l + 1] = cur
    return walk


def walk_pairs(walk):
    """main.py:243-253"""
    L = walk.size(1) - 1
    pairs, weights = [], []
    for j in range(L):
        pairs.append(walk[:, [0, j + 1]])
        weights.append(torch.ones((walk.size(0),), dtype=torch.float) / (j + 1))
    pairs, weights = torch.cat(pairs, 0), torch.cat(weights, 0)
    mask = (pairs[:, 0] - pairs[:, 1]) != 0
    return torch.masked_select(pairs, mask.view(-1, 1)).view(-1, 2), torch.masked_select(weights, mask)
