"""Negative samplers with the reference's signatures (/root/reference/plnlp/negative_sample.py),
running as GPU kernels (csrc/sample.cu) instead of host python.

Both return the reference layout: int64 ``[E, num_neg, 2]`` where the negatives of positive ``i``
are ``out[i, :, :]``.  RNG: Philox4x32-10 keyed by a seed drawn from torch's CPU generator, so
``torch.manual_seed`` makes a run reproducible.
"""
from __future__ import annotations

import math

import torch

from . import _ops

_EDGE_ID_CACHE = {}


def _sorted_edge_ids(edge_index, num_nodes):
    """sorted ids ``edge_index[0]*N + edge_index[1]`` of the existing edges (cached per tensor)."""
    key = (edge_index.data_ptr(), edge_index.size(1), edge_index._version, int(num_nodes))
    hit = _EDGE_ID_CACHE.get(key)
    if hit is None:
        ids = edge_index[0].to(torch.int64) * int(num_nodes) + edge_index[1].to(torch.int64)
        hit = torch.sort(ids)[0].contiguous()
        _EDGE_ID_CACHE.clear()
        _EDGE_ID_CACHE[key] = hit
    return hit


def global_neg_sample(edge_index, num_nodes, num_samples, num_neg, method='sparse'):
    """negative_sample.py:6-20.  ``num_samples * num_neg`` DISTINCT uniformly drawn cells (r, c)
    that are neither self pairs (the reference adds self loops before sampling, :8) nor existing
    edges; if fewer distinct cells survive, the shortfall is filled with randomly chosen
    duplicates (:14-18)."""
    if not edge_index.is_cuda:
        raise RuntimeError("plnlp_b200 samplers run on the GPU; edge_index must be a CUDA tensor")
    N, want = int(num_nodes), int(num_samples) * int(num_neg)
    ids = _sorted_edge_ids(edge_index, N)
    free = N * N - ids.numel() - N
    if free <= 0:
        raise RuntimeError("graph has no non-edges to sample")
    # oversample for rejections (edges, self pairs) and duplicate draws (the role of PyG's alpha
    # factor): v valid draws leave free*(1 - exp(-v/free)) distinct cells in expectation
    frac_valid = free / float(N * N)
    if want < 0.95 * free:
        v = -free * math.log1p(-want / float(free))
    else:
        v = 3.0 * free
    n_cand = int(v / frac_valid * 1.05) + 4096
    cand, keep = _ops.global_neg_candidates_raw(ids, N, n_cand, _ops.new_seed())
    got = cand[keep.to(torch.bool)][:want]            # candidate order = random order
    if got.numel() < want:
        extra = torch.randint(0, max(got.numel(), 1), (want - got.numel(),), device=got.device)
        got = torch.cat([got, got[extra]])
    src = torch.div(got, N, rounding_mode="floor")
    dst = got - src * N
    return torch.stack([src, dst], dim=-1).reshape(-1, int(num_neg), 2)


def local_neg_sample(pos_edges, num_nodes, num_neg, random_src=False):
    """negative_sample.py:31-43: keep the source of every positive, draw ``num_neg`` uniform
    destinations in [0, num_nodes); nothing is filtered."""
    if random_src:
        raise NotImplementedError("random_src=True is never used by the reference's call sites (utils.py:17-20)")
    if not pos_edges.is_cuda:
        raise RuntimeError("plnlp_b200 samplers run on the GPU; pos_edges must be a CUDA tensor")
    return _ops.local_neg_sample_raw(pos_edges, num_nodes, num_neg, _ops.new_seed())


def global_perm_neg_sample(*args, **kwargs):
    raise NotImplementedError("global_perm_neg_sample (negative_sample.py:23-28) is outside the hot-path "
                              "scope of plnlp_b200 (SURVEY.md section 8f)")


def sample_perm_copy(*args, **kwargs):
    raise NotImplementedError("sample_perm_copy (negative_sample.py:61-76) is outside the hot-path scope")
