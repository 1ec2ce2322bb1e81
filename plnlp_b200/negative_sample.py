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
    N, want = int(num_nodes), int(num_samples) * int(num_neg)
    got = _distinct_non_edges(edge_index, N, want)     # candidate order = random order
    if got.numel() < want:
        extra = torch.randint(0, max(got.numel(), 1), (want - got.numel(),), device=got.device)
        got = torch.cat([got, got[extra]])
    src = torch.div(got, N, rounding_mode="floor")
    dst = got - src * N
    return torch.stack([src, dst], dim=-1).reshape(-1, int(num_neg), 2)


def local_neg_sample(pos_edges, num_nodes, num_neg, random_src=False):
    """negative_sample.py:31-43: keep the source of every positive, draw ``num_neg`` uniform
    destinations in [0, num_nodes); nothing is filtered."""
    if not pos_edges.is_cuda:
        raise RuntimeError("plnlp_b200 samplers run on the GPU; pos_edges must be a CUDA tensor")
    if random_src:
        # negative_sample.py:32-34 (no call site of the reference uses it): the kept endpoint of every positive is
        # drawn uniformly from its two ends.  Index plumbing on the device; the kernel reads column 0 as the source.
        side = torch.randint(0, 2, (pos_edges.size(0), 1), dtype=torch.long, device=pos_edges.device)
        src = pos_edges.gather(1, side)
        pos_edges = torch.cat([src, src], dim=1).contiguous()
    return _ops.local_neg_sample_raw(pos_edges, num_nodes, num_neg, _ops.new_seed())


def _distinct_non_edges(edge_index, num_nodes, want):
    """``want`` distinct uniformly drawn non-edge, non-self cells as linear ids r*N + c (fewer if the rejection
    step came up short), in random order; the common core of the two global samplers"""
    if not edge_index.is_cuda:
        raise RuntimeError("plnlp_b200 samplers run on the GPU; edge_index must be a CUDA tensor")
    N = int(num_nodes)
    ids = _sorted_edge_ids(edge_index, N)
    free = N * N - ids.numel() - N
    if free <= 0:
        raise RuntimeError("graph has no non-edges to sample")
    # oversample for rejections (edges, self pairs) and duplicate draws (the role of PyG's alpha
    # factor): v valid draws leave free*(1 - exp(-v/free)) distinct cells in expectation
    frac_valid = free / float(N * N)
    v = -free * math.log1p(-want / float(free)) if want < 0.95 * free else 3.0 * free
    n_cand = int(v / frac_valid * 1.05) + 4096
    cand, keep = _ops.global_neg_candidates_raw(ids, N, n_cand, _ops.new_seed())
    return cand[keep.to(torch.bool)][:want]


def global_perm_neg_sample(edge_index, num_nodes, num_samples, num_neg, method='sparse'):
    """negative_sample.py:23-28: ``num_samples`` distinct negatives, then ``num_neg - 1`` shuffled copies of
    the same set (``sample_perm_copy``), so every positive sees each negative set member once per copy."""
    N = int(num_nodes)
    got = _distinct_non_edges(edge_index, N, int(num_samples))
    src = torch.div(got, N, rounding_mode="floor")
    return sample_perm_copy(torch.stack([src, got - src * N]), int(num_samples), int(num_neg))


def sample_perm_copy(edge_index, target_num_sample, num_perm_copy):
    """negative_sample.py:61-76 on the device: pad with randomly chosen duplicates up to ``target_num_sample``,
    then stack ``num_perm_copy`` copies, every copy after the first independently permuted -> [target, k, 2]
    laid out exactly as the reference's reshape of the concatenated copies."""
    dev = edge_index.device
    src, dst = edge_index[0], edge_index[1]
    if edge_index.size(1) < target_num_sample:
        k = target_num_sample - edge_index.size(1)
        rand_index = torch.randperm(edge_index.size(1), device=dev)[:k]
        src, dst = torch.cat((src, src[rand_index])), torch.cat((dst, dst[rand_index]))
    tmp_src, tmp_dst = src, dst
    for _ in range(num_perm_copy - 1):
        rand_index = torch.randperm(target_num_sample, device=dev)
        src, dst = torch.cat((src, tmp_src[rand_index])), torch.cat((dst, tmp_dst[rand_index]))
    return torch.reshape(torch.stack((src, dst), dim=-1), (-1, num_perm_copy, 2))
