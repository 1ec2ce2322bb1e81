"""A/B of the narrow-row SpMM (two rows per warp) on the citation2-shape graph: the [N, 50] operand on its own
pitch (generic warp-per-row kernel) vs a 16-byte aligned pitch of 52 / 64 floats (narrow kernel).
PLNLP_SPMM_NARROW=0 / PLNLP_SPMM_NB are read once per process by the library: run one process per setting."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph, structure_of  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402
from tools.microbench import HBM, powerlaw_graph, timeit  # noqa: E402

N, E = 2927963, 30561187
adj = gcn_normalization(CSRGraph.from_edge_index(powerlaw_graph(N, E, 1), None, N).to_symmetric())
plan = structure_of(adj).fwd
_ops.NARROW_SPMM = False                  # no automatic re-pitching: the layouts below are what the kernel sees
tag = f"narrow={os.environ.get('PLNLP_SPMM_NARROW', '1')} NB={os.environ.get('PLNLP_SPMM_NB', 'default')}"
for F, pitches in ((50, (50, 52, 64)), (64, (64,)), (32, (32,)), (16, (16,))):
    x = torch.randn(N, F, device="cuda")
    ref = None
    for pitch in pitches:
        xp = torch.zeros(N, pitch, device="cuda")[:, :F]
        xp.copy_(x)
        ms = timeit(lambda: _ops.spmm_raw(plan, xp, use_val=True, div_rows=False))
        y = _ops.spmm_raw(plan, xp, use_val=True, div_rows=False)
        ref = y if ref is None else ref
        alg = plan.alg_bytes(F, 4)
        print(f"{tag} F={F:3d} pitch={pitch:3d} {ms:7.3f} ms {alg / ms / 1e6:8.1f} GB/s {alg / ms / 1e6 / HBM:6.1%} of HBM"
              f"  bit-equal to first layout: {bool(torch.equal(y, ref))}", flush=True)
