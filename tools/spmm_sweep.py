"""SpMM timing sweep on the citation2-shape graph (HBM-bound regime): feature widths x dtypes.
Usage: python tools/spmm_sweep.py [F ...]      (run under gpurun; PLNLP_SPMM_NB overrides the load depth)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph, structure_of  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402
from tools.microbench import HBM, powerlaw_graph, timeit  # noqa: E402

N, E = 2927963, 30561187
widths = [int(a) for a in sys.argv[1:]] or [50, 64, 128, 200, 256]
adj = gcn_normalization(CSRGraph.from_edge_index(powerlaw_graph(N, E, 1), None, N).to_symmetric())
st = structure_of(adj)
plan = st.fwd
print(f"nnz {adj.nnz()} items {plan.n_items} fix {plan.n_fix} NB={os.environ.get('PLNLP_SPMM_NB', 'default')}")
for F in widths:
    x = torch.randn(N, F, device="cuda")
    for dt in (torch.float32, torch.bfloat16):
        xx = x.to(dt)
        ms = timeit(lambda: _ops.spmm_raw(plan, xx, use_val=True, div_rows=False))
        alg = plan.alg_bytes(F, xx.element_size())
        line = f"F={F:4d} {str(dt)[6:]:9s} {ms:8.3f} ms  {alg / ms / 1e6:8.1f} GB/s  {alg / ms / 1e6 / HBM:6.1%} of HBM"
        if dt == torch.bfloat16:
            y = _ops.spmm_raw(plan, xx, use_val=True, div_rows=False).float()
            ref = _ops.spmm_raw(plan, xx.float(), use_val=True, div_rows=False)
            err = float((y - ref).abs().max() / ref.abs().max())
            line += f"  max|bf16 - f32(bf16 inputs)|/max = {err:.2e}"
        print(line, flush=True)
