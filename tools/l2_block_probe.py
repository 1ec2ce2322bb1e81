"""Feasibility probe for an L2-blocked SpMM: gather only the entries whose column falls in one block of the source
matrix (a block that fits in L2) and time the existing kernel on that sub-matrix.  If the gathers of a resident
block run much faster per entry than the full kernel's, column blocking pays."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph, build_plan, structure_of  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402
from tools.microbench import powerlaw_graph, timeit  # noqa: E402

N, E, F = 2927963, 30561187, int(sys.argv[1]) if len(sys.argv) > 1 else 50
adj = gcn_normalization(CSRGraph.from_edge_index(powerlaw_graph(N, E, 1), None, N).to_symmetric())
rowptr, col, val = adj.csr()
row = adj.coo()[0]
x = torch.randn(N, F, device="cuda")
full = structure_of(adj).fwd
ms_full = timeit(lambda: _ops.spmm_raw(full, x, use_val=True, div_rows=False))
print(f"F={F} full: nnz {col.numel()} {ms_full:.3f} ms = {ms_full / col.numel() * 1e6:.3f} ns/entry")
for rows_per_block in (80000, 160000, 320000, 640000):
    for lo in (0, N // 2):
        hi = min(lo + rows_per_block, N)
        keep = (col >= lo) & (col < hi)
        r, c, v = row[keep], col[keep], val[keep]
        rp = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
        rp[1:] = torch.cumsum(torch.bincount(r, minlength=N), 0)
        plan = build_plan(rp, c, v, N, N)
        ms = timeit(lambda: _ops.spmm_raw(plan, x, use_val=True, div_rows=False))
        n = int(keep.sum())
        # rows that actually have entries in the block (a blocked kernel would only visit those)
        live = int((rp[1:] > rp[:-1]).sum())
        print(f"  block rows [{lo}, {hi}) = {(hi - lo) * F * 4 / 1e6:6.1f} MB: nnz {n:9d} live rows {live:8d}  {ms:.3f} ms = "
              f"{ms / max(n, 1) * 1e6:.3f} ns/entry (full kernel: {ms_full / col.numel() * 1e6:.3f}); "
              f"out write alone = {N * F * 4 / 6.5e9:.3f} ms", flush=True)
