"""one SpMM launch sequence on the citation2-shape graph for ncu (HBM-bound regime)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph, structure_of  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402
from tools.microbench import powerlaw_graph  # noqa: E402

N, E, F = 2927963, 30561187, int(sys.argv[1]) if len(sys.argv) > 1 else 200
PITCH = int(sys.argv[2]) if len(sys.argv) > 2 else F          # leading dimension of the operand
adj = gcn_normalization(CSRGraph.from_edge_index(powerlaw_graph(N, E, 1), None, N).to_symmetric())
st = structure_of(adj)
x = torch.zeros(N, PITCH, device="cuda")[:, :F]
x.copy_(torch.randn(N, F, device="cuda"))
for _ in range(3):
    y = _ops.spmm_raw(st.fwd, x, use_val=True, div_rows=False)
torch.cuda.synchronize()
print("ok", adj.nnz(), float(y[0, 0]))
