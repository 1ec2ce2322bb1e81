"""The layer-1 aggregation exactly as the citation2-shape model issues it (operand = the nn.Embedding weight on its own
50-float pitch, output = the first 50 columns of the [N, 180] aggregate buffer; backward operand on a 64-float pitch)
under the SpMM tuning knobs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _lib, _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph, structure_of  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402
from tools.microbench import HBM, powerlaw_graph, timeit  # noqa: E402

lib = _lib.load()
N, E = 2927963, 30561187
adj = gcn_normalization(CSRGraph.from_edge_index(powerlaw_graph(N, E, 1), None, N).to_symmetric())
plan = structure_of(adj).fwd
emb = torch.nn.Embedding(N, 50).cuda()
torch.nn.init.xavier_uniform_(emb.weight)
buf = torch.zeros(N, 180, device="cuda")
gu = _ops._rows_for_spmm(N, 50, "cuda")
gu.copy_(torch.randn(N, 50, device="cuda"))
for name, pf, staged, warps in (("base", 0, 0, 4), ("default", 3, 12, 4), ("staged9w4", 0, 9, 4), ("staged10w4", 0, 10, 4)):
    assert lib.plnlp_spmm_tune(pf, staged, warps, 0) == 0
    with torch.no_grad():
        a = timeit(lambda: _ops.spmm_raw(plan, emb.weight, use_val=True, div_rows=False, out=buf[:, :50]), iters=7)
        b = timeit(lambda: _ops.spmm_raw(plan, gu, use_val=True, div_rows=False), iters=7)
        c = timeit(lambda: _ops.spmm_raw(plan, emb.weight, use_val=True, div_rows=False), iters=7)
    alg = plan.alg_bytes(50, 4)
    print(f"{name:12s} fwd into buf[:, :50] {a:6.3f} ms ({alg / a / 1e6 / HBM:5.1%})  fwd fresh out {c:6.3f} ms  "
          f"bwd (pitch 64) {b:6.3f} ms ({alg / b / 1e6 / HBM:5.1%})", flush=True)
_ops.apply_spmm_defaults()
