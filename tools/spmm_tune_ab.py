"""A/B of the SpMM gather knobs (plnlp_spmm_tune) on the citation2-shape graph, one process: L2 row prefetch
(prefetch.global.L2 / cp.async.bulk.prefetch.L2), the shared-memory staged kernel (cp.async) in its pipeline shapes,
and the device's L2 fetch granularity.  Every configuration is checked bit for bit against the first one.
Usage: python tools/spmm_tune_ab.py [quick]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _lib, _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph, structure_of  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402
from tools.microbench import HBM, powerlaw_graph, timeit  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
lib = _lib.load()
N, E = 2927963, 30561187
adj = gcn_normalization(CSRGraph.from_edge_index(powerlaw_graph(N, E, 1), None, N).to_symmetric())
plan = structure_of(adj).fwd
print(f"graph: N={N} nnz={plan.nnz} items={plan.n_items} fix={plan.n_fix}", flush=True)


def operand(F, pitch):
    xp = torch.zeros(N, pitch, device="cuda")[:, :F]
    xp.copy_(torch.randn(N, F, device="cuda", generator=torch.Generator(device="cuda").manual_seed(F)))
    return xp


cases = [(50, 50), (50, 64), (64, 64), (32, 32), (18, 18), (200, 200)]
if quick:
    cases = [(50, 50), (50, 64), (64, 64)]
ops = {c: operand(*c) for c in cases}
ref = {}

configs = [("base", 0, 0, 4)]
configs += [(f"staged{m}w{w}", 0, m, w) for m, w in ((9, 4), (9, 2), (9, 8), (10, 4), (10, 2), (10, 8), (10, 16), (11, 4),
                                                      (11, 8), (6, 4), (5, 4))]
configs += [("default", -1, -1, 0)]
for name, pf, staged, warps in configs:
    if name == "default":
        _ops.apply_spmm_defaults()
    else:
        assert lib.plnlp_spmm_tune(pf, staged, warps, 0) == 0
    for (F, pitch), xp in ops.items():
        if staged > 0 and F > 64:
            continue
        ms = timeit(lambda: _ops.spmm_raw(plan, xp, use_val=True, div_rows=False), warm=2, iters=7)
        y = _ops.spmm_raw(plan, xp, use_val=True, div_rows=False)
        ref.setdefault((F, pitch), y)
        alg = plan.alg_bytes(F, 4)
        print(f"{name:14s} F={F:3d} pitch={pitch:3d} {ms:7.3f} ms {alg / ms / 1e6:8.1f} GB/s "
              f"{alg / ms / 1e6 / HBM:6.1%} of HBM  bit-equal: {bool(torch.equal(y, ref[(F, pitch)]))}", flush=True)
        del y
_ops.apply_spmm_defaults()
