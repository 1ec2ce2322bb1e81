"""How fast can this GPU fetch RANDOM rows of a table that does not fit in L2?  The ceiling the narrow SpMM gathers
(F <= 64) run against: below ~256 B per row the cost of a fetch is the DRAM row activation, not the bytes, so "per
cent of the copy bandwidth" under-reads how close such a kernel is to what the memory system can do.

Measured with the repo's own fused gather + dot kernel (two random rows per pair in, 4 bytes out: no write traffic to
speak of) on a [2 927 963, F] table, uniform random row ids (no reuse beyond chance) and with the citation2-shape
graph's own column-index stream (power-law reuse, what the SpMM actually sees)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from plnlp_b200.graph import CSRGraph  # noqa: E402
from plnlp_b200.utils import gcn_normalization  # noqa: E402
from tools.microbench import HBM, powerlaw_graph, timeit  # noqa: E402

N, E, P = 2927963, 30561187, 30 * 1024 * 1024
g = torch.Generator(device="cuda").manual_seed(0)
uni = torch.randint(0, N, (P, 2), generator=g, device="cuda")
adj = gcn_normalization(CSRGraph.from_edge_index(powerlaw_graph(N, E, 1), None, N).to_symmetric())
col = adj.csr()[1]
real = col[: 2 * P].reshape(P, 2).contiguous()                 # the SpMM's own gather stream, in its own order
del adj
for F, pitch in ((16, 16), (32, 32), (50, 50), (50, 64), (64, 64), (128, 128), (200, 200), (256, 256)):
    h = torch.randn(N, pitch, device="cuda")[:, :F]
    for name, e in (("uniform", uni), ("citation2 col stream", real)):
        ms = timeit(lambda: _ops.edge_dot_raw(h, e))
        rows = 2 * P
        print(f"F={F:4d} pitch={pitch:4d} {name:22s} {ms:7.3f} ms  {rows / ms / 1e6:6.2f} G rows/s  "
              f"{rows * F * 4 / ms / 1e6:7.1f} GB/s = {rows * F * 4 / ms / 1e6 / HBM:5.1%} of the HBM copy peak", flush=True)
