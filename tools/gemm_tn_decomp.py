"""timing decomposition of the TMA-fed weight-gradient GEMM (csrc/gemm_tma_tn.cu) at the citation2 shape: run one
process per PLNLP_TN_DEBUG value (16 = no MMAs, 32 = no convert work, 48 = TMA ring only)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

M = 2927963
g = torch.randn(M, 200, device="cuda")
ext = torch.randn(M, 180, device="cuda")[:, :179]
byt = (g.numel() + M * 179) * 4
for backend in ("tf32x3c2", "tf32c2"):
    for tn in ("auto", "0"):
        _ops.GEMM_TMA_TN = tn
        ms = timeit(lambda: _ops.gemm_raw(g, ext, transa=True, backend=backend))
        print(f"dbg={os.environ.get('PLNLP_TN_DEBUG', '0'):3s} {backend:9s} tma_tn={tn:4s}: {ms:.3f} ms  {byt / ms / 1e6:7.1f} GB/s  "
              f"{2.0 * 200 * 179 * M / ms / 1e9:6.1f} TFLOP/s", flush=True)
    if os.environ.get("PLNLP_TN_DEBUG", "0") != "0":
        break
