"""the layer-1 forward GEMM of the citation2-shape encoder on the TMA-fed kernel, for ncu / timing experiments
(PLNLP_TMA_DEBUG, PLNLP_TMA_SA).  Usage: python tools/gemm_tma_one.py [M] [N] [K] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2927963
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
K = int(sys.argv[3]) if len(sys.argv) > 3 else 178
_ops.GEMM_TMA = "1"
A = torch.randn(M, (K + 3) // 4 * 4, device="cuda")[:, :K]
W = torch.randn(N, K, device="cuda")
b = torch.randn(N, device="cuda")
C = torch.empty(M, N, device="cuda")
f = lambda: _ops.gemm_raw(A, W, transb=True, C=C, bias=b, act=_ops.ACT_RELU)  # noqa: E731
ms = timeit(f, warm=2, iters=int(sys.argv[4]) if len(sys.argv) > 4 else 10)
print(f"debug={os.environ.get('PLNLP_TMA_DEBUG', '0')} sa={os.environ.get('PLNLP_TMA_SA', 'auto')} {M}x{N}x{K}: {ms:.3f} ms", flush=True)
