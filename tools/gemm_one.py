"""one tensor-core GEMM launch sequence for ncu (small M so the ~40 replays stay short)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 37888          # 296 M-tiles x 2 N-tiles = 4 full waves
backend = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 512
K = int(sys.argv[4]) if len(sys.argv) > 4 else 512
A = torch.randn(M, K, device="cuda")
W = torch.randn(N, K, device="cuda")
for _ in range(3):
    C = _ops.gemm_raw(A, W, transb=True, backend=backend)
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
