"""one fused edge-scoring forward (ddi bench shape) for ncu"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402

N, H, P = 4267, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 262144
h = torch.randn(N, H, device="cuda")
e = torch.randint(0, N, (P, 2), device="cuda")
W1, b1 = torch.randn(H, H, device="cuda"), torch.randn(H, device="cuda")
w2, b2 = torch.randn(1, H, device="cuda"), torch.randn(1, device="cuda")
for _ in range(2):
    s, a = _ops.edge_mlp_fwd_raw(h, e, W1, b1, w2, b2, 0.3, 5)
torch.cuda.synchronize()
print("ok", float(s[0]))
