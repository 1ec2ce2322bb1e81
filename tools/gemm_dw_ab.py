"""the layer-1 weight-gradient GEMM at citation2 shape (dW = dY^T [A x | 1], 200 x 179 x 2.9 M, both operands
MN-major, split-k): the CTA-pair kernel vs the one-CTA kernel, and the K per accumulator."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

M = 2927963
g = torch.randn(M, 200, device="cuda")
ext = torch.randn(M, 180, device="cuda")[:, :179]
ref = None
for backend in ("tf32x3c2", "tf32x3"):
    for kcap in (1088, 2176):
        _ops.TF32X3_KCAP = kcap
        ms = timeit(lambda: _ops.gemm_raw(g, ext, transa=True, backend=backend))
        out = _ops.gemm_raw(g, ext, transa=True, backend=backend)
        if ref is None:
            ref = (g.double().t() @ ext.double())
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        byt = (g.numel() + M * 179) * 4
        print(f"{backend:9s} K per accumulator <= {kcap}: {ms:.3f} ms  {byt / ms / 1e6:7.1f} GB/s  "
              f"{2.0 * 200 * 179 * M / ms / 1e9:6.1f} TFLOP/s  err vs fp64 {err:.2e}", flush=True)
