"""A/B of the tall-skinny encoder GEMMs at citation2 shape: TMA-fed persistent kernel (csrc/gemm_tma.cu) vs the
CTA-pair register-path kernel.  HBM roof = (A read + C write) / measured copy bandwidth; tensor floor = 3 passes at
the tf32 rate (half the measured bf16 peak)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _ops  # noqa: E402
from tools.microbench import HBM, timeit  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2927963
shapes = [("layer-1 forward  [A emb | A x] W1^T + b, relu", M, 200, 178, True, True),
          ("layer-1 backward dZ1 W1[:, :50]              ", M, 50, 200, False, False),
          ("last conv        agg W2^T + b (T rows)       ", 290000, 200, 200, True, True),
          ("ddi-like         h W^T, 512-wide             ", 262144, 256, 512, True, False)]
for name, m, n, k, tb, epi in shapes:
    kp = (k + 3) // 4 * 4
    A = torch.randn(m, kp, device="cuda")[:, :k]
    B = torch.randn((n, k) if tb else (k, n), device="cuda")
    bias = torch.randn(n, device="cuda") if epi else None
    C = torch.empty(m, n, device="cuda")
    res = {}
    for mode in ("1", "0"):
        _ops.GEMM_TMA = mode
        f = lambda: _ops.gemm_raw(A, B, transb=tb, C=C, bias=bias, act=_ops.ACT_RELU if epi else _ops.ACT_NONE)  # noqa: E731
        ms = timeit(f)
        res[mode] = (ms, C.clone())
    byt = (m * k + m * n) * 4
    fl = 2.0 * m * n * k
    err = float((res["1"][1] - res["0"][1]).abs().max() / res["0"][1].abs().max())
    for mode, tag in (("1", "tma "), ("0", "pair")):
        ms = res[mode][0]
        print(f"{name} {m}x{n}x{k} {tag} {ms:7.3f} ms  {byt / ms / 1e6:7.1f} GB/s = {byt / ms / 1e6 / HBM:5.1%} of HBM copy peak,"
              f" {fl / ms / 1e9:6.1f} TFLOP/s", flush=True)
    print(f"    max|tma - pair| / max = {err:.2e}", flush=True)
