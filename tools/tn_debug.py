"""index-permutation probe of the TMA-fed weight-gradient GEMM (csrc/gemm_tma_tn.cu): A and B are zero except for one
K-row, so C[m, n] = (m + 1)(n + 1) shows which operand element every output row / column actually read.  It is how the
MN-major layout (PLNLP_TN_DEBUG=8: plain SWIZZLE_128B returns zeros) and the TMEM lane-quarter mapping were found."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _lib, _ops
lib = _lib.load()
_ops.GEMM_TMA_TN = "1"
torch.set_printoptions(linewidth=220, precision=0, sci_mode=False, threshold=100000)
for M, N in ((128, 128), (256, 192)):
    K = 2048
    A = torch.zeros(K, M); B = torch.zeros(K, N)
    A[5, :] = torch.arange(M).float() + 1      # only K-row 5: C[m, n] = (m + 1) * (n + 1)
    B[5, :] = torch.arange(N).float() + 1
    C = _ops.gemm_raw(A.cuda(), B.cuda(), transa=True, backend="tf32c2").cpu()
    print("dbg", os.environ.get("PLNLP_TN_DEBUG"), M, N, "max err", float((C - A.t() @ B).abs().max()))
    print("m index seen by D row m (column 0):", (C[:, 0] - 1).long().tolist())
    print("n index seen by D col n (row of m=?):", (C[0, :] / C[0, 0] - 1).round().long().tolist())
