"""EXPERIMENT (not product): consumer-side proxy fence variant of the 2-CTA GEMM.
Measures speed and checks bitwise run-to-run / cross-variant equality over many repetitions."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plnlp_b200 import _lib, _ops
from tools.microbench import timeit
M, N, K = 262144, 512, 512
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda")
ref = _ops.gemm_raw(A, W, transb=True, backend="tf32x3c2").clone()
ms0 = timeit(lambda: _ops.gemm_raw(A, W, transb=True, backend="tf32x3c2"))
exp = ctypes.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), "libplnlp_b200_exp.so"))
fn = exp.plnlp_gemm_tf32_2cta
fn.restype, fn.argtypes = _lib.SIGNATURES["plnlp_gemm_tf32_2cta"]
C = torch.empty(M, N, device="cuda")
def run():
    rc = fn(3, 0, 1, M, N, K, A.data_ptr(), K, W.data_ptr(), K, C.data_ptr(), N, 0.0, None, 0, None, 0, 0.0, 0, None, 0, 1,
            torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
ms1 = timeit(run)
bad = 0
for i in range(100):
    C.zero_(); run()
    bad += int(not torch.equal(C, ref))
print(f"writer-side fence {ms0:.3f} ms; consumer-side fence {ms1:.3f} ms; mismatching repetitions {bad}/100")
