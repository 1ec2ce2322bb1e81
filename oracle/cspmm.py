"""ctypes front end of oracle/spmm_ref.c (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_spmm.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def spmm(rowptr, col, val, x, reduce="sum", f64=False, threads=0):
    """In-order CPU SpMM; rowptr/col int64, val fp32 or None, x fp32 [N, F]."""
    rowptr = rowptr.to(torch.int64).contiguous()
    col = col.to(torch.int64).contiguous()
    x = x.to(torch.float32).contiguous()
    if val is not None:
        val = val.to(torch.float32).contiguous()
    M, F = rowptr.numel() - 1, x.size(1)
    out = torch.empty(M, F, dtype=torch.float64 if f64 else torch.float32)
    fn = _lib().plnlp_oracle_spmm_f64 if f64 else _lib().plnlp_oracle_spmm_f32
    p = ctypes.c_void_p
    fn(p(rowptr.data_ptr()), p(col.data_ptr()), p(val.data_ptr() if val is not None else None),
       p(x.data_ptr()), ctypes.c_int64(x.stride(0)), p(out.data_ptr()), ctypes.c_int64(F),
       ctypes.c_int64(M), ctypes.c_int64(F), ctypes.c_int(1 if reduce == "mean" else 0), ctypes.c_int(threads))
    return out
