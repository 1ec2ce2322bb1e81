"""CPU-side checks of the C-ABI boundary (-m "not gpu"): the shared library builds, loads, exports
every symbol include/plnlp_b200.h declares, and the ctypes table in plnlp_b200/_lib.py matches the
header declaration by declaration.  No kernel is launched here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "plnlp_b200.h")


def _declarations():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|int64_t)\s+(plnlp_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        decls[name] = (ret, params)
    return decls


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build()
    from plnlp_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_something():
    d = _declarations()
    assert len(d) >= 20 and "plnlp_spmm_csr_f32" in d and "plnlp_gemm_f32" in d


def test_library_exports_every_declared_symbol(built_lib):
    for name in _declarations():
        assert hasattr(built_lib, name), f"{name} declared in the header but not exported"


def _ctype_of(param):
    if "*" in param:
        return ctypes.c_void_p
    t = param.split()
    if "float" in t:
        return ctypes.c_float
    if "uint64_t" in t:
        return ctypes.c_uint64
    if "int64_t" in t:
        return ctypes.c_int64
    if "int" in t:
        return ctypes.c_int
    raise AssertionError(f"unhandled parameter type: {param}")


def test_ctypes_table_matches_header(built_lib):
    from plnlp_b200 import _lib
    decls = _declarations()
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, (ret, params) in decls.items():
        res, args = _lib.SIGNATURES[name]
        assert res is (ctypes.c_int64 if ret == "int64_t" else ctypes.c_int), name
        assert [_ctype_of(p) for p in params] == list(args), name


def test_abi_version_and_pure_helpers(built_lib):
    from plnlp_b200 import _lib
    lib = _lib.load()
    assert lib.plnlp_abi_version() == 1
    assert lib.plnlp_pair_loss_workspace_bytes(65536) >= 8 * 256
    assert lib.plnlp_colsum_workspace_bytes(1000, 512) >= 2 * 512 * 4
    assert lib.plnlp_mlp_out_bwd_workspace_bytes(262144, 512) >= 1024 * 512 * 4


def test_product_has_no_cpu_path():
    """the package must fail loudly instead of computing on the CPU"""
    import torch
    from plnlp_b200 import _ops
    from plnlp_b200.graph import CSRGraph
    from plnlp_b200.model import BaseModel
    with pytest.raises(RuntimeError):
        BaseModel(lr=0.01, dropout=0.0, grad_clip_norm=1.0, gnn_num_layers=1, mlp_num_layers=1,
                  emb_hidden_channels=8, gnn_hidden_channels=8, mlp_hidden_channels=8, num_nodes=10,
                  num_node_feats=0, gnn_encoder_name="SAGE", predictor_name="DOT", loss_func="AUC",
                  optimizer_name="Adam", device="cpu", use_node_feats=False, train_node_emb=True)
    g = CSRGraph.from_edge_index(torch.tensor([[0, 1], [1, 0]]), None, 2)
    with pytest.raises(RuntimeError):
        _ops.spmm(g, torch.randn(2, 4))
    with pytest.raises(RuntimeError):
        _ops.pair_loss("AUC", torch.randn(4), torch.randn(4), 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "plnlp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
