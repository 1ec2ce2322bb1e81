"""GPU parity tests (-m gpu): every CUDA kernel, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Bars (BASELINE.json north_star): index / integer work bit-exact;
fp32 results within 1e-5 relative (max|a-b| / max|b|)."""
import numpy as np
import pytest
import torch

from oracle import cspmm, ogb_eval, plnlp_ref, sparse
from tests.helpers import rand_graph, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    from plnlp_b200 import _lib, _ops
    assert _lib.load().plnlp_check_device() == 0
    return _ops


def _to_gpu_graph(o):
    from plnlp_b200.graph import CSRGraph
    rowptr, col, val = o.csr()
    return CSRGraph(rowptr.cuda(), col.cuda(), None if val is None else val.cuda(), o.sparse_sizes())


# ------------------------------------------------------------------ SpMM
@pytest.mark.parametrize("F", [1, 3, 50, 64, 128, 178, 200, 256, 512, 1024])
@pytest.mark.parametrize("weighted,reduce", [(False, "mean"), (True, "sum"), (True, "mean"), (False, "sum")])
def test_spmm_forward(ops, F, weighted, reduce):
    N = 211                                  # not a multiple of 32
    ei, w = rand_graph(N, 1500, seed=F, weighted=weighted, hub=True)
    o = sparse.to_sparse_tensor(ei, w, N)
    g = _to_gpu_graph(o)
    x = torch.randn(N, F)
    got = ops.spmm(g, x.cuda(), reduce).cpu()
    oref = o.set_value(None) if reduce == "mean" else o
    rowptr, col, val = oref.csr()
    want = cspmm.spmm(rowptr, col, val, x, reduce)       # in-order fp32 loop
    assert rel_err(got, want) < TOL
    empty = (rowptr[1:] - rowptr[:-1]) == 0
    assert empty.any() and torch.all(got[empty] == 0)


def test_spmm_unsplit_rows_are_bit_exact(ops):
    """rows that fit one work item are accumulated strictly in CSR order: identical bits to the
    in-order CPU loop, valued and value-less"""
    from plnlp_b200.graph import Structure
    N, F = 300, 200
    ei, w = rand_graph(N, 4000, seed=77, weighted=True)
    for weights, reduce in ((w, "sum"), (None, "mean"), (None, "sum")):
        o = sparse.to_sparse_tensor(ei, weights, N)
        g = _to_gpu_graph(o)
        st = Structure(g, chunk=1024)
        assert st.fwd.n_fix == 0
        x = torch.randn(N, F)
        got = ops.spmm_raw(st.fwd, x.cuda(), use_val=weights is not None, div_rows=(reduce == "mean")).cpu()
        rowptr, col, val = o.csr()
        assert torch.equal(got, cspmm.spmm(rowptr, col, val, x, reduce))


@pytest.mark.parametrize("F", [1, 3, 17, 32, 50, 52, 63, 64])
@pytest.mark.parametrize("layout", ["auto_pad", "pitch64", "slice_of_wide"])
def test_spmm_narrow_rows_two_per_warp(ops, F, layout):
    """operands of <= 64 columns run on the two-rows-per-warp kernel (16 lanes x 16-byte loads per row): rows that
    fit one work item keep the bits of the in-order CPU loop (valued, value-less, mean), hub rows are combined in
    fixed order, the epilogue (bias / relu / dropout) is the generic kernel's, and the output may be a column slice
    of a wider buffer (the [A emb | A x] aggregate buffer of the GCN layer)"""
    from plnlp_b200 import _lib
    from plnlp_b200.graph import Structure
    N = 333                                                   # odd item count: the last warp has one idle half
    ei, w = rand_graph(N, 5000, seed=100 + F, weighted=True, hub=True)
    gen = torch.Generator().manual_seed(F)
    x = torch.randn(N, F, generator=gen)
    if layout == "auto_pad":
        xg = x.cuda()                                         # contiguous [N, F]: narrow kernel iff F % 4 == 0
    elif layout == "pitch64":
        xg = torch.full((N, 64), float("nan")).cuda()[:, :F]   # pitch padding is garbage and must never leak
        xg.copy_(x)
    else:
        xg = torch.full((N, 136), float("nan")).cuda()[:, 8:8 + F]    # 32-byte column offset inside a wide matrix
        xg.copy_(x)
    for weights, reduce, chunk in ((w, "sum", 1024), (None, "mean", 1024), (None, "sum", 1024), (w, "sum", 64)):
        o = sparse.to_sparse_tensor(ei, weights, N)
        st = Structure(_to_gpu_graph(o), chunk=chunk)
        rowptr, col, val = o.csr()
        wide = torch.full((N, 96), -7.0).cuda()
        out = wide[:, 5:5 + F]                                # unaligned column slice as the output
        n0 = _lib.launch_count()
        ops.spmm_raw(st.fwd, xg, use_val=weights is not None, div_rows=(reduce == "mean"), out=out)
        assert _lib.launch_count() - n0 == (1 if st.fwd.n_fix == 0 else 2)
        want = cspmm.spmm(rowptr, col, val, x, reduce)
        if st.fwd.n_fix == 0:
            assert torch.equal(out.cpu(), want)
        else:
            assert rel_err(out.cpu(), cspmm.spmm(rowptr, col, val, x, reduce, f64=True)) < TOL
            fixed = st.fwd.fix_row.cpu().long()
            keep = torch.ones(N, dtype=torch.bool)
            keep[fixed] = False
            assert torch.equal(out.cpu()[keep], want[keep])  # unsplit rows stay bit-exact next to split ones
        assert torch.all(wide[:, :5] == -7.0) and torch.all(wide[:, 5 + F:] == -7.0)     # nothing outside the slice
    # fused epilogue = the generic kernel's (same Philox indexing): compare with the wide kernel on the same plan
    o = sparse.to_sparse_tensor(ei, w, N)
    st = Structure(_to_gpu_graph(o), chunk=1024)
    bias = torch.randn(F, generator=gen).cuda()
    a = ops.spmm_raw(st.fwd, xg, use_val=True, div_rows=False, bias=bias, relu=True, drop_p=0.3, seed=99)
    x1 = torch.zeros(N, F + 1).cuda()[:, :F]                 # odd pitch: the generic warp-per-row kernel
    x1.copy_(x)
    b = ops.spmm_raw(st.fwd, x1, use_val=True, div_rows=False, bias=bias, relu=True, drop_p=0.3, seed=99)
    assert torch.equal(a, b) and not torch.isnan(a).any()


@pytest.mark.parametrize("F,layout", [(50, "own"), (50, "pitch64"), (64, "own"), (32, "offset"), (18, "own"), (2, "own")])
def test_spmm_tuning_modes_keep_every_bit(ops, F, layout):
    """plnlp_spmm_tune: the L2 row prefetch (modes 1, 2) and the shared-memory staged kernel (cp.async row copies,
    every pipeline shape, 8- and 16-byte pieces) give the bits of the plain register kernels -- valued / value-less /
    mean, hub rows cut into items (partial slots), empty rows, the fused epilogue, a row-subset plan and a column
    slice as the output; wide operands (F = 200, prefetch only) and the row-sparse operand as well"""
    from plnlp_b200 import _lib
    from plnlp_b200.graph import Structure, build_subset_plan
    lib = _lib.load()
    N = 1500
    ei, w = rand_graph(N, 30000, seed=300 + F, weighted=True, hub=True)
    gen = torch.Generator().manual_seed(F)
    x = torch.randn(N, F, generator=gen)
    if layout == "own":
        xg = x.cuda()
    elif layout == "pitch64":
        xg = torch.full((N, 64), float("nan")).cuda()[:, :F]
        xg.copy_(x)
    else:
        xg = torch.full((N, 136), float("nan")).cuda()[:, 8:8 + F]
        xg.copy_(x)
    bias = torch.randn(F, generator=gen).cuda()
    mask = torch.randn(N, F, generator=gen).cuda()
    xw = torch.randn(N, 200, generator=gen).cuda()
    xi = torch.arange(N, dtype=torch.int32)
    xi[torch.rand(N, generator=gen) < 0.7] = -1
    xi = xi.cuda()

    def run_all():
        outs = []
        for weights, reduce, chunk in ((w, "sum", 1024), (None, "mean", 64), (w, "sum", 32)):
            o = sparse.to_sparse_tensor(ei, weights, N)
            st = Structure(_to_gpu_graph(o), chunk=chunk)
            wide = torch.full((N, 96), -7.0).cuda()
            ops.spmm_raw(st.fwd, xg, use_val=weights is not None, div_rows=(reduce == "mean"), out=wide[:, 6:6 + F])
            outs.append(wide)
            outs.append(ops.spmm_raw(st.fwd, xg, use_val=weights is not None, div_rows=False, bias=bias, relu=True,
                                     drop_p=0.25, seed=5, mask=mask, mask_scale=1.25))
            rows = torch.randperm(N, generator=gen)[:400].sort().values.cuda()
            sub = build_subset_plan(st.fwd, o.csr()[0].cuda(), rows)
            outs.append(ops.spmm_raw(sub, xg, use_val=weights is not None, div_rows=False))
            outs.append(ops.spmm_raw(st.fwd, xw, use_val=weights is not None, div_rows=False))
            outs.append(ops.spmm_raw(st.fwd, xw, use_val=weights is not None, div_rows=False, x_index=xi))
        return outs

    try:
        assert lib.plnlp_spmm_tune(0, 0, 4, 0) == 0
        gen.manual_seed(F)
        base = run_all()
        assert not any(torch.isnan(t).any() for t in base)
        for pf, staged, warps in ((1, 0, 4), (2, 0, 4), (3, 0, 4), (0, 1, 4), (0, 2, 3), (0, 3, 16), (0, 4, 1), (0, 5, 8),
                                  (0, 6, 4), (0, 7, 5), (0, 8, 2), (0, 9, 4), (0, 10, 4), (0, 11, 7), (1, 1, 2), (3, 12, 4)):
            assert lib.plnlp_spmm_tune(pf, staged, warps, 0) == 0
            gen.manual_seed(F)
            got = run_all()
            for a, b in zip(base, got):
                assert torch.equal(a, b), (pf, staged, warps)
    finally:
        lib.plnlp_spmm_tune(-1, -1, 0, 0)          # leave the knobs as the package set them
        ops.apply_spmm_defaults()


@pytest.mark.parametrize("chunk", [32, 64])
def test_spmm_split_rows(ops, chunk):
    from plnlp_b200.graph import Structure
    N, F = 150, 96
    ei, w = rand_graph(N, 3000, seed=5, weighted=True, hub=True)
    o = sparse.to_sparse_tensor(ei, w, N)
    st = Structure(_to_gpu_graph(o), chunk=chunk)
    assert st.fwd.n_fix > 0
    x = torch.randn(N, F)
    got = ops.spmm_raw(st.fwd, x.cuda(), use_val=True, div_rows=False).cpu()
    rowptr, col, val = o.csr()
    assert rel_err(got, cspmm.spmm(rowptr, col, val, x, "sum", f64=True)) < TOL
    got2 = ops.spmm_raw(st.fwd, x.cuda(), use_val=True, div_rows=False).cpu()
    assert torch.equal(got, got2)            # deterministic


@pytest.mark.parametrize("F", [3, 50, 200, 512])
def test_spmm_row_sparse_operand_mask(ops, F, monkeypatch):
    """x_index < 0 skips the gathers of all-zero rows of x: identical bits to the dense kernel (the skipped terms are
    exact zeros, the surviving ones keep their CSR order), valued and value-less, hub rows included; and the
    autograd path (spmm(..., sparse_grad=True): SpMM.backward measures the live rows of its own incoming gradient)
    gives the same gradient as the plain path"""
    from plnlp_b200 import graph
    from plnlp_b200.graph import Structure
    monkeypatch.setattr(graph, "DENSE_SPMM", False)       # CSR kernels also for the autograd part below
    N = 260
    ei, w = rand_graph(N, 4000, seed=F, weighted=True, hub=True)
    g = _to_gpu_graph(sparse.to_sparse_tensor(ei, w, N))
    st = Structure(g, chunk=64)
    assert st.fwd.n_fix > 0
    gen = torch.Generator().manual_seed(F)
    x = torch.randn(N, F, generator=gen)
    keep = torch.rand(N, generator=gen) < 0.15
    keep[2] = True                                     # the hub column stays live
    x[~keep] = 0
    xg = x.cuda()
    mask = ops.row_nonzero_index_raw(xg)
    assert torch.equal(mask.cpu() >= 0, keep) and torch.equal(mask.cpu()[keep], torch.nonzero(keep).reshape(-1).int())
    for use_val, div in ((True, False), (False, True)):
        plan = st.fwd if use_val else st.fwd_noval
        dense = ops.spmm_raw(plan, xg, use_val=use_val, div_rows=div)
        sparse_ = ops.spmm_raw(plan, xg, use_val=use_val, div_rows=div, x_index=mask)
        assert torch.equal(dense, sparse_)
    none = ops.spmm_raw(st.fwd, xg, use_val=True, div_rows=False, x_index=torch.full_like(mask, -1))
    assert torch.all(none == 0)
    # autograd: gradient that is non-zero only at a few rows
    idx = torch.nonzero(keep).reshape(-1).cuda()
    wgt = torch.randn(idx.numel(), F, generator=gen).cuda()
    grads = []
    for hinted in (False, True):
        z = torch.randn(N, F, generator=torch.Generator().manual_seed(1)).cuda().requires_grad_(True)
        y = ops.spmm(g, z, "sum", relu=True, sparse_grad=hinted)
        (y[idx] * wgt).sum().backward()
        grads.append(z.grad.clone())
    assert torch.equal(grads[0], grads[1])


@pytest.mark.parametrize("F", [3, 50, 200])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_spmm_row_subset_matches_full_product(ops, F, reduce, monkeypatch):
    """spmm_rows: the selected output rows carry the bits of the full product (hub rows cut into items, empty
    rows, epilogue), and its backward equals the backward of 'full product, then pick the rows'"""
    from plnlp_b200 import graph
    monkeypatch.setattr(graph, "DENSE_SPMM", False)
    N = 260
    ei, w = rand_graph(N, 4000, seed=F + 1, weighted=(reduce == "sum"), hub=True)
    g = _to_gpu_graph(sparse.to_sparse_tensor(ei, w, N))
    graph._CACHE[id(g)] = (g, graph.Structure(g, chunk=64))          # hub rows split into several items
    gen = torch.Generator().manual_seed(F)
    rows = torch.unique(torch.cat([torch.randint(0, N, (40,), generator=gen), torch.tensor([2, N - 1])])).cuda()
    b = torch.randn(F, generator=gen).cuda()
    wgt = torch.randn(rows.numel(), F, generator=gen).cuda()
    outs, grads = [], []
    for restricted in (False, True):
        x = torch.randn(N, F, generator=torch.Generator().manual_seed(3)).cuda().requires_grad_(True)
        bb = b.clone().requires_grad_(True)
        if restricted:
            y = ops.spmm_rows(g, x, rows, reduce, bias=bb, relu=True)
        else:
            y = ops.spmm(g, x, reduce, bias=bb, relu=True)[rows]
        (y * wgt).sum().backward()
        outs.append(y.detach().clone())
        grads.append((x.grad.clone(), bb.grad.clone()))
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(grads[0][0], grads[1][0])
    assert rel_err(grads[1][1], grads[0][1]) < TOL            # bias gradient: column sums over different row sets


def test_spmm_all_rows_empty(ops):
    from plnlp_b200.graph import CSRGraph
    N, F = 70, 8
    g = CSRGraph(torch.zeros(N + 1, dtype=torch.int64).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), None, (N, N))
    b = torch.randn(F).cuda()
    out = ops.spmm(g, torch.randn(N, F).cuda(), "sum", bias=b)
    assert torch.equal(out, b.expand(N, F))


@pytest.mark.parametrize("F", [2, 7, 50, 64, 128, 200, 256, 520])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_spmm_bf16_storage(ops, F, reduce):
    """the separately stated bf16 path: bf16 feature storage, fp32 accumulation in CSR order, ONE rounding at
    the store.  Oracle: the in-order fp32 loop on the bf16-rounded inputs, rounded to bf16 -> identical bits
    for unsplit rows (allowing 1 bf16 ulp where a hub row's partial sums are combined in a different order)."""
    from plnlp_b200.graph import Structure
    N = 211
    ei, w = rand_graph(N, 1500, seed=F, weighted=(reduce == "sum"), hub=True)
    o = sparse.to_sparse_tensor(ei, w, N)
    st = Structure(_to_gpu_graph(o))
    x = torch.randn(N, F).to(torch.bfloat16)
    plan = st.fwd if reduce == "sum" else st.fwd_noval
    got = ops.spmm_raw(plan, x.cuda(), use_val=(reduce == "sum"), div_rows=(reduce == "mean"))
    assert got.dtype == torch.bfloat16
    oref = o.set_value(None) if reduce == "mean" else o
    rowptr, col, val = oref.csr()
    want32 = cspmm.spmm(rowptr, col, val, x.float(), reduce)
    want = want32.to(torch.bfloat16)
    deg = rowptr[1:] - rowptr[:-1]
    unsplit = deg <= plan.chunk
    assert torch.equal(got.cpu()[unsplit], want[unsplit])
    assert rel_err(got.cpu().float(), want32) < 2.0 ** -8


@pytest.mark.parametrize("weighted,reduce", [(False, "mean"), (True, "sum")])
def test_spmm_backward_nonsymmetric(ops, weighted, reduce):
    N, F = 97, 40
    ei, w = rand_graph(N, 700, seed=3, weighted=weighted, hub=True)   # directed: A^T != A
    o = sparse.to_sparse_tensor(ei, w, N)
    g = _to_gpu_graph(o)
    x = torch.randn(N, F)
    gout = torch.randn(N, F)
    xg = x.cuda().requires_grad_(True)
    ops.spmm(g, xg, reduce).backward(gout.cuda())
    xc = x.clone().requires_grad_(True)
    sparse.matmul(o.set_value(None) if reduce == "mean" else o, xc, reduce).backward(gout)
    assert rel_err(xg.grad.cpu(), xc.grad) < TOL


def test_spmm_fused_epilogue_and_grad(ops):
    """GCN epilogue: relu(A z + bias), backward through the mask, bias gradient"""
    N, F = 120, 64
    ei, _ = rand_graph(N, 900, seed=4)
    o = sparse.gcn_normalization(sparse.to_sparse_tensor(ei, None, N).to_symmetric())
    g = _to_gpu_graph(o)
    z, b, gout = torch.randn(N, F), torch.randn(F), torch.randn(N, F)
    zg, bg = z.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.spmm(g, zg, "sum", bias=bg, relu=True)
    out.backward(gout.cuda())
    zc, bc = z.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.relu(sparse.matmul(o, zc, "sum") + bc)
    ref.backward(gout)
    assert rel_err(out.cpu(), ref) < TOL
    assert rel_err(zg.grad.cpu(), zc.grad) < TOL and rel_err(bg.grad.cpu(), bc.grad) < TOL


def test_dropout_statistics_and_backward(ops):
    N, F, p = 400, 128, 0.3
    ei, _ = rand_graph(N, 6000, seed=6)
    g = _to_gpu_graph(sparse.to_sparse_tensor(ei, None, N))
    x = (torch.rand(N, F) + 0.5).cuda().requires_grad_(True)
    out = ops.spmm(g, x, "sum", relu=True, drop_p=p, seed=1234)
    base = ops.spmm(g, x.detach(), "sum", relu=True)
    nz = base > 0
    kept = (out != 0) & nz
    frac = kept.sum().item() / nz.sum().item()
    assert abs(frac - (1 - p)) < 0.02
    assert rel_err(out[kept], base[kept] / (1 - p)) < 1e-6
    out2 = ops.spmm(g, x.detach(), "sum", relu=True, drop_p=p, seed=1234)
    assert torch.equal(out.detach(), out2)                      # same seed, same mask
    out.sum().backward()
    # gradient flows only through kept entries, scaled by 1/(1-p): d/dx = A^T (mask/(1-p))
    gref = ops.spmm_raw(ops.structure_of(g).bwd, (kept.float() / (1 - p)), use_val=False, div_rows=False)
    assert rel_err(x.grad, gref) < TOL


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 19, 23), (128, 128, 16), (4267, 512, 512), (300, 200, 178),
                                   (129, 257, 50), (512, 512, 4100)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_layouts(ops, M, N, K, ta, tb):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    got = ops.gemm_raw(A.cuda(), B.cuda(), transa=ta, transb=tb).cpu()
    want = ((A.t() if ta else A).double() @ (B.t() if tb else B).double())
    assert rel_err(got, want) < TOL


def test_gemm_splitk_deterministic_and_beta_bias(ops):
    g = torch.Generator().manual_seed(1)
    A, B = torch.randn(20000, 96, generator=g), torch.randn(20000, 130, generator=g)
    C0, bias = torch.randn(96, 130, generator=g), torch.randn(130, generator=g)
    outs = []
    for _ in range(2):
        C = C0.clone().cuda()
        ops.gemm_raw(A.cuda(), B.cuda(), transa=True, C=C, beta=1.0, bias=bias.cuda(), split_k=7)
        outs.append(C.cpu())
    assert torch.equal(outs[0], outs[1])
    want = A.double().t() @ B.double() + C0.double() + bias.double()
    assert rel_err(outs[0], want) < TOL


def test_gemm_views_with_leading_dimension(ops):
    """column slices of a wider weight (concat-free [emb | x] @ W^T) and misaligned bases"""
    g = torch.Generator().manual_seed(2)
    W = torch.randn(200, 178, generator=g)
    e, x = torch.randn(333, 50, generator=g), torch.randn(333, 128, generator=g)
    Wg = W.cuda()
    y = ops.gemm_raw(e.cuda(), Wg[:, :50], transb=True)
    y = ops.gemm_raw(x.cuda(), Wg[:, 50:], transb=True, C=y, beta=1.0)
    want = torch.cat([e, x], 1).double() @ W.double().t()
    assert rel_err(y.cpu(), want) < TOL


def test_fused_linear_autograd(ops):
    g = torch.Generator().manual_seed(3)
    a, x = torch.randn(90, 24, generator=g), torch.randn(90, 24, generator=g)
    wl, wr, b = torch.randn(40, 24, generator=g), torch.randn(40, 24, generator=g), torch.randn(40, generator=g)
    gout = torch.randn(90, 40, generator=g)
    cu = [t.cuda().requires_grad_(True) for t in (a, x, wl, wr, b)]
    y = ops.fused_linear([cu[0], cu[1]], [cu[2], cu[3]], cu[4], act=ops.ACT_RELU)
    y.backward(gout.cuda())
    cp = [t.clone().requires_grad_(True) for t in (a, x, wl, wr, b)]
    yr = torch.relu(cp[0] @ cp[2].t() + cp[4] + cp[1] @ cp[3].t())
    yr.backward(gout)
    assert rel_err(y.cpu(), yr) < TOL
    for u, v in zip(cu, cp):
        assert rel_err(u.grad.cpu(), v.grad) < TOL


# ------------------------------------------------------------------ edge scoring
@pytest.mark.parametrize("H", [1, 12, 50, 200, 256, 512])
def test_gather_hadamard_and_dot(ops, H):
    g = torch.Generator().manual_seed(H)
    N, P = 77, 501
    h = torch.randn(N, H, generator=g)
    edges = torch.randint(0, N, (P, 2), generator=g)
    edges[5] = torch.tensor([-1, 3])                      # python-style negative index (model.py:191-194)
    edges[6] = torch.tensor([9, 9])                       # self pair
    want = h[edges[:, 0]] * h[edges[:, 1]]
    assert torch.equal(ops.gather_hadamard_raw(h.cuda(), edges.cuda()).cpu(), want)
    assert rel_err(ops.edge_dot_raw(h.cuda(), edges.cuda()).cpu(), want.double().sum(-1)) < TOL


@pytest.mark.parametrize("mode", ["sorted", "atomic"])
@pytest.mark.parametrize("head", ["mlp", "dot"])
def test_edge_scatter(ops, mode, head):
    g = torch.Generator().manual_seed(11)
    N, P, H = 60, 700, 36
    h = torch.randn(N, H, generator=g)
    edges = torch.randint(0, N, (P, 2), generator=g)
    edges[:40, 0] = 7                                       # a hot node
    edges[3] = torch.tensor([5, 5])
    hc = h.clone().requires_grad_(True)
    if head == "mlp":
        da = torch.randn(P, H, generator=g)
        (hc[edges[:, 0]] * hc[edges[:, 1]] * da).sum().backward()
        got = ops.edge_scatter_raw(h.cuda(), edges.cuda(), da=da.cuda(), mode=mode).cpu()
    else:
        ds = torch.randn(P, generator=g)
        ((hc[edges[:, 0]] * hc[edges[:, 1]]).sum(-1) * ds).sum().backward()
        got = ops.edge_scatter_raw(h.cuda(), edges.cuda(), dscore=ds.cuda(), mode=mode).cpu()
    assert rel_err(got, hc.grad) < TOL
    if mode == "sorted":
        again = ops.edge_scatter_raw(h.cuda(), edges.cuda(), da=da.cuda() if head == "mlp" else None,
                                     dscore=None if head == "mlp" else ds.cuda(), mode=mode).cpu()
        assert torch.equal(got, again)


def test_mlp_out_layer(ops):
    g = torch.Generator().manual_seed(12)
    P, H = 1000, 200
    a = torch.relu(torch.randn(P, H, generator=g))
    w, b, ds = torch.randn(1, H, generator=g), torch.randn(1, generator=g), torch.randn(P, generator=g)
    s = ops.mlp_out_fwd_raw(a.cuda(), w.cuda(), b.cuda()).cpu()
    assert rel_err(s, (a.double() @ w.double().t()).reshape(-1) + b.double()) < TOL
    dz, dw, db = ops.mlp_out_bwd_raw(a.cuda(), w.cuda(), ds.cuda(), mask_a=True, drop_scale=1.0)
    assert rel_err(dz.cpu(), ds[:, None] * w * (a > 0)) < TOL
    assert rel_err(dw.cpu(), (ds.double()[:, None] * a.double()).sum(0)) < TOL
    assert rel_err(db.cpu(), ds.double().sum().reshape(1)) < TOL


@pytest.mark.parametrize("name", ["AUC", "HingeAUC", "WeightedHingeAUC"])
@pytest.mark.parametrize("k", [1, 3])
def test_pair_loss(ops, name, k):
    g = torch.Generator().manual_seed(13)
    B = 5000
    pos, neg, w = torch.randn(B, generator=g), torch.randn(B * k, generator=g), torch.rand(B, generator=g) + 0.1
    p, n = pos.cuda().requires_grad_(True), neg.cuda().requires_grad_(True)
    loss = ops.pair_loss(name, p, n, k, w.cuda() if name == "WeightedHingeAUC" else None)
    (loss * 0.5).backward()
    want = plnlp_ref.pair_loss(name, pos.double(), neg.double(), k, w.double())
    gp, gn = plnlp_ref.pair_loss_grad(name, pos, neg, k, w)
    assert rel_err(loss.cpu(), want) < TOL
    assert rel_err(p.grad.cpu(), 0.5 * gp) < TOL and rel_err(n.grad.cpu(), 0.5 * gn.reshape(-1)) < TOL


def test_losses_against_reference_golden(ops, golden_dir):
    import os
    from plnlp_b200 import loss as L
    G = torch.load(os.path.join(golden_dir, "losses.pt"))
    for key, rec in G.items():
        k = int(key[-1])
        for name, fn in (("AUC", L.auc_loss), ("HingeAUC", L.hinge_auc_loss),
                         ("WeightedHingeAUC", L.weighted_hinge_auc_loss), ("WeightedAUC", L.weighted_auc_loss),
                         ("AdaAUC", L.adaptive_auc_loss), ("AdaHingeAUC", L.adaptive_hinge_auc_loss),
                         ("LogRank", L.log_rank_loss), ("CE", L.ce_loss), ("InfoNCE", L.info_nce_loss)):
            p, n = rec["pos"].cuda().requires_grad_(True), rec["neg"].cuda().requires_grad_(True)
            if name == "CE":
                args = (p, n)
            elif name in ("WeightedHingeAUC", "WeightedAUC", "AdaAUC", "AdaHingeAUC"):
                args = (p, n, k, rec["weight"].cuda())
            else:
                args = (p, n, k)
            loss = fn(*args)
            loss.backward()
            assert loss.dim() == 0
            assert rel_err(loss.cpu(), rec[name]["loss"]) < TOL
            assert rel_err(p.grad.cpu(), rec[name]["gpos"]) < TOL
            assert rel_err(n.grad.cpu(), rec[name]["gneg"]) < TOL


def test_relu_bwd_and_colsum(ops):
    g = torch.Generator().manual_seed(14)
    y, dy = torch.randn(1234, 178, generator=g), torch.randn(1234, 178, generator=g)
    assert torch.equal(ops.relu_drop_bwd_raw(y.cuda(), dy.cuda(), 2.0).cpu(), torch.where(y > 0, dy * 2.0, 0.0))
    assert rel_err(ops.colsum_raw(y.cuda(), 0.5).cpu(), 0.5 * y.double().sum(0)) < TOL


# ------------------------------------------------------------------ samplers
def test_local_neg_sample(ops):
    from plnlp_b200.negative_sample import local_neg_sample
    torch.manual_seed(0)
    N, E, k = 1000, 4001, 3
    pos = torch.randint(0, N, (E, 2))
    out = local_neg_sample(pos.cuda(), N, k)
    assert out.shape == (E, k, 2) and out.dtype == torch.int64
    out = out.cpu()
    assert torch.equal(out[:, :, 0], pos[:, :1].expand(E, k))      # sources kept: bit-exact
    dst = out[:, :, 1].reshape(-1)
    assert dst.min() >= 0 and dst.max() < N
    counts = torch.bincount(dst, minlength=N).double()
    chi2 = ((counts - counts.mean()) ** 2 / counts.mean()).sum().item()
    assert chi2 < N + 6 * (2 * N) ** 0.5                            # uniform within 6 sigma
    torch.manual_seed(0)
    pos2 = torch.randint(0, N, (E, 2))
    assert torch.equal(local_neg_sample(pos2.cuda(), N, k).cpu(), out)   # reproducible under manual_seed


def test_local_neg_sample_random_src(ops):
    """negative_sample.py:32-34: with random_src the kept endpoint of every positive is one of its two ends, drawn
    uniformly; the k negatives of a positive share it; destinations as in the plain sampler"""
    from plnlp_b200.negative_sample import local_neg_sample
    torch.manual_seed(1)
    N, E, k = 500, 6000, 2
    pos = torch.stack([torch.randint(0, N, (E,)), torch.randint(0, N, (E,)) + N], 1)      # ends are distinguishable
    out = local_neg_sample(pos.cuda(), 2 * N, k, random_src=True).cpu()
    assert out.shape == (E, k, 2) and out.dtype == torch.int64
    src = out[:, :, 0]
    assert torch.equal(src[:, 0], src[:, 1])                        # one draw per positive
    from_first, from_second = src[:, 0] == pos[:, 0], src[:, 0] == pos[:, 1]
    assert bool((from_first ^ from_second).all())                   # always one of the two ends
    n1 = int(from_first.sum())
    assert abs(n1 - E / 2) < 6 * (E / 4) ** 0.5                     # a fair coin within 6 sigma
    dst = out[:, :, 1].reshape(-1)
    assert dst.min() >= 0 and dst.max() < 2 * N


def test_global_neg_sample(ops):
    from plnlp_b200.negative_sample import global_neg_sample
    torch.manual_seed(1)
    N = 300
    ei, _ = rand_graph(N, 9000, seed=15)
    E, k = 5000, 3
    out = global_neg_sample(ei.cuda(), N, E, k).cpu()
    assert out.shape == (E, k, 2) and out.dtype == torch.int64
    src, dst = out[..., 0].reshape(-1), out[..., 1].reshape(-1)
    assert (src != dst).all()                                       # self loops excluded (negative_sample.py:8)
    ids = src * N + dst
    assert not np.isin(ids.numpy(), (ei[0] * N + ei[1]).numpy()).any()   # never an existing edge
    assert ids.unique().numel() == ids.numel()                      # distinct
    # roughly uniform over rows
    counts = torch.bincount(src, minlength=N).double()
    assert counts.min() > 0.3 * counts.mean() and counts.max() < 2.0 * counts.mean()


def test_global_neg_sample_pads_when_short(ops):
    from plnlp_b200.negative_sample import global_neg_sample
    torch.manual_seed(2)
    N = 12
    ei, _ = rand_graph(N, 60, seed=16)
    out = global_neg_sample(ei.cuda(), N, 100, 2).cpu()              # asks for more than exist
    assert out.shape == (100, 2, 2)
    ids = (out[..., 0] * N + out[..., 1]).reshape(-1)
    assert not np.isin(ids.numpy(), (ei[0] * N + ei[1]).numpy()).any()
    assert (out[..., 0] != out[..., 1]).all()


def test_global_perm_neg_sample(ops):
    """negative_sample.py:23-28: num_samples distinct non-edges, then num_neg - 1 permuted copies of the SAME set,
    in the reference's [E, k, 2] reshape of the concatenated copies"""
    from plnlp_b200.negative_sample import global_perm_neg_sample
    torch.manual_seed(3)
    N = 300
    ei, _ = rand_graph(N, 9000, seed=17)
    E, k = 4000, 3
    out = global_perm_neg_sample(ei.cuda(), N, E, k).cpu()
    assert out.shape == (E, k, 2) and out.dtype == torch.int64
    flat = out.reshape(-1, 2)                         # = the k concatenated copies, copy c = rows c*E .. (c+1)*E
    ids = flat[:, 0] * N + flat[:, 1]
    assert not np.isin(ids.numpy(), (ei[0] * N + ei[1]).numpy()).any()
    assert (flat[:, 0] != flat[:, 1]).all()
    base = ids[:E]
    assert base.unique().numel() == E                 # the first copy is E distinct negatives
    for c in range(1, k):                             # every further copy is a permutation of the first
        assert torch.equal(torch.sort(ids[c * E:(c + 1) * E])[0], torch.sort(base)[0])
        assert not torch.equal(ids[c * E:(c + 1) * E], base)


@pytest.mark.parametrize("H", [1, 6, 50, 128, 200])
def test_gather_rows_and_sorted_row_scatter(ops, H):
    """x = h[edges[:, side]] read in place from the [P, 2] tensor (negative index = last rows), and its backward
    against index_add_ in fp64; deterministic"""
    N, P = 97, 1500
    g = torch.Generator().manual_seed(H)
    h = torch.randn(N, H, generator=g)
    edges = torch.randint(0, N - 5, (P, 2), generator=g)          # last rows never referenced
    edges[::50, 1] = -1
    gout = torch.randn(P, H, generator=g)
    for side in (0, 1):
        hg = h.cuda().requires_grad_(True)
        x = ops.GatherRows.apply(hg, edges.cuda(), side)
        assert torch.equal(x.detach().cpu(), h[edges[:, side]])
        x.backward(gout.cuda())
        idx = torch.where(edges[:, side] < 0, edges[:, side] + N, edges[:, side])
        want = torch.zeros(N, H, dtype=torch.float64).index_add_(0, idx, gout.double())
        assert rel_err(hg.grad.cpu(), want) < TOL
        assert torch.all(hg.grad[N - 5:N - 1] == 0)
        again = ops.row_scatter_raw(gout.cuda(), edges[:, side].cuda(), N)
        assert torch.equal(again, hg.grad)


def test_segment_softmax_forward_and_backward(ops):
    """per-row softmax over CSR entries (rows of length 0, 1, 33, a hub) against torch in fp64"""
    lens = torch.tensor([0, 1, 5, 33, 0, 64, 700, 2, 0])
    rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lens, 0)])
    nnz = int(rowptr[-1])
    g = torch.Generator().manual_seed(4)
    s = (torch.randn(nnz, generator=g) * 3)
    gout = torch.randn(nnz, generator=g)
    scale = 0.37
    sg = s.cuda().requires_grad_(True)
    alpha = ops.SegmentSoftmax.apply(sg, rowptr.cuda(), scale)
    alpha.backward(gout.cuda())
    sc = s.double().requires_grad_(True)
    parts = [torch.softmax(sc[rowptr[i]:rowptr[i + 1]] * scale, 0) for i in range(lens.numel())]
    ref = torch.cat(parts)
    ref.backward(gout.double())
    assert rel_err(alpha.detach().cpu(), ref.detach()) < TOL
    assert rel_err(sg.grad.cpu(), sc.grad) < TOL
    for i in range(lens.numel()):
        if lens[i] > 0:
            assert abs(float(alpha.detach()[rowptr[i]:rowptr[i + 1]].sum()) - 1.0) < 1e-5


# ------------------------------------------------------------------ ranking
@pytest.mark.parametrize("n,K", [(1, 1), (50, 20), (1000, 100), (101882, 20), (101882, 50), (300000, 100)])
def test_kth_largest_and_hits(ops, n, K):
    from plnlp_b200.utils import hits_at_k
    g = torch.Generator().manual_seed(n + K)
    neg = torch.randn(n, generator=g)
    neg[: n // 3] = neg[: n // 3].round(decimals=1)               # plenty of exact ties
    if n > 10:
        neg[3], neg[4] = 0.0, -0.0
    pos = torch.randn(5000, generator=g).round(decimals=1)
    kth = ops.kth_largest_raw(neg.cuda(), K).cpu()
    assert torch.equal(kth, torch.topk(neg, K)[0][-1:])             # bit-exact
    assert hits_at_k(pos.cuda(), neg.cuda(), K) == ogb_eval.hits_at_k(pos, neg, K)


def test_hits_fewer_negatives_than_k(ops):
    from plnlp_b200.utils import hits_at_k
    assert hits_at_k(torch.randn(10).cuda(), torch.randn(5).cuda(), 20) == 1.0


def test_mrr_counts(ops):
    from plnlp_b200.utils import mrr_list
    g = torch.Generator().manual_seed(17)
    S, K = 997, 1000
    pos, neg = torch.randn(S, generator=g), torch.randn(S, K, generator=g)
    neg[:, 5] = pos                                              # an exact tie in every row
    gt, ge = ops.mrr_counts_raw(pos.cuda(), neg.cuda())
    ogt, oge = ogb_eval.mrr_ranks(pos, neg)
    assert torch.equal(gt.cpu().long() + 1, ogt) and torch.equal(ge.cpu().long() + 1, oge)
    # ties: the mean of the optimistic and the pessimistic rank (current ogb), not the optimistic one
    assert torch.equal(mrr_list(pos.cuda(), neg.cuda()).cpu(), 1.0 / (0.5 * (ogt + oge).float()))
    pos2, neg2 = torch.randn(S, generator=g), torch.randn(S, K, generator=g)       # no ties: ogb 1.3.2's rank
    o2, p2 = ogb_eval.mrr_ranks(pos2, neg2)
    assert torch.equal(o2, p2) and torch.equal(mrr_list(pos2.cuda(), neg2.cuda()).cpu(), 1.0 / o2.float())
    flat = torch.zeros(S)                                        # a collapsed model does not report MRR = 1
    assert float(mrr_list(flat.cuda(), torch.zeros(S, K).cuda()).max()) < 0.01


def test_eval_glue_against_reference_golden(ops, golden_dir):
    import os
    from plnlp_b200.utils import evaluate_hits, evaluate_mrr, get_pos_neg_edges
    G = torch.load(os.path.join(golden_dir, "edges_eval.pt"))
    h = G["hits"]
    assert evaluate_hits(None, h["pv"], h["nv"], h["pt"], h["nt"]) == h["res"]
    m = G["mrr"]
    got = evaluate_mrr(None, m["pv"], m["nv"], m["pt"], m["nt"])["MRR"]
    assert abs(got[0] - m["res"]["MRR"][0]) < 1e-6 and abs(got[1] - m["res"]["MRR"][1]) < 1e-6
    c = G["citation_style"]
    pos, neg = get_pos_neg_edges("valid", c["split"], device=torch.device("cuda"))
    assert torch.equal(pos.cpu(), c["pos"]) and torch.equal(neg.cpu(), c["neg"])


# ------------------------------------------------------------------ tensor-core GEMM (tcgen05)
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 256, 64), (256, 512, 512), (4267, 512, 512),
                                   (300, 200, 178), (129, 257, 50), (512, 512, 4100), (1000, 72, 96)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tf32x3_layouts(ops, M, N, K, ta, tb):
    """3xTF32 on tcgen05 must sit inside the fp32 parity bar for every operand layout"""
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    got = ops.gemm_raw(A.cuda(), B.cuda(), transa=ta, transb=tb, backend="tf32x3").cpu()
    want = ((A.t() if ta else A).double() @ (B.t() if tb else B).double())
    assert rel_err(got, want) < TOL, rel_err(got, want)


def test_gemm_tf32_plain_and_epilogues(ops):
    g = torch.Generator().manual_seed(4)
    A, W = torch.randn(1000, 512, generator=g), torch.randn(384, 512, generator=g)
    bias, C0 = torch.randn(384, generator=g), torch.randn(1000, 384, generator=g)
    want = A.double() @ W.double().t()
    fast = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, backend="tf32").cpu()
    assert 1e-5 < rel_err(fast, want) < 5e-3                   # really TF32, stated separately
    # bias + beta + relu epilogue, and the relu-grad epilogue
    C = C0.clone().cuda()
    ops.gemm_raw(A.cuda(), W.cuda(), transb=True, C=C, beta=1.0, bias=bias.cuda(), act=ops.ACT_RELU,
                 backend="tf32x3")
    ref = torch.relu(want + C0.double() + bias.double())
    assert rel_err(C.cpu(), ref) < TOL
    aux = torch.randn(1000, 384, generator=g)
    G = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU_GRAD, aux=aux.cuda(), backend="tf32x3")
    assert rel_err(G.cpu(), want * (aux > 0)) < TOL
    # dropout epilogue uses the same Philox stream as the FFMA kernel
    d1 = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU, drop_p=0.3, seed=99, backend="tf32x3")
    d2 = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU, drop_p=0.3, seed=99, backend="ffma")
    assert torch.equal(d1 == 0, d2 == 0) and rel_err(d1, d2) < TOL


def test_gemm_tf32x3_splitk(ops):
    g = torch.Generator().manual_seed(5)
    A, B = torch.randn(30000, 512, generator=g), torch.randn(30000, 512, generator=g)
    outs = [ops.gemm_raw(A.cuda(), B.cuda(), transa=True, split_k=9, backend="tf32x3").cpu() for _ in range(2)]
    assert torch.equal(outs[0], outs[1])
    assert rel_err(outs[0], A.double().t() @ B.double()) < TOL


@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (128, 256, 64), (512, 512, 512), (4267, 512, 512),
                                   (300, 200, 178), (129, 257, 50), (512, 512, 4100), (1000, 72, 96)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tf32x3_2cta_layouts(ops, M, N, K, ta, tb):
    """CTA-pair (cta_group::2) variant: same bar, every layout, ragged M / N / K"""
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    got = ops.gemm_raw(A.cuda(), B.cuda(), transa=ta, transb=tb, backend="tf32x3c2").cpu()
    want = ((A.t() if ta else A).double() @ (B.t() if tb else B).double())
    assert rel_err(got, want) < TOL, rel_err(got, want)


@pytest.mark.parametrize("M,N,K", [(300, 200, 178), (1000, 50, 200), (257, 512, 96), (129, 16, 32), (4096, 256, 512),
                                   (130, 208, 1000), (20000, 200, 178), (70000, 64, 200), (5000, 300, 40)])
@pytest.mark.parametrize("tb", [True, False])
def test_gemm_tma_tall_skinny(ops, M, N, K, tb, monkeypatch):
    """TMA-fed persistent kernel (csrc/gemm_tma.cu): ragged M / N / K (zero fill by the tensor map), both layouts of
    the weight, more tiles than SMs (persistent loop, both TMEM accumulators, ring wrap-around), two column tiles;
    the same 3xTF32 bar as the other tensor-core kernels, and bit-identical to the CTA-pair kernel (same split, same
    MMA order per k-step)"""
    monkeypatch.setattr(ops, "GEMM_TMA", "1")
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn(M, (K + 3) // 4 * 4 + 4, generator=g)          # rows on a 16-byte pitch, wider than K
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    Ag = A.cuda()[:, :K]
    n0 = _launches()
    got = ops.gemm_raw(Ag, B.cuda(), transb=tb, backend="tf32x3c2")
    if M * N * K >= (1 << 20):                                    # (tiny products go to the FFMA kernel)
        assert _launches() - n0 == 2                               # weight split + the GEMM: the TMA path ran
    want = A[:, :K].double() @ (B.t() if tb else B).double()
    assert rel_err(got.cpu(), want) < TOL, rel_err(got.cpu(), want)
    monkeypatch.setattr(ops, "GEMM_TMA", "0")
    old = ops.gemm_raw(Ag, B.cuda(), transb=tb, backend="tf32x3c2")
    assert rel_err(got, old) < 2e-6


def _launches():
    from plnlp_b200 import _lib
    return _lib.launch_count()


def test_gemm_tma_epilogues(ops, monkeypatch):
    monkeypatch.setattr(ops, "GEMM_TMA", "1")
    g = torch.Generator().manual_seed(44)
    M, N, K = 3000, 200, 180
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    bias, C0 = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    want = A.double() @ W.double().t()
    n0 = _launches()
    C = C0.clone().cuda()
    ops.gemm_raw(A.cuda(), W.cuda(), transb=True, C=C, beta=1.0, bias=bias.cuda(), act=ops.ACT_RELU)
    assert rel_err(C.cpu(), torch.relu(want + C0.double() + bias.double())) < TOL
    aux = torch.randn(M, N, generator=g)
    G = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU_GRAD, aux=aux.cuda())
    assert rel_err(G.cpu(), want * (aux > 0)) < TOL
    d1 = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU, drop_p=0.3, seed=99)
    assert _launches() - n0 == 6
    d2 = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU, drop_p=0.3, seed=99, backend="ffma")
    assert torch.equal(d1 == 0, d2 == 0) and rel_err(d1, d2) < TOL
    # output into a column slice of a wider matrix (unaligned leading dimension: scalar epilogue stores)
    wide = torch.full((M, N + 7), -3.0).cuda()
    ops.gemm_raw(A.cuda(), W.cuda(), transb=True, C=wide[:, 3:3 + N])
    assert rel_err(wide[:, 3:3 + N].cpu(), want) < TOL and torch.all(wide[:, :3] == -3.0) and torch.all(wide[:, 3 + N:] == -3.0)
    # plain TF32 (single pass), stated separately
    fast = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, backend="tf32c2").cpu()
    assert 1e-5 < rel_err(fast, want) < 5e-3


@pytest.mark.parametrize("M,N,K,lda,ldb", [(200, 179, 40000, 200, 180), (128, 64, 20011, 128, 64), (256, 256, 5000, 256, 256),
                                           (72, 8, 3000, 72, 8), (200, 180, 1088 * 3 + 5, 204, 184), (33, 250, 1088, 36, 252),
                                           (130, 17, 600, 132, 20)])
def test_gemm_tma_tn_weight_gradient(ops, M, N, K, lda, ldb, monkeypatch):
    """C = A^T B over a huge row count on the TMA-fed kernel (MN-major operands straight from the TMA boxes, hi / lo split
    by convert warps, one accumulator per 1088 rows, CTA partials added in order): against fp64, deterministic, operands
    with leading dimensions, partial units / slabs / 32-column atoms, one and two 128-row halves"""
    from plnlp_b200 import _lib
    monkeypatch.setattr(ops, "GEMM_TMA_TN", "1")
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(K, lda, generator=g)[:, :M]
    B = torch.randn(K, ldb, generator=g)[:, :N]
    Ag, Bg = torch.full((K, lda), float("nan")).cuda()[:, :M], torch.full((K, ldb), float("nan")).cuda()[:, :N]
    Ag.copy_(A)
    Bg.copy_(B)
    want = A.double().t() @ B.double()
    n0 = _lib.launch_count()
    got = ops.gemm_raw(Ag, Bg, transa=True, backend="tf32x3c2")
    assert _lib.launch_count() - n0 == 2                      # the TMA kernel + the reduction of the CTA partials
    assert rel_err(got.cpu(), want) < TOL
    assert torch.equal(got, ops.gemm_raw(Ag, Bg, transa=True, backend="tf32x3c2"))
    wide = torch.full((M, N + 5), -2.0).cuda()
    ops.gemm_raw(Ag, Bg, transa=True, C=wide[:, 2:2 + N], backend="tf32x3c2")
    assert torch.equal(wide[:, 2:2 + N], got) and torch.all(wide[:, :2] == -2.0) and torch.all(wide[:, 2 + N:] == -2.0)
    fast = ops.gemm_raw(Ag, Bg, transa=True, backend="tf32c2")
    assert rel_err(fast.cpu(), want) < 5e-3


def test_gemm_tf32x3_2cta_epilogues_and_splitk(ops):
    g = torch.Generator().manual_seed(6)
    A, W = torch.randn(1000, 512, generator=g), torch.randn(384, 512, generator=g)
    bias, C0 = torch.randn(384, generator=g), torch.randn(1000, 384, generator=g)
    want = A.double() @ W.double().t()
    C = C0.clone().cuda()
    ops.gemm_raw(A.cuda(), W.cuda(), transb=True, C=C, beta=1.0, bias=bias.cuda(), act=ops.ACT_RELU, backend="tf32x3c2")
    assert rel_err(C.cpu(), torch.relu(want + C0.double() + bias.double())) < TOL
    d1 = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU, drop_p=0.3, seed=99, backend="tf32x3c2")
    d2 = ops.gemm_raw(A.cuda(), W.cuda(), transb=True, act=ops.ACT_RELU, drop_p=0.3, seed=99, backend="ffma")
    assert torch.equal(d1 == 0, d2 == 0) and rel_err(d1, d2) < TOL
    X, Y = torch.randn(30000, 512, generator=g), torch.randn(30000, 512, generator=g)
    outs = [ops.gemm_raw(X.cuda(), Y.cuda(), transa=True, backend="tf32x3c2").cpu() for _ in range(2)]
    assert torch.equal(outs[0], outs[1]) and rel_err(outs[0], X.double().t() @ Y.double()) < TOL


# ------------------------------------------------------------------ random-walk augmentation
def test_random_walk_bit_exact_with_supplied_uniforms(ops):
    from oracle import rw
    from plnlp_b200 import augment
    N = 500
    ei, _ = rand_graph(N, 3000, seed=31, hub=True)                 # has isolated nodes (deg 0: stay put)
    adj = sparse.to_sparse_tensor(ei, None, N)
    rowptr, col, _ = adj.csr()
    g = torch.Generator().manual_seed(4)
    start = torch.randint(0, N, (4000,), generator=g)
    rand = torch.rand(4000, 10, generator=g)
    want = rw.random_walk(rowptr, col, start, 10, rand)
    got = augment.random_walk(None, col.cuda(), start.cuda(), 10, rowptr=rowptr.cuda(), rand=rand.cuda())
    assert torch.equal(got.cpu(), want)                             # index work: bit-exact
    pw, ww = rw.walk_pairs(want)
    pg, wg = augment.walk_pairs(got)
    assert torch.equal(pg.cpu(), pw) and torch.equal(wg.cpu(), ww)


def test_random_walk_philox_statistics(ops):
    from plnlp_b200 import augment
    from plnlp_b200.graph import CSRGraph
    torch.manual_seed(5)
    N = 64
    src = torch.arange(N).repeat_interleave(N - 1)
    dst = torch.stack([torch.cat([torch.arange(i), torch.arange(i + 1, N)]) for i in range(N)]).reshape(-1)
    adj = CSRGraph.from_edge_index(torch.stack([src, dst]).cuda(), None, N)       # complete graph
    start = torch.zeros(200000, dtype=torch.int64, device="cuda")
    walk = augment.random_walk(None, adj.csr()[1], start, 3, rowptr=adj.csr()[0])
    assert walk.shape == (200000, 4) and (walk[:, 0] == 0).all()
    assert (walk[:, 1:] != walk[:, :-1]).all()                      # no self loops in the graph -> always moves
    counts = torch.bincount(walk[:, 1], minlength=N).double().cpu()[1:]
    chi2 = ((counts - counts.mean()) ** 2 / counts.mean()).sum().item()
    assert chi2 < (N - 1) + 6 * (2 * (N - 1)) ** 0.5               # first step uniform over the 63 neighbours
    edges, w = augment.random_walk_pairs(adj, start[:1000], 5)
    assert edges.shape[1] == 2 and (edges[:, 0] != edges[:, 1]).all() and w.numel() == edges.size(0)


# ------------------------------------------------------------------ fused edge scoring (MLP head)
@pytest.mark.parametrize("N,H,P", [(77, 64, 300), (4267, 512, 5000), (999, 200, 1234), (500, 256, 257)])
def test_fused_edge_mlp_forward(ops, N, H, P):
    """gather + Hadamard + Linear/relu + Linear(->1) in one kernel vs the oracle predictor"""
    g = torch.Generator().manual_seed(N + H)
    h = torch.randn(N, H, generator=g)
    edges = torch.randint(0, N, (P, 2), generator=g)
    edges[1] = torch.tensor([-1, 3])
    W1, b1 = torch.randn(H, H, generator=g) / H ** 0.5, torch.randn(H, generator=g)
    w2, b2 = torch.randn(1, H, generator=g) / H ** 0.5, torch.randn(1, generator=g)
    score, a1 = ops.edge_mlp_fwd_raw(h.cuda(), edges.cuda(), W1.cuda(), b1.cuda(), w2.cuda(), b2.cuda())
    a0 = (h[edges[:, 0]] * h[edges[:, 1]]).double()
    ref_a1 = torch.relu(a0 @ W1.double().t() + b1.double())
    ref_s = (ref_a1 @ w2.double().t()).reshape(-1) + b2.double()
    assert rel_err(a1.cpu(), ref_a1) < TOL
    assert rel_err(score.cpu(), ref_s) < TOL
    s2, none = ops.edge_mlp_fwd_raw(h.cuda(), edges.cuda(), W1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), need_a1=False)
    assert none is None and torch.equal(s2, score)
    # dropout: same Philox stream as the unfused GEMM epilogue
    sd, ad = ops.edge_mlp_fwd_raw(h.cuda(), edges.cuda(), W1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), 0.3, 77)
    a0g = ops.gather_hadamard_raw(h.cuda(), edges.cuda())
    au = ops.gemm_raw(a0g, W1.cuda(), transb=True, bias=b1.cuda(), act=ops.ACT_RELU, drop_p=0.3, seed=77)
    assert torch.equal(ad == 0, au == 0) and rel_err(ad, au) < TOL
    assert rel_err(sd, ops.mlp_out_fwd_raw(au, w2.cuda(), b2.cuda())) < TOL


@pytest.mark.parametrize("N,H,N1,B,k,drop", [(300, 128, 128, 200, 3, 0.0), (4267, 512, 512, 1500, 3, 0.3),
                                             (999, 200, 200, 1234, 1, 0.0), (500, 64, 192, 257, 2, 0.5),
                                             (77, 256, 320, 4100, 3, 0.0)])
def test_fused_and_unfused_step_agree(ops, N, H, N1, B, k, drop):
    """EdgeScoreLoss three ways -- fused forward + fused backward (dZ1 formed in the GEMM loaders, Hadamard product
    re-gathered inside the weight-gradient GEMM), fused forward + unfused backward, everything unfused: same loss and
    gradients; the gradients of the fused backward also against fp64 autograd on the same dropout mask"""
    from plnlp_b200 import _lib
    g = torch.Generator().manual_seed(21 + N)
    h0 = torch.randn(N, H, generator=g)
    pos, neg = torch.randint(0, N, (B, 2), generator=g), torch.randint(0, N, (B * k, 2), generator=g)
    pos[0] = torch.tensor([-1, 3])                       # negative index = counted from the end
    params0 = [torch.randn(N1, H, generator=g) / H ** 0.5, torch.randn(N1, generator=g),
               torch.randn(1, N1, generator=g) / N1 ** 0.5, torch.randn(1, generator=g)]
    out = {}
    for mode in ("fused", "fused_fwd", "unfused"):
        ops.FUSED_EDGE_MLP = mode != "unfused"
        ops.FUSED_EDGE_BWD = "1" if mode == "fused" else "0"
        try:
            h = h0.cuda().requires_grad_(True)
            params = [p.cuda().requires_grad_(True) for p in params0]
            n0 = _lib.launch_count()
            loss = ops.edge_score_loss(h, pos.cuda(), neg.cuda(), k, "AUC", head="MLP", params=params, drop_p=drop, seed=5)
            loss.backward()
            out[mode] = [loss.detach(), h.grad] + [p.grad for p in params]
            out[mode + "_launches"] = _lib.launch_count() - n0
        finally:
            ops.FUSED_EDGE_MLP = True
            ops.FUSED_EDGE_BWD = "auto"
    assert out["fused_launches"] < out["fused_fwd_launches"] <= out["unfused_launches"]
    floor = TOL * max(float(t.abs().max()) for t in out["unfused"][1:])
    for mode in ("fused", "fused_fwd"):
        for a, b in zip(out[mode], out["unfused"]):
            assert rel_err(a, b) < 2e-5 or float((a - b).abs().max()) <= floor, mode
    # fp64 autograd with the kernel's own dropout mask (read off the stored activation)
    edges = torch.cat([pos, neg]).cuda()
    _, a1 = ops.edge_mlp_fwd_raw(h0.cuda(), edges, params0[0].cuda(), params0[1].cuda(), params0[2].cuda(),
                                 params0[3].cuda(), drop, 5 + 7919)
    keep = (a1 > 0).double().cpu()
    hd = h0.double().requires_grad_(True)
    pd = [p.double().requires_grad_(True) for p in params0]
    e = torch.cat([pos, neg])
    a0 = hd[e[:, 0]] * hd[e[:, 1]]
    z1 = a0 @ pd[0].t() + pd[1]
    act = z1 * keep / (1.0 - drop)                        # keep already holds relu and dropout: a1 > 0
    score = (act @ pd[2].t()).reshape(-1) + pd[3]
    sp, sn = score[:B], score[B:].reshape(B, k)
    ((1 - (sp.unsqueeze(1) - sn)) ** 2).sum().backward()
    want = [hd.grad] + [p.grad for p in pd]
    for a, b in zip(out["fused"][1:], want):
        assert rel_err(a.cpu(), b) < 2e-5 or float((a.cpu().double() - b).abs().max()) <= floor


# ------------------------------------------------------------------ dense tensor-core path of the aggregation
@pytest.mark.parametrize("weighted,reduce", [(False, "mean"), (True, "sum")])
def test_dense_adjacency_path(ops, weighted, reduce):
    """small dense-ish graphs (ddi-shape) aggregate through a dense tcgen05 GEMM: forward, backward and
    the fused epilogue must match the oracle like the gather kernel does"""
    from plnlp_b200.graph import structure_of
    N, F = 300, 96
    ei, w = rand_graph(N, 14000, seed=12, weighted=weighted, hub=True)          # ~15 % dense
    o = sparse.to_sparse_tensor(ei, w, N)
    g = _to_gpu_graph(o)
    st = structure_of(g)
    assert st.dense_ok and st.density > 0.06
    x, gout, b = torch.randn(N, F), torch.randn(N, F), torch.randn(F)
    xg, bg = x.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.spmm(g, xg, reduce, bias=bg, relu=True)
    out.backward(gout.cuda())
    xc, bc = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.relu(sparse.matmul(o.set_value(None) if reduce == "mean" else o, xc, reduce) + bc)
    ref.backward(gout)
    assert rel_err(out.cpu(), ref) < TOL
    assert rel_err(xg.grad.cpu(), xc.grad) < TOL and rel_err(bg.grad.cpu(), bc.grad) < TOL
    # and it agrees with the gather kernel it replaces
    plan = st.fwd_noval if reduce == "mean" else st.fwd
    gk = ops.spmm_raw(plan, x.cuda(), use_val=reduce != "mean", div_rows=reduce == "mean")
    assert rel_err(ops.spmm(g, x.cuda(), reduce), gk) < TOL


@pytest.mark.parametrize("chunk", [4, 32, 1024])
def test_subset_plan_kernels_equal_host_construction(ops, chunk):
    """csrc/graph_build.cu plnlp_subset_plan_count / _fill build the row-subset SpMM plan of the last conv on the
    device: every array equals the torch construction on the host (index work: bit-exact), incl. split hub rows,
    empty rows, a single row, all rows"""
    from plnlp_b200.graph import build_plan, build_subset_plan
    N = 500
    ei, w = rand_graph(N, 6000, seed=31, weighted=True, hub=True)
    o = sparse.to_sparse_tensor(ei, w, N)
    rowptr, col, val = o.csr()
    parent_c = build_plan(rowptr, col, val, N, N, chunk)
    parent_g = build_plan(rowptr.cuda(), col.cuda(), val.cuda(), N, N, chunk)
    g = torch.Generator().manual_seed(chunk)
    for rows in (torch.tensor([2]), torch.arange(N), torch.unique(torch.randint(0, N, (120,), generator=g)),
                 torch.tensor([0, 2, N - 2, N - 1])):
        a = build_subset_plan(parent_c, rowptr, rows)
        b = build_subset_plan(parent_g, rowptr.cuda(), rows.cuda())
        assert (a.n_rows, a.n_items, a.n_fix, a.n_partial, a.nnz) == (b.n_rows, b.n_items, b.n_fix, b.n_partial, b.nnz)
        for name in ("item_ptr", "item_end", "item_row", "item_slot", "fix_ptr", "fix_row", "row_cnt"):
            assert torch.equal(getattr(a, name), getattr(b, name).cpu()), name
