"""CPU tests (-m "not gpu") of the host-side logic of plnlp_b200: the CSRGraph adjacency holder
against the oracle's torch_sparse restatement, the SpMM work plan (emulated in python), and the
small pure-host pieces (adjust_lr, Logger, loss-name dispatch)."""
import io

import pytest
import torch

from oracle import sparse
from plnlp_b200.graph import CSRGraph, build_plan
from tests.helpers import rand_graph, rel_err


def _same(a: CSRGraph, b: sparse.SparseTensor):
    ra, ca, va = a.csr()
    rb, cb, vb = b.csr()
    assert torch.equal(ra, rb) and torch.equal(ca, cb)        # index work: bit-exact
    assert (va is None) == (vb is None)
    if va is not None:
        assert torch.equal(va, vb)


@pytest.mark.parametrize("weighted", [False, True])
def test_csrgraph_matches_oracle_sparse(weighted):
    N = 41
    ei, w = rand_graph(N, 220, seed=8, weighted=weighted, hub=True)
    g, o = CSRGraph.from_edge_index(ei, w, N), sparse.to_sparse_tensor(ei, w, N)
    _same(g, o)
    _same(g.to_symmetric(), o.to_symmetric())
    _same(g.set_diag(), o.set_diag())
    _same(g.t(), o.t())
    assert torch.equal(g.sum(dim=1), o.sum(dim=1))
    r1, c1, _ = g.coo()
    r2, c2, _ = o.coo()
    assert torch.equal(r1, r2) and torch.equal(c1, c2)


def test_gcn_normalization_matches_oracle():
    from plnlp_b200.utils import gcn_normalization
    N = 37
    ei, _ = rand_graph(N, 150, seed=9)
    a = gcn_normalization(CSRGraph.from_edge_index(ei, None, N).to_symmetric())
    b = sparse.gcn_normalization(sparse.to_sparse_tensor(ei, None, N).to_symmetric())
    _same(a, b)


def _emulate(plan, x, use_val, div, x_index=None):
    """python emulation of csrc/spmm.cu driven by the plan arrays (explicit item ends of row-subset plans and the
    x_index of a row-sparse operand included)"""
    F = x.size(1)
    out = torch.full((plan.n_rows, F), float("nan"))
    partial = torch.zeros(max(plan.n_partial, 1), F)
    for i in range(plan.n_items):
        b = int(plan.item_ptr[i])
        e = int(plan.item_end[i]) if plan.item_end is not None else int(plan.item_ptr[i + 1])
        acc = torch.zeros(F)
        for p in range(b, e):
            v = plan.val[p] if (use_val and plan.val is not None) else 1.0
            c = int(plan.col[p])
            if x_index is not None:
                c = int(x_index[c])
                if c < 0:
                    continue
            acc = acc + v * x[c]
        s = int(plan.item_slot[i])
        if s >= 0:
            partial[s] = acc
        else:
            out[int(plan.item_row[i])] = acc
    for j in range(plan.n_fix):
        acc = torch.zeros(F)
        for s in range(int(plan.fix_ptr[j]), int(plan.fix_ptr[j + 1])):
            acc = acc + partial[s]
        out[int(plan.fix_row[j])] = acc
    if div:
        out = out / plan.row_cnt[:, None]
    return out


@pytest.mark.parametrize("chunk", [4, 32, None])
def test_spmm_plan_covers_every_row_once(chunk):
    N = 53
    ei, w = rand_graph(N, 300, seed=10, weighted=True, hub=True)
    o = sparse.to_sparse_tensor(ei, w, N)
    rowptr, col, val = o.csr()
    plan = build_plan(rowptr, col, val, N, N, chunk)
    assert int(plan.item_ptr[0]) == 0 and int(plan.item_ptr[-1]) == col.numel()
    assert torch.all(plan.item_ptr[1:] >= plan.item_ptr[:-1])
    if chunk is not None:
        assert int((plan.item_ptr[1:] - plan.item_ptr[:-1]).max()) <= chunk
        assert plan.n_fix > 0                      # the hub row is split
    x = torch.randn(N, 6)
    for reduce in ("sum", "mean"):
        got = _emulate(plan, x, use_val=(reduce == "sum"), div=(reduce == "mean"))
        adj = o if reduce == "sum" else o.set_value(None)
        assert not torch.isnan(got).any()
        assert rel_err(got, sparse.matmul(adj, x, reduce)) < 1e-5


def test_subset_plan_covers_exactly_the_selected_rows():
    """build_subset_plan: items tile the stored entries of the selected rows (in order, chunked like the parent
    plan), empty rows get one empty item, split rows get consecutive partial slots"""
    from plnlp_b200.graph import build_subset_plan
    N = 120
    ei, _ = rand_graph(N, 900, seed=12, hub=True)
    o = sparse.to_sparse_tensor(ei, None, N)
    rowptr, col, _ = o.csr()
    parent = build_plan(rowptr, col, None, N, N, chunk=32)
    rows = torch.tensor([0, 2, 5, 17, 60, N - 2, N - 1])           # row 2 is the hub, the last rows are empty
    p = build_subset_plan(parent, rowptr, rows)
    assert p.subset and p.n_rows == rows.numel() and p.n_cols == N and p.col is parent.col
    beg, end, irow, slot = (t.long() for t in (p.item_ptr, p.item_end, p.item_row, p.item_slot))
    assert p.n_items == beg.numel() == end.numel()
    for t, r in enumerate(rows.tolist()):
        sel = irow == t
        assert sel.any()
        b, e = beg[sel], end[sel]
        assert int(b[0]) == int(rowptr[r]) and int(e[-1]) == int(rowptr[r + 1])
        assert torch.equal(b[1:], e[:-1]) and bool(((e - b) <= 32).all())
        if sel.sum() > 1:
            assert bool((slot[sel] >= 0).all()) and torch.equal(slot[sel], slot[sel][0] + torch.arange(int(sel.sum())))
        else:
            assert int(slot[sel][0]) == -1
    assert p.nnz == int((rowptr[rows + 1] - rowptr[rows]).sum())
    assert p.n_fix == int(((rowptr[rows + 1] - rowptr[rows]) > 32).sum()) and p.n_partial == int((slot >= 0).sum())


def test_subset_plan_and_row_sparse_operand_emulated():
    """the two plan-level tricks behind the last conv, emulated on the CPU from the plan
    arrays exactly as csrc/spmm.cu walks them: a row-subset plan yields the selected rows of the full product,
    and an x_index with -1 entries yields the full product of the operand with those rows zeroed"""
    from plnlp_b200.graph import build_subset_plan
    N, F = 90, 5
    ei, w = rand_graph(N, 700, seed=21, weighted=True, hub=True)
    o = sparse.to_sparse_tensor(ei, w, N)
    rowptr, col, val = o.csr()
    parent = build_plan(rowptr, col, val, N, N, chunk=16)
    x = torch.randn(N, F)
    full = _emulate(parent, x, True, False)
    assert rel_err(full, sparse.matmul(o, x, "sum")) < 1e-5
    rows = torch.unique(torch.cat([torch.randint(0, N, (25,)), torch.tensor([2, N - 1])]))
    sub = build_subset_plan(parent, rowptr, rows)
    assert torch.equal(_emulate(sub, x, True, False), full[rows])            # same per-row order: same bits
    keep = torch.rand(N) < 0.3
    x_index = torch.where(keep, torch.arange(N), torch.full((N,), -1)).to(torch.int32)
    xz = x.clone()
    xz[~keep] = 0
    assert torch.equal(_emulate(parent, x, True, False, x_index=x_index), _emulate(parent, xz, True, False))
    # compact operand: x_index maps a source row to its position in a [T, F] matrix
    live = torch.nonzero(keep).reshape(-1)
    compact_index = torch.full((N,), -1, dtype=torch.int32)
    compact_index[live] = torch.arange(live.numel(), dtype=torch.int32)
    assert torch.equal(_emulate(parent, x[live], True, False, x_index=compact_index),
                       _emulate(parent, xz, True, False))


def test_adjust_lr_and_loss_dispatch():
    from plnlp_b200.model import BaseModel, adjust_lr
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    assert adjust_lr(opt, 0.25, 0.01) == pytest.approx(0.0075)
    assert opt.param_groups[0]["lr"] == pytest.approx(0.0075)
    assert adjust_lr(opt, 1.0, 0.01) == pytest.approx(1e-6)
    m = BaseModel.__new__(BaseModel)
    m.loss_func_name = "WeightedHingeAUC"
    assert m._loss_name(True) == "WeightedHingeAUC" and m._loss_name(False) == "AUC"
    m.loss_func_name = "HingeAUC"
    assert m._loss_name(False) == "HingeAUC"
    m.loss_func_name = "something-else"
    assert m._loss_name(False) == "AUC"
    m.loss_func_name = "CE"
    assert m._loss_name(False) == "CE"
    m.loss_func_name = "AdaAUC"
    assert m._loss_name(True) == "AdaAUC" and m._loss_name(False) == "AUC"


def test_logger_statistics():
    from plnlp_b200.logger import Logger
    lg = Logger(2)
    for r, seq in enumerate([[(0.1, 0.2), (0.5, 0.3), (0.5, 0.4)], [(0.7, 0.6), (0.2, 0.9)]]):
        for res in seq:
            lg.add_result(r, res)
    buf = io.StringIO()
    lg.print_statistics(0, f=buf)
    assert "Highest Eval Point: 2" in buf.getvalue() and "Final Test: 30.00" in buf.getvalue()
    buf = io.StringIO()
    lg.print_statistics(0, f=buf, last_best=True)
    assert "Highest Eval Point: 3" in buf.getvalue() and "Final Test: 40.00" in buf.getvalue()
    buf = io.StringIO()
    lg.print_statistics(f=buf)
    assert "Highest Valid: 60.00" in buf.getvalue() and "Final Test: 45.00" in buf.getvalue()


def test_every_reference_layer_and_loss_name_exists():
    import plnlp_b200.layer as L
    import plnlp_b200.loss as S
    for name in ("MLPCatPredictor", "MLPDotPredictor", "MLPBilPredictor", "BilinearPredictor", "MLPPredictor",
                 "DotPredictor", "SAGE", "GCN", "WSAGE", "Transformer"):
        assert isinstance(getattr(L, name), type)
    for name in ("weighted_auc_loss", "adaptive_auc_loss", "adaptive_hinge_auc_loss", "log_rank_loss",
                 "ce_loss", "info_nce_loss", "auc_loss", "hinge_auc_loss", "weighted_hinge_auc_loss"):
        assert callable(getattr(S, name))


def test_extra_predictor_parameter_names_match_reference(golden_dir):
    """state_dict keys and shapes of BIL / MLPDOT / MLPBIL / MLPCAT equal the real reference modules'
    (tests/golden/predictors_extra.pt), so reference checkpoints load unchanged"""
    import os
    import plnlp_b200.layer as L
    G = torch.load(os.path.join(golden_dir, "predictors_extra.pt"))
    H = 20
    made = {"bil": L.BilinearPredictor(H), "mlpdot_L2": L.MLPDotPredictor(H, H, 2, 0.0),
            "mlpbil_L2": L.MLPBilPredictor(H, H, 2, 0.0), "mlpcat_L3": L.MLPCatPredictor(H, H, 1, 3, 0.0)}
    for key, m in made.items():
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        ref = {k: tuple(v.shape) for k, v in G[key]["state"].items()}
        assert ours == ref, key


def test_sample_perm_copy_replays_reference(golden_dir):
    """negative_sample.py:61-76 is index plumbing on torch's generator: with the generator in the state the
    reference run had, the output is bit-identical (padding with random duplicates, permuted copies, and the
    reference's reshape layout)"""
    import os
    from plnlp_b200.negative_sample import sample_perm_copy
    G = torch.load(os.path.join(golden_dir, "predictors_extra.pt"))
    torch.manual_seed(16)
    e = torch.randint(0, 50, (2, 30))
    for key in ("perm_copy_30_3", "perm_copy_40_2"):
        rec = G[key]
        assert torch.equal(e, rec["edge_index"])
        out = sample_perm_copy(e, rec["target"], rec["k"])
        assert out.dtype == torch.int64 and torch.equal(out, rec["out"])


# ---------------------------------------------------------------------------------------------------------------
# python restatements of the index arithmetic of the shared-memory staged SpMM (csrc/spmm.cu,
# spmm_csr_staged_kernel) and of the launch geometry of the TMA-fed weight-gradient GEMM (csrc/gemm_tma_tn.cu):
# the GPU tests prove the kernels, these keep the reasoning behind them checkable without a GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_items", [1, 31, 32, 33, 1000, 2928161])
def test_staged_spmm_strided_item_assignment_covers_every_item_once(n_items):
    """warp w owns items w, w + W, w + 2W, ... with W = ceil(n_items / 32): every item exactly once, at most 32 per
    warp, and the consecutive chunk items of a hub row land in different warps"""
    W = -(-n_items // 32)
    seen = torch.zeros(n_items, dtype=torch.int32)
    w = torch.arange(W)
    cnt = torch.clamp(-(-(n_items - w) // W), max=32)            # ceil((n_items - w) / W), the kernel's `cnt`
    assert int(cnt.min()) >= 1 and int(cnt.max()) <= 32 and int(cnt.sum()) == n_items
    for lane in range(32):
        item = w + lane * W
        ok = lane < cnt
        assert bool((item[ok] < n_items).all()) and bool((item[~ok] >= n_items).all())
        seen[item[ok]] += 1
    assert bool((seen == 1).all())
    if n_items >= 64 and W >= 25:
        hub = torch.arange(7, 7 + 25)                            # 25 consecutive chunk items of one hub row
        assert torch.unique(hub % W).numel() == 25               # 25 different warps


@pytest.mark.parametrize("cap", [8, 16, 32])
def test_staged_spmm_batches_of_an_item(cap):
    """every item yields max(1, ceil(len / CAP)) batches, the last one flagged; an empty item yields one empty batch
    (its row still gets bias / zeros written)"""
    for length in (0, 1, cap - 1, cap, cap + 1, 3 * cap, 3 * cap + 5, 1024):
        beg, end = 100, 100 + length
        base, batches = beg, []
        while True:                                              # the kernel's iterator
            n = max(0, min(cap, end - base))
            last = base + cap >= end
            batches.append((n, last))
            base += cap
            if base >= end:
                break
        assert len(batches) == max(1, -(-length // cap))
        assert sum(n for n, _ in batches) == length
        assert [last for _, last in batches] == [False] * (len(batches) - 1) + [True]


@pytest.mark.parametrize("F,ldx", [(50, 50), (18, 18), (2, 2), (62, 62), (50, 54), (30, 38)])
def test_staged_spmm_aligned_window_arithmetic(F, ldx):
    """rows on a pitch of 4k + 2 floats, table 16-byte aligned: the copy takes the 16-byte aligned window around the row;
    the entry record says where the row's first float landed; only the window of an EVEN last row leaves the table, and
    cutting its last piece to 8 bytes keeps every copied byte inside"""
    assert ldx % 4 == 2 and F % 2 == 0 and ldx >= F
    lpr = -(-(F * 4 + 8) // 16)                                  # 16-byte pieces per row slot
    for rows in (1, 2, 7, 8):
        table_bytes = ((rows - 1) * ldx + F) * 4                 # the last row has no pitch padding behind it
        for c in range(rows):
            eoff = c * ldx                                       # element offset of the row (fits 32 bits: staged_ok)
            start = (eoff & ~3) * 4                              # window start, bytes from the 16-byte aligned base
            off = (eoff & 2) << 2                                # where the row's first float sits in its slot
            assert start % 16 == 0 and off in (0, 8) and start + off == eoff * 4
            assert off + F * 4 <= lpr * 16                       # the slot holds the whole row
            for part in range(lpr):                              # the pieces the issue loop copies
                lo = start + part * 16
                cut = part == lpr - 1 and c == rows - 1 and not (eoff & 2)
                size = 8 if cut else 16
                needed = lo < eoff * 4 + F * 4                   # piece overlaps the row's data
                if lo + size > table_bytes:
                    # a piece may only stick out of the table if the row does not need it ... and then only for
                    # the last row's tail piece, which the kernel cuts (so this must never trigger with the cut)
                    assert not needed and c == rows - 1, (F, ldx, rows, c, part)
                    assert False, "an uncut piece leaves the table"
                if cut:                                          # the cut piece still covers what the row needs of it
                    assert eoff * 4 + F * 4 <= lo + 8


def test_weight_gradient_gemm_launch_geometry():
    """gemm_tma_tn.cu: stage layout [4 or 8 atoms of A | atoms of B], ring depth chosen to fit 227 KB, units of 1088
    K-rows spread contiguously over at most 148 persistent CTAs"""
    ATOM, UNIT, SM = 16 * 128, 1088, 148
    for M, N in ((200, 179), (128, 64), (256, 256), (72, 8), (33, 250), (130, 17), (1, 8)):
        ma_live, ma = -(-M // 32), (8 if M > 128 else 4)
        n_mma = -(-N // 16) * 16
        na = -(-n_mma // 32)
        assert ma_live <= ma and n_mma <= 256 and 32 * na >= n_mma
        assert (M // 32) <= ma_live and (N // 32) <= na          # atoms fetched by the 3-D box are a prefix
        stages = min(4, (227 * 1024 - 2048) // (2 * (ma + na) * ATOM))
        assert stages >= 2
        assert stages * 2 * (ma + na) * ATOM + 1024 <= 227 * 1024 - 1024
        assert 2 * 256 <= 512 and n_mma <= 256                   # two accumulators of <= 256 TMEM columns
    for K in (16, 1088, 1089, 40000, 2927963):
        units = -(-K // UNIT)
        grid = min(units, SM)
        bounds = [units * b // grid for b in range(grid + 1)]
        assert bounds[0] == 0 and bounds[-1] == units
        sizes = [b - a for a, b in zip(bounds, bounds[1:])]
        assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1  # every CTA has work, balanced to within one unit
