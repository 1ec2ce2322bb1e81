"""CPU tests of the oracle itself (-m "not gpu").

Layer 1 (third-party restatements, parity UNPINNED by the reference) is cross-checked
against scipy.sparse, an independent index_add formulation, the plain-C in-order loop
and fp64 gradcheck.  Layer 2 (the reference's own files) is checked against the golden
fixtures produced by the real reference (tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import cspmm, ogb_eval, plnlp_ref, pyg, sparse
from tests.helpers import map_encoder_state, map_predictor_state, rand_graph, rel_err


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_spmm_vs_scipy_and_c(weighted, reduce):
    N, F = 97, 19
    ei, w = rand_graph(N, 600, seed=1, weighted=weighted, hub=True)
    adj = sparse.to_sparse_tensor(ei, w, N)
    rowptr, col, val = adj.csr()
    x = torch.randn(N, F)
    out = sparse.matmul(adj, x, reduce)
    A = sp.csr_matrix((np.ones(col.numel(), np.float64) if val is None else val.double().numpy(),
                       col.numpy(), rowptr.numpy()), shape=(N, N))
    ref = A @ x.double().numpy()
    if reduce == "mean":
        cnt = np.maximum(np.diff(rowptr.numpy()), 1)[:, None]
        ref = ref / cnt
    assert rel_err(out, torch.from_numpy(ref)) < 2e-6
    assert rel_err(sparse.matmul_indexadd(adj, x, reduce), torch.from_numpy(ref)) < 2e-6
    c32 = cspmm.spmm(rowptr, col, val, x, reduce)
    c64 = cspmm.spmm(rowptr, col, val, x, reduce, f64=True)
    assert rel_err(c64, torch.from_numpy(ref)) < 1e-12
    assert rel_err(c32, c64) < 2e-6
    # empty rows give exactly 0 (not NaN) under mean
    empty = (rowptr[1:] - rowptr[:-1]) == 0
    assert empty.any() and torch.all(out[empty] == 0) and torch.all(c32[empty] == 0)


def test_spmm_c_is_order_exact():
    """the C loop is the in-order definition: compare with a python loop bit for bit."""
    N, F = 23, 5
    ei, w = rand_graph(N, 90, seed=2, weighted=True)
    adj = sparse.to_sparse_tensor(ei, w, N)
    rowptr, col, val = adj.csr()
    x = torch.randn(N, F)
    out = cspmm.spmm(rowptr, col, val, x, "sum", threads=3)
    ref = torch.zeros(N, F)
    for m in range(N):
        for p in range(int(rowptr[m]), int(rowptr[m + 1])):
            ref[m] = ref[m] + val[p] * x[col[p]]
    assert torch.equal(out, ref)


@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_spmm_gradcheck(reduce):
    N, F = 12, 3
    ei, w = rand_graph(N, 40, seed=3, weighted=True)
    adj = sparse.to_sparse_tensor(ei, w.double(), N)
    x = torch.randn(N, F, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: sparse.matmul(adj, t, reduce), (x,))
    g1 = torch.autograd.grad(sparse.matmul(adj, x, reduce).square().sum(), x)[0]
    g2 = torch.autograd.grad(sparse.matmul_indexadd(adj, x, reduce).square().sum(), x)[0]
    assert rel_err(g1, g2) < 1e-12


def test_structure_ops_bruteforce():
    N = 17
    ei, _ = rand_graph(N, 60, seed=4)
    adj = sparse.to_sparse_tensor(ei, None, N)
    dense = torch.zeros(N, N)
    for s, d in ei.t().tolist():
        dense[d, s] += 1           # row = dst, col = src, duplicates kept
    assert torch.equal(adj.to_dense(), dense)
    rowptr, col, _ = adj.csr()
    for m in range(N):             # sorted by column inside each row
        c = col[rowptr[m]:rowptr[m + 1]]
        assert torch.all(c[1:] >= c[:-1])
    sym = adj.to_symmetric()
    assert torch.equal(sym.to_dense(), ((dense + dense.t()) > 0).float())
    sd_ = sym.set_diag()
    expect = ((dense + dense.t()) > 0).float()
    expect.fill_diagonal_(1.0)
    assert torch.equal(sd_.to_dense(), expect)
    norm = sparse.gcn_normalization(sym)
    deg = expect.sum(1)
    dis = deg.pow(-0.5)
    assert rel_err(norm.to_dense(), dis[:, None] * expect * dis[None, :]) < 1e-6
    # transposition
    assert torch.equal(adj.t().to_dense(), dense.t())


def test_negative_sampling_properties():
    import random
    random.seed(0)
    N = 40
    ei, _ = rand_graph(N, 200, seed=5)
    ei2, _ = pyg.add_self_loops(ei, num_nodes=N)
    neg = pyg.negative_sampling(ei2, num_nodes=N, num_neg_samples=300)
    ids = neg[0] * N + neg[1]
    assert ids.unique().numel() == ids.numel()
    assert not np.isin(ids.numpy(), (ei2[0] * N + ei2[1]).numpy()).any()
    assert neg.size(1) <= 300


def test_evaluator_bruteforce():
    g = torch.Generator().manual_seed(6)
    pos, neg = torch.randn(50, generator=g), torch.randn(80, generator=g)
    for K in (1, 20, 80):
        kth = sorted(neg.tolist(), reverse=True)[K - 1]
        assert ogb_eval.hits_at_k(pos, neg, K) == sum(p > kth for p in pos.tolist()) / 50
    assert ogb_eval.hits_at_k(pos, neg[:10], 20) == 1.0
    negm = torch.randn(50, 30, generator=g)
    d = ogb_eval.mrr_dict(pos, negm)
    opt, pes = ogb_eval.mrr_ranks(pos, negm)
    assert torch.equal(opt, pes)
    assert torch.allclose(d["mrr_list"], 1.0 / opt.float())


# ------------------------- layer 2: against the real reference ---------------------
def test_losses_match_reference(golden_dir):
    G = torch.load(os.path.join(golden_dir, "losses.pt"))
    for key, rec in G.items():
        k = int(key[-1])
        for name in ("AUC", "HingeAUC", "WeightedHingeAUC"):
            loss = plnlp_ref.pair_loss(name, rec["pos"], rec["neg"], k, rec["weight"])
            gp, gn = plnlp_ref.pair_loss_grad(name, rec["pos"], rec["neg"], k, rec["weight"])
            assert rel_err(loss, rec[name]["loss"]) < 1e-6
            assert rel_err(gp, rec[name]["gpos"].reshape(-1)) < 1e-6
            assert rel_err(gn.reshape(-1), rec[name]["gneg"].reshape(-1)) < 1e-6
        for name in ("WeightedAUC", "AdaAUC", "AdaHingeAUC", "LogRank", "CE", "InfoNCE"):
            loss, gp, gn = plnlp_ref.pair_loss_autograd(name, rec["pos"], rec["neg"], k, rec["weight"])
            assert rel_err(loss, rec[name]["loss"]) < 1e-6
            assert rel_err(gp, rec[name]["gpos"].reshape(-1)) < 1e-6
            assert rel_err(gn, rec[name]["gneg"].reshape(-1)) < 1e-6


def test_predictors_match_reference(golden_dir):
    G = torch.load(os.path.join(golden_dir, "predictors.pt"))
    for key, rec in G.items():
        if key == "dot":
            assert torch.equal(plnlp_ref.dot_score(rec["xi"], rec["xj"]), rec["out"])
            continue
        L = int(key.split("_L")[1])
        lins = [(rec["state"][f"lins.{i}.weight"], rec["state"][f"lins.{i}.bias"]) for i in range(L)]
        assert rel_err(plnlp_ref.mlp_score(lins, rec["xi"], rec["xj"]), rec["out"]) < 1e-6


def test_encoders_match_reference(golden_dir):
    G = torch.load(os.path.join(golden_dir, "encoders.pt"))
    gr = G["graph"]
    adj = sparse.to_sparse_tensor(gr["edge_index"], gr["edge_weight"], gr["num_nodes"])
    adj_gcn = sparse.gcn_normalization(sparse.to_sparse_tensor(gr["edge_index"], None, gr["num_nodes"]))
    assert torch.equal(adj_gcn.csr()[0], gr["gcn_rowptr"]) and torch.equal(adj_gcn.csr()[1], gr["gcn_col"])
    for key, rec in G.items():
        if key == "graph":
            continue
        kind, L = key.split("_L")
        layers = []
        for i in range(int(L)):
            pre = f"convs.{i}."
            layers.append({k[len(pre):]: v for k, v in rec["state"].items() if k.startswith(pre)})
        out = plnlp_ref.encoder_forward(kind, layers, rec["x"], adj if kind == "SAGE" else adj_gcn)
        assert rel_err(out, rec["out"]) < 1e-6, key


def test_edges_and_eval_match_reference(golden_dir):
    G = torch.load(os.path.join(golden_dir, "edges_eval.pt"))
    c = G["citation_style"]
    pos, neg = plnlp_ref.eval_edges("valid", c["split"])
    assert torch.equal(pos, c["pos"]) and torch.equal(neg, c["neg"])
    loc = G["local"]
    torch.manual_seed(70)
    mine = plnlp_ref.local_neg_sample(loc["pos"], loc["num_nodes"], 3)
    assert torch.equal(mine, loc["neg"])
    h = G["hits"]
    assert plnlp_ref.evaluate_hits(h["pv"], h["nv"], h["pt"], h["nt"]) == h["res"]
    m = G["mrr"]
    assert plnlp_ref.evaluate_mrr(m["pv"], m["nv"], m["pt"], m["nt"]) == m["res"]


@pytest.mark.parametrize("tag", ["ddi_like", "collab_like", "citation_like", "hinge_like", "sgd_like",
                                 "sgd_gcn_like"])
def test_train_replay_matches_reference(golden_dir, tag):
    R = torch.load(os.path.join(golden_dir, "train_runs.pt"))[tag]
    cfg = R["cfg"]
    m = plnlp_ref.OracleModel(num_nodes=cfg["num_nodes"], emb_hidden=cfg["emb"], gnn_hidden=cfg["hid"],
                              mlp_hidden=cfg["hid"], gnn_layers=cfg["gnn_layers"], mlp_layers=cfg["mlp_layers"],
                              encoder=cfg["encoder"], predictor=cfg["predictor"], loss=cfg["loss"], lr=cfg["lr"],
                              clip_norm=cfg["clip"], num_node_feats=cfg["feats"], use_node_feats=cfg["use_feats"],
                              optimizer=cfg.get("optimizer", "Adam"))
    st = map_encoder_state(R["init"]["encoder"])
    st.update(map_predictor_state(R["init"]["predictor"]))
    st["emb"] = R["init"]["emb"]
    m.load(st)
    adj = sparse.SparseTensor(rowptr=R["adj_rowptr"], col=R["adj_col"], value=R["adj_val"],
                              sparse_sizes=(cfg["num_nodes"], cfg["num_nodes"]), is_sorted=True)
    pos = plnlp_ref.train_pos_edges(R["split"])
    w = R["split"]["train"].get("weight")
    for ep, (neg, perms) in enumerate(zip(R["negs"], R["perms"])):
        loss = m.train_epoch(R["x"], adj, pos, neg, perms, cfg["num_neg"], w)
        assert abs(loss - R["losses"][ep]) <= 2e-5 * abs(R["losses"][ep]), (ep, loss, R["losses"][ep])
    fin = map_encoder_state(R["final"]["encoder"])
    fin.update(map_predictor_state(R["final"]["predictor"]))
    fin["emb"] = R["final"]["emb"]
    got = m.state()
    for k, v in fin.items():
        assert rel_err(got[k], v) < 5e-5, (k, rel_err(got[k], v))
    pv, nv = plnlp_ref.eval_edges("valid", R["split"])
    assert rel_err(m.predict(R["x"], adj, pv), R["scores"]["pos_valid"]) < 5e-5
    assert rel_err(m.predict(R["x"], adj, nv), R["scores"]["neg_valid"]) < 5e-5


def test_random_walk_oracle_bruteforce():
    from oracle import rw
    N = 30
    ei, _ = rand_graph(N, 120, seed=8)
    adj = sparse.to_sparse_tensor(ei, None, N)
    rowptr, col, _ = adj.csr()
    g = torch.Generator().manual_seed(3)
    start = torch.randint(0, N, (50,), generator=g)
    rand = torch.rand(50, 6, generator=g)
    walk = rw.random_walk(rowptr, col, start, 6, rand)
    for n in range(50):                      # python loop over every step
        cur = int(start[n])
        assert int(walk[n, 0]) == cur
        for l in range(6):
            b, e = int(rowptr[cur]), int(rowptr[cur + 1])
            if e > b:
                cur = int(col[b + min(int(float(rand[n, l]) * (e - b)), e - b - 1)])
            assert int(walk[n, l + 1]) == cur
    pairs, w = rw.walk_pairs(walk)
    assert (pairs[:, 0] != pairs[:, 1]).all() and pairs.size(0) == w.numel()
    assert set(w.tolist()) <= {float(torch.tensor(1.0) / (j + 1)) for j in range(6)}


def test_transformer_conv_restatement_matches_dense_masked_attention():
    """oracle.pyg.TransformerConv (PyG 2.0.1 defaults restated from memory, parity unpinned upstream) against an
    independent dense formulation: masked softmax(Q K^T / sqrt(C)) V + skip; isolated nodes get the skip only"""
    from oracle import pyg
    torch.manual_seed(9)
    N, Fi, C = 40, 7, 6
    key = torch.unique(torch.randint(0, N * N, (300,)))
    src, dst = key // N, key % N
    src, dst = src[dst < N - 3], dst[dst < N - 3]            # the last rows have no incoming entries
    ei = torch.stack([src, dst])
    adj = sparse.to_sparse_tensor(ei, None, N)
    conv = pyg.TransformerConv(Fi, C).double()
    x = torch.randn(N, Fi, dtype=torch.float64)
    got = conv(x, adj)
    mask = torch.zeros(N, N, dtype=torch.bool)
    mask[dst, src] = True                                   # adj_t[target, source]
    q, k, v = conv.lin_query(x), conv.lin_key(x), conv.lin_value(x)
    sc = (q @ k.t()) / C ** 0.5
    att = torch.softmax(sc.masked_fill(~mask, float("-inf")), 1)
    att = torch.where(mask.any(1, keepdim=True), att, torch.zeros_like(att))
    want = att @ v + conv.lin_skip(x)
    assert rel_err(got, want) < 1e-12


def test_extra_predictors_and_encoders_match_reference(golden_dir):
    """the restatements of layer.py:48-63, 90-189 (WSAGE, Transformer stacking; BIL / MLPDOT / MLPBIL / MLPCAT) against
    the outputs and gradients of the REAL reference modules (tests/golden/predictors_extra.pt)"""
    G = torch.load(os.path.join(golden_dir, "predictors_extra.pt"))

    def lins_of(state):
        n = len([k for k in state if k.startswith("lins.") and k.endswith(".weight")])
        return [(state[f"lins.{i}.weight"].clone().requires_grad_(True), state[f"lins.{i}.bias"].clone().requires_grad_(True))
                for i in range(n)]

    for key, rec in G.items():
        if key.startswith(("perm_copy", "wsage", "transformer")):
            continue
        xi, xj = rec["xi"].clone().requires_grad_(True), rec["xj"].clone().requires_grad_(True)
        st = rec["state"]
        lins = lins_of(st)
        W = st["bilin.weight"].clone().requires_grad_(True) if "bilin.weight" in st else None
        if key == "bil":
            out = plnlp_ref.bil_score(W, xi, xj)
        elif key.startswith("mlpdot"):
            out = plnlp_ref.mlpdot_score(lins, xi, xj)
        elif key.startswith("mlpbil"):
            out = plnlp_ref.mlpbil_score(lins, W, xi, xj)
        else:
            out = plnlp_ref.mlpcat_score(lins, xi, xj)
        assert out.shape == rec["out"].shape and rel_err(out.detach(), rec["out"]) < 1e-6, key
        out.backward(rec["g"])
        assert rel_err(xi.grad, rec["gxi"]) < 1e-5 and rel_err(xj.grad, rec["gxj"]) < 1e-5, key
        for i, (Wl, bl) in enumerate(lins):
            assert rel_err(Wl.grad, rec["gparams"][f"lins.{i}.weight"]) < 1e-5, key
        if W is not None:
            assert rel_err(W.grad, rec["gparams"]["bilin.weight"]) < 1e-5, key
    for kind, prefix, weighted in (("WSAGE", "wsage", True), ("TRANSFORMER", "transformer", False)):
        for L in (1, 2):
            rec = G[f"{prefix}_L{L}"]
            adj = sparse.to_sparse_tensor(rec["edge_index"], rec.get("edge_weight") if weighted else None, rec["num_nodes"])
            layers = []
            for i in range(L):
                pre = f"convs.{i}."
                layers.append({k[len(pre):]: v for k, v in rec["state"].items() if k.startswith(pre)})
            out = plnlp_ref.encoder_forward(kind, layers, rec["x"], adj)
            assert rel_err(out, rec["out"]) < 1e-6, (kind, L)
