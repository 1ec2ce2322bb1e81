"""GPU parity tests (-m gpu) at module / trainer level.

The golden fixtures under tests/golden/ hold outputs of the REAL reference package
(tests/golden/make_golden.py): module forwards, and whole ``BaseModel.train`` epochs with the
shuffles and negatives it used.  Here the same parameters, shuffles and negatives are replayed
through plnlp_b200 on the GPU.  fp32 bar: 1e-5 relative per tensor for single ops / one step;
multi-step trajectories (2 epochs of Adam) are checked at 2e-4 because Adam's 1/sqrt(v) amplifies
rounding differences of near-zero gradients.
"""
import os

import pytest
import torch

from oracle import plnlp_ref, sparse
from tests.helpers import fp32_close, rand_graph, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = torch.device("cuda")


class Data:
    pass


def _gpu_graph(rowptr, col, val, n):
    from plnlp_b200.graph import CSRGraph
    return CSRGraph(rowptr.cuda(), col.cuda(), None if val is None else val.cuda(), (n, n))


def _load_module(module, state):
    module.load_state_dict({k: v.cuda() for k, v in state.items()})


def test_predictors_against_reference_golden(golden_dir):
    from plnlp_b200.layer import DotPredictor, MLPPredictor
    G = torch.load(os.path.join(golden_dir, "predictors.pt"))
    for key, rec in G.items():
        if key == "dot":
            out = DotPredictor()(rec["xi"].cuda(), rec["xj"].cuda())
            assert out.shape == rec["out"].shape and rel_err(out.cpu(), rec["out"]) < TOL
            continue
        H, L = int(key.split("_")[1][1:]), int(key.split("_L")[1])
        m = MLPPredictor(H, H, 1, L, 0.0).cuda()
        _load_module(m, rec["state"])
        m.eval()
        out = m(rec["xi"].cuda(), rec["xj"].cuda())
        assert out.shape == rec["out"].shape               # [B, 1]
        assert rel_err(out.cpu(), rec["out"]) < TOL


def test_encoders_against_reference_golden(golden_dir):
    from plnlp_b200.graph import CSRGraph
    from plnlp_b200.layer import GCN, SAGE
    from plnlp_b200.utils import gcn_normalization
    G = torch.load(os.path.join(golden_dir, "encoders.pt"))
    gr = G["graph"]
    N = gr["num_nodes"]
    adj = CSRGraph.from_edge_index(gr["edge_index"].cuda(), gr["edge_weight"].cuda(), N)
    adj_gcn = gcn_normalization(CSRGraph.from_edge_index(gr["edge_index"].cuda(), None, N))
    # index work of graph preparation: bit-exact with what the reference pipeline produced
    assert torch.equal(adj_gcn.csr()[0].cpu(), gr["gcn_rowptr"]) and torch.equal(adj_gcn.csr()[1].cpu(), gr["gcn_col"])
    assert rel_err(adj_gcn.csr()[2].cpu(), gr["gcn_val"]) < 1e-6
    for key, rec in G.items():
        if key == "graph":
            continue
        kind, L = key.split("_L")
        cls = SAGE if kind == "SAGE" else GCN
        m = cls(rec["x"].size(1), 12, 12, int(L), 0.0).cuda()
        _load_module(m, rec["state"])
        m.eval()
        out = m(rec["x"].cuda(), adj if kind == "SAGE" else adj_gcn)
        assert rel_err(out.cpu(), rec["out"]) < TOL, key
        # the concat-free tuple input gives the same result
        out2 = m((rec["x"][:, :5].contiguous().cuda(), rec["x"][:, 5:].contiguous().cuda()),
                 adj if kind == "SAGE" else adj_gcn)
        assert rel_err(out2.cpu(), rec["out"]) < TOL, key


def _build_model(cfg):
    from plnlp_b200.model import BaseModel
    return BaseModel(lr=cfg["lr"], dropout=0.0, grad_clip_norm=cfg["clip"], gnn_num_layers=cfg["gnn_layers"],
                     mlp_num_layers=cfg["mlp_layers"], emb_hidden_channels=cfg["emb"],
                     gnn_hidden_channels=cfg["hid"], mlp_hidden_channels=cfg["hid"], num_nodes=cfg["num_nodes"],
                     num_node_feats=cfg["feats"], gnn_encoder_name=cfg["encoder"], predictor_name=cfg["predictor"],
                     loss_func=cfg["loss"], optimizer_name=cfg.get("optimizer", "Adam"), device=DEV, use_node_feats=cfg["use_feats"],
                     train_node_emb=True)


def _setup_run(R):
    cfg = R["cfg"]
    model = _build_model(cfg)
    _load_module(model.encoder, R["init"]["encoder"])
    _load_module(model.predictor, R["init"]["predictor"])
    with torch.no_grad():
        model.emb.weight.copy_(R["init"]["emb"].cuda())
    data = Data()
    data.adj_t = _gpu_graph(R["adj_rowptr"], R["adj_col"], R["adj_val"], cfg["num_nodes"])
    data.edge_index = R["edge_index"].cuda()
    data.x = None if R["x"] is None else R["x"].cuda()
    return cfg, model, data


@pytest.mark.parametrize("tag", ["ddi_like", "collab_like", "citation_like", "hinge_like"])
@pytest.mark.parametrize("scatter", ["sorted", "atomic", "sorted+sparse_rows"])
def test_first_step_gradients_match_reference(golden_dir, tag, scatter, monkeypatch):
    """one step from the reference's initial state with its first shuffle + negatives: loss and every
    parameter gradient against the oracle (itself pinned to the reference by tests/test_oracle.py).

    ``+sparse_rows``: the model is told its node set is huge, so the batch counts as touching a small part of it
    and the CSR kernels run the row-restricted last conv (GCN) or the full conv with a row-sparse backward
    (SAGE: ddi_like is the 2-layer SAGE whose FIRST layer then receives a dense gradient through the root weight
    plus A^T g -- the case a pointer-keyed sparsity hint got wrong in round 1)."""
    from plnlp_b200 import _ops
    R = torch.load(os.path.join(golden_dir, "train_runs.pt"))[tag]
    cfg, model, data = _setup_run(R)
    if scatter.endswith("+sparse_rows"):
        from plnlp_b200 import graph
        monkeypatch.setattr(graph, "DENSE_SPMM", False)
        model.num_nodes = 10 ** 9
        scatter = "sorted"
    _ops.SCATTER_MODE = scatter
    try:
        pos = plnlp_ref.train_pos_edges(R["split"])
        perm, neg = R["perms"][0][0], R["negs"][0]
        w = R["split"]["train"].get("weight")
        model.encoder.train(); model.predictor.train()
        model.clip_norm = -1.0                                    # look at raw gradients
        loss = model.train_batch(data, pos[perm].cuda(), neg[perm].reshape(-1, 2).cuda(), cfg["num_neg"],
                                 None if w is None else w[perm].cuda())
    finally:
        _ops.SCATTER_MODE = "sorted"
    st = {"enc." + k[len("convs."):]: v for k, v in R["init"]["encoder"].items()}
    st.update({"pred." + k[len("lins."):]: v for k, v in R["init"]["predictor"].items()})
    st["emb"] = R["init"]["emb"]
    refs = {}
    for dt in (torch.float32, torch.float64):
        ref = plnlp_ref.OracleModel(num_nodes=cfg["num_nodes"], emb_hidden=cfg["emb"], gnn_hidden=cfg["hid"],
                                    mlp_hidden=cfg["hid"], gnn_layers=cfg["gnn_layers"],
                                    mlp_layers=cfg["mlp_layers"], encoder=cfg["encoder"],
                                    predictor=cfg["predictor"], loss=cfg["loss"], lr=cfg["lr"], clip_norm=-1.0,
                                    num_node_feats=cfg["feats"], use_node_feats=cfg["use_feats"], dtype=dt)
        ref.load({k: v.to(dt) for k, v in st.items()})
        val = R["adj_val"]
        adj = sparse.SparseTensor(rowptr=R["adj_rowptr"], col=R["adj_col"], value=None if val is None else val.to(dt),
                                  sparse_sizes=(cfg["num_nodes"],) * 2, is_sorted=True)
        rloss, _ = ref.step(None if R["x"] is None else R["x"].to(dt), adj, pos[perm], neg[perm], cfg["num_neg"],
                            None if w is None else w[perm].to(dt), do_update=False)
        refs[dt] = (rloss, ref)
    (l32, r32), (l64, r64) = refs[torch.float32], refs[torch.float64]
    assert rel_err(loss.cpu(), l64) < TOL

    floor = TOL * max(float(v.grad.abs().max()) for v in r64.params.values())

    def check(name, got, key):
        ok, msg = fp32_close(got, r32.params[key].grad, r64.params[key].grad, TOL, floor=floor)
        assert ok, f"{name}: {msg}"

    check("emb", model.emb.weight.grad, "emb")
    for name, p in model.encoder.named_parameters():
        check(name, p.grad, "enc." + name[len("convs."):])
    for name, p in model.predictor.named_parameters():
        check(name, p.grad, "pred." + name[len("lins."):])


@pytest.mark.parametrize("mode", ["buffer", "reassoc_only", "reference_order", "restricted_last_layer",
                                  "row_sparse_grad_only"])
def test_gcn_feature_plus_embedding_layer_orders(golden_dir, mode, monkeypatch):
    """GCNConv on [emb | x] (citation2 recipe): the three evaluation orders -- aggregate-first into one buffer
    with the constant-feature aggregate cached (default on sparse graphs), aggregate-first per block, and the
    reference's linear-first order -- all match the oracle's first-step loss and gradients."""
    from plnlp_b200 import graph, layer
    monkeypatch.setattr(graph, "DENSE_SPMM", False)          # the golden graph is tiny: force the CSR kernels
    if mode == "reassoc_only":
        monkeypatch.setattr(layer, "_agg_buffer_ok", lambda adj, parts: False)
    if mode == "reference_order":
        monkeypatch.setattr(layer, "REASSOCIATE", False)
    R = torch.load(os.path.join(golden_dir, "train_runs.pt"))["citation_like"]
    cfg, model, data = _setup_run(R)
    pos = plnlp_ref.train_pos_edges(R["split"])
    st = {"enc." + k[len("convs."):]: v for k, v in R["init"]["encoder"].items()}
    st.update({"pred." + k[len("lins."):]: v for k, v in R["init"]["predictor"].items()})
    st["emb"] = R["init"]["emb"]
    ref = plnlp_ref.OracleModel(num_nodes=cfg["num_nodes"], emb_hidden=cfg["emb"], gnn_hidden=cfg["hid"],
                                mlp_hidden=cfg["hid"], gnn_layers=cfg["gnn_layers"], mlp_layers=cfg["mlp_layers"],
                                encoder=cfg["encoder"], predictor=cfg["predictor"], loss=cfg["loss"], lr=cfg["lr"],
                                clip_norm=-1.0, num_node_feats=cfg["feats"], use_node_feats=cfg["use_feats"],
                                dtype=torch.float64)
    ref.load({k: v.double() for k, v in st.items()})
    adj = sparse.SparseTensor(rowptr=R["adj_rowptr"], col=R["adj_col"], value=R["adj_val"].double(),
                              sparse_sizes=(cfg["num_nodes"],) * 2, is_sorted=True)
    if mode in ("restricted_last_layer", "row_sparse_grad_only"):
        # what train_batch does on citation2-shape, where a batch touches ~10 % of the nodes: the last conv
        # computes only the endpoint rows (or, where it cannot, its backward skips the all-zero gradient rows)
        model.num_nodes = 10 ** 9
        if mode == "row_sparse_grad_only":
            from plnlp_b200.layer import GCNConv
            monkeypatch.setattr(GCNConv, "can_restrict", lambda self, x, adj_t: False)
    model.encoder.train(); model.predictor.train()
    model.clip_norm = -1.0
    model.optimizer.param_groups[0]["lr"] = 0.0           # two batches from the same parameters
    for b in range(min(2, len(R["perms"][0]))):           # the second pass reuses the cached aggregate / buffer
        perm, neg = R["perms"][0][b], R["negs"][0]
        loss = model.train_batch(data, pos[perm].cuda(), neg[perm].reshape(-1, 2).cuda(), cfg["num_neg"], None)
        rloss, _ = ref.step(R["x"].double(), adj, pos[perm], neg[perm], cfg["num_neg"], None, do_update=False)
        assert rel_err(loss.cpu(), rloss) < TOL
        assert rel_err(model.emb.weight.grad.cpu(), ref.params["emb"].grad) < 2 * TOL
        for name, p in model.encoder.named_parameters():
            assert rel_err(p.grad.cpu(), ref.params["enc." + name[len("convs."):]].grad) < 2 * TOL, name
    used_buffer = "_plnlp_agg_buffer" in data.adj_t.__dict__
    assert used_buffer == (mode in ("buffer", "restricted_last_layer", "row_sparse_grad_only"))


def test_prepared_batches_equal_unprepared(golden_dir, monkeypatch):
    """BaseModel.run_batches prepares the index work of batch i + 1 (endpoint ids, renumbered edges, row-subset plan,
    backward index) on a side stream while batch i runs: same bits as train_batch doing it inline, step after step"""
    from plnlp_b200 import graph, model as M
    monkeypatch.setattr(graph, "DENSE_SPMM", False)
    R = torch.load(os.path.join(golden_dir, "train_runs.pt"))["citation_like"]
    pos = plnlp_ref.train_pos_edges(R["split"])
    outs = []
    for ahead in (False, True):
        monkeypatch.setattr(M, "PREPARE_AHEAD", ahead)
        cfg, model, data = _setup_run(R)
        model.num_nodes = 10 ** 9                                  # every batch counts as row-sparse
        model.encoder.train(); model.predictor.train()
        batches = [(pos[p].cuda(), R["negs"][0][p].reshape(-1, 2).cuda(), None) for p in R["perms"][0][:3]]
        calls = []
        orig = model.prepare_batch
        model.prepare_batch = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        tot, n_ex, n_b = model.run_batches(data, iter(batches), cfg["num_neg"])
        assert n_b == len(batches) and len(calls) == (len(batches) if ahead else 0)
        outs.append((float(tot), [p.detach().clone() for p in model.para_list]))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, b)


def test_gcn_aggregate_buffer_survives_interleaved_forwards(monkeypatch):
    """the aggregate buffer is shared by every forward on one adjacency: a backward that runs after a LATER
    forward must first restore its own live block (AggLinear's stamp check)"""
    from plnlp_b200 import graph
    from plnlp_b200.layer import GCNConv, mark_constant
    monkeypatch.setattr(graph, "DENSE_SPMM", False)
    N, Fe, Ff, H = 150, 6, 10, 24
    ei, _ = rand_graph(N, 900, seed=11)
    o = sparse.gcn_normalization(sparse.to_sparse_tensor(ei, None, N).to_symmetric())
    rowptr, col, val = o.csr()
    g = _gpu_graph(rowptr, col, val, N)
    torch.manual_seed(3)
    conv = GCNConv(Fe + Ff, H).cuda()
    feat = mark_constant(torch.randn(N, Ff).cuda())
    e1 = torch.randn(N, Fe).cuda().requires_grad_(True)
    e2 = torch.randn(N, Fe).cuda().requires_grad_(True)
    y1 = conv((e1, feat), g)
    y2 = conv((e2, feat), g)                         # overwrites the live columns of the shared buffer
    gout = torch.randn(N, H).cuda()
    y1.backward(gout)
    A = o.to_dense().double()
    W, bias = conv.lin.weight.detach().double().cpu(), conv.bias.detach().double().cpu()
    e1c = e1.detach().double().cpu().requires_grad_(True)
    Wc = W.clone().requires_grad_(True)
    yref = A @ (torch.cat([e1c, feat.double().cpu()], 1) @ Wc.t()) + bias
    yref.backward(gout.double().cpu())
    assert rel_err(y1.detach().cpu(), yref.detach()) < TOL
    assert rel_err(e1.grad.cpu(), e1c.grad) < TOL
    assert rel_err(conv.lin.weight.grad.cpu(), Wc.grad) < TOL
    assert e2.grad is None and y2.shape == y1.shape


def test_wsage_against_reference_golden(golden_dir):
    """WSAGE (layer.py:48-54; PyG GraphConv = weighted-sum SAGE): the reference's stacking over the restated
    conv, forward and every gradient, 1 and 2 layers; sparse and dense aggregation paths"""
    from plnlp_b200 import graph
    from plnlp_b200.layer import WSAGE
    G = torch.load(os.path.join(golden_dir, "predictors_extra.pt"))
    for dense in (True, False):
        graph.DENSE_SPMM = dense
        try:
            for L in (1, 2):
                rec = G[f"wsage_L{L}"]
                N = rec["num_nodes"]
                o = sparse.to_sparse_tensor(rec["edge_index"], rec["edge_weight"], N)
                rowptr, col, val = o.csr()
                g = _gpu_graph(rowptr, col, val, N)
                m = WSAGE(12, 16, 16, L, 0.0).cuda()
                _load_module(m, rec["state"])
                m.eval()
                x = rec["x"].cuda().requires_grad_(True)
                y = m(x, g)
                assert rel_err(y.detach().cpu(), rec["out"]) < TOL
                y.backward(rec["g"].cuda())
                assert rel_err(x.grad.cpu(), rec["gx"]) < TOL
                for name, p in m.named_parameters():
                    assert rel_err(p.grad.cpu(), rec["gparams"][name]) < 2 * TOL, name
        finally:
            graph.DENSE_SPMM = True


def test_transformer_against_reference_golden(golden_dir):
    """Transformer (layer.py:57-63; PyG TransformerConv = per-destination softmax attention): the reference's
    stacking over the restated conv, forward and every gradient, 1 and 2 layers.

    lin_key.bias shifts every score of a destination row by the same <q_i, b_k>, which the softmax ignores: its
    exact gradient is 0 and the reference's own fp32 value is rounding noise (~1e-7 next to gradients of ~10), so
    parameter gradients are compared with an absolute floor of 2e-5 x the largest gradient tensor, like the
    predictor biases under the AUC loss elsewhere in this file."""
    from plnlp_b200.layer import Transformer
    G = torch.load(os.path.join(golden_dir, "predictors_extra.pt"))
    for L in (1, 2):
        rec = G[f"transformer_L{L}"]
        N = rec["num_nodes"]
        o = sparse.to_sparse_tensor(rec["edge_index"], None, N)
        rowptr, col, val = o.csr()
        g = _gpu_graph(rowptr, col, val, N)
        m = Transformer(12, 16, 16, L, 0.0).cuda()
        assert sorted(k for k, _ in m.named_parameters()) == sorted(rec["state"])
        _load_module(m, rec["state"])
        m.eval()
        x = rec["x"].cuda().requires_grad_(True)
        y = m(x, g)
        assert rel_err(y.detach().cpu(), rec["out"]) < TOL
        y.backward(rec["g"].cuda())
        assert rel_err(x.grad.cpu(), rec["gx"]) < 2 * TOL
        floor = 2 * TOL * max(float(v.abs().max()) for v in rec["gparams"].values())
        for name, p in m.named_parameters():
            want = rec["gparams"][name]
            err = float((p.grad.cpu() - want).abs().max())
            assert err <= 2 * TOL * float(want.abs().max()) or err <= floor, (name, err)
    # training mode: relu + dropout fused into the skip GEMM's epilogue, deterministic given the seed stream
    m.train(); m.dropout = 0.5
    torch.manual_seed(1)
    a = m(x.detach(), g)
    torch.manual_seed(1)
    b = m(x.detach(), g)
    assert torch.equal(a, b) and a.shape == y.shape


def test_extra_predictors_against_reference_golden(golden_dir):
    """BIL / MLPDOT / MLPBIL / MLPCAT (layer.py:90-189): forward and every gradient against
    outputs of the real reference modules; the edge-level entry ``score_edges`` (node-level transform where no
    dropout is active) must agree with the pair-level ``forward``"""
    from plnlp_b200 import layer
    G = torch.load(os.path.join(golden_dir, "predictors_extra.pt"))
    H = 20
    for key, rec in G.items():
        if key.startswith(("perm_copy", "wsage", "transformer")):
            continue
        L = int(key.split("_L")[1]) if "_L" in key else 0
        if key == "bil":
            m = layer.BilinearPredictor(H)
        elif key.startswith("mlpdot"):
            m = layer.MLPDotPredictor(H, H, L, 0.0)
        elif key.startswith("mlpbil"):
            m = layer.MLPBilPredictor(H, H, L, 0.0)
        else:
            m = layer.MLPCatPredictor(H, H, 1, L, 0.0)
        m = m.cuda()
        assert sorted(k for k, _ in m.named_parameters()) == sorted(rec["state"])
        _load_module(m, rec["state"])
        m.eval()
        xi = rec["xi"].cuda().requires_grad_(True)
        xj = rec["xj"].cuda().requires_grad_(True)
        out = m(xi, xj)
        assert out.shape == rec["out"].shape, key
        assert rel_err(out.detach().cpu(), rec["out"]) < TOL, key
        out.backward(rec["g"].cuda())
        assert rel_err(xi.grad.cpu(), rec["gxi"]) < TOL and rel_err(xj.grad.cpu(), rec["gxj"]) < TOL, key
        for name, p in m.named_parameters():
            assert rel_err(p.grad.cpu(), rec["gparams"][name]) < 2 * TOL, (key, name)
        # edge-level entry on a node table: rows of xi and xj stacked, pair p = (p, P + p)
        P = xi.size(0)
        h = torch.cat([rec["xi"], rec["xj"]], 0).cuda().requires_grad_(True)
        ar = torch.arange(P, device=DEV)
        edges = torch.stack([ar, ar + P], 1)
        m.zero_grad()
        s2 = m.score_edges(h, edges)
        assert rel_err(s2.detach().reshape(-1).cpu(), rec["out"].reshape(-1)) < TOL, key
        s2.backward(rec["g"].cuda().reshape(s2.shape))
        assert rel_err(h.grad[:P].cpu(), rec["gxi"]) < TOL and rel_err(h.grad[P:].cpu(), rec["gxj"]) < TOL, key
        for name, p in m.named_parameters():
            assert rel_err(p.grad.cpu(), rec["gparams"][name]) < 2 * TOL, (key, name)
        # dropout active: pair-level path with independent masks per side still runs and keeps the shape
        if hasattr(m, "dropout"):
            m.dropout = 0.5
            m.train()
            assert m.score_edges(h.detach(), edges).shape == s2.shape


@pytest.mark.parametrize("predictor", ["BIL", "MLPDOT", "MLPBIL", "MLPCAT"])
def test_train_step_with_extra_predictors(predictor):
    """BaseModel.train_batch through the generic scoring path: loss and gradients against the same modules
    evaluated pair-level in fp64 torch (CPU) from the model's own parameters"""
    from plnlp_b200.model import BaseModel
    torch.manual_seed(4)
    N, H, k, B = 120, 16, 2, 64
    ei, _ = rand_graph(N, 700, seed=8)
    o = sparse.to_sparse_tensor(torch.cat([ei, ei.flip(0)], 1), None, N)
    rowptr, col, val = o.csr()
    data = Data()
    data.adj_t, data.x, data.edge_index = _gpu_graph(rowptr, col, val, N), None, ei.cuda()
    model = BaseModel(lr=0.0, dropout=0.0, grad_clip_norm=-1.0, gnn_num_layers=1, mlp_num_layers=2,
                      emb_hidden_channels=H, gnn_hidden_channels=H, mlp_hidden_channels=H, num_nodes=N,
                      num_node_feats=0, gnn_encoder_name="SAGE", predictor_name=predictor, loss_func="AUC",
                      optimizer_name="SGD", device=DEV, use_node_feats=False, train_node_emb=True)
    model.param_init()
    model.encoder.train(); model.predictor.train()
    pos = ei.t()[:B].contiguous()
    neg = torch.randint(0, N, (B * k, 2))
    loss = model.train_batch(data, pos.cuda(), neg.cuda(), k)
    # fp64 reference: same encoder through the oracle's SAGE restatement, predictor pair-level in torch
    import copy
    pred = copy.deepcopy(model.predictor).cpu().double()
    emb = model.emb.weight.detach().cpu().double().requires_grad_(True)
    conv = model.encoder.convs[0]
    Wl, bl, Wr = (t.detach().cpu().double() for t in (conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight))
    hh = torch.relu(sparse.matmul(o, emb, "mean") @ Wl.t() + bl + emb @ Wr.t())
    edges = torch.cat([pos, neg], 0)
    xi, xj = hh[edges[:, 0]], hh[edges[:, 1]]

    def lin(mod, x, relu=False):
        y = x @ mod.weight.t() + (mod.bias if mod.bias is not None else 0)
        return torch.relu(y) if relu else y
    if predictor == "BIL":
        s = (lin(pred.bilin, xi) * xj).sum(-1)
    elif predictor in ("MLPDOT", "MLPBIL"):
        for l in pred.lins:
            xi, xj = lin(l, xi, True), lin(l, xj, True)
        s = ((lin(pred.bilin, xi) if predictor == "MLPBIL" else xi) * xj).sum(-1)
    else:
        x1, x2 = torch.cat([xi, xj], -1), torch.cat([xj, xi], -1)
        for l in pred.lins[:-1]:
            x1, x2 = lin(l, x1, True), lin(l, x2, True)
        s = ((lin(pred.lins[-1], x1) + lin(pred.lins[-1], x2)) / 2).reshape(-1)
    rloss = plnlp_ref.pair_loss("AUC", s[:B], s[B:], k)
    rloss.backward()
    assert rel_err(loss.cpu(), rloss.detach()) < TOL
    assert rel_err(model.emb.weight.grad.cpu(), emb.grad) < 2 * TOL
    for (name, p), (_, q) in zip(model.predictor.named_parameters(), pred.named_parameters()):
        ok, msg = fp32_close(p.grad.cpu(), q.grad.float(), q.grad, 2 * TOL,
                             floor=2 * TOL * float(emb.grad.abs().max()))
        assert ok, (name, msg)


def _check_update(name, got, init, final, lr, steps, adam):
    """compare the parameter UPDATE (final - init) of a whole replayed trajectory"""
    got, init, final = got.detach().double().cpu(), init.double(), final.double()
    d_got, d_ref = got - init, final - init
    scale = max(float(d_ref.abs().max()), 1e-30)
    err = (d_got - d_ref).abs()
    if not adam:
        # 1e-7 absolute floor: parameters whose exact update is 0 (bias under sum(dscore) == 0)
        assert float(err.max()) <= 2e-4 * scale + 1e-7, \
            f"{name}: update err {float(err.max()):.3e} vs scale {scale:.3e}"
        return
    # Adam divides by sqrt(v): an element whose gradient is pure rounding noise moves by +-lr per step
    # in an arbitrary direction IN THE REFERENCE TOO.  The predictor biases are exactly that under an
    # AUC-family loss (sum(d loss/d score) == 0, so d loss/d bias is a sum that cancels to rounding
    # noise): for them only the largest drift Adam can produce is bounded.  Every other tensor must
    # follow the reference in bulk as well.
    assert float(err.max()) <= 2.0 * lr * steps + 1e-6, f"{name}: drift {float(err.max()):.3e}"
    if name.startswith("lins.") and name.endswith(".bias"):
        return
    frac_ok = float((err <= 2e-3 * scale).double().mean())
    assert frac_ok >= 0.8, f"{name}: only {frac_ok:.2f} of the elements follow the reference update"


@pytest.mark.parametrize("tag", ["ddi_like", "collab_like", "citation_like", "hinge_like", "sgd_like",
                                 "sgd_gcn_like"])
def test_train_epochs_replay_reference(golden_dir, tag):
    """BaseModel.train for the reference's epochs (its shuffles, its negatives): reported loss,
    parameter updates, validation scores and metrics.  The SGD runs are compared element by element;
    the Adam runs robustly (see _check_update)."""
    R = torch.load(os.path.join(golden_dir, "train_runs.pt"))[tag]
    cfg, model, data = _setup_run(R)
    adam = cfg.get("optimizer", "Adam") == "Adam"
    ltol = 2e-4 if adam else 2e-5
    steps = 0
    for ep in range(len(R["losses"])):
        loss = model.train(data, R["split"], batch_size=cfg["batch_size"], neg_sampler_name=cfg["sampler"],
                           num_neg=cfg["num_neg"], perms=R["perms"][ep], neg_edges=R["negs"][ep])
        steps += len(R["perms"][ep])
        assert abs(loss - R["losses"][ep]) <= ltol * abs(R["losses"][ep]), (ep, loss, R["losses"][ep])
    for name, p in model.encoder.named_parameters():
        _check_update(name, p, R["init"]["encoder"][name], R["final"]["encoder"][name], cfg["lr"], steps, adam)
    for name, p in model.predictor.named_parameters():
        _check_update(name, p, R["init"]["predictor"][name], R["final"]["predictor"][name], cfg["lr"], steps, adam)
    _check_update("emb", model.emb.weight, R["init"]["emb"], R["final"]["emb"], cfg["lr"], steps, adam)
    stol = 5e-3 if adam else 2e-4
    # scoring path (model.py:175-194)
    model.encoder.eval(); model.predictor.eval()
    h = model.encode_for_test(data)
    assert rel_err(h.cpu(), R["scores"]["h"]) < stol
    from plnlp_b200.utils import get_pos_neg_edges
    pv, nv = get_pos_neg_edges("valid", R["split"], device=DEV)
    pos_s = model.batch_predict(h, pv, cfg["batch_size"]).cpu()
    neg_s = model.batch_predict(h, nv, cfg["batch_size"]).cpu()
    assert pos_s.shape == R["scores"]["pos_valid"].shape and neg_s.shape == R["scores"]["neg_valid"].shape
    res = model.test(data, R["split"], batch_size=cfg["batch_size"], evaluator=None, eval_metric=cfg["metric"])
    assert set(res) == set(R["test"])
    # after an Adam trajectory with an MLP head the scores ride on the noise-driven predictor biases
    # (see _check_update), so score VALUES are only compared for the well-conditioned runs
    if not adam or cfg["predictor"] == "DOT":
        assert rel_err(pos_s, R["scores"]["pos_valid"]) < stol
        assert rel_err(neg_s, R["scores"]["neg_valid"]) < stol
        for k in res:                 # ranks can move by one when scores differ in the last bits
            for a, b in zip(res[k], R["test"][k]):
                assert abs(a - b) <= 0.03, (k, res[k], R["test"][k])


def test_train_end_to_end_with_gpu_samplers():
    """the public call a user makes: BaseModel.train with the GPU samplers and shuffle; loss must
    decrease over a few epochs; dropout > 0 exercised"""
    from plnlp_b200.graph import CSRGraph
    from plnlp_b200.model import BaseModel
    torch.manual_seed(0)
    N = 500
    ei, _ = rand_graph(N, 4000, seed=21)
    ei = ei[:, ei[0] != ei[1]]
    und = torch.cat([ei, ei.flip(0)], 1)
    data = Data()
    data.adj_t = CSRGraph.from_edge_index(und.cuda(), None, N)
    row, col, _ = data.adj_t.coo()
    data.edge_index = torch.stack([col, row], 0)
    data.x = None
    split = {"train": {"edge": ei.t().contiguous()},
             "valid": {"edge": torch.randint(0, N, (200, 2)), "edge_neg": torch.randint(0, N, (300, 2))},
             "test": {"edge": torch.randint(0, N, (200, 2)), "edge_neg": torch.randint(0, N, (300, 2))}}
    for sampler, pred, loss in (("global", "MLP", "AUC"), ("local", "DOT", "HingeAUC")):
        m = BaseModel(lr=0.005, dropout=0.3, grad_clip_norm=2.0, gnn_num_layers=2, mlp_num_layers=2,
                      emb_hidden_channels=64, gnn_hidden_channels=64, mlp_hidden_channels=64, num_nodes=N,
                      num_node_feats=0, gnn_encoder_name="SAGE", predictor_name=pred, loss_func=loss,
                      optimizer_name="Adam", device=DEV, use_node_feats=False, train_node_emb=True)
        m.param_init()
        losses = [m.train(data, split, batch_size=1024, neg_sampler_name=sampler, num_neg=3) for _ in range(6)]
        assert all(l == l for l in losses) and losses[-1] < losses[0], losses
        res = m.test(data, split, batch_size=1024, evaluator=None, eval_metric="hits")
        assert set(res) == {"Hits@20", "Hits@50", "Hits@100"}
        assert all(0.0 <= v <= 1.0 for pair in res.values() for v in pair)


def test_full_size_ddi_shape_properties():
    """BASELINE config 2 shape (N=4267, nnz~2.1M, F=512): size-independent properties instead of a
    slow CPU comparison -- linearity of the SpMM, adjointness <A x, y> = <x, A^T y> of forward vs
    backward kernels, mean-aggregation of a constant is the constant, determinism."""
    from plnlp_b200 import _ops
    from plnlp_b200.graph import CSRGraph
    g = torch.Generator().manual_seed(5)
    N, E, F = 4267, 1067911, 512
    lo = torch.randint(0, N, (int(E * 1.3),), generator=g)
    hi = torch.randint(0, N, (int(E * 1.3),), generator=g)
    key = torch.unique(torch.minimum(lo, hi) * N + torch.maximum(lo, hi))
    key = key[(key // N) != (key % N)][:E]
    ei = torch.stack([key // N, key % N])
    adj = CSRGraph.from_edge_index(torch.cat([ei, ei.flip(0)], 1).cuda(), None, N)
    x, y = torch.randn(N, F, generator=g).cuda(), torch.randn(N, F, generator=g).cuda()
    ax, ay = _ops.spmm(adj, x, "mean"), _ops.spmm(adj, y, "mean")
    assert rel_err(_ops.spmm(adj, 2.0 * x - 3.0 * y, "mean"), 2.0 * ax - 3.0 * ay) < TOL
    assert torch.equal(ax, _ops.spmm(adj, x, "mean"))
    ones = torch.ones(N, F, device=DEV)
    deg = adj.sum(dim=1)
    # (this graph is 11.7 % dense, so the aggregation runs on the dense 3xTF32 tensor-core path)
    assert rel_err(_ops.spmm(adj, ones, "mean")[deg > 0], ones[deg > 0]) < TOL
    xr = x.clone().requires_grad_(True)
    _ops.spmm(adj, xr, "mean").backward(y)
    lhs = (ax.double() * y.double()).sum()
    rhs = (x.double() * xr.grad.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < TOL
    # one sampled row against the in-order definition
    rowptr, col, _ = adj.csr()
    r = 1234
    nb = col[rowptr[r]:rowptr[r + 1]]
    want = x[nb].double().sum(0) / max(nb.numel(), 1)
    assert rel_err(ax[r], want) < TOL
