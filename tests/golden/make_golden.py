"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Run (in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

The reference's own files are imported unmodified from /root/reference.  Its
third-party imports that cannot be installed here (torch_geometric) are satisfied
by the oracle restatements in ``oracle/pyg.py`` / ``oracle/sparse.py`` /
``oracle/ogb_eval.py``; everything in ``plnlp/*.py`` (layer stacking, predictors,
losses, samplers, edge assembly, evaluation glue, the whole ``BaseModel.train`` /
``test`` loop with clip + Adam) is the reference's code executing.  The only
instrumentation is recording the shuffles ``DataLoader`` yields and the negatives
the sampler returns, so the same step can be replayed elsewhere.

The fixtures are small ``.pt`` files (a few hundred kB in total) and are committed;
nothing on the GPU box reads /root/reference.
"""
import os
import random
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ogb_eval, pyg, sparse  # noqa: E402

# --- satisfy the reference's third-party imports with the oracle layer -------
tg = types.ModuleType("torch_geometric")
tg_nn = types.ModuleType("torch_geometric.nn")
tg_utils = types.ModuleType("torch_geometric.utils")
for name in ("SAGEConv", "GCNConv", "GraphConv", "TransformerConv"):
    setattr(tg_nn, name, getattr(pyg, name))
tg_utils.negative_sampling = pyg.negative_sampling
tg_utils.add_self_loops = pyg.add_self_loops
tg.nn, tg.utils = tg_nn, tg_utils
sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn,
                    "torch_geometric.utils": tg_utils})
for k in [k for k in sys.modules if k == "plnlp" or k.startswith("plnlp.")]:
    del sys.modules[k]
sys.path.insert(0, "/root/reference")
import plnlp.layer as ref_layer  # noqa: E402
import plnlp.loss as ref_loss  # noqa: E402
import plnlp.model as ref_model  # noqa: E402
import plnlp.negative_sample as ref_ns  # noqa: E402
import plnlp.utils as ref_utils  # noqa: E402

assert ref_model.__file__.startswith("/root/reference/"), ref_model.__file__


class Data:
    pass


def make_graph(N, E_und, seed, weighted=False, directed=False):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, N, (4 * E_und,), generator=g)
    dst = torch.randint(0, N, (4 * E_und,), generator=g)
    keep = src != dst
    lo, hi = torch.minimum(src, dst)[keep], torch.maximum(src, dst)[keep]
    key = torch.unique(lo * N + hi)
    key = key[torch.randperm(key.numel(), generator=g)[:E_und]]
    lo, hi = key // N, key % N
    if directed:
        flip = torch.rand(lo.numel(), generator=g) < 0.5
        s = torch.where(flip, hi, lo)
        d = torch.where(flip, lo, hi)
        return torch.stack([s, d]), None
    ei = torch.cat([torch.stack([lo, hi]), torch.stack([hi, lo])], 1)
    w = None
    if weighted:
        w1 = torch.randint(1, 6, (lo.numel(),), generator=g).float()
        w = torch.cat([w1, w1])
    return ei, w


def sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def golden_losses():
    g = torch.Generator().manual_seed(11)
    out = {}
    for num_neg in (1, 3):
        B = 37
        pos = torch.randn(B, 1, generator=g)
        neg = torch.randn(B * num_neg, 1, generator=g)
        w = torch.rand(B, generator=g) + 0.2
        rec = {"pos": pos, "neg": neg, "weight": w}
        for name, fn, use_w in (("AUC", ref_loss.auc_loss, False),
                                ("HingeAUC", ref_loss.hinge_auc_loss, False),
                                ("WeightedHingeAUC", ref_loss.weighted_hinge_auc_loss, True),
                                ("WeightedAUC", ref_loss.weighted_auc_loss, True),
                                ("AdaAUC", ref_loss.adaptive_auc_loss, True),
                                ("AdaHingeAUC", ref_loss.adaptive_hinge_auc_loss, True),
                                ("LogRank", ref_loss.log_rank_loss, False),
                                ("CE", ref_loss.ce_loss, None),
                                ("InfoNCE", ref_loss.info_nce_loss, False)):
            p = pos.clone().requires_grad_(True)
            n = neg.clone().requires_grad_(True)
            loss = fn(p, n) if use_w is None else (fn(p, n, num_neg, w) if use_w else fn(p, n, num_neg))
            loss.backward()
            rec[name] = {"loss": loss.detach(), "gpos": p.grad.clone(), "gneg": n.grad.clone()}
        out[f"num_neg{num_neg}"] = rec
    return out


def golden_predictors():
    torch.manual_seed(5)
    out = {}
    for H, L in ((16, 2), (24, 3), (20, 1)):
        m = ref_layer.MLPPredictor(H, H, 1, L, 0.0)
        xi, xj = torch.randn(41, H), torch.randn(41, H)
        out[f"mlp_H{H}_L{L}"] = {"state": sd(m), "xi": xi, "xj": xj, "out": m(xi, xj).detach()}
    d = ref_layer.DotPredictor()
    xi, xj = torch.randn(33, 12), torch.randn(33, 12)
    out["dot"] = {"xi": xi, "xj": xj, "out": d(xi, xj).detach()}
    return out


def golden_encoders():
    torch.manual_seed(6)
    out = {}
    N = 50
    ei, w = make_graph(N, 140, seed=21, weighted=True)
    adj = sparse.to_sparse_tensor(ei, w, N)
    adj_gcn = ref_utils.gcn_normalization(sparse.to_sparse_tensor(ei, None, N))
    rowptr, col, val = adj_gcn.csr()
    out["graph"] = {"edge_index": ei, "edge_weight": w, "num_nodes": N,
                    "gcn_rowptr": rowptr, "gcn_col": col, "gcn_val": val}
    for kind, cls, a in (("SAGE", ref_layer.SAGE, adj), ("GCN", ref_layer.GCN, adj_gcn)):
        for L in (1, 2, 3):
            fin, H = 14, 12
            m = cls(fin, H, H, L, 0.0)
            m.eval()
            x = torch.randn(N, fin)
            out[f"{kind}_L{L}"] = {"state": sd(m), "x": x, "out": m(x, a).detach()}
    return out


def golden_edges_eval():
    torch.manual_seed(7)
    out = {}
    # citation2-style split (utils.py:10-13, 36-40)
    S, K, N = 9, 5, 40
    split = {"train": {"source_node": torch.randint(0, N, (20,)), "target_node": torch.randint(0, N, (20,))},
             "valid": {"source_node": torch.randint(0, N, (S,)), "target_node": torch.randint(0, N, (S,)),
                       "target_node_neg": torch.randint(0, N, (S, K))}}
    pos, neg = ref_utils.get_pos_neg_edges("valid", split)
    out["citation_style"] = {"split": split, "pos": pos, "neg": neg}
    # local sampler structure (negative_sample.py:31-43)
    pe = torch.randint(0, N, (13, 2))
    torch.manual_seed(70)
    ns = ref_ns.local_neg_sample(pe, N, 3)
    out["local"] = {"pos": pe, "neg": ns, "num_nodes": N}
    # hits / mrr glue through the oracle evaluator
    ev = ogb_eval.Evaluator("ogbl-ddi")
    pv, nv, pt, nt = torch.randn(300), torch.randn(500), torch.randn(280), torch.randn(450)
    out["hits"] = {"pv": pv, "nv": nv, "pt": pt, "nt": nt,
                   "res": ref_utils.evaluate_hits(ev, pv, nv, pt, nt)}
    ev = ogb_eval.Evaluator("ogbl-citation2")
    pv, nv, pt, nt = torch.randn(60), torch.randn(60 * 40), torch.randn(50), torch.randn(50 * 40)
    out["mrr"] = {"pv": pv, "nv": nv, "pt": pt, "nt": nt,
                  "res": ref_utils.evaluate_mrr(ev, pv, nv, pt, nt)}
    return out


class _RecordingLoader:
    """Stands in for torch.utils.data.DataLoader inside plnlp.model only to RECORD
    the index batches the real DataLoader yields."""
    log = []

    def __init__(self, dataset, batch_size, shuffle=False):
        from torch.utils.data import DataLoader
        self._it = DataLoader(dataset, batch_size, shuffle=shuffle)
        self._shuffle = shuffle

    def __iter__(self):
        for perm in self._it:
            if self._shuffle:
                _RecordingLoader.log.append(perm.clone())
            yield perm


def golden_train(tag, *, encoder, predictor, loss, num_neg, sampler, gnn_layers, mlp_layers,
                 emb, hid, feats, use_feats, weighted, clip, directed_sym=False, epochs=2,
                 optimizer="Adam", lr=0.01):
    torch.manual_seed(100)
    random.seed(100)
    N, B = 64, 96
    ei, w = make_graph(N, 130, seed=31, weighted=weighted, directed=directed_sym)
    data = Data()
    adj = sparse.to_sparse_tensor(ei, w, N)
    row, col, _ = adj.coo()
    data.edge_index = torch.stack([col, row], 0)         # main.py:82-83
    if directed_sym:
        adj = adj.to_symmetric()                           # main.py:109-110
    if encoder == "GCN":
        adj = ref_utils.gcn_normalization(adj)             # main.py:177-179
    data.adj_t = adj
    data.x = torch.randn(N, feats) if feats else None
    if directed_sym:
        split = {"train": {"source_node": ei[0].clone(), "target_node": ei[1].clone()}}
        S = 12
        for sp in ("valid", "test"):
            split[sp] = {"source_node": torch.randint(0, N, (S,)), "target_node": torch.randint(0, N, (S,)),
                         "target_node_neg": torch.randint(0, N, (S, 7))}
    else:
        und = ei[:, : ei.size(1) // 2].t().contiguous()
        split = {"train": {"edge": und}}
        if weighted:
            split["train"]["weight"] = torch.rand(und.size(0)) + 0.1
        for sp in ("valid", "test"):
            split[sp] = {"edge": torch.randint(0, N, (40, 2)), "edge_neg": torch.randint(0, N, (120, 2))}

    model = ref_model.BaseModel(
        lr=lr, dropout=0.0, grad_clip_norm=clip, gnn_num_layers=gnn_layers, mlp_num_layers=mlp_layers,
        emb_hidden_channels=emb, gnn_hidden_channels=hid, mlp_hidden_channels=hid, num_nodes=N,
        num_node_feats=feats, gnn_encoder_name=encoder, predictor_name=predictor, loss_func=loss,
        optimizer_name=optimizer, device=torch.device("cpu"), use_node_feats=use_feats, train_node_emb=True)
    model.param_init()
    init = {"encoder": sd(model.encoder), "predictor": sd(model.predictor), "emb": model.emb.weight.detach().clone()}

    negs = []
    real_gpne = ref_utils.get_pos_neg_edges

    def recording_gpne(split_name, *a, **k):
        p, n = real_gpne(split_name, *a, **k)
        if split_name == "train":
            negs.append(n.clone())
        return p, n

    ref_model.get_pos_neg_edges = recording_gpne
    ref_model.DataLoader = _RecordingLoader
    _RecordingLoader.log = []
    losses, perms = [], []
    try:
        for _ in range(epochs):
            _RecordingLoader.log = []
            losses.append(model.train(data, split, batch_size=B, neg_sampler_name=sampler, num_neg=num_neg))
            perms.append(list(_RecordingLoader.log))
    finally:
        ref_model.get_pos_neg_edges = real_gpne
        from torch.utils.data import DataLoader
        ref_model.DataLoader = DataLoader
    final = {"encoder": sd(model.encoder), "predictor": sd(model.predictor), "emb": model.emb.weight.detach().clone()}

    metric = "mrr" if directed_sym else "hits"
    ev = ogb_eval.Evaluator("ogbl-citation2" if directed_sym else "ogbl-ddi")
    res = model.test(data, split, batch_size=B, evaluator=ev, eval_metric=metric)
    # the scores themselves, for a tolerance-based comparison (ranks can flip on near ties)
    h = model.encoder(model.create_input_feat(data), data.adj_t)
    h = torch.cat([h, h.mean(0, keepdim=True)], 0).detach()
    pv, nv = ref_utils.get_pos_neg_edges("valid", split)
    scores = {"pos_valid": model.batch_predict(h, pv, B), "neg_valid": model.batch_predict(h, nv, B), "h": h}

    rowptr, colv, val = adj.csr()
    return {"tag": tag, "cfg": dict(encoder=encoder, predictor=predictor, loss=loss, num_neg=num_neg,
                                    sampler=sampler, gnn_layers=gnn_layers, mlp_layers=mlp_layers, emb=emb,
                                    hid=hid, feats=feats, use_feats=use_feats, clip=clip, lr=lr, optimizer=optimizer,
                                    batch_size=B, num_nodes=N, metric=metric),
            "edge_index": data.edge_index, "adj_rowptr": rowptr, "adj_col": colv, "adj_val": val,
            "x": data.x, "split": split, "init": init, "final": final, "negs": negs, "perms": perms,
            "losses": losses, "test": res, "scores": scores}


def golden_predictors_extra():
    """the predictors of SURVEY 8f rank 3 (layer.py:90-189): BIL, MLPDOT, MLPBIL, MLPCAT.  Outputs in eval
    mode plus the gradients of <out, g> w.r.t. both inputs and every parameter, from the real reference."""
    torch.manual_seed(15)
    out = {}
    H, P = 20, 45

    def record(key, m):
        m.eval()
        xi = torch.randn(P, H).requires_grad_(True)
        xj = torch.randn(P, H).requires_grad_(True)
        y = m(xi, xj)
        g = torch.randn_like(y)
        y.backward(g)
        out[key] = {"state": sd(m), "xi": xi.detach().clone(), "xj": xj.detach().clone(), "out": y.detach(),
                    "g": g, "gxi": xi.grad.clone(), "gxj": xj.grad.clone(),
                    "gparams": {k: v.grad.clone() for k, v in m.named_parameters()}}

    record("bil", ref_layer.BilinearPredictor(H))
    for L in (1, 2):
        record(f"mlpdot_L{L}", ref_layer.MLPDotPredictor(H, H, L, 0.0))
        record(f"mlpbil_L{L}", ref_layer.MLPBilPredictor(H, H, L, 0.0))
    for L in (1, 2, 3):
        record(f"mlpcat_L{L}", ref_layer.MLPCatPredictor(H, H, 1, L, 0.0))
    # WSAGE (layer.py:48-54): the reference's layer stacking over the GraphConv restatement, weighted graph
    torch.manual_seed(17)
    N = 50
    ei, w = make_graph(N, 140, seed=23, weighted=True)
    adj = sparse.to_sparse_tensor(ei, w, N)
    for L in (1, 2):
        m = ref_layer.WSAGE(12, 16, 16, L, 0.0)
        m.eval()
        x = torch.randn(N, 12).requires_grad_(True)
        y = m(x, adj)
        g = torch.randn_like(y)
        y.backward(g)
        out[f"wsage_L{L}"] = {"state": sd(m), "x": x.detach().clone(), "out": y.detach(), "g": g,
                              "gx": x.grad.clone(), "gparams": {k: v.grad.clone() for k, v in m.named_parameters()},
                              "edge_index": ei, "edge_weight": w, "num_nodes": N}
    # Transformer (layer.py:57-63): the reference's stacking over the TransformerConv restatement, value-less graph
    torch.manual_seed(18)
    adj_nv = sparse.to_sparse_tensor(ei, None, N)
    for L in (1, 2):
        m = ref_layer.Transformer(12, 16, 16, L, 0.0)
        m.eval()
        x = torch.randn(N, 12).requires_grad_(True)
        y = m(x, adj_nv)
        g = torch.randn_like(y)
        y.backward(g)
        out[f"transformer_L{L}"] = {"state": sd(m), "x": x.detach().clone(), "out": y.detach(), "g": g,
                                    "gx": x.grad.clone(), "gparams": {k: v.grad.clone() for k, v in m.named_parameters()},
                                    "edge_index": ei, "num_nodes": N}
    # sample_perm_copy (negative_sample.py:61-76): shapes and the multiset property the copies keep
    torch.manual_seed(16)
    e = torch.randint(0, 50, (2, 30))
    for target, k in ((30, 3), (40, 2)):
        r = ref_ns.sample_perm_copy(e, target, k)
        out[f"perm_copy_{target}_{k}"] = {"edge_index": e, "target": target, "k": k, "out": r}
    return out


def main():
    torch.set_num_threads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "extra":       # later additions: leave the earlier fixtures untouched
        torch.save(golden_predictors_extra(), os.path.join(HERE, "predictors_extra.pt"))
        return
    torch.save(golden_losses(), os.path.join(HERE, "losses.pt"))
    torch.save(golden_predictors(), os.path.join(HERE, "predictors.pt"))
    torch.save(golden_encoders(), os.path.join(HERE, "encoders.pt"))
    torch.save(golden_edges_eval(), os.path.join(HERE, "edges_eval.pt"))
    runs = [
        golden_train("ddi_like", encoder="SAGE", predictor="MLP", loss="AUC", num_neg=3, sampler="global",
                     gnn_layers=2, mlp_layers=2, emb=16, hid=16, feats=0, use_feats=False, weighted=False,
                     clip=2.0),
        golden_train("collab_like", encoder="SAGE", predictor="DOT", loss="WeightedHingeAUC", num_neg=1,
                     sampler="global", gnn_layers=1, mlp_layers=2, emb=12, hid=12, feats=0, use_feats=False,
                     weighted=True, clip=1.0),
        golden_train("citation_like", encoder="GCN", predictor="MLP", loss="AUC", num_neg=3, sampler="local",
                     gnn_layers=2, mlp_layers=2, emb=6, hid=20, feats=9, use_feats=True, weighted=False,
                     clip=1.0, directed_sym=True),
        golden_train("hinge_like", encoder="SAGE", predictor="MLP", loss="HingeAUC", num_neg=2, sampler="local",
                     gnn_layers=2, mlp_layers=3, emb=10, hid=18, feats=0, use_feats=False, weighted=False,
                     clip=-1.0),
        # SGD (model.py:87-88): the update is linear in the gradient, so the whole trajectory is
        # well conditioned and can be compared element by element
        golden_train("sgd_like", encoder="SAGE", predictor="MLP", loss="AUC", num_neg=3, sampler="local",
                     gnn_layers=2, mlp_layers=2, emb=16, hid=16, feats=0, use_feats=False, weighted=False,
                     clip=2.0, optimizer="SGD", lr=1e-4, epochs=3),
        golden_train("sgd_gcn_like", encoder="GCN", predictor="DOT", loss="HingeAUC", num_neg=2, sampler="local",
                     gnn_layers=2, mlp_layers=2, emb=6, hid=20, feats=9, use_feats=True, weighted=False,
                     clip=1.0, directed_sym=True, optimizer="SGD", lr=1e-4, epochs=3),
    ]
    torch.save({r["tag"]: r for r in runs}, os.path.join(HERE, "train_runs.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
