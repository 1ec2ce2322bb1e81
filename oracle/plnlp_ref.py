"""Oracle restatement of the reference's own hot-path files (functional style).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pinned by the fixtures under
``tests/golden/`` which were produced by running the REAL reference package
(``/root/reference/plnlp``) here -- see ``tests/golden/make_golden.py``.

Each function cites the reference lines it follows.  Everything is plain torch
on CPU (fp32 by default, fp64 when the inputs are fp64).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import pyg, sparse
from .ogb_eval import hits_at_k, mrr_dict

# ---------------------------------------------------------------------------
# encoders  (plnlp/layer.py:7-45)
# ---------------------------------------------------------------------------


def conv_forward(kind, p, x, adj_t):
    """One conv.  ``p`` is a dict of tensors with PyG's parameter names.
    SAGE: lin_l.weight, lin_l.bias, lin_r.weight; GCN: lin.weight, bias."""
    if kind == "SAGE":
        agg = sparse.matmul(adj_t.set_value(None), x, reduce="mean")
        return F.linear(agg, p["lin_l.weight"], p["lin_l.bias"]) + F.linear(x, p["lin_r.weight"])
    if kind == "GCN":
        z = F.linear(x, p["lin.weight"])
        return sparse.matmul(adj_t, z, reduce="sum") + p["bias"]
    if kind == "WSAGE":          # layer.py:48-54: PyG GraphConv, weighted-sum aggregation
        agg = sparse.matmul(adj_t, x, reduce="sum")
        return F.linear(agg, p["lin_rel.weight"], p["lin_rel.bias"]) + F.linear(x, p["lin_root.weight"])
    if kind == "TRANSFORMER":    # layer.py:57-63: PyG TransformerConv with its defaults (see oracle/pyg.py)
        row, col, _ = adj_t.coo()
        q = F.linear(x, p["lin_query.weight"], p["lin_query.bias"])
        k = F.linear(x, p["lin_key.weight"], p["lin_key.bias"])
        v = F.linear(x, p["lin_value.weight"], p["lin_value.bias"])
        score = (q[row] * k[col]).sum(-1) / math.sqrt(q.size(1))
        n = x.size(0)
        mx = torch.full((n,), float("-inf"), dtype=score.dtype).scatter_reduce(0, row, score, "amax", include_self=True)
        e = torch.exp(score - mx[row])
        alpha = e / torch.zeros(n, dtype=score.dtype).index_add_(0, row, e)[row]
        agg = torch.zeros(n, q.size(1), dtype=x.dtype).index_add_(0, row, alpha.unsqueeze(-1) * v[col])
        return agg + F.linear(x, p["lin_skip.weight"], p["lin_skip.bias"])
    raise NotImplementedError(kind)


def encoder_forward(kind, layers, x, adj_t, keep_masks=None, p_drop=0.0):
    """layer.py:18-27.  Every conv but the last is followed by relu + dropout; the
    last conv is bare unless the net has exactly one layer, in which case it too is
    followed by relu + dropout.  ``keep_masks[i]`` (0/1 tensor, optional) replaces
    the RNG: activation *= mask / (1 - p_drop)."""
    L = len(layers)
    for i, p in enumerate(layers):
        x = conv_forward(kind, p, x, adj_t)
        if i < L - 1 or L == 1:
            x = torch.relu(x)
            if keep_masks is not None and p_drop > 0:
                x = x * keep_masks[i].to(x.dtype) / (1.0 - p_drop)
    return x


# ---------------------------------------------------------------------------
# predictors  (plnlp/layer.py:66-87, 167-176)
# ---------------------------------------------------------------------------
def mlp_score(lins, x_i, x_j, keep_masks=None, p_drop=0.0):
    """layer.py:80-87.  ``lins`` = [(W, b), ...]; output [P, out]."""
    x = x_i * x_j
    for i, (W, b) in enumerate(lins[:-1]):
        x = torch.relu(F.linear(x, W, b))
        if keep_masks is not None and p_drop > 0:
            x = x * keep_masks[i].to(x.dtype) / (1.0 - p_drop)
    W, b = lins[-1]
    return F.linear(x, W, b)


def dot_score(x_i, x_j):
    """layer.py:174-176; output [P]."""
    return (x_i * x_j).sum(-1)


def _node_mlp(lins, x):
    """the shared loop of MLPDotPredictor / MLPBilPredictor (layer.py:132-137, 157-162) in eval mode: EVERY linear
    is followed by relu (dropout is the identity without training)"""
    for W, b in lins:
        x = torch.relu(F.linear(x, W, b))
    return x


def bil_score(W, x_i, x_j):
    """BilinearPredictor, layer.py:179-189: sum(bilin(x_i) * x_j, -1); output [P]."""
    return (F.linear(x_i, W) * x_j).sum(-1)


def mlpdot_score(lins, x_i, x_j):
    """MLPDotPredictor, layer.py:119-139; output [P]."""
    return (_node_mlp(lins, x_i) * _node_mlp(lins, x_j)).sum(-1)


def mlpbil_score(lins, W, x_i, x_j):
    """MLPBilPredictor, layer.py:142-164; output [P]."""
    return (F.linear(_node_mlp(lins, x_i), W) * _node_mlp(lins, x_j)).sum(-1)


def mlpcat_score(lins, x_i, x_j):
    """MLPCatPredictor, layer.py:90-116: the MLP on [x_i | x_j] and on [x_j | x_i], averaged; output [P, out]."""
    x1, x2 = torch.cat([x_i, x_j], -1), torch.cat([x_j, x_i], -1)
    for W, b in lins[:-1]:
        x1, x2 = torch.relu(F.linear(x1, W, b)), torch.relu(F.linear(x2, W, b))
    W, b = lins[-1]
    return (F.linear(x1, W, b) + F.linear(x2, W, b)) / 2


# ---------------------------------------------------------------------------
# losses  (plnlp/loss.py:5-14, 31-35) + closed-form gradients
# ---------------------------------------------------------------------------
def pair_loss(kind, pos_out, neg_out, num_neg, weight=None):
    p = pos_out.reshape(-1, 1)
    n = neg_out.reshape(-1, num_neg)
    if kind == "AUC":
        return (1 - (p - n)).square().sum()
    if kind == "HingeAUC":
        return (1 - (p - n)).clamp(min=0).square().sum()
    if kind == "WeightedHingeAUC":
        w = weight.reshape(-1, 1)
        return (w * (w - (p - n)).clamp(min=0).square()).sum()
    if kind == "WeightedAUC":                      # loss.py:17-21
        return (weight.reshape(-1, 1) * (1 - (p - n)).square()).sum()
    if kind == "AdaAUC":                           # loss.py:24-28
        return (weight.reshape(-1, 1) - (p - n)).square().sum()
    if kind == "AdaHingeAUC":                      # loss.py:38-42
        return (weight.reshape(-1, 1) - (p - n)).clamp(min=0).square().sum()
    if kind == "LogRank":                          # loss.py:45-48
        return -torch.log(torch.sigmoid(p - n) + 1e-15).mean()
    if kind == "CE":                               # loss.py:51-54
        return -torch.log(torch.sigmoid(pos_out) + 1e-15).mean() - torch.log(1 - torch.sigmoid(neg_out) + 1e-15).mean()
    if kind == "InfoNCE":                          # loss.py:57-62
        pe, ne = torch.exp(p), torch.exp(n).sum(1, keepdim=True)
        return -torch.log(pe / (pe + ne) + 1e-15).mean()
    raise NotImplementedError(kind)


def pair_loss_autograd(kind, pos_out, neg_out, num_neg, weight=None):
    """(loss, d loss/d pos [B], d loss/d neg [B*num_neg]) through torch autograd, any kind"""
    p = pos_out.detach().clone().reshape(-1).requires_grad_(True)
    n = neg_out.detach().clone().reshape(-1).requires_grad_(True)
    loss = pair_loss(kind, p, n, num_neg, weight)
    gp, gn = torch.autograd.grad(loss, (p, n))
    return loss.detach(), gp, gn


def pair_loss_grad(kind, pos_out, neg_out, num_neg, weight=None):
    """Closed-form d loss / d pos [B], d loss / d neg [B, num_neg]."""
    p = pos_out.reshape(-1, 1)
    n = neg_out.reshape(-1, num_neg)
    if kind == "AUC":
        t = 1 - (p - n)
        g = 2 * t
    elif kind == "HingeAUC":
        t = (1 - (p - n)).clamp(min=0)
        g = 2 * t
    elif kind == "WeightedHingeAUC":
        w = weight.reshape(-1, 1)
        t = (w - (p - n)).clamp(min=0)
        g = 2 * w * t
    else:
        raise NotImplementedError(kind)
    return -g.sum(1), g


def select_loss(name, has_weight):
    """model.py:107-126 dispatch restricted to the in-scope losses: unknown names
    and weighted losses without a weight fall back to AUC."""
    if name in ("CE", "InfoNCE", "LogRank", "HingeAUC"):
        return name
    if name in ("AdaAUC", "WeightedAUC", "AdaHingeAUC", "WeightedHingeAUC") and has_weight:
        return name
    return "AUC"


# ---------------------------------------------------------------------------
# negative samplers  (plnlp/negative_sample.py:6-20, 31-43)
# ---------------------------------------------------------------------------
def global_neg_sample(edge_index, num_nodes, num_samples, num_neg):
    ei, _ = pyg.add_self_loops(edge_index, num_nodes=num_nodes)
    neg = pyg.negative_sampling(ei, num_nodes=num_nodes, num_neg_samples=num_samples * num_neg)
    src, dst = neg[0], neg[1]
    want = num_samples * num_neg
    if neg.size(1) < want:
        extra = torch.randperm(neg.size(1))[: want - neg.size(1)]
        src, dst = torch.cat([src, src[extra]]), torch.cat([dst, dst[extra]])
    return torch.stack([src, dst], -1).reshape(-1, num_neg, 2)


def local_neg_sample(pos_edges, num_nodes, num_neg):
    E = pos_edges.size(0)
    src = pos_edges[:, 0].reshape(-1, 1).repeat(1, num_neg).reshape(-1)
    dst = torch.randint(0, num_nodes, (num_neg * E,), dtype=torch.long)
    return torch.stack([src, dst], -1).reshape(-1, num_neg, 2)


# ---------------------------------------------------------------------------
# edge assembly  (plnlp/utils.py:7-41)
# ---------------------------------------------------------------------------
def eval_edges(split, split_edge):
    tr = split_edge["train"]
    if "edge" in tr:
        return split_edge[split]["edge"], split_edge[split]["edge_neg"]
    s, t = split_edge[split]["source_node"], split_edge[split]["target_node"]
    tn = split_edge[split]["target_node_neg"]
    pos = torch.stack([s, t]).t()
    neg = torch.stack([s.repeat_interleave(tn.size(1)), tn.reshape(-1)]).t()
    return pos, neg


def train_pos_edges(split_edge):
    tr = split_edge["train"]
    if "edge" in tr:
        return tr["edge"]
    return torch.stack([tr["source_node"], tr["target_node"]]).t()


# ---------------------------------------------------------------------------
# evaluation glue  (plnlp/utils.py:44-80)
# ---------------------------------------------------------------------------
def evaluate_hits(pos_val, neg_val, pos_test, neg_test):
    return {f"Hits@{K}": (hits_at_k(pos_val, neg_val, K), hits_at_k(pos_test, neg_test, K))
            for K in (20, 50, 100)}


def evaluate_mrr(pos_val, neg_val, pos_test, neg_test):
    v = mrr_dict(pos_val, neg_val.view(pos_val.shape[0], -1))["mrr_list"].mean().item()
    t = mrr_dict(pos_test, neg_test.view(pos_test.shape[0], -1))["mrr_list"].mean().item()
    return {"MRR": (v, t)}


# ---------------------------------------------------------------------------
# the model: a CPU trainer with the reference's step semantics (model.py)
# ---------------------------------------------------------------------------
class OracleModel:
    """Functional restatement of ``BaseModel`` (model.py:45-226) for the in-scope
    configurations.  Parameters live in plain leaf tensors so tests can inject and
    read them by name:

      emb                       [N, emb_hidden]          (model.py:229-249)
      enc.{i}.<pyg name>        per conv                 (layer.py:30-45)
      pred.{i}.weight/.bias     MLP head only            (layer.py:69-74)
    """

    def __init__(self, *, num_nodes, emb_hidden, gnn_hidden, mlp_hidden, gnn_layers, mlp_layers,
                 encoder="SAGE", predictor="MLP", loss="AUC", lr=1e-3, dropout=0.0,
                 clip_norm=2.0, num_node_feats=0, use_node_feats=False, train_node_emb=True,
                 dtype=torch.float32, optimizer="Adam"):
        self.kind, self.pred_kind, self.loss_name = encoder.upper(), predictor.upper(), loss
        self.dropout, self.clip_norm, self.lr = dropout, clip_norm, lr
        self.optimizer_name = optimizer
        self.use_node_feats, self.train_node_emb = use_node_feats, train_node_emb
        self.num_nodes = num_nodes
        in_dim = 0
        self.params = {}
        g = torch.Generator().manual_seed(1234)

        def uni(shape, bound):
            return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

        if use_node_feats:
            in_dim += num_node_feats
        if (not use_node_feats) or train_node_emb:
            self.params["emb"] = uni((num_nodes, emb_hidden), math.sqrt(6.0 / (num_nodes + emb_hidden)))
            in_dim += emb_hidden
        self.in_dim = in_dim
        for i in range(gnn_layers):
            fi = in_dim if i == 0 else gnn_hidden
            b = 1.0 / math.sqrt(fi)
            if self.kind == "SAGE":
                self.params[f"enc.{i}.lin_l.weight"] = uni((gnn_hidden, fi), b)
                self.params[f"enc.{i}.lin_l.bias"] = uni((gnn_hidden,), b)
                self.params[f"enc.{i}.lin_r.weight"] = uni((gnn_hidden, fi), b)
            else:
                self.params[f"enc.{i}.lin.weight"] = uni((gnn_hidden, fi), math.sqrt(6.0 / (fi + gnn_hidden)))
                self.params[f"enc.{i}.bias"] = torch.zeros(gnn_hidden, dtype=dtype)
        self.gnn_layers = gnn_layers
        self.mlp_layers = mlp_layers if self.pred_kind == "MLP" else 0
        for i in range(self.mlp_layers):
            fi = mlp_hidden
            fo = 1 if i == mlp_layers - 1 else mlp_hidden
            b = 1.0 / math.sqrt(fi)
            self.params[f"pred.{i}.weight"] = uni((fo, fi), b)
            self.params[f"pred.{i}.bias"] = uni((fo,), b)
        for v in self.params.values():
            v.requires_grad_(True)
        self._make_optimizer()

    def _make_optimizer(self):
        # parameter order of model.py:81-83: encoder, predictor, embedding
        order = [k for k in self.params if k.startswith("enc.")] + \
                [k for k in self.params if k.startswith("pred.")] + \
                [k for k in self.params if k == "emb"]
        plist = [self.params[k] for k in order]
        if self.optimizer_name == "SGD":   # model.py:87-88
            self.optimizer = torch.optim.SGD(plist, lr=self.lr, momentum=0.9, weight_decay=1e-5, nesterov=True)
        else:
            self.optimizer = torch.optim.Adam(plist, lr=self.lr)

    def load(self, state):
        with torch.no_grad():
            for k, v in state.items():
                self.params[k].copy_(v)

    def state(self):
        return {k: v.detach().clone() for k, v in self.params.items()}

    # -- pieces --------------------------------------------------------------
    def input_feat(self, x):
        """model.py:98-105"""
        if self.use_node_feats:
            if self.train_node_emb:
                return torch.cat([self.params["emb"], x], dim=-1)
            return x
        return self.params["emb"]

    def enc_layers(self):
        out = []
        for i in range(self.gnn_layers):
            pre = f"enc.{i}."
            out.append({k[len(pre):]: v for k, v in self.params.items() if k.startswith(pre)})
        return out

    def encode(self, x, adj_t, keep_masks=None):
        return encoder_forward(self.kind, self.enc_layers(), self.input_feat(x), adj_t,
                               keep_masks, self.dropout)

    def score(self, h, edges, keep_masks=None):
        """edges [P, 2] -> scores [P]"""
        xi, xj = h[edges[:, 0]], h[edges[:, 1]]
        if self.pred_kind == "DOT":
            return dot_score(xi, xj)
        lins = [(self.params[f"pred.{i}.weight"], self.params[f"pred.{i}.bias"])
                for i in range(self.mlp_layers)]
        return mlp_score(lins, xi, xj, keep_masks, self.dropout).reshape(-1)

    # -- one optimisation step (model.py:148-171) ------------------------------
    def step(self, x, adj_t, pos_edge, neg_edge, num_neg, weight=None, do_update=True):
        """pos_edge [B,2]; neg_edge [B,num_neg,2].  Returns (loss, h)."""
        self.optimizer.zero_grad()
        h = self.encode(x, adj_t)
        pos_out = self.score(h, pos_edge)
        neg_out = self.score(h, neg_edge.reshape(-1, 2))
        kind = select_loss(self.loss_name, weight is not None)
        loss = pair_loss(kind, pos_out, neg_out, num_neg, weight)
        loss.backward()
        if do_update:
            if self.clip_norm >= 0:
                torch.nn.utils.clip_grad_norm_(
                    [v for k, v in self.params.items() if k.startswith("enc.")], self.clip_norm)
                pp = [v for k, v in self.params.items() if k.startswith("pred.")]
                if pp:
                    torch.nn.utils.clip_grad_norm_(pp, self.clip_norm)
            self.optimizer.step()
        return loss.detach(), h.detach()

    def train_epoch(self, x, adj_t, pos_train, neg_train, perms, num_neg, weight=None):
        """model.py:128-173 with the shuffles (``perms``: list of index tensors) and
        the negatives supplied.  Returns the reference's reported loss."""
        tot = cnt = 0
        for perm in perms:
            w = weight[perm] if weight is not None else None
            loss, _ = self.step(x, adj_t, pos_train[perm], neg_train[perm], num_neg, w)
            tot += loss.item() * perm.numel()
            cnt += perm.numel()
        return tot / cnt

    @torch.no_grad()
    def predict(self, x, adj_t, edges, batch_size=65536):
        """model.py:175-194 (encoder once; mean row appended so index -1 resolves)."""
        h = self.encode(x, adj_t)
        h = torch.cat([h, h.mean(0, keepdim=True)], 0)
        return torch.cat([self.score(h, edges[i:i + batch_size])
                          for i in range(0, edges.size(0), batch_size)])
