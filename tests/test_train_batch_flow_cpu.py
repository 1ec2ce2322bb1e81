"""CPU test (-m "not gpu") of BaseModel.train_batch's CONTROL FLOW with the kernels replaced by torch stand-ins:
which branch runs (plain / row-restricted last conv / full output with a row-sparse gradient), that edges are renumbered consistently
with the compact table, and that the loss and gradients equal the plain path's.  The arithmetic itself is covered
by the GPU parity tests; this pins the Python glue that sits between them."""
import pytest
import torch

from plnlp_b200 import _ops, model as M


class _Enc(torch.nn.Module):
    """h = relu(x W): supports out_rows when ``restrict`` is set (returns the compact rows)"""

    def __init__(self, n, f, restrict):
        super().__init__()
        self.w = torch.nn.Parameter(torch.randn(f, f) * 0.3)
        self.restrict, self.calls = restrict, []

    def forward(self, x, adj_t, out_rows=None, sparse_grad=False):
        h = torch.relu(x[0] @ self.w)
        if out_rows is None:
            self.calls.append("full+sparse_grad" if sparse_grad else "full")
            return h
        self.calls.append("rows" if self.restrict else "full+flag")
        return (h[out_rows], True) if self.restrict else (h, False)


class _Pred(torch.nn.Module):
    dropout = 0.0

    def flat_params(self):
        return []


def _fake_edge_score_loss(h, pos_edge, neg_edge, num_neg, loss_name, weight=None, head="MLP", params=(), drop_p=0.0, seed=0):
    s = lambda e: (h[e[:, 0]] * h[e[:, 1]]).sum(-1)
    p, n = s(pos_edge).reshape(-1, 1), s(neg_edge).reshape(-1, num_neg)
    return (1 - (p - n)).square().sum()


def _model(n, f, restrict, num_nodes):
    m = object.__new__(M.BaseModel)
    m.encoder, m.predictor = _Enc(n, f, restrict), _Pred()
    m.predictor.__class__ = type("DotPredictor", (M.DotPredictor,), {"flat_params": lambda self: []})
    m.emb = torch.nn.Embedding(n, f)
    m.use_node_feats, m.train_node_emb, m.num_nodes, m.clip_norm = False, True, num_nodes, -1.0
    m.world_size, m.rank, m.partitioned, m.loss_func_name, m.device = 1, 0, False, "AUC", torch.device("cpu")
    m.para_list = list(m.encoder.parameters()) + list(m.emb.parameters())
    m.optimizer = torch.optim.SGD(m.para_list, lr=0.0)
    return m


@pytest.mark.parametrize("restrict", [True, False])
def test_train_batch_branches_agree(restrict, monkeypatch):
    torch.manual_seed(0)
    n, f, B, k = 200, 6, 8, 2
    monkeypatch.setattr(_ops, "edge_score_loss", _fake_edge_score_loss)

    class D:
        adj_t, x, edge_index = None, None, None

    pos, neg = torch.randint(0, n, (B, 2)), torch.randint(0, n, (B * k, 2))
    results = []
    # first the model claims a tiny node set (every node counts as touched -> plain path), then a huge one
    # (the batch touches "a small part of the nodes" -> restricted last conv, or the row-sparse hint)
    for claimed in (10, 10 ** 9):
        m = _model(n, f, restrict, claimed)
        if results:
            m.encoder.load_state_dict(results[0][2]); m.emb.load_state_dict(results[0][3])
        enc_state = {k_: v.clone() for k_, v in m.encoder.state_dict().items()}
        emb_state = {k_: v.clone() for k_, v in m.emb.state_dict().items()}
        loss = m.train_batch(D, pos.clone(), neg.clone(), k)
        results.append((float(loss), [p.grad.clone() for p in m.para_list], enc_state, emb_state, list(m.encoder.calls)))
    (l0, g0, _, _, c0), (l1, g1, _, _, c1) = results
    assert c0 == ["full"] and c1 == (["rows"] if restrict else ["full+flag"])
    assert abs(l0 - l1) <= 1e-5 * abs(l0)
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
