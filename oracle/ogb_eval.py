"""Oracle restatement of ``ogb.linkproppred.Evaluator`` (ogb 1.3.2).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  ogb is pinned by
/root/reference/README.md:18 but not vendored / installable here.  Call sites:
/root/reference/main.py:215 (construction), plnlp/utils.py:47-56 (hits) and
plnlp/utils.py:67-76 (mrr).

Assumption (10): Hits@K = fraction of positives STRICTLY greater than the K-th
largest negative score, 1.0 when there are fewer than K negatives.  MRR (1.3.2) =
``1 / (1 + position of the positive in a descending argsort of [pos, negs])``; under
exact ties that position is implementation defined, so the oracle also exposes the
optimistic (``#neg > pos``) and pessimistic (``#neg >= pos``) ranks; parity tests
assert they coincide on the compared inputs.
"""
from __future__ import annotations

import torch

_METRIC = {"ogbl-ddi": ("hits", 20), "ogbl-collab": ("hits", 50), "ogbl-ppa": ("hits", 100),
           "ogbl-citation2": ("mrr", None)}


class Evaluator:
    def __init__(self, name):
        self.name = name
        self.eval_metric, self.K = _METRIC.get(name, ("hits", 20))

    def eval(self, input_dict):
        pos, neg = input_dict["y_pred_pos"], input_dict["y_pred_neg"]
        if neg.dim() == 1:
            return {f"hits@{self.K}": hits_at_k(pos, neg, self.K)}
        return mrr_dict(pos, neg)


def hits_at_k(pos, neg, K):
    if neg.numel() < K:
        return 1.0
    kth = torch.topk(neg, K)[0][-1]
    return float(torch.sum(pos > kth).cpu()) / pos.numel()


def mrr_ranks(pos, neg):
    """(optimistic, pessimistic) 1-based integer ranks of each positive among its
    row of negatives."""
    gt = (neg > pos.view(-1, 1)).sum(1)
    ge = (neg >= pos.view(-1, 1)).sum(1)
    return gt + 1, ge + 1


def mrr_dict(pos, neg):
    y = torch.cat([pos.view(-1, 1), neg], dim=1)
    order = torch.argsort(y, dim=1, descending=True)
    rank = torch.nonzero(order == 0, as_tuple=False)[:, 1] + 1
    return {"mrr_list": 1.0 / rank.to(torch.float),
            "hits@1_list": (rank <= 1).to(torch.float),
            "hits@3_list": (rank <= 3).to(torch.float),
            "hits@10_list": (rank <= 10).to(torch.float)}
