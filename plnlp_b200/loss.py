"""Pairwise losses with the reference's call signatures (/root/reference/plnlp/loss.py), each one
a single fused forward+gradient kernel (csrc/loss.cu)."""
from __future__ import annotations

from . import _ops


def auc_loss(pos_out, neg_out, num_neg):
    """loss.py:5-8: sum (1 - (pos - neg))^2"""
    return _ops.pair_loss("AUC", pos_out, neg_out, num_neg)


def hinge_auc_loss(pos_out, neg_out, num_neg):
    """loss.py:11-14: sum max(0, 1 - (pos - neg))^2"""
    return _ops.pair_loss("HingeAUC", pos_out, neg_out, num_neg)


def weighted_hinge_auc_loss(pos_out, neg_out, num_neg, weight):
    """loss.py:31-35: sum w * max(0, w - (pos - neg))^2 (w is weight and margin)"""
    return _ops.pair_loss("WeightedHingeAUC", pos_out, neg_out, num_neg, weight)


def _out_of_scope(name, where):
    def f(*args, **kwargs):
        raise NotImplementedError(f"{name} ({where}) is outside the hot-path scope of plnlp_b200 "
                                  "(SURVEY.md section 8f)")
    f.__name__ = name
    return f


weighted_auc_loss = _out_of_scope("weighted_auc_loss", "loss.py:17-21")
adaptive_auc_loss = _out_of_scope("adaptive_auc_loss", "loss.py:24-28")
adaptive_hinge_auc_loss = _out_of_scope("adaptive_hinge_auc_loss", "loss.py:38-42")
log_rank_loss = _out_of_scope("log_rank_loss", "loss.py:45-48")
ce_loss = _out_of_scope("ce_loss", "loss.py:51-54")
info_nce_loss = _out_of_scope("info_nce_loss", "loss.py:57-62")
