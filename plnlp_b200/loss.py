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


def weighted_auc_loss(pos_out, neg_out, num_neg, weight):
    """loss.py:17-21: sum w * (1 - (pos - neg))^2"""
    return _ops.pair_loss("WeightedAUC", pos_out, neg_out, num_neg, weight)


def adaptive_auc_loss(pos_out, neg_out, num_neg, margin):
    """loss.py:24-28: sum (margin - (pos - neg))^2"""
    return _ops.pair_loss("AdaAUC", pos_out, neg_out, num_neg, margin)


def weighted_hinge_auc_loss(pos_out, neg_out, num_neg, weight):
    """loss.py:31-35: sum w * max(0, w - (pos - neg))^2 (w is weight and margin)"""
    return _ops.pair_loss("WeightedHingeAUC", pos_out, neg_out, num_neg, weight)


def adaptive_hinge_auc_loss(pos_out, neg_out, num_neg, weight):
    """loss.py:38-42: sum max(0, margin - (pos - neg))^2"""
    return _ops.pair_loss("AdaHingeAUC", pos_out, neg_out, num_neg, weight)


def log_rank_loss(pos_out, neg_out, num_neg):
    """loss.py:45-48: mean -log(sigmoid(pos - neg) + 1e-15)"""
    return _ops.pair_loss("LogRank", pos_out, neg_out, num_neg)


def ce_loss(pos_out, neg_out):
    """loss.py:51-54: mean -log(sigmoid(pos) + 1e-15) + mean -log(1 - sigmoid(neg) + 1e-15); the number of
    negatives per positive is inferred from the shapes"""
    k = max(neg_out.numel() // max(pos_out.numel(), 1), 1)
    return _ops.pair_loss("CE", pos_out, neg_out, k)


def info_nce_loss(pos_out, neg_out, num_neg):
    """loss.py:57-62: mean -log(e^pos / (e^pos + sum_j e^neg_j) + 1e-15)"""
    return _ops.pair_loss("InfoNCE", pos_out, neg_out, num_neg)
