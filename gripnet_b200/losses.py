"""Fused losses of the reference training scripts (one kernel pair each).

``link_prediction_loss`` = ``GripNet-pose.py:140-142``;
``node_classification_loss`` = ``GripNet-aminer.py:133``.
The decoders still return plain scores, so the scripts' own torch expressions
keep working; these are the fused, deterministic equivalents.
"""
from . import ops

EPS = ops.EPS


def link_prediction_loss(pos_score, neg_score):
    """``-mean(log(pos + EPS)) - mean(log(1 - neg + EPS))``"""
    return ops.LinkPredLoss.apply(pos_score, neg_score)


def node_classification_loss(score, labels):
    """``-mean(log(score[i, labels[i]] + EPS))``"""
    return ops.NodeClassLoss.apply(score, labels)
