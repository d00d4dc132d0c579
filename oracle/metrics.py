"""Oracle (TEST INFRASTRUCTURE ONLY) for the device-side evaluation metrics (``gn_lp_metrics``,
``gn_nc_metrics``).

The reference computes them with scikit-learn (``gripnet/utils.py:28-52``: ``roc_auc_score``,
``average_precision_score``, ``auc(precision_recall_curve)``, ``f1_score`` micro/macro, ``accuracy_score``) —
a third-party dependency with no pin in the reference (``setup.py:3-10``); this image has scikit-learn 1.9.0.
The functions below restate the published definitions in numpy float64 and are pinned against sklearn itself
in ``tests/test_metrics.py`` (sklearn is importable here and on the GPU box, so the GPU tests also compare
with it directly).
"""
import numpy as np


def binary_curve(y_true, y_score):
    """Cumulative (fps, tps) at every DISTINCT score, scores descending (sklearn's
    ``_binary_clf_curve`` / ``confusion_matrix_at_thresholds``)."""
    y_true = np.asarray(y_true, dtype=np.float64)
    y_score = np.asarray(y_score, dtype=np.float64)
    order = np.argsort(-y_score, kind="stable")
    s, t = y_score[order], y_true[order]
    ends = np.r_[np.nonzero(np.diff(s))[0], s.size - 1]
    tps = np.cumsum(t)[ends]
    fps = (1 + ends) - tps
    return fps, tps


def auprc_auroc_ap(y_true, y_score):
    """(auprc, auroc, ap) of one binary problem, in the reference's return order (utils.py:28-35)."""
    fps, tps = binary_curve(y_true, y_score)
    P, N = tps[-1], fps[-1]
    nan = float("nan")
    if P <= 0:
        return nan, nan, nan
    prec = tps / (tps + fps)
    rec = tps / P
    rec_prev = np.r_[0.0, rec[:-1]]
    prec_prev = np.r_[1.0, prec[:-1]]                       # the curve is closed at (recall 0, precision 1)
    ap = float(np.sum((rec - rec_prev) * prec))
    auprc = float(np.sum((rec - rec_prev) * (prec + prec_prev) * 0.5))
    if N <= 0:
        return auprc, nan, ap
    tpr, fpr = np.r_[0.0, tps / P], np.r_[0.0, fps / N]
    auroc = float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) * 0.5))
    return auprc, auroc, ap


def lp_record(pos_score, neg_score, pos_range, neg_range=None):
    """``record[3, R]`` of the reference's evaluation loop (GripNet-pose.py:148-164)."""
    neg_range = pos_range if neg_range is None else neg_range
    rec = np.zeros((3, len(pos_range)))
    for r, ((ps, pe), (ns, ne)) in enumerate(zip(pos_range, neg_range)):
        p, n = np.asarray(pos_score[ps:pe]), np.asarray(neg_score[ns:ne])
        rec[:, r] = auprc_auroc_ap(np.r_[np.ones(p.size), np.zeros(n.size)], np.r_[p, n])
    return rec


def micro_macro_acc(target, pred):
    """(micro-F1, macro-F1, accuracy) of single-label multiclass predictions (utils.py:38-52)."""
    target, pred = np.asarray(target), np.asarray(pred)
    labels = np.union1d(target, pred)
    f1 = []
    for c in labels:
        tp = np.sum((target == c) & (pred == c))
        fp = np.sum((target != c) & (pred == c))
        fn = np.sum((target == c) & (pred != c))
        f1.append(2.0 * tp / (2.0 * tp + fp + fn) if (2 * tp + fp + fn) > 0 else 0.0)
    accuracy = float(np.mean(target == pred))
    return accuracy, float(np.mean(f1)), accuracy
