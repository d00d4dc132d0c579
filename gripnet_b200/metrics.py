"""Evaluation metrics of the training scripts, computed on the device.

The reference moves every relation's scores to the host and calls sklearn once per relation
(``GripNet-pose.py:148-164``, ``:188-199`` around ``gripnet/utils.py:28-35``); ``lp_metrics`` ranks all
relations with one radix sort and returns the whole ``record[3, R]`` (rows auprc, auroc, ap) as a device
tensor — no device->host round trip per relation.  ``auprc_auroc_ap`` / ``micro_macro`` / ``acc`` keep the
reference's ``gripnet.utils`` signatures for CUDA tensors.  Values follow sklearn's definitions (ties form
one threshold; AUPRC is the trapezoid over the precision-recall points closed at (recall 0, precision 1); AP
is the step-wise sum); sums are float64 like sklearn's.
"""
import torch

from . import _lib
from .graph import _Cache, _ptr, _stream, _ws, require_cuda

_range_cache = _Cache(capacity=8)


def _ranges(range_list, n, device, name):
    """Validated device copy of a ``[R, 2]`` range list; tensors are validated once per (tensor, version), so a
    repeated call with the same range tensor does no device<->host traffic (CUDA-graph capturable)."""
    if torch.is_tensor(range_list):
        key = (_Cache.tkey(range_list), int(n), str(device))
        return _range_cache.get(key, (range_list,), lambda: _ranges_checked(range_list, n, device, name))
    return _ranges_checked(range_list, n, device, name)


def _ranges_checked(range_list, n, device, name):
    rl = torch.as_tensor(range_list)
    if rl.dim() != 2 or rl.size(1) != 2:
        raise RuntimeError(f"{name} must have shape [R, 2]")
    host = rl.to(torch.int64).cpu()
    flat = host.flatten().tolist()
    ok = all(0 <= flat[2 * r] <= flat[2 * r + 1] <= n for r in range(host.size(0))) and \
        all(flat[2 * r + 1] <= flat[2 * r + 2] for r in range(host.size(0) - 1))
    if not ok:
        raise RuntimeError(f"{name} must hold ascending, non-overlapping [start, end) slices within [0, {n}]")
    return host.to(device).contiguous()


def lp_metrics(pos_score, neg_score, range_list, neg_range_list=None, out=None):
    """float64 ``[3, R]`` device tensor: rows auprc, auroc, ap; column r scores
    ``pos_score[range_list[r,0]:range_list[r,1]]`` against the negatives of ``neg_range_list`` (default: the
    same slices).  NaN where sklearn would refuse (a relation with a single class)."""
    lib = _lib.load()
    require_cuda(pos_score, "pos_score", torch.float32)
    require_cuda(neg_score, "neg_score", torch.float32)
    pos, neg = pos_score.detach().contiguous().view(-1), neg_score.detach().contiguous().view(-1)
    dev = pos.device
    pr = _ranges(range_list, pos.numel(), dev, "range_list")
    nr = pr if neg_range_list is None else _ranges(neg_range_list, neg.numel(), dev, "neg_range_list")
    if neg_range_list is None and neg.numel() < pos.numel():
        _ranges(range_list, neg.numel(), dev, "range_list (applied to neg_score)")
    if nr.size(0) != pr.size(0):
        raise RuntimeError("neg_range_list must have one slice per relation")
    n_rel = int(pr.size(0))
    if out is None:
        out = torch.empty((3, n_rel), dtype=torch.float64, device=dev)
    elif out.shape != (3, n_rel) or out.dtype != torch.float64 or not out.is_contiguous() or out.device != dev:
        raise RuntimeError("out must be a contiguous float64 [3, R] tensor on the scores' device")
    if n_rel == 0:
        return out
    nbytes = int(lib.gn_lp_metrics_workspace_bytes(pos.numel(), neg.numel(), n_rel))
    ws = _ws(nbytes, dev)
    _lib.check(lib.gn_lp_metrics(_ptr(pos) if pos.numel() else None, pos.numel(), _ptr(neg) if neg.numel() else None,
                                 neg.numel(), _ptr(pr), _ptr(nr), n_rel, _ptr(out), _ptr(ws), nbytes, _stream()),
               "gn_lp_metrics")
    return out


def auprc_auroc_ap(target_tensor, score_tensor):
    """``gripnet.utils.auprc_auroc_ap`` (utils.py:28-35) for CUDA tensors: ``(auprc, auroc, ap)`` floats of one
    binary problem (targets 1 = positive, 0 = negative)."""
    require_cuda(score_tensor, "score_tensor")
    t = target_tensor.to(score_tensor.device)
    s = score_tensor.detach().to(torch.float32).view(-1)
    pos, neg = s[t.view(-1) > 0.5].contiguous(), s[t.view(-1) <= 0.5].contiguous()
    rec = lp_metrics(pos, neg, [[0, pos.numel()]], [[0, neg.numel()]]).cpu()
    return float(rec[0, 0]), float(rec[1, 0]), float(rec[2, 0])


def argmax_rows(score):
    """``torch.argmax(score, dim=1)`` (GripNet-aminer.py:131) as one kernel; int64 ``[n]``."""
    lib = _lib.load()
    require_cuda(score, "score", torch.float32)
    if score.dim() != 2 or score.stride(1) != 1:
        raise RuntimeError("score must be a 2-D tensor with unit column stride")
    n, c = score.shape
    out = torch.empty(n, dtype=torch.int64, device=score.device)
    _lib.check(lib.gn_argmax_rows(_ptr(score) if n else None, score.stride(0), n, c, _ptr(out) if n else None, _stream()),
               "gn_argmax_rows")
    return out


def nc_metrics(target, pred, num_classes, out=None):
    """float64 ``[3]`` device tensor: micro-F1, macro-F1, accuracy of integer class predictions."""
    lib = _lib.load()
    require_cuda(target, "target", torch.int64)
    require_cuda(pred, "pred", torch.int64)
    t, p = target.contiguous().view(-1), pred.contiguous().view(-1)
    if t.numel() != p.numel():
        raise RuntimeError("target and pred must have the same length")
    if out is None:
        out = torch.empty(3, dtype=torch.float64, device=t.device)
    elif out.shape != (3,) or out.dtype != torch.float64 or not out.is_contiguous() or out.device != t.device:
        raise RuntimeError("out must be a contiguous float64 [3] tensor on the labels' device")
    nbytes = int(lib.gn_nc_metrics_workspace_bytes(int(num_classes)))
    ws = _ws(nbytes, t.device)
    n = t.numel()
    _lib.check(lib.gn_nc_metrics(_ptr(t) if n else None, _ptr(p) if n else None, n, int(num_classes), _ptr(out),
                                 _ptr(ws), nbytes, _stream()), "gn_nc_metrics")
    return out


def _num_classes(target, pred):
    return int(max(int(target.max()), int(pred.max()))) + 1 if target.numel() else 1


def micro_macro(target_tensor, score_tensor, num_classes=None):
    """``gripnet.utils.micro_macro`` (utils.py:38-46): ``(micro_f1, macro_f1)`` of class predictions."""
    c = _num_classes(target_tensor, score_tensor) if num_classes is None else num_classes
    m = nc_metrics(target_tensor, score_tensor, c).cpu()
    return float(m[0]), float(m[1])


def acc(target_tensor, score_tensor, num_classes=None):
    """``gripnet.utils.acc`` (utils.py:49-52)."""
    c = _num_classes(target_tensor, score_tensor) if num_classes is None else num_classes
    return float(nc_metrics(target_tensor, score_tensor, c)[2].cpu())
