"""Device-side evaluation metrics (gn_lp_metrics / gn_nc_metrics): the numpy oracle is pinned against scikit-learn
(the library the reference calls, gripnet/utils.py:28-52); the CUDA kernels are compared with both."""
import warnings

import numpy as np
import pytest
import torch
from sklearn import metrics as skm

from oracle import metrics as om


def _sk_auprc_auroc_ap(y, pred):
    """gripnet/utils.py:28-35, verbatim semantics."""
    auroc, ap = skm.roc_auc_score(y, pred), skm.average_precision_score(y, pred)
    prec, rec, _ = skm.precision_recall_curve(y, pred)
    return skm.auc(rec, prec), auroc, ap


def _scores(rs, n, kind):
    if kind == "continuous":
        return rs.rand(n).astype(np.float32)
    if kind == "ties":                       # few distinct values: every threshold is a tie group
        return (rs.randint(0, 7, n) / 8.0).astype(np.float32)
    if kind == "saturated":                  # sigmoid saturation: many exact 0.0 / 1.0
        return np.clip(rs.randn(n) * 2 + 0.5, 0, 1).astype(np.float32)
    if kind == "signed":                     # raw (non-sigmoid) scores incl. negatives and -0.0
        s = rs.randn(n).astype(np.float32)
        s[::17] = -0.0
        s[1::17] = 0.0
        return s
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["continuous", "ties", "saturated", "signed"])
def test_oracle_matches_sklearn_binary(kind):
    rs = np.random.RandomState(0)
    for n_pos, n_neg in [(1, 1), (5, 3), (100, 200), (1000, 1000)]:
        p = _scores(rs, n_pos, kind) + (0.2 if kind == "continuous" else 0)
        n = _scores(rs, n_neg, kind)
        y, s = np.r_[np.ones(n_pos), np.zeros(n_neg)], np.r_[p, n]
        np.testing.assert_allclose(om.auprc_auroc_ap(y, s), _sk_auprc_auroc_ap(y, s), rtol=1e-12, atol=1e-14)


def test_oracle_single_class_is_nan():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = om.auprc_auroc_ap(np.ones(4), np.arange(4.0))
        assert np.isnan(a[1]) and a[0] == pytest.approx(1.0) and a[2] == pytest.approx(1.0)
        assert all(np.isnan(om.auprc_auroc_ap(np.zeros(4), np.arange(4.0))))


def test_oracle_matches_sklearn_multiclass():
    rs = np.random.RandomState(1)
    for c, n in [(2, 50), (8, 1000), (5, 7)]:
        t, p = rs.randint(0, c, n), rs.randint(0, c, n)
        if c == 8:
            p[p == 3] = 2                   # a class never predicted
            t[t == 6] = 5                   # a class absent from the targets (and 6 still predicted)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = (skm.f1_score(t, p, average="micro"), skm.f1_score(t, p, average="macro"), skm.accuracy_score(t, p))
        np.testing.assert_allclose(om.micro_macro_acc(t, p), want, rtol=1e-12)


def _ranges(sizes):
    b = np.cumsum([0] + list(sizes))
    return np.stack([b[:-1], b[1:]], axis=1).astype(np.int64)


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["continuous", "ties", "saturated", "signed"])
@pytest.mark.parametrize("sizes", [[1], [5, 3, 9], [1025, 1, 4096, 333, 2], [2500] * 16, [40000, 7]])   # last: > 1 tile per CTA
def test_device_lp_metrics_match_oracle_and_sklearn(kind, sizes):
    from gripnet_b200.metrics import lp_metrics
    rs = np.random.RandomState(len(sizes))
    n = int(sum(sizes))
    pos = _scores(rs, n, kind) + (0.1 if kind == "continuous" else 0)
    neg = _scores(rs, n, kind)
    rl = _ranges(sizes)
    rec = lp_metrics(torch.tensor(pos, device="cuda"), torch.tensor(neg, device="cuda"), torch.tensor(rl)).cpu().numpy()
    want = om.lp_record(pos, neg, rl)
    np.testing.assert_allclose(rec, want, rtol=1e-12, atol=1e-14)
    for r, (s, e) in enumerate(rl):                    # the reference's own loop (GripNet-pose.py:148-164)
        y, sc = np.r_[np.ones(e - s), np.zeros(e - s)], np.r_[pos[s:e], neg[s:e]]
        np.testing.assert_allclose(rec[:, r], _sk_auprc_auroc_ap(y, sc), rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
def test_device_lp_metrics_separate_negative_ranges_gaps_and_empty_relations():
    from gripnet_b200.metrics import lp_metrics
    rs = np.random.RandomState(7)
    pos, neg = rs.rand(900).astype(np.float32), rs.rand(1500).astype(np.float32)
    pr = np.array([[0, 100], [150, 150], [200, 900]], dtype=np.int64)        # a gap and an empty slice
    nr = np.array([[0, 700], [700, 800], [800, 800]], dtype=np.int64)        # relation 2 has no negatives
    rec = lp_metrics(torch.tensor(pos, device="cuda"), torch.tensor(neg, device="cuda"), pr, nr).cpu().numpy()
    want = om.lp_record(pos, neg, pr, nr)
    np.testing.assert_allclose(rec[:, 0], want[:, 0], rtol=1e-12)
    assert np.isnan(rec[:, 1]).all() and np.isnan(want[:, 1]).all()           # no positives
    assert np.isnan(rec[1, 2]) and rec[0, 2] == pytest.approx(1.0) and rec[2, 2] == pytest.approx(1.0)


@pytest.mark.gpu
def test_device_lp_metrics_pose_size_and_graph_capture():
    """pose-0 evaluation size: 16 relations x 25 000 positives + negatives; replayed from a CUDA graph."""
    from gripnet_b200.metrics import lp_metrics
    rs = np.random.RandomState(11)
    sizes = [25000] * 16
    n = sum(sizes)
    rl = torch.tensor(_ranges(sizes), device="cuda")
    pos = torch.empty(n, device="cuda")
    neg = torch.empty(n, device="cuda")
    out = torch.empty(3, 16, dtype=torch.float64, device="cuda")
    lp_metrics(pos, neg, rl, out=out)                         # validates + caches the ranges outside capture
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            lp_metrics(pos, neg, rl, out=out)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(2):
        p = (1 / (1 + np.exp(-(rs.randn(n) + 1)))).astype(np.float32)
        q = (1 / (1 + np.exp(-(rs.randn(n) - 1)))).astype(np.float32)
        pos.copy_(torch.tensor(p))
        neg.copy_(torch.tensor(q))
        g.replay()
        np.testing.assert_allclose(out.cpu().numpy(), om.lp_record(p, q, _ranges(sizes)), rtol=1e-12)


@pytest.mark.gpu
def test_utils_signatures_on_cuda_tensors():
    from gripnet_b200 import utils
    rs = np.random.RandomState(3)
    y = (rs.rand(500) > 0.4).astype(np.float32)
    s = rs.rand(500).astype(np.float32)
    got = utils.auprc_auroc_ap(torch.tensor(y, device="cuda"), torch.tensor(s, device="cuda"))
    np.testing.assert_allclose(got, _sk_auprc_auroc_ap(y, s), rtol=1e-12)
    t, p = rs.randint(0, 6, 2000), rs.randint(0, 6, 2000)
    tt, pp = torch.tensor(t, device="cuda"), torch.tensor(p, device="cuda")
    micro, macro = utils.micro_macro(tt, pp)
    assert micro == pytest.approx(skm.f1_score(t, p, average="micro"), rel=1e-12)
    assert macro == pytest.approx(skm.f1_score(t, p, average="macro"), rel=1e-12)
    assert utils.acc(tt, pp) == pytest.approx(skm.accuracy_score(t, p), rel=1e-12)


@pytest.mark.gpu
def test_device_nc_metrics_and_argmax():
    from gripnet_b200.metrics import argmax_rows, nc_metrics
    rs = np.random.RandomState(5)
    score = rs.rand(3000, 8).astype(np.float32)
    score[10] = 0.5                                   # an all-tie row: first index wins
    score[11, 3] = score[11, 6] = 2.0
    t = rs.randint(0, 8, 3000)
    t[t == 6] = 5
    ds = torch.tensor(score, device="cuda")
    pred = argmax_rows(ds)
    assert torch.equal(pred.cpu(), torch.tensor(score.argmax(1)))
    wide = torch.zeros(3000, 12, device="cuda")       # a column slice of a wider buffer (row stride 12)
    wide[:, 2:10] = ds
    assert torch.equal(argmax_rows(wide[:, 2:10]), pred)
    got = nc_metrics(torch.tensor(t, device="cuda"), pred, 8).cpu().numpy()
    np.testing.assert_allclose(got, om.micro_macro_acc(t, score.argmax(1)), rtol=1e-12)
    bad = nc_metrics(torch.tensor(t, device="cuda") + 8, pred, 8).cpu().numpy()
    assert np.isnan(bad).all()


# ------------------------------------------------------------------------------------------------
# property tests of the oracle against scikit-learn (CPU)
# ------------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.booleans(), st.integers(0, 12)), min_size=2, max_size=60))
def test_oracle_equals_sklearn_on_arbitrary_tie_patterns(items):
    """Scores drawn from 13 values force every kind of tie group (all-positive, all-negative, mixed, first, last)."""
    y = np.array([1.0 if a else 0.0 for a, _ in items])
    if y.min() == y.max():
        return                                         # sklearn refuses single-class problems
    s = np.array([b / 12.0 for _, b in items], dtype=np.float32)
    np.testing.assert_allclose(om.auprc_auroc_ap(y, s), _sk_auprc_auroc_ap(y, s), rtol=1e-12, atol=1e-14)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 5), st.integers(0, 5)), min_size=1, max_size=80))
def test_oracle_f1_equals_sklearn_on_arbitrary_label_sets(pairs):
    t = np.array([a for a, _ in pairs])
    p = np.array([b for _, b in pairs])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = (skm.f1_score(t, p, average="micro"), skm.f1_score(t, p, average="macro"), skm.accuracy_score(t, p))
    np.testing.assert_allclose(om.micro_macro_acc(t, p), want, rtol=1e-12, atol=1e-15)
