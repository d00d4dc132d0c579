"""Independent float64 dense-matrix restatement of the hot path (numpy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Small graphs only
(builds N x N matrices).  It shares no code with ``oracle/port.py``: where the
port gathers/scatters per edge like PyG, this builds adjacency matrices, so an
error in one formulation does not silently carry into the other.  Used to
arbitrate fp32 tolerance questions (error of the CUDA path vs float64 must not
exceed the reference's own fp32 error by more than a small factor).

Reference lines restated: gripnet/layers.py:52-69 (normalisation), :71-100
(GCN), :165-197 (RGCN, joint mean), :252-318 (homoGraph), :362-387
(interGraph), gripnet/decoder.py:19-23, :38-45.
"""
import numpy as np


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def gcn_adj(edge_index, n, edge_weight=None, improved=False):
    """A_hat[t, s] = D^-1/2 (A + I) D^-1/2 with D = weighted in-degree by target.

    Duplicated edges add up; an existing self-loop replaces the added loop's
    weight (the last one listed wins), matching layers.py:59-62.
    """
    ei = _np(edge_index).astype(np.int64)
    w = np.ones(ei.shape[1]) if edge_weight is None else _np(edge_weight).astype(np.float64)
    a = np.zeros((n, n))
    loop = np.full(n, 2.0 if improved else 1.0)
    for k in range(ei.shape[1]):
        s, t = ei[0, k], ei[1, k]
        if s == t:
            loop[s] = w[k]
        else:
            a[t, s] += w[k]
    a[np.arange(n), np.arange(n)] += loop
    deg = a.sum(axis=1)
    with np.errstate(divide="ignore"):
        dis = np.where(deg > 0, deg ** -0.5, 0.0)
    return dis[:, None] * a * dis[None, :]


def gcn(x, weight, bias, a_hat):
    out = a_hat @ (_np(x).astype(np.float64) @ _np(weight).astype(np.float64))
    return out if bias is None else out + _np(bias).astype(np.float64)


def rgcn(x, basis, att, root, bias, edge_index, range_list):
    x = _np(x).astype(np.float64)
    basis, att, root = (_np(v).astype(np.float64) for v in (basis, att, root))
    n = x.shape[0]
    ei = _np(edge_index).astype(np.int64)
    w = np.einsum("rb,bio->rio", att, basis)
    acc = np.zeros((n, root.shape[1]))
    cnt = np.zeros(n)
    for r, (s, e) in enumerate(_np(range_list).tolist()):
        a_r = np.zeros((n, n))
        np.add.at(a_r, (ei[1, s:e], ei[0, s:e]), 1.0)
        acc += a_r @ (x @ w[r])
        cnt += a_r.sum(axis=1)
    out = acc / np.maximum(cnt, 1.0)[:, None] + x @ root
    return out if bias is None else out + _np(bias).astype(np.float64)


def homo(p, x, edge_index, edge_weight=None, range_list=None, if_catout=False, multi_relational=False):
    if "embedding" in p:
        x = p["embedding"]
    x = _np(x).astype(np.float64)
    outs = [x]
    a_hat = None
    l = 0
    while f"conv_list.{l}.weight" in p or f"conv_list.{l}.basis" in p:
        k = f"conv_list.{l}."
        if multi_relational:
            x = rgcn(x, p[k + "basis"], p[k + "att"], p[k + "root"], p.get(k + "bias"), edge_index, range_list)
        else:
            if a_hat is None:
                a_hat = gcn_adj(edge_index, x.shape[0], edge_weight)
            x = gcn(x, p[k + "weight"], p.get(k + "bias"), a_hat)
        x = np.maximum(x, 0.0)
        outs.append(x)
        l += 1
    return np.concatenate(outs, axis=1) if if_catout else x


def inter(p, x, inter_edge_index, n_target, edge_weight=None, if_relu=True, mod="cat", if_one_external=True):
    """Closed form of interGraph: h_t = (1 + sum_w_t)^-1/2 * sum_{s->t} w (x_s W) + b."""
    x = _np(x).astype(np.float64)
    ei = _np(inter_edge_index).astype(np.int64)
    w = np.ones(ei.shape[1]) if edge_weight is None else _np(edge_weight).astype(np.float64)
    b = np.zeros((n_target, x.shape[0]))
    np.add.at(b, (ei[1], ei[0]), w)
    scale = (1.0 + b.sum(axis=1)) ** -0.5
    h = scale[:, None] * (b @ (x @ _np(p["conv.weight"]).astype(np.float64)))
    if p.get("conv.bias") is not None:
        h = h + _np(p["conv.bias"]).astype(np.float64)
    if if_relu:
        h = np.maximum(h, 0.0)
    if not if_one_external:
        return h
    tf = _np(p["target_feat"]).astype(np.float64)
    if mod == "cat":
        return np.concatenate([h, np.abs(tf)], axis=1)
    if h.shape[1] == tf.shape[1]:
        return (h + np.abs(tf)) / 2
    return (h + np.maximum(tf @ _np(p["target_feat_down"]).astype(np.float64), 0.0)) / 2


def distmult(z, weight, edge_index, edge_type, sigmoid=True):
    z = _np(z).astype(np.float64)
    w = _np(weight).astype(np.float64)
    ei = _np(edge_index)
    s = np.einsum("ek,ek,ek->e", z[ei[0]], z[ei[1]], w[_np(edge_type)])
    return 1.0 / (1.0 + np.exp(-s)) if sigmoid else s


def multiclass(z, weight, node_list, softmax=True):
    logits = _np(z).astype(np.float64)[_np(node_list)] @ _np(weight).astype(np.float64)
    if not softmax:
        return logits
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)
