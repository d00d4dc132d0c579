"""CPU restatement of the GripNet hot path (torch-CPU, functional style).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the
CUDA path and the timed "port" CPU baseline of ``bench.py``.  Never imported by
``gripnet_b200``.

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  Parameters are passed as plain dicts that use the
reference's ``state_dict`` key names, so one ``state_dict``
drives the reference modules, this port and the CUDA modules alike.

The execution strategy deliberately mirrors what the reference runs through
PyG 1.x on CPU — ``index_select`` gather → per-edge message → ``index_add_``
scatter, differentiated by autograd — because this module is also the CPU
baseline; the independent formulation lives in ``oracle/dense64.py``.

Parity pin: checked against ``tests/golden/*.npz`` (outputs of the unmodified
reference run over ``oracle/pyg_shim.py`` in the build container; generator
``tests/golden/make_golden.py``).
"""
import numpy as np
import torch

EPS = 1e-13  # gripnet/utils.py:10


# --------------------------------------------------------------------------
# a1  graph preprocessing            gripnet/layers.py:52-69  (+ PyG add_remaining_self_loops)
# --------------------------------------------------------------------------
def add_self_loops_keep_existing(edge_index, edge_weight, fill_value, num_nodes):
    """Self-loop rewrite used by ``myGCN.norm`` (layers.py:59-62).

    Existing (i,i) edges are removed, the remaining edges keep their order, and
    one loop per node is appended at the tail in node order.  A removed loop
    donates its weight to the appended loop of its node (last one wins);
    other nodes get ``fill_value``.
    """
    src, dst = edge_index[0], edge_index[1]
    not_loop = src != dst
    ids = torch.arange(num_nodes, dtype=edge_index.dtype)
    out_index = torch.cat([edge_index[:, not_loop], torch.stack([ids, ids])], dim=1)
    loop_w = torch.full((num_nodes,), float(fill_value), dtype=edge_weight.dtype)
    is_loop = ~not_loop
    if bool(is_loop.any()):
        loop_nodes = src[is_loop].numpy()
        loop_vals = edge_weight[is_loop].numpy()
        lw = loop_w.numpy()
        lw[loop_nodes] = loop_vals  # numpy fancy assignment: last occurrence wins
    return out_index, torch.cat([edge_weight[not_loop], loop_w])


def gcn_norm(edge_index, num_nodes, edge_weight=None, improved=False, dtype=torch.float32):
    """``myGCN.norm`` (layers.py:52-69): returns (edge_index', norm).

    deg is the weighted in-degree by TARGET (``col``) including the loop;
    ``norm_e = (deg^-1/2[row_e] * w_e) * deg^-1/2[col_e]`` with inf -> 0.
    """
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype)
    fill = 2 if improved else 1
    edge_index, edge_weight = add_self_loops_keep_existing(edge_index, edge_weight, fill, num_nodes)
    src, dst = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=edge_weight.dtype).index_add_(0, dst, edge_weight)
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    return edge_index, dis[src] * edge_weight * dis[dst]


# --------------------------------------------------------------------------
# a2  GCN conv                       gripnet/layers.py:71-100
# --------------------------------------------------------------------------
def gcn_conv(x, weight, bias, edge_index_aug, norm):
    """``myGCN.forward/message/update``: ``out_i = sum_e norm_e (xW)[row_e] + b``."""
    y = x @ weight                                              # :73
    msg = norm.view(-1, 1) * y.index_select(0, edge_index_aug[0])  # :94-95
    out = torch.zeros_like(y).index_add_(0, edge_index_aug[1], msg)  # propagate(aggr="add") :92
    return out + bias if bias is not None else out             # :97-100


# --------------------------------------------------------------------------
# a5  RGCN conv                      gripnet/layers.py:165-197
# --------------------------------------------------------------------------
def rgcn_conv(x, basis, att, root, bias, edge_index, range_list):
    """``myRGCN``: basis-decomposed relation weights, ONE joint mean over all in-edges.

    The relation of edge ``e`` is defined only by which ``range_list`` slice
    contains ``e`` (layers.py:178-186; ``edge_type`` is ignored there).
    """
    n_rel, n_base = att.shape
    in_c, out_c = root.shape
    w = (att @ basis.reshape(n_base, in_c * out_c)).reshape(n_rel, in_c, out_c)  # :172-173
    x_j = x.index_select(0, edge_index[0])
    ranges = range_list.tolist()
    ws = w.unbind(0)
    if ranges and ranges[0][0] == 0 and ranges[-1][1] == x_j.size(0) and \
            all(ranges[i][1] == ranges[i + 1][0] for i in range(len(ranges) - 1)):
        # the same slices as below; split/unbind keep autograd from materialising one full-size zero
        # tensor per relation in backward (R ~ 10^3 at config 4)
        xs = torch.split(x_j, [e - s for s, e in ranges])
        pieces = [xs[r] @ ws[r] for r in range(len(ranges))]
    else:
        pieces = [x_j[int(s):int(e)] @ ws[r] for r, (s, e) in enumerate(ranges)]        # :178-186
    msg = torch.cat(pieces)                                                      # :189
    dst = edge_index[1]
    agg = torch.zeros(x.size(0), out_c, dtype=x.dtype).index_add_(0, dst, msg)
    cnt = torch.zeros(x.size(0), dtype=x.dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=x.dtype))
    agg = agg / cnt.clamp(min=1).unsqueeze(1)                                    # aggr="mean" :131
    out = agg + x @ root                                                         # :193
    return out + bias if bias is not None else out                              # :195-196


def _relu(x, on=None):
    """``relu`` -- or, when a test pins the activation pattern (``on``: bool tensor, True where the
    unit is active), the same piecewise-linear branch chosen by that pattern.  ReLU makes the
    gradient discontinuous in the pre-activations: at full size a few of ~10^7 pre-activations
    sit within fp32 rounding of zero, so two correct fp32 implementations (or fp32 vs fp64) pick
    different branches there and their gradients differ by O(1e-3).  Pinning the branch to the one
    the implementation under test took makes the backward comparison exact; the test checks
    separately that the patterns only disagree where the pre-activation is ~0."""
    if on is None:
        return torch.relu(x)
    return x * on.to(x.dtype)


# --------------------------------------------------------------------------
# a3 / a6  homoGraph                 gripnet/layers.py:252-318
# --------------------------------------------------------------------------
def homo_forward(p, x, edge_index, edge_weight=None, edge_type=None, range_list=None,
                 if_catout=False, multi_relational=False, norm_cache=None, prefix="", pattern=None):
    """``homoGraph.forward``.  ``p`` uses keys ``embedding`` / ``conv_list.{i}.*``.

    ``norm_cache`` (a dict) plays the role of ``myGCN.cached_result``
    (layers.py:83-90): per layer index -> (edge_index', norm).
    ReLU follows EVERY layer, the last one included (:279, :305).
    ``pattern`` (tests only, see ``_relu``): this module's output as computed by the implementation
    under test; its sign pattern replaces the ReLU branch decisions.
    """
    if prefix + "embedding" in p:                                   # start_graph, :261-262
        x = p[prefix + "embedding"]
    n_layers = 0
    while (prefix + f"conv_list.{n_layers}.weight" in p) or (prefix + f"conv_list.{n_layers}.basis" in p):
        n_layers += 1
    outs = [x]
    for i in range(n_layers):
        k = prefix + f"conv_list.{i}."
        if multi_relational:
            x = rgcn_conv(x, p[k + "basis"], p[k + "att"], p[k + "root"], p.get(k + "bias"),
                          edge_index, range_list)
        else:
            if norm_cache is not None and i in norm_cache:
                ei, nrm = norm_cache[i]
            else:
                ei, nrm = gcn_norm(edge_index, x.size(0), edge_weight, False, x.dtype)
                if norm_cache is not None:
                    norm_cache[i] = (ei, nrm)
            x = gcn_conv(x, p[k + "weight"], p.get(k + "bias"), ei, nrm)
        on = None
        if pattern is not None:
            if if_catout:
                off = sum(o.shape[1] for o in outs)
                on = pattern[:, off:off + x.shape[1]] > 0
            elif i == n_layers - 1:
                on = pattern > 0
            else:
                raise ValueError("pinned pattern of a hidden layer needs if_catout")
        x = _relu(x, on)
        outs.append(x)
    return torch.cat(outs, dim=1) if if_catout else x               # :307-309


# --------------------------------------------------------------------------
# a4  interGraph                     gripnet/layers.py:362-387
# --------------------------------------------------------------------------
def inter_forward(p, x, inter_edge_index, n_target, edge_weight=None, if_relu=True, mod="cat",
                  if_one_external=True, norm_cache=None, prefix="", pattern=None):
    """``interGraph.forward``: bipartite parent -> child propagation.

    Runs the GCN over the stacked (n_source + n_target)-node graph exactly as
    the reference does (:363-368), then the tail (:369-384).
    """
    n_source = x.shape[0]
    ei = inter_edge_index.clone()
    ei[1] += n_source                                               # :364-365
    xs = torch.cat([x, torch.zeros(n_target, x.shape[1], dtype=x.dtype)], dim=0)  # :367
    if norm_cache is not None and "inter" in norm_cache:
        ei_aug, nrm = norm_cache["inter"]
    else:
        ei_aug, nrm = gcn_norm(ei, n_source + n_target, edge_weight, False, x.dtype)
        if norm_cache is not None:
            norm_cache["inter"] = (ei_aug, nrm)
    h = gcn_conv(xs, p[prefix + "conv.weight"], p.get(prefix + "conv.bias"), ei_aug, nrm)[n_source:]  # :368
    if if_relu:
        on = None
        if pattern is not None:
            if if_one_external and mod != "cat":
                raise ValueError("pinned pattern needs h to be visible in the output")
            on = pattern[:, :h.shape[1]] > 0
        h = _relu(h, on)                                            # :369-370
    if not if_one_external:
        return h                                                    # :372-373
    tf = p[prefix + "target_feat"]
    if mod == "cat":
        return torch.cat([h, tf.abs()], dim=1)                      # :375-376
    if h.shape[1] == tf.shape[1]:
        return (h + tf.abs()) / 2                                   # :378-379
    return (h + torch.relu(tf @ p[prefix + "target_feat_down"])) / 2  # :382-384


# --------------------------------------------------------------------------
# a7  DistMult decoder               gripnet/decoder.py:19-23
# --------------------------------------------------------------------------
def distmult(z, weight, edge_index, edge_type, sigmoid=True):
    s = (z.index_select(0, edge_index[0]) * z.index_select(0, edge_index[1])
         * weight.index_select(0, edge_type)).sum(dim=1)
    return torch.sigmoid(s) if sigmoid else s


# --------------------------------------------------------------------------
# a9  multi-class decoder            gripnet/decoder.py:38-45
# --------------------------------------------------------------------------
def multiclass(z, weight, node_list, softmax=True):
    logits = z.index_select(0, node_list) @ weight
    return torch.softmax(logits, dim=1) if softmax else logits


# --------------------------------------------------------------------------
# a8 / a9 losses                     GripNet-pose.py:140-142, GripNet-aminer.py:133
# --------------------------------------------------------------------------
def lp_loss(pos_score, neg_score):
    return -torch.log(pos_score + EPS).mean() - torch.log(1 - neg_score + EPS).mean()


def nc_loss(score, labels):
    return -torch.log(score[torch.arange(score.shape[0]), labels] + EPS).mean()


# --------------------------------------------------------------------------
# model wiring                       GripNet-pose.py:94-99,117-142; GripNet-aminer.py:103-133;
#                                    GripNet-freebase-d.py:103-137,151-166
# --------------------------------------------------------------------------
def _sub(p, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def pose_forward(p, g, cache=None, patterns=None):
    """One pose-shaped forward: gg -> gd -> dd -> DistMult(pos, neg) -> loss.

    ``p``: flat dict with prefixes ``gg.``, ``gd.``, ``dd.``, ``dmt.``;
    ``g``: dict from ``oracle.synth.pose_graph``.  Returns (loss, z, pos, neg).
    """
    cache = {} if cache is None else cache
    pat = patterns or {}
    z = homo_forward(_sub(p, "gg."), None, g["gg_edge_index"], g.get("gg_edge_weight"), if_catout=True,
                     norm_cache=cache.setdefault("gg", {}), pattern=pat.get("gg"))
    z = inter_forward(_sub(p, "gd."), z, g["gd_edge_index"], g["n_d"], mod="cat", if_relu=True,
                      norm_cache=cache.setdefault("gd", {}), pattern=pat.get("gd"))
    z = homo_forward(_sub(p, "dd."), z, g["dd_edge_index"], edge_type=g["dd_edge_type"],
                     range_list=g["dd_range_list"], if_catout=True, multi_relational=True, pattern=pat.get("dd"))
    w = p["dmt.weight"]
    pos = distmult(z, w, g["dd_edge_index"], g["dd_edge_type"])
    neg = distmult(z, w, g["neg_edge_index"], g["dd_edge_type"])
    return lp_loss(pos, neg), z, pos, neg


def aminer_forward(p, g, cache=None, patterns=None):
    """aminer-shaped NC forward: pp -> pa -> aa -> softmax decoder -> loss."""
    cache = {} if cache is None else cache
    pat = patterns or {}
    z = homo_forward(_sub(p, "pp."), None, g["pp_edge_index"], g.get("pp_edge_weight"), if_catout=True,
                     norm_cache=cache.setdefault("pp", {}), pattern=pat.get("pp"))
    z = inter_forward(_sub(p, "pa."), z, g["pa_edge_index"], g["n_a"], mod="cat", if_relu=True,
                      norm_cache=cache.setdefault("pa", {}), pattern=pat.get("pa"))
    z = homo_forward(_sub(p, "aa."), z, g["aa_edge_index"], g.get("aa_edge_weight"), if_catout=True,
                     norm_cache=cache.setdefault("aa", {}), pattern=pat.get("aa"))
    score = multiclass(z, p["mcip.weight"], g["train_node_idx"])
    return nc_loss(score, g["train_node_class"]), z, score


def freebase_d_forward(p, g, cache=None, patterns=None):
    """freebase-d-shaped NC forward: (pp->pa) + (qq->qa) + learned emb, mean, aa, decoder."""
    cache = {} if cache is None else cache
    pat = patterns or {}
    z = homo_forward(_sub(p, "pp."), None, g["pp_edge_index"], g.get("pp_edge_weight"), if_catout=True,
                     norm_cache=cache.setdefault("pp", {}), pattern=pat.get("pp"))
    z = inter_forward(_sub(p, "pa."), z, g["pa_edge_index"], g["n_a"], mod="add", if_relu=True,
                      if_one_external=False, norm_cache=cache.setdefault("pa", {}), pattern=pat.get("pa"))
    z1 = homo_forward(_sub(p, "qq."), None, g["qq_edge_index"], g.get("qq_edge_weight"), if_catout=True,
                      norm_cache=cache.setdefault("qq", {}), pattern=pat.get("qq"))
    z1 = inter_forward(_sub(p, "qa."), z1, g["qa_edge_index"], g["n_a"], mod="add", if_relu=True,
                       if_one_external=False, norm_cache=cache.setdefault("qa", {}), pattern=pat.get("qa"))
    z = homo_forward(_sub(p, "aa."), (z + z1 + p["aa_embeddings"]) / 3, g["aa_edge_index"],
                     g.get("aa_edge_weight"), norm_cache=cache.setdefault("aa", {}), pattern=pat.get("aa"))
    score = multiclass(z, p["mcip.weight"], g["train_node_idx"])
    return nc_loss(score, g["train_node_class"]), z, score


# --------------------------------------------------------------------------
# bit-exact integer oracles for the CSR contract (numpy)
# --------------------------------------------------------------------------
def csr_from_edges(dst, n_rows):
    """Stable counting sort by ``dst``: returns (rowptr int64[n+1], perm int64[E]).

    ``perm[k]`` is the original position of the edge stored in CSR slot ``k``;
    within a row the original relative order is preserved.
    """
    dst = np.asarray(dst, dtype=np.int64)
    perm = np.argsort(dst, kind="stable")
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n_rows), out=rowptr[1:])
    return rowptr, perm


def gcn_csr_oracle(edge_index, num_nodes, edge_weight=None, improved=False):
    """The graph-prep contract of SURVEY.md §8 a1, as plain numpy.

    Returns a dict with the augmented edge list (reference order), the
    dst-sorted CSR and the src-sorted (transpose) CSR over that list, integer
    in-degree (edge count per target incl. the loop) and fp32 ``deg``/``norm``.
    """
    ei = torch.as_tensor(edge_index)
    w = None if edge_weight is None else torch.as_tensor(edge_weight, dtype=torch.float32)
    ei_aug, norm = gcn_norm(ei, num_nodes, w, improved, torch.float32)
    src = ei_aug[0].numpy()
    dst = ei_aug[1].numpy()
    rowptr, perm = csr_from_edges(dst, num_nodes)
    rowptr_t, perm_t = csr_from_edges(src, num_nodes)
    return {
        "edge_index_aug": ei_aug.numpy(),
        "norm": norm.numpy(),
        "rowptr": rowptr, "perm": perm, "col": src[perm], "val": norm.numpy()[perm],
        "rowptr_t": rowptr_t, "perm_t": perm_t, "col_t": dst[perm_t], "val_t": norm.numpy()[perm_t],
        "indeg": np.diff(rowptr),
    }
