"""Seeded synthetic supergraphs shaped like the reference's datasets.

The real datasets are not in the reference checkout (``/root/reference/.gitignore:133-141``)
and there is no network, so benchmarks and tests run on synthetic graphs whose shapes
follow SURVEY.md §8d (pose-0: ``n_g=19081, E_gg=1431224, n_d=645, E_gd=18596, R=16,
E_dd=400000``).  Generated with ``numpy.random.RandomState`` so one seed gives the same
tensors on every machine.  Tensor conventions are the reference's: ``edge_index`` int64
``[2,E]`` (row 0 = source, row 1 = target), ``range_list`` int64 ``[R,2]`` half-open.
"""
import numpy as np
import torch


def _mirror(pairs):
    """``to_bidirection`` (gripnet/utils.py:132-138): cat([e, flipped e], dim=1)."""
    return np.concatenate([pairs, pairs[::-1]], axis=1)


def _neg_pairs(rs, pos, n_nodes, count):
    """Uniform (src,dst) pairs that are not positive edges (semantics of
    gripnet/utils.py:98-112; not its RNG stream)."""
    taken = np.unique(pos[0].astype(np.int64) * n_nodes + pos[1])
    out = rs.randint(0, n_nodes * n_nodes, size=count).astype(np.int64)
    bad = np.isin(out, taken)
    while bad.any():
        out[bad] = rs.randint(0, n_nodes * n_nodes, size=int(bad.sum()))
        bad = np.isin(out, taken)
    return np.stack([out // n_nodes, out % n_nodes])


def pose_graph(n_g=19081, gg_pairs=715612, n_d=645, e_gd=18596, n_rel=16, dd_pairs_per_rel=12500,
               seed=1111, weighted=False, rel_sizes=None):
    """pose-shaped supergraph (config 1; config 4 via ``n_rel`` / ``rel_sizes``)."""
    rs = np.random.RandomState(seed)
    gg = _mirror(rs.randint(0, n_g, size=(2, gg_pairs)).astype(np.int64))
    gd = np.stack([rs.randint(0, n_g, size=e_gd), rs.randint(0, n_d, size=e_gd)]).astype(np.int64)
    if rel_sizes is None:
        rel_sizes = [dd_pairs_per_rel] * n_rel
    chunks, ranges, start = [], [], 0
    for k in rel_sizes:
        e = _mirror(rs.randint(0, n_d, size=(2, int(k))).astype(np.int64))
        chunks.append(e)
        ranges.append((start, start + e.shape[1]))   # get_range_list, gripnet/utils.py:141-148
        start += e.shape[1]
    dd = np.concatenate(chunks, axis=1)
    et = np.concatenate([np.full(b - a, r, dtype=np.int64) for r, (a, b) in enumerate(ranges)])
    neg = _neg_pairs(rs, dd, n_d, dd.shape[1])
    g = {
        "n_g": n_g, "n_d": n_d, "n_rel": len(rel_sizes),
        "gg_edge_index": torch.from_numpy(gg), "gd_edge_index": torch.from_numpy(gd),
        "dd_edge_index": torch.from_numpy(dd), "dd_edge_type": torch.from_numpy(et),
        "dd_range_list": torch.tensor(ranges, dtype=torch.int64),
        "neg_edge_index": torch.from_numpy(neg),
    }
    if weighted:
        g["gg_edge_weight"] = torch.from_numpy(rs.uniform(0.5, 1.5, gg.shape[1]).astype(np.float32))
    return g


def pose_edges_per_epoch(g, gg_layers=2, dd_layers=1):
    """E_epoch of SURVEY.md §8d: L_gg*E_gg + E_gd + L_dd*E_dd + 2*E_dd."""
    e_dd = g["dd_edge_index"].shape[1]
    return (gg_layers * g["gg_edge_index"].shape[1] + g["gd_edge_index"].shape[1]
            + dd_layers * e_dd + 2 * e_dd)


def nc_graph(n_p, e_pp, n_a, e_pa, e_aa, n_class=8, train_frac=0.2, seed=1111, n_q=None, e_qq=None,
             e_qa=None):
    """aminer-shaped (config 2) or, with ``n_q``, freebase-d-shaped (config 3) NC supergraph."""
    rs = np.random.RandomState(seed)

    def homo(n, e):
        return torch.from_numpy(_mirror(rs.randint(0, n, size=(2, e // 2)).astype(np.int64)))

    def bip(ns, nt, e):
        return torch.from_numpy(np.stack([rs.randint(0, ns, size=e), rs.randint(0, nt, size=e)]).astype(np.int64))

    g = {"n_p": n_p, "n_a": n_a, "n_class": n_class,
         "pp_edge_index": homo(n_p, e_pp), "pa_edge_index": bip(n_p, n_a, e_pa),
         "aa_edge_index": homo(n_a, e_aa)}
    if n_q is not None:
        g.update({"n_q": n_q, "qq_edge_index": homo(n_q, e_qq), "qa_edge_index": bip(n_q, n_a, e_qa)})
    n_train = max(1, int(n_a * train_frac))
    g["train_node_idx"] = torch.from_numpy(np.sort(rs.permutation(n_a)[:n_train]).astype(np.int64))
    g["train_node_class"] = torch.from_numpy(rs.randint(0, n_class, size=n_train).astype(np.int64))
    return g


# presets -------------------------------------------------------------------
def pose_small(seed=1111, weighted=False):
    return pose_graph(n_g=300, gg_pairs=1500, n_d=40, e_gd=200, n_rel=5, dd_pairs_per_rel=60,
                      seed=seed, weighted=weighted)


def pose_medium(seed=1111):
    return pose_graph(n_g=4000, gg_pairs=60000, n_d=200, e_gd=3000, n_rel=8, dd_pairs_per_rel=2500, seed=seed)


def pose2_rel_sizes(n_rel=1097, total_pairs=4_150_000, min_pairs=450, seed=1111):
    """Power-law relation sizes for the pose-2-shaped config 4 (E_dd ~ 8.3 M directed)."""
    rs = np.random.RandomState(seed)
    w = 1.0 / np.arange(1, n_rel + 1) ** 0.8
    rs.shuffle(w)
    sizes = np.maximum(min_pairs, (w / w.sum() * total_pairs).astype(np.int64))
    return sizes.tolist()


def aminer_small(seed=1111):
    return nc_graph(n_p=500, e_pp=5000, n_a=300, e_pa=1500, e_aa=3000, n_class=5, seed=seed)


def aminer_full(seed=1111):
    return nc_graph(n_p=200_000, e_pp=2_000_000, n_a=150_000, e_pa=600_000, e_aa=1_500_000, n_class=8, seed=seed)


def freebase_d_small(seed=1111):
    return nc_graph(n_p=400, e_pp=4000, n_a=250, e_pa=1200, e_aa=2000, n_class=4, seed=seed,
                    n_q=350, e_qq=3000, e_qa=1000)


def freebase_d_full(seed=1111):
    return nc_graph(n_p=300_000, e_pp=3_000_000, n_a=100_000, e_pa=1_000_000, e_aa=1_000_000, n_class=8,
                    seed=seed, n_q=300_000, e_qq=3_000_000, e_qa=1_000_000)
