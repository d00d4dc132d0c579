"""Seeded synthetic supergraphs shaped like the reference's datasets.

NEUTRAL module: numpy / torch only.  It imports neither ``gripnet_b200`` (so ``bench.py --impl reference``
and ``oracle/`` can build their inputs without loading the CUDA library) nor ``oracle`` (so the product's
own benchmark arm never touches test infrastructure).

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
               seed=1111, weighted=False, rel_sizes=None, pair_pool=None):
    """pose-shaped supergraph (config 1; config 4 via ``n_rel`` / ``rel_sizes`` / ``pair_pool``).

    ``pair_pool``: every relation draws its drug pairs from one fixed pool of that many random
    pairs instead of from all ``n_d**2`` (the pose datasets hold ~63 k interacting drug pairs and
    each side effect occurs on a subset of them).  Needed when ``E_dd >> n_d**2``: uniform pairs
    would make every pair a positive and leave no negative to sample."""
    rs = np.random.RandomState(seed)
    gg = _mirror(rs.randint(0, n_g, size=(2, gg_pairs)).astype(np.int64))
    gd = np.stack([rs.randint(0, n_g, size=e_gd), rs.randint(0, n_d, size=e_gd)]).astype(np.int64)
    if rel_sizes is None:
        rel_sizes = [dd_pairs_per_rel] * n_rel
    chunks, ranges, start = [], [], 0
    pool = None if pair_pool is None else rs.randint(0, n_d, size=(2, int(pair_pool))).astype(np.int64)
    for k in rel_sizes:
        if pool is None:
            e = _mirror(rs.randint(0, n_d, size=(2, int(k))).astype(np.int64))
        else:
            e = _mirror(pool[:, rs.randint(0, pool.shape[1], size=int(k))])
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


def pose_graph_scaled(scale, seed=1111):
    """pose-0-shaped supergraph with ``scale`` times the nodes and edges (weak scaling over ``scale``
    GPUs: per-GPU work stays that of pose-0).  ``scale == 1`` is exactly ``pose_graph()``."""
    if scale == 1:
        return pose_graph(seed=seed)
    return pose_graph(n_g=19081 * scale, gg_pairs=715612 * scale, n_d=645 * scale, e_gd=18596 * scale, n_rel=16,
                      dd_pairs_per_rel=12500 * scale, seed=seed)


# ---------------------------------------------------------------------------
# config 5: scaled three-supervertex chain with power-law (R-MAT) degrees, generated on the device
# ---------------------------------------------------------------------------
def rmat_edge_chunks(log2_n, n_edges, device, gen, probs=(0.57, 0.19, 0.19, 0.05), n_nodes=None, chunk=1 << 26):
    """Generator of the R-MAT edge list over ``2**log2_n`` ids in chunks of at most ``chunk`` edges (recursive
    quadrant choice with probabilities a,b,c,d), ids scattered by a fixed multiplicative hash so hubs do not
    cluster in one partition block and folded into ``[0, n_nodes)``.  Yields int64 ``[2, m]`` tensors on
    ``device``; the random stream is consumed chunk by chunk, so streaming and materialising
    (``rmat_edges``) give the same edges."""
    a, b, c, _ = probs
    n = 1 << log2_n
    n_nodes = n if n_nodes is None else n_nodes
    for s in range(0, n_edges, chunk):
        m = min(chunk, n_edges - s)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(log2_n):
            r = torch.rand(m, device=device, generator=gen)
            src = src * 2 + (r >= a + b).to(torch.int64)                 # quadrants c, d -> lower half (src bit 1)
            dst = dst * 2 + (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)   # b, d -> dst bit 1
        # scatter ids: odd multiplier mod 2^k is a bijection
        src = (src * 0x9E3779B1 + 0x7F4A7C15) & (n - 1)
        dst = (dst * 0x85EBCA6B + 0x2545F491) & (n - 1)
        yield torch.stack([src % n_nodes, dst % n_nodes])


def rmat_edges(log2_n, n_edges, device, gen, probs=(0.57, 0.19, 0.19, 0.05), n_nodes=None, chunk=1 << 26):
    """The whole R-MAT edge list, int64 ``[2, n_edges]`` on ``device``."""
    out = torch.empty((2, n_edges), dtype=torch.int64, device=device)
    s = 0
    for part in rmat_edge_chunks(log2_n, n_edges, device, gen, probs, n_nodes, chunk):
        out[:, s:s + part.size(1)] = part
        s += part.size(1)
    return out


def _log2_ceil(n):
    k = 1
    while (1 << k) < n:
        k += 1
    return k


def chain_graph(n_a, n_b, n_c, e_aa, e_ab, e_bb, e_bc, e_cc, device, n_class=8, train_frac=0.2, seed=1111,
                homo_sink=None, bip_sink=None):
    """Config 5: supervertex chain A -> B -> C, intra graphs R-MAT and mirrored
    (undirected), inter graphs R-MAT sources x uniform targets; labels on ``train_frac`` of C.

    ``homo_sink(name, n, e_directed, half_chunks)`` / ``bip_sink(name, n_src, n_tgt, edge_index)`` decide what
    is stored under ``g[name]``: by default the materialised global ``edge_index``; a partitioned run passes
    sinks that keep only this rank's shard of each chunk (``pipelines.shard_chain_streamed``), so no rank
    holds a global edge list.  Either way the random stream is consumed identically: every rank, and the
    single-GPU run, see the same global graph."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    def homo(name, n, e):
        chunks = rmat_edge_chunks(_log2_ceil(n), e // 2, device, gen, n_nodes=n)
        if homo_sink is not None:
            return homo_sink(name, n, 2 * (e // 2), chunks)
        half = torch.cat(list(chunks), dim=1)
        return torch.cat([half, half.flip(0)], dim=1)

    def bip(name, ns, nt, e):
        src = rmat_edges(_log2_ceil(ns), e, device, gen, n_nodes=ns)[0]
        dst = torch.randint(0, nt, (e,), device=device, generator=gen)
        ei = torch.stack([src, dst])
        return ei if bip_sink is None else bip_sink(name, ns, nt, ei)

    g = {"n_a": n_a, "n_b": n_b, "n_c": n_c, "n_class": n_class}
    g["aa_edge_index"] = homo("aa_edge_index", n_a, e_aa)
    g["ab_edge_index"] = bip("ab_edge_index", n_a, n_b, e_ab)
    g["bb_edge_index"] = homo("bb_edge_index", n_b, e_bb)
    g["bc_edge_index"] = bip("bc_edge_index", n_b, n_c, e_bc)
    g["cc_edge_index"] = homo("cc_edge_index", n_c, e_cc)
    g["edge_counts"] = {"aa": 2 * (e_aa // 2), "ab": e_ab, "bb": 2 * (e_bb // 2), "bc": e_bc, "cc": 2 * (e_cc // 2)}
    n_train = max(1, int(n_c * train_frac))
    g["train_node_idx"] = torch.randperm(n_c, device=device, generator=gen)[:n_train].sort().values
    g["train_node_class"] = torch.randint(0, n_class, (n_train,), device=device, generator=gen)
    return g


CHAIN_FULL = dict(n_a=4_000_000, n_b=4_000_000, n_c=2_000_000, e_aa=200_000_000, e_ab=50_000_000, e_bb=200_000_000,
                  e_bc=50_000_000, e_cc=20_000_000)
CHAIN_SMALL = dict(n_a=3000, n_b=2500, n_c=1500, e_aa=40_000, e_ab=9_000, e_bb=30_000, e_bc=8_000, e_cc=12_000,
                   n_class=5)


def chain_full(device, seed=1111, **sinks):
    """~10 M nodes, ~520 M directed edges."""
    return chain_graph(device=device, seed=seed, **CHAIN_FULL, **sinks)


def chain_small(device, seed=1111, **sinks):
    return chain_graph(device=device, seed=seed, **CHAIN_SMALL, **sinks)


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


def pose2_graph(seed=1111):
    """BASELINE config 4 (pose-2-shaped): R = 1097 relations with power-law sizes, E_dd ~ 8.3 M directed
    edges over a pool of 63 473 drug pairs."""
    sizes = pose2_rel_sizes(seed=seed)
    return pose_graph(n_rel=len(sizes), rel_sizes=sizes, pair_pool=63473, seed=seed)


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
