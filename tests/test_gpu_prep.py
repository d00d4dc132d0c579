"""GPU: K1 graph preprocessing is BIT-EXACT against the numpy / reference-order oracle.

Integer results (row pointers, column indices, permutations, degrees, the
augmented edge list) must match exactly; fp32 normalisation coefficients within
1e-6 relative (they are products of correctly rounded deg^-1/2 values).
"""
import numpy as np
import pytest
import torch

from golden_util import load_grouped, rel_err
from oracle import port

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _csr_from_keys(keys_np, n_rows):
    import gripnet_b200 as gb
    from gripnet_b200.graph import _ptr, _stream, _ws
    lib = gb._lib.load()
    keys = torch.from_numpy(keys_np.astype(np.int32)).to(_dev())
    n = keys.numel()
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=_dev())
    perm = torch.empty(max(n, 1), dtype=torch.int32, device=_dev())
    ws = _ws(lib.gn_csr_from_keys_workspace_bytes(n, n_rows), _dev())
    gb._lib.check(lib.gn_csr_from_keys(_ptr(keys) if n else None, n, n_rows, _ptr(rowptr), _ptr(perm), _ptr(ws),
                                       ws.numel(), _stream()))
    torch.cuda.synchronize()
    return rowptr.cpu().numpy(), perm[:n].cpu().numpy()


@pytest.mark.parametrize("n,n_rows", [(0, 5), (1, 1), (31, 2), (33, 300), (4096, 7), (4097, 300), (100_003, 70_000),
                                      (1_000_000, 645), (300_000, 20_000_000 // 16)])
def test_stable_sort_matches_numpy(n, n_rows):
    rs = np.random.RandomState(n + n_rows)
    keys = rs.randint(0, n_rows, size=n)
    rowptr, perm = _csr_from_keys(keys, n_rows)
    rp_ref, perm_ref = port.csr_from_edges(keys, n_rows)
    assert np.array_equal(rowptr, rp_ref)
    assert np.array_equal(perm, perm_ref)


def test_sort_skewed_keys():
    """Hub rows: one key holds half of the entries (exercises the per-warp digit counters)."""
    rs = np.random.RandomState(3)
    keys = rs.randint(0, 1000, size=200_000)
    keys[rs.rand(200_000) < 0.5] = 77
    rowptr, perm = _csr_from_keys(keys, 1000)
    rp_ref, perm_ref = port.csr_from_edges(keys, 1000)
    assert np.array_equal(rowptr, rp_ref) and np.array_equal(perm, perm_ref)


def _check_gcn_graph(ei_np, n, w_np=None, improved=False):
    from gripnet_b200.graph import GcnGraph
    ei = torch.from_numpy(np.asarray(ei_np, dtype=np.int64)).to(_dev())
    w = None if w_np is None else torch.from_numpy(w_np).to(_dev())
    g = GcnGraph(ei, n, n, w, improved, bipartite=False, want_aug=True)
    o = port.gcn_csr_oracle(ei_np, n, w_np, improved)
    torch.cuda.synchronize()
    assert g.nnz == o["edge_index_aug"].shape[1]
    assert np.array_equal(g.aug_edge_index.cpu().numpy(), o["edge_index_aug"])
    for mine, ref in ((g.fwd.rowptr, o["rowptr"]), (g.fwd.col[: g.nnz], o["col"]), (g.perm, o["perm"]),
                      (g.bwd.rowptr, o["rowptr_t"]), (g.bwd.col[: g.nnz], o["col_t"]), (g.perm_t, o["perm_t"]),
                      (g.indeg, o["indeg"])):
        assert np.array_equal(mine.cpu().numpy().astype(np.int64), ref.astype(np.int64))
    assert rel_err(g.aug_norm.cpu(), o["norm"]) < 1e-6
    assert rel_err(g.fwd.val[: g.nnz].cpu(), o["val"]) < 1e-6
    assert rel_err(g.bwd.val[: g.nnz].cpu(), o["val_t"]) < 1e-6
    return g, o


@pytest.mark.parametrize("name", ["kat6", "loops_w", "loops_improved", "isolated", "empty"])
def test_gcn_prep_golden(name):
    c = load_grouped("gcn_norm")[name]
    ei = c["edge_index"].numpy()
    w = c["edge_weight"].numpy() if "edge_weight" in c else None
    g, _ = _check_gcn_graph(ei, int(c["num_nodes"]), w, bool(c["improved"]))
    assert np.array_equal(g.aug_edge_index.cpu().numpy(), c["out_edge_index"].numpy())   # vs the reference itself
    assert rel_err(g.aug_norm.cpu(), c["out_norm"]) < 1e-6


def test_gcn_prep_random_large():
    rs = np.random.RandomState(11)
    n, e = 5000, 120_000
    ei = rs.randint(0, n, size=(2, e))
    ei[:, ::97] = ei[0, ::97]                         # self loops
    ei[:, 1000:1100] = ei[:, 2000:2100]               # duplicates
    _check_gcn_graph(ei, n)
    _check_gcn_graph(ei, n, rs.uniform(0.5, 1.5, e).astype(np.float32))
    _check_gcn_graph(ei, n, None, improved=True)


def test_gcn_unit_degree_exact():
    from gripnet_b200.graph import GcnGraph
    rs = np.random.RandomState(5)
    n, e = 3000, 60_000
    ei = rs.randint(0, n, size=(2, e))
    g = GcnGraph(torch.from_numpy(ei).to(_dev()), n, n)
    src, dst = ei
    keep = src != dst
    deg_ref = np.bincount(dst[keep], minlength=n).astype(np.float32) + 1.0
    assert np.array_equal(g.deg.cpu().numpy(), deg_ref)          # integer-valued fp32: bit-exact


def test_bipartite_prep_matches_stacked_reference_graph():
    """interGraph closed form == myGCN.norm over the stacked (n_src+n_tgt)-node graph (layers.py:363-368)."""
    from gripnet_b200.graph import GcnGraph
    rs = np.random.RandomState(2)
    n_s, n_t, e = 700, 90, 5000
    ei = np.stack([rs.randint(0, n_s, e), rs.randint(0, n_t, e)])
    for w in (None, rs.uniform(0.5, 1.5, e).astype(np.float32)):
        g = GcnGraph(torch.from_numpy(ei).to(_dev()), n_s, n_t,
                     None if w is None else torch.from_numpy(w).to(_dev()), bipartite=True)
        stacked = ei.copy()
        stacked[1] += n_s
        ei_aug, nrm = port.gcn_norm(torch.from_numpy(stacked), n_s + n_t,
                                    None if w is None else torch.from_numpy(w))
        nrm = nrm.numpy()[:e]                                     # the real edges come first, in order
        rp, perm = port.csr_from_edges(ei[1], n_t)
        rp_t, perm_t = port.csr_from_edges(ei[0], n_s)
        assert g.nnz == e
        assert np.array_equal(g.fwd.rowptr.cpu().numpy(), rp) and np.array_equal(g.perm.cpu().numpy(), perm)
        assert np.array_equal(g.fwd.col[:e].cpu().numpy(), ei[0][perm])
        assert np.array_equal(g.bwd.rowptr.cpu().numpy(), rp_t) and np.array_equal(g.perm_t.cpu().numpy(), perm_t)
        assert np.array_equal(g.bwd.col[:e].cpu().numpy(), ei[1][perm_t])
        assert rel_err(g.fwd.val[:e].cpu(), nrm[perm]) < 1e-6
        assert rel_err(g.bwd.val[:e].cpu(), nrm[perm_t]) < 1e-6


def test_rgcn_prep():
    from gripnet_b200.graph import RgcnGraph
    rs = np.random.RandomState(4)
    n, sizes = 200, [500, 0, 1200, 37, 0, 900]
    n_rel = len(sizes)
    ei = rs.randint(0, n, size=(2, sum(sizes)))
    bounds = np.cumsum([0] + sizes)
    rl = torch.tensor(np.stack([bounds[:-1], bounds[1:]], 1), dtype=torch.int64)
    rel = np.concatenate([np.full(k, r) for r, k in enumerate(sizes)])
    g = RgcnGraph(torch.from_numpy(ei).to(_dev()), rl.to(_dev()), n, n_rel)
    rp, perm = port.csr_from_edges(ei[1], n)
    e = ei.shape[1]
    assert np.array_equal(g.fwd.rowptr.cpu().numpy(), rp)
    assert np.array_equal(g.perm[:e].cpu().numpy(), perm)
    assert np.array_equal(g.fwd.col[:e].cpu().numpy(), (ei[0] * n_rel + rel)[perm])
    cnt = np.maximum(np.diff(rp), 1).astype(np.float32)
    assert np.array_equal(g.inv_cnt.cpu().numpy(), (1.0 / cnt).astype(np.float32))
    key_t = ei[0] * n_rel + rel
    rp_t, perm_t = port.csr_from_edges(key_t, n * n_rel)
    assert np.array_equal(g.bwd.rowptr.cpu().numpy(), rp_t)
    assert np.array_equal(g.perm_t[:e].cpu().numpy(), perm_t)
    assert np.array_equal(g.bwd.col[:e].cpu().numpy(), ei[1][perm_t])
    assert np.array_equal(g.bwd.val[:e].cpu().numpy(), (1.0 / cnt).astype(np.float32)[ei[1][perm_t]])
    with pytest.raises(RuntimeError):
        RgcnGraph(torch.from_numpy(ei).to(_dev()), torch.tensor([[0, 10], [20, e]]).to(_dev()), n, 2)


@pytest.mark.parametrize("n,r,e", [(120, 7, 4000), (645, 16, 50000), (1024, 3, 9000), (1500, 300, 20000),
                                   (70000, 2, 30000), (5, 1, 0)])
def test_index_prep(n, r, e):
    """Index CSR (relation lists, labelled-node lists), bit-exact vs a stable argsort.  Sizes cover the
    single-pass path (<= 1024 rows, 8- and 10-bit digits) and the multi-pass radix sort."""
    from gripnet_b200.graph import IndexStruct
    rs = np.random.RandomState(9)
    et = rs.randint(0, r, size=e)
    rel = IndexStruct(torch.from_numpy(et).to(_dev()), r)
    rp_r, perm_r = port.csr_from_edges(et, r)
    assert np.array_equal(rel.csr.rowptr.cpu().numpy(), rp_r)
    assert np.array_equal(rel.perm[:e].cpu().numpy(), perm_r)
    m = min(300, e)
    idx = rs.randint(0, n, size=m)
    st = IndexStruct(torch.from_numpy(idx).to(_dev()), n)
    rp_i, perm_i = port.csr_from_edges(idx, n)
    assert np.array_equal(st.csr.rowptr.cpu().numpy(), rp_i) and np.array_equal(st.perm[:m].cpu().numpy(), perm_i)
    assert int(st.csr.row_counter.abs().sum()) == 0


@pytest.mark.parametrize("n_rows", [400, 3000, 20000, 50000])   # one-block builder (<= 32768 rows) and the scan path
@pytest.mark.parametrize("chunk_len", [32, 64, 1024])
def test_chunk_list_covers_every_entry_once(chunk_len, n_rows):
    from gripnet_b200.graph import Csr
    rs = np.random.RandomState(1)
    lens = np.concatenate([rs.randint(0, 50, n_rows), [0, 0, 5000, 33, 32, 31, 1]])
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    nnz = int(rowptr[-1])
    csr = Csr(torch.from_numpy(rowptr).to(_dev()), torch.zeros(nnz, dtype=torch.int32, device=_dev()), None,
              len(lens), 1, nnz, chunk_len=chunk_len)
    cp = csr.chunk_ptr.cpu().numpy()
    crow = csr.chunk_row[: csr.n_chunks].cpu().numpy()
    cbeg = csr.chunk_beg[: csr.n_chunks].cpu().numpy()
    expect = np.maximum(1, -(-lens // chunk_len))
    assert np.array_equal(np.diff(cp), expect) and csr.n_chunks == expect.sum()
    assert np.array_equal(crow, np.repeat(np.arange(len(lens)), expect))
    seen = np.zeros(nnz, dtype=np.int32)
    for c in range(csr.n_chunks):
        end = min(cbeg[c] + chunk_len, rowptr[crow[c] + 1])
        seen[cbeg[c]:end] += 1
    assert (seen == 1).all()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("case", ["square", "square_weighted_improved", "bipartite", "bipartite_weighted"])
def test_partitioned_prep_equals_slices_of_the_global_csr(case, world):
    """Per-rank graph prep of the destination-partitioned path (gn_edge_filter + gn_gcn_part_structure +
    gn_gcn_part_values): every rank's rows, built ONLY from the edges that touch its block plus the all-gathered
    deg^-1/2 vector, are BIT-IDENTICAL to the matching rows of the single-GPU CSR pair (columns, coefficients,
    degrees) — self-loops, duplicate edges, empty rows and an uneven last block included."""
    from gripnet_b200.graph import GcnGraph, filter_edges, part_structure, part_values
    from gripnet_b200.parallel import block_bounds, block_size
    rs = np.random.RandomState(world)
    d = _dev()
    bip = case.startswith("bipartite")
    n_src, n_dst = (1003, 157) if bip else (1003, 1003)
    e = 20_000
    ei = np.stack([rs.randint(0, n_src, e), rs.randint(0, n_dst, e)])
    if not bip:
        ei[1, :300] = ei[0, :300]                 # self-loops (some nodes several times)
        ei[:, 300:600] = ei[:, 600:900]           # duplicate edges
    ei[1, ei[1] == 5] = 6                          # an empty target row
    w = rs.uniform(0.5, 1.5, e).astype(np.float32) if "weighted" in case else None
    improved = "improved" in case
    ei_t = torch.from_numpy(ei).to(d)
    w_t = None if w is None else torch.from_numpy(w).to(d)
    g = GcnGraph(ei_t, n_src, n_dst, w_t, improved, bipartite=bip)
    fill, loops = (2.0 if improved else 1.0), (0 if bip else 1)
    b_dst = block_size(n_dst, world)
    parts, dis_all = [], torch.zeros(world * b_dst, device=d)
    for r in range(world):
        d0, d1 = block_bounds(n_dst, world, r)
        shard, ws = filter_edges(ei_t, w_t, False, d0, d1)
        keep = (ei[1] >= d0) & (ei[1] < d1)
        assert np.array_equal(shard.cpu().numpy(), ei[:, keep])            # order-preserving filter
        rowptr, col, val, deg, dis, nnz = part_structure(shard.contiguous(), ws, 1, d0, d1 - d0, loops, fill, True)
        dis_all[r * b_dst: r * b_dst + (d1 - d0)] = dis[: d1 - d0]
        parts.append((d0, d1, rowptr, col, val, deg, nnz))
    g_rp, g_col, g_val = g.fwd.rowptr.cpu().numpy(), g.fwd.col.cpu().numpy(), g.fwd.val.cpu().numpy()
    for d0, d1, rowptr, col, val, deg, nnz in parts:
        part_values(rowptr, col, val, d1 - d0, d0, None if bip else dis_all, dis_all, False)
        a, b = g_rp[d0], g_rp[d1]
        assert nnz == b - a
        assert np.array_equal(rowptr[: d1 - d0 + 1].cpu().numpy(), g_rp[d0:d1 + 1] - a)
        assert np.array_equal(col[:nnz].cpu().numpy(), g_col[a:b])
        assert np.array_equal(val[:nnz].cpu().numpy(), g_val[a:b])         # bit-identical fp32 coefficients
        assert np.array_equal(deg[: d1 - d0].cpu().numpy(), g.deg[d0:d1].cpu().numpy())
    t_rp, t_col, t_val = g.bwd.rowptr.cpu().numpy(), g.bwd.col.cpu().numpy(), g.bwd.val.cpu().numpy()
    for r in range(world):
        s0, s1 = block_bounds(n_src, world, r)
        shard, ws = filter_edges(ei_t, w_t, True, s0, s1)
        rowptr, col, val, _, _, nnz = part_structure(shard.contiguous(), ws, 0, s0, s1 - s0, loops, fill, False)
        part_values(rowptr, col, val, s1 - s0, s0, None if bip else dis_all, dis_all, True)
        a, b = t_rp[s0], t_rp[s1]
        assert nnz == b - a
        assert np.array_equal(rowptr[: s1 - s0 + 1].cpu().numpy(), t_rp[s0:s1 + 1] - a)
        assert np.array_equal(col[:nnz].cpu().numpy(), t_col[a:b])
        assert np.array_equal(val[:nnz].cpu().numpy(), t_val[a:b])


@pytest.mark.parametrize("n,r,e", [(120, 7, 4000), (645, 16, 50000), (1500, 300, 20000), (5, 1, 0)])
def test_pair_prep(n, r, e):
    """(node, relation) pair CSR of the one-pass decoder backward: bit-exact vs a stable argsort of
    node * n_rel + rel over the 2E endpoint entries."""
    from gripnet_b200.graph import PairStruct
    rs = np.random.RandomState(11)
    ei = rs.randint(0, n, size=(2, e))
    et = rs.randint(0, r, size=e)
    ps = PairStruct(torch.from_numpy(ei).to(_dev()), torch.from_numpy(et).to(_dev()), n, r, exact=True)
    keys = np.concatenate([ei[0], ei[1]]) * r + np.concatenate([et, et])
    rp, perm = port.csr_from_edges(keys, n * r)
    assert np.array_equal(ps.csr.rowptr.cpu().numpy(), rp)
    assert np.array_equal(ps.ent_other[: 2 * e].cpu().numpy(), np.concatenate([ei[1], ei[0]])[perm])
    assert np.array_equal(ps.ent_eid[: 2 * e].cpu().numpy(), np.where(perm < e, perm, perm - e))
