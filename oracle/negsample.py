"""CPU restatement of the device negative sampler (``gripnet_b200/csrc/negsample.cu``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference semantics (``gripnet/utils.py:98-119``): every positive edge gets one uniformly random node
pair ``code = row * N + col`` drawn from ``[0, N^2)`` WITH replacement (``np.random.choice(N**2, E)``);
draws that hit a positive pair are redrawn until none is left.  ``negative_sampling`` rejects against all
positives, ``typed_negative_sampling`` (:115-119) against the positives of the same ``range_list`` slice.
Rejection per element is the same distribution as the reference's redraw-the-rejected loop: i.i.d. uniform
over the non-positive pairs.  The RNG STREAM is this implementation's own (counter-based Philox4x32-10, so
that the device kernel needs no state): parity with the reference is distributional, parity between this
file and the CUDA kernel is bit-exact.

Draw ``a`` (attempt 0, 1, ...) of edge ``e`` in epoch ``t`` with seed ``s``:
    (x0, x1, x2, x3) = philox4x32_10(counter = (e_lo, e_hi, a, t_lo), key = (s_lo, s_hi))
    u = x0 | x1 << 32                       # 64 random bits
    code = (u * N^2) >> 64                   # multiply-high: uniform on [0, N^2) up to 2^-64 * N^2 bias
    row, col = code // N, code % N           # exact integer division (the reference's float `perm / N`
                                             # (:111) is exact only while N^2 < 2^24)
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  uint32 arrays in, four uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _mulhi64(u, n2):
    """floor(u * n2 / 2^64) for uint64 arrays u and a Python int n2 < 2^63, exactly."""
    u = u.astype(np.uint64)
    a_hi, a_lo = u >> np.uint64(32), u & MASK32
    b_hi, b_lo = np.uint64(n2 >> 32), np.uint64(n2 & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        ll = a_lo * b_lo
        lh = a_lo * b_hi
        hl = a_hi * b_lo
        hh = a_hi * b_hi
        mid = (ll >> np.uint64(32)) + (lh & MASK32) + (hl & MASK32)
        return hh + (lh >> np.uint64(32)) + (hl >> np.uint64(32)) + (mid >> np.uint64(32))


def draw_codes(edge_ids, attempts, epoch, seed, n_nodes):
    e = np.asarray(edge_ids, dtype=np.uint64)
    a = np.asarray(attempts, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10((e & MASK32).astype(np.uint32), (e >> np.uint64(32)).astype(np.uint32),
                                 (a & MASK32).astype(np.uint32), np.uint32(epoch & 0xFFFFFFFF),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = x0.astype(np.uint64) | (x1.astype(np.uint64) << np.uint64(32))
    return _mulhi64(u, int(n_nodes) * int(n_nodes)).astype(np.int64)


def negative_sampling(pos_edge_index, n_nodes, seed, epoch, range_list=None):
    """int64 ``[2, E]`` negatives.  ``range_list`` (``[R, 2]`` half-open slices) switches to the typed rule."""
    pos = np.asarray(pos_edge_index, dtype=np.int64)
    n_edges = pos.shape[1]
    n2 = int(n_nodes) * int(n_nodes)
    pos_code = pos[0] * n_nodes + pos[1]
    rel = np.zeros(n_edges, dtype=np.int64)
    if range_list is not None:
        for r, (s, t) in enumerate(np.asarray(range_list, dtype=np.int64).tolist()):
            rel[s:t] = r
    taken = np.unique(rel * n2 + pos_code)                 # (relation, pair) keys; relation 0 when untyped
    out = np.empty(n_edges, dtype=np.int64)
    todo = np.arange(n_edges, dtype=np.int64)
    attempt = 0
    while todo.size:
        code = draw_codes(todo, np.full(todo.size, attempt), epoch, seed, n_nodes)
        bad = np.isin(rel[todo] * n2 + code, taken)
        out[todo[~bad]] = code[~bad]
        todo = todo[bad]
        attempt += 1
    return np.stack([out // n_nodes, out % n_nodes])
