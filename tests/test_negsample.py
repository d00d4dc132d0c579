"""Negative sampling (reference gripnet/utils.py:98-119).

CPU: the oracle's Philox4x32-10 against the published Random123 known-answer vectors, the oracle's draws
against plain Python integer arithmetic, and the reference's contract (one non-positive pair per positive,
uniform).  GPU: the device sampler against the oracle, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import negsample as ns


def test_philox4x32_10_known_answer_vectors():
    """Random123 kat_vectors (Salmon et al.): counter, key -> output."""
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        got = ns.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_draw_is_multiply_high_of_the_first_two_words():
    e = np.array([0, 1, 2 ** 33 + 5, 123456789], dtype=np.int64)
    a = np.array([0, 3, 1, 70000], dtype=np.int64)
    seed, epoch, n = 0x1234567887654321, 7, 3_000_000
    x0, x1, _, _ = ns.philox4x32_10((e & 0xFFFFFFFF).astype(np.uint32), (e >> 32).astype(np.uint32),
                                    a.astype(np.uint32), np.uint32(epoch), seed & 0xFFFFFFFF, seed >> 32)
    want = [((int(lo) | (int(hi) << 32)) * n * n) >> 64 for lo, hi in zip(x0, x1)]
    assert ns.draw_codes(e, a, epoch, seed, n).tolist() == want


@pytest.mark.parametrize("typed", [False, True])
def test_oracle_contract(typed):
    rs = np.random.RandomState(3)
    n, sizes = 40, [300, 500, 100]
    pos = np.concatenate([rs.randint(0, n, (2, k)) for k in sizes], axis=1)
    bounds = np.cumsum([0] + sizes)
    rl = np.stack([bounds[:-1], bounds[1:]], axis=1) if typed else None
    neg = ns.negative_sampling(pos, n, seed=11, epoch=0, range_list=rl)
    assert neg.shape == pos.shape and neg.dtype == np.int64 and neg.min() >= 0 and neg.max() < n
    code_p, code_n = pos[0] * n + pos[1], neg[0] * n + neg[1]
    if typed:
        for s, t in rl:
            assert not np.isin(code_n[s:t], code_p[s:t]).any()
        assert np.isin(code_n, code_p).any()          # pairs of OTHER relations are allowed (and occur)
    else:
        assert not np.isin(code_n, code_p).any()
    # epochs and seeds give different draws; the same (seed, epoch) is reproducible
    assert not np.array_equal(neg, ns.negative_sampling(pos, n, 11, 1, rl))
    assert not np.array_equal(neg, ns.negative_sampling(pos, n, 12, 0, rl))
    assert np.array_equal(neg, ns.negative_sampling(pos, n, 11, 0, rl))


def test_oracle_is_uniform_over_the_free_pairs():
    n = 12
    pos = np.stack([np.repeat(np.arange(n), 4), np.tile(np.arange(4), n)])      # 48 of 144 pairs taken
    counts = np.zeros(n * n)
    for epoch in range(400):
        neg = ns.negative_sampling(pos, n, seed=5, epoch=epoch)
        np.add.at(counts, neg[0] * n + neg[1], 1)
    taken = pos[0] * n + pos[1]
    assert counts[taken].sum() == 0
    free = np.setdiff1d(np.arange(n * n), taken)
    exp = counts.sum() / free.size
    chi2 = ((counts[free] - exp) ** 2 / exp).sum()
    assert chi2 < 160, chi2          # 95 dof: mean 95, sd 13.8 -> > 4.7 sd would fail


def test_sampling_utils_have_no_cpu_path():
    """``gripnet.utils.negative_sampling`` / ``typed_negative_sampling`` keep the reference's signatures but run
    on the device only: CPU tensors are refused like everywhere else in the package."""
    import gripnet_b200.utils as u
    pos = torch.from_numpy(np.random.RandomState(0).randint(0, 30, (2, 200)))
    with pytest.raises(RuntimeError, match="CUDA"):
        u.negative_sampling(pos, 30)
    with pytest.raises(RuntimeError, match="CUDA"):
        u.typed_negative_sampling(pos, 30, torch.tensor([[0, 120], [120, 200]]))


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,sizes,typed", [(645, [25000] * 16, False), (645, [25000] * 16, True),
                                           (40, [700, 900], False), (40, [700, 900], True),
                                           (100000, [5000], False), (7, [3, 0, 5], True)])
def test_device_sampler_matches_oracle_bit_for_bit(n, sizes, typed):
    from gripnet_b200.utils import NegativeSampler
    d = torch.device("cuda:0")
    rs = np.random.RandomState(n + len(sizes))
    if n == 7:
        pos = np.concatenate([rs.randint(0, 3, (2, k)) for k in sizes], axis=1)    # leaves free pairs
    else:
        pos = np.concatenate([rs.randint(0, n, (2, k)) for k in sizes], axis=1)
    bounds = np.cumsum([0] + sizes)
    rl = np.stack([bounds[:-1], bounds[1:]], axis=1)
    seed = 0x9E3779B97F4A7C15
    s = NegativeSampler(torch.from_numpy(pos).to(d), n, torch.from_numpy(rl) if typed else None, seed=seed)
    for epoch in range(3):
        assert s.epoch == epoch
        got = s.sample().cpu().numpy()
        want = ns.negative_sampling(pos, n, seed, epoch, rl if typed else None)
        assert np.array_equal(got, want), (epoch, int((got != want).sum()))


@pytest.mark.gpu
def test_device_sampler_in_a_cuda_graph_draws_new_negatives_every_replay():
    from gripnet_b200.utils import NegativeSampler
    d = torch.device("cuda:0")
    rs = np.random.RandomState(1)
    pos = rs.randint(0, 200, (2, 5000))
    s = NegativeSampler(torch.from_numpy(pos).to(d), 200, seed=42)
    out = torch.empty(2, 5000, dtype=torch.int64, device=d)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        s.sample(out)                                  # epoch 0, eager
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s.sample(out)
    for epoch in (1, 2, 3):
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ns.negative_sampling(pos, 200, 42, epoch))


@pytest.mark.gpu
def test_utils_negative_sampling_on_cuda_tensors():
    import gripnet_b200.utils as u
    d = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    pos = torch.from_numpy(rs.randint(0, 300, (2, 4000))).to(d)
    a, b = u.negative_sampling(pos, 300), u.negative_sampling(pos, 300)
    assert a.is_cuda and a.dtype == torch.int64 and a.shape == pos.shape and not torch.equal(a, b)
    code_p = (pos[0] * 300 + pos[1]).cpu().numpy()
    for t in (a, b):
        assert not np.isin((t[0] * 300 + t[1]).cpu().numpy(), code_p).any()
    rl = torch.tensor([[0, 1000], [1000, 4000]])
    t = u.typed_negative_sampling(pos, 300, rl)
    for s, e in rl.tolist():
        assert not np.isin((t[0, s:e] * 300 + t[1, s:e]).cpu().numpy(), code_p[s:e]).any()


@pytest.mark.gpu
def test_utils_negative_sampling_cache_misses_do_not_repeat_a_stream():
    """A fresh clone / slice of the positives on every call misses the sampler cache: each miss must still
    draw from its own stream (the reference draws fresh numpy randoms per call, utils.py:104-110), and more
    live edge tensors than the cache holds must keep working (LRU eviction)."""
    import gripnet_b200.utils as u
    d = torch.device("cuda:0")
    rs = np.random.RandomState(1)
    pos = torch.from_numpy(rs.randint(0, 300, (2, 4000))).to(d)
    draws = [u.negative_sampling(pos.clone(), 300) for _ in range(4)]
    for i in range(4):
        for j in range(i):
            assert not torch.equal(draws[i], draws[j])
    slices = [pos[:, i * 50:(i + 1) * 50].contiguous() for i in range(u._SAMPLER_CAPACITY + 6)]
    first = [u.negative_sampling(s, 300) for s in slices]
    again = [u.negative_sampling(s, 300) for s in slices]
    assert len(u._samplers) <= u._SAMPLER_CAPACITY
    assert sum(torch.equal(a, b) for a, b in zip(first, again)) == 0
