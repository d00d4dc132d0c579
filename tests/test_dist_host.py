"""CPU (gloo, world_size 2): host-side logic of the destination-partitioned path — block partition,
in-place slot all-gather, reduce-scatter backward, loss weighting, replicated-gradient all-reduce — and the
partition MATH itself: a GCN layer computed from row slices of the global dst-CSR plus an all-gather of
the narrowed operand equals the global layer (the oracle does the arithmetic here; kernels are not
involved)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fn_name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn_name, port, world=2):
    mp.spawn(_worker, args=(world, port, fn_name), nprocs=world, join=True)


# ------------------------------------------------------------------------------------------------
def _collectives(rank, world):
    from gripnet_b200 import parallel
    ctx = parallel.DistContext()
    assert (ctx.rank, ctx.world) == (rank, world)
    n = 7                                           # uneven: blocks of 4 and 3
    b = ctx.block(n)
    assert b == 4 and [parallel.block_bounds(n, world, r) for r in range(world)] == [(0, 4), (4, 7)]
    glob = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
    mine = ctx.shard_rows(glob)
    full = torch.full((world * b, 3), -1.0)
    r0, r1 = ctx.bounds(n)
    full[rank * b: rank * b + (r1 - r0)] = mine
    ctx.all_gather_slots(full)
    assert torch.equal(full[:n], glob)               # gathered row index == global node id
    # differentiable gather: backward is the reduce-scatter of the partial gradients
    z = mine.clone().requires_grad_(True)
    zf = parallel.all_gather_rows(z, ctx, n)
    assert torch.equal(zf[:n], glob) and float(zf[n:].abs().sum()) == 0.0
    wgt = torch.arange(world * b * 3, dtype=torch.float32).view(world * b, 3) * (rank + 1)
    (zf * wgt).sum().backward()
    expect = sum(torch.arange(world * b * 3, dtype=torch.float32).view(world * b, 3) * (r + 1) for r in range(world))
    assert torch.equal(z.grad, expect[rank * b: rank * b + (r1 - r0)])
    # replicated parameter: gradient all-reduced
    w = torch.ones(2, 2, requires_grad=True)
    (parallel.replicated(w, ctx) * (rank + 1)).sum().backward()
    assert torch.equal(w.grad, torch.full((2, 2), float(sum(range(1, world + 1)))))
    # global mean from per-rank means
    local = torch.tensor(float(rank + 1), requires_grad=True)
    n_local = [3, 1][rank]
    loss = parallel.global_mean_loss(local, n_local, 4, ctx)
    assert abs(float(loss) - (1 * 3 / 4 + 2 * 1 / 4)) < 1e-6
    loss.backward()
    assert abs(float(local.grad) - n_local / 4) < 1e-7
    # registry
    ei = torch.randint(0, n, (2, 10))
    assert parallel.lookup(ei) is None
    parallel.distribute_edges(ei, ctx, n)
    spec = parallel.lookup(ei)
    assert spec.n_src == n and spec.n_dst == n and spec.ctx is ctx
    parallel.clear_registry()


def _partition_math(rank, world):
    from gripnet_b200 import parallel
    from oracle import port
    ctx = parallel.DistContext()
    rs = np.random.RandomState(3)
    n, e, k, f = 53, 400, 6, 4
    ei = torch.from_numpy(rs.randint(0, n, size=(2, e)).astype(np.int64))
    x = torch.from_numpy(rs.randn(n, k).astype(np.float32))
    w = torch.from_numpy(rs.randn(k, f).astype(np.float32))
    bias = torch.from_numpy(rs.randn(f).astype(np.float32))
    st = port.gcn_csr_oracle(ei, n)                                      # global dst-CSR (oracle)
    want = port.gcn_conv(x, w, bias, torch.from_numpy(st["edge_index_aug"]), torch.from_numpy(st["norm"]))
    r0, r1 = ctx.bounds(n)
    b = ctx.block(n)
    rowptr, col, val = st["rowptr"], st["col"], st["val"]
    # local operand slot -> in-place all-gather -> local rows of the CSR
    y = torch.zeros(world * b, f)
    y[rank * b: rank * b + (r1 - r0)] = x[r0:r1] @ w
    ctx.all_gather_slots(y)
    out = torch.zeros(r1 - r0, f)
    for i in range(r0, r1):
        a, z = int(rowptr[i]), int(rowptr[i + 1])
        out[i - r0] = (torch.from_numpy(val[a:z])[:, None] * y[torch.from_numpy(col[a:z])]).sum(0) + bias
    err = float((out - want[r0:r1]).abs().max() / want.abs().max())
    assert err < 1e-5, err


def test_collectives_and_registry_world2():
    _spawn("_collectives", 29561)


def test_partitioned_gcn_layer_equals_global_world2():
    _spawn("_partition_math", 29563)


def test_block_partition_properties():
    sys.path.insert(0, ROOT)
    from gripnet_b200.parallel import block_bounds, block_size
    for n in [0, 1, 5, 645, 19081, 4_000_000]:
        for world in [1, 2, 3, 4, 8]:
            b = block_size(n, world)
            spans = [block_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert all(0 <= hi - lo <= b for lo, hi in spans) and world * b >= n


# ------------------------------------------------------------------------------------------------
def _halo_plan(rank, world):
    """HaloPlan (graph.py) on CPU tensors over gloo: the packed operand [own rows | referenced rows of each peer]
    read through the remapped columns equals the full gathered operand read through the global columns, for an
    uneven partition with rows that reference nothing remote and peers that are not referenced at all."""
    from types import SimpleNamespace
    from gripnet_b200 import graph as G
    from gripnet_b200 import parallel
    G.HALO_MODE = "force"
    ctx = parallel.DistContext()
    n = 23                                            # blocks of 12 and 11 (world 2)
    b = ctx.block(n)
    r0, r1 = ctx.bounds(n)
    rs = np.random.RandomState(7)                     # same global graph on every rank
    cols_all = [rs.choice(n, size=rs.randint(0, 6), replace=True) for _ in range(n)]
    mine = cols_all[r0:r1]
    col = torch.tensor(np.concatenate(mine + [np.array([r0])]).astype(np.int32))       # + one local column
    csr = SimpleNamespace(col=col.clone(), nnz=col.numel())
    plan = G.HaloPlan.build(csr, ctx, b, n, r0, r1)
    assert plan is not None and plan.rows >= b
    feat = torch.arange(world * b * 3, dtype=torch.float32).view(world * b, 3)          # row g = global node g
    packed = torch.full((plan.rows, 3), -1.0)
    packed[: r1 - r0] = feat[r0:r1]
    ctx.halo_gather(packed, plan)                     # gloo all-to-all of the packed rows
    assert torch.equal(packed[csr.col.long()], feat[col.long()])
    # every peer learnt exactly which of its rows this rank needs
    need = sorted(set(int(c) for c in col.tolist() if not (r0 <= c < r1)))
    assert sum(plan.recv_counts) == len(need)


def test_halo_plan_packs_and_remaps_over_gloo():
    _spawn("_halo_plan", 29635)
