"""Multi-rank parity worker for the destination-partitioned path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/dist_worker.py

Every rank (1) runs the partitioned PoseModel / ChainModel on its block of rows, (2) runs the SAME global
model on its own GPU through the single-GPU path, and checks that the partitioned outputs / gradients are
the row slices (row-partitioned tensors) or equal (replicated parameters, loss) of the global ones within
1e-5 relative; the pose case is also checked against the CPU oracle on rank 0.  With WORLD_SIZE unset it
runs as a single rank (degenerate partition).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

TOL = 1e-5


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.shape != b.shape:
        return float("inf")
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check(name, a, b, tol=TOL):
    e = rel(a, b)
    assert e < tol, f"{name}: rel err {e:.3e}"
    return e


def run_pose(dctx, dev, with_oracle):
    from gripnet_b200 import graph as G
    from gripnet_b200.pipelines import PoseModel, load_flat_params, shard_pose, shard_pose_params, to_device
    from oracle import synth
    g = synth.pose_medium()
    p = synth.pose_params(g)
    # ---- global model, single-GPU path
    ref = load_flat_params(PoseModel(g["n_g"], g["n_d"], g["n_rel"]), p).to(dev)
    loss_g, z_g, pos_g, neg_g = ref(to_device(g, dev))
    loss_g.backward()
    # ---- partitioned model
    G.clear_cache()
    data = shard_pose(g, dctx, dev)
    m = load_flat_params(PoseModel(data["n_g"], data["n_d"], g["n_rel"]), shard_pose_params(p, g, dctx)).to(dev)
    m.dmt.dist_ctx = dctx
    loss, z, pos, neg = m(data)
    loss.backward()
    if dctx.defer_grad_reduce:          # partial sums in .grad (and this rank's loss share) until the reduction
        dctx.reduce_gradients([v for k, v in m.named_parameters() if k not in ("gg.embedding", "gd.target_feat")])
        loss = dctx.loss_value.view(())
    worst = 0.0
    r0, r1 = dctx.bounds(g["n_d"])
    e0, e1 = data["edge_slice"]
    worst = max(worst, check("z", z, z_g[r0:r1]), check("loss", loss, loss_g),
                check("pos", pos, pos_g[e0:e1]), check("neg", neg, neg_g[e0:e1]))
    named_g = dict(ref.named_parameters())
    for k, v in m.named_parameters():
        gg = named_g[k].grad
        if gg is None:
            assert v.grad is None, k
            continue
        if k == "gg.embedding":
            gg = dctx.shard_rows(gg, g["n_g"])
        elif k == "gd.target_feat":
            gg = dctx.shard_rows(gg, g["n_d"])
        worst = max(worst, check("grad." + k, v.grad, gg, 2e-5))
    if with_oracle:
        from oracle import port
        pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        loss_o, z_o, _, _ = port.pose_forward(pl, g)
        worst = max(worst, check("z vs oracle", z, z_o[r0:r1]), check("loss vs oracle", loss, loss_o))
    return worst


def run_chain(dctx, dev, streamed=False, halo="auto"):
    """``streamed``: every rank keeps only its shard of each generated chunk (no global edge list, no global
    CSR on any rank) — the way bench.py builds BASELINE config 5.  ``halo="force"``: every partitioned CSR uses
    the halo-packed operand layout (only referenced remote rows are exchanged), whatever its density."""
    from gripnet_b200 import graph as G
    G.HALO_MODE = halo
    from gripnet_b200.pipelines import ROW_PARTITIONED, ChainModel, shard_chain, shard_chain_streamed
    from synthdata import chain_small
    g = chain_small(dev)
    torch.manual_seed(5)
    ref = ChainModel(g["n_a"], g["n_b"], g["n_c"], g["n_class"], hid=16, out=8).to(dev)
    loss_g, z_g, _ = ref(g)
    loss_g.backward()
    G.clear_cache()
    data = shard_chain_streamed(chain_small, dctx, dev) if streamed else shard_chain(g, dctx, dev)
    m = ChainModel(data["n_a"], data["n_b"], data["n_c"], g["n_class"], hid=16, out=8).to(dev)
    m.mcip.dist_ctx = dctx
    sd = {}
    for k, v in ref.state_dict().items():
        sd[k] = dctx.shard_rows(v, g[ROW_PARTITIONED[k]]).clone() if k in ROW_PARTITIONED else v.clone()
    m.load_state_dict(sd)
    loss, z, _ = m(data)
    loss.backward()
    r0, r1 = dctx.bounds(g["n_c"])
    G.HALO_MODE = "auto"
    worst = max(check("chain z", z, z_g[r0:r1]), check("chain loss", loss, loss_g))
    named_g = dict(ref.named_parameters())
    for k, v in m.named_parameters():
        gg = named_g[k].grad
        if gg is None:
            assert v.grad is None, k
            continue
        if k in ROW_PARTITIONED:
            gg = dctx.shard_rows(gg, g[ROW_PARTITIONED[k]])
        worst = max(worst, check("chain grad." + k, v.grad, gg, 2e-5))
    return worst


def run_chain_peer(dctx, dev):
    """ChainModel with every exchange halo-packed: step 1 runs the NCCL all-to-all fallback, later steps
    gn_peer_halo_push over the symmetric arena; eager and captured steps must agree bit for bit on z and to
    rounding on the reduced gradients, and with the single-GPU model within 1e-5."""
    from gripnet_b200 import graph as G
    from gripnet_b200.capture import CapturedStep
    from gripnet_b200.pipelines import ROW_PARTITIONED, ChainModel, shard_chain_streamed
    from synthdata import chain_small
    g = chain_small(dev)
    torch.manual_seed(5)
    ref = ChainModel(g["n_a"], g["n_b"], g["n_c"], g["n_class"], hid=16, out=8).to(dev)
    loss_g, z_g, _ = ref(g)
    G.clear_cache()
    G.HALO_MODE = "force"
    data = shard_chain_streamed(chain_small, dctx, dev)
    m = ChainModel(data["n_a"], data["n_b"], data["n_c"], g["n_class"], hid=16, out=8).to(dev)
    m.mcip.dist_ctx = dctx
    sd = {k: (dctx.shard_rows(v, g[ROW_PARTITIONED[k]]).clone() if k in ROW_PARTITIONED else v.clone())
          for k, v in ref.state_dict().items()}
    m.load_state_dict(sd)
    replicated = [v for k, v in m.named_parameters() if k not in ROW_PARTITIONED]
    snaps = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for step in range(3):
            m.zero_grad(set_to_none=True)
            out = m(data)
            out[0].backward()
            dctx.reduce_gradients(replicated)
            torch.cuda.synchronize()
            snaps.append([dctx.loss_value.detach().clone(), out[1].detach().clone()])
    torch.cuda.synchronize()
    G.HALO_MODE = "auto"
    assert not dctx.peer_failed()
    assert any(getattr(c._graph, "fwd_halo", None) is not None for c in m.aa.conv_list), "halo plan was not built"
    assert torch.equal(snaps[0][1], snaps[2][1]), "halo push over peer memory differs from the NCCL all-to-all"
    r0, r1 = dctx.bounds(g["n_c"])
    worst = max(check("halo chain z", snaps[2][1], z_g[r0:r1]), check("halo chain loss", snaps[2][0].view(()), loss_g))
    step = CapturedStep(lambda: m(data), m.parameters(), warmup=1, post_backward=lambda: dctx.reduce_gradients(replicated))
    for _ in range(3):
        out = step.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[1].detach(), snaps[2][1])
    return worst


def run_peer(dctx, dev):
    """Peer-memory exchange (gn_peer_allgather / gn_peer_push + gn_slot_sum over the symmetric arena) against
    the NCCL exchange: step 1 runs on NCCL (the arena is sized from it), steps 2.. move everything over NVLink
    peer memory.  Gathers are copies and every kernel is deterministic, so forward results must agree BIT FOR
    BIT; reductions are summed in rank order instead of NCCL's order, so gradients agree to rounding (1e-6) —
    and bit for bit between two peer-memory steps, eagerly and when the step is replayed from a CUDA graph."""
    from gripnet_b200 import graph as G
    from gripnet_b200.capture import CapturedStep
    from gripnet_b200.pipelines import PoseModel, load_flat_params, shard_pose, shard_pose_params
    from oracle import synth
    g = synth.pose_medium()
    p = synth.pose_params(g)
    G.clear_cache()
    data = shard_pose(g, dctx, dev)
    m = load_flat_params(PoseModel(data["n_g"], data["n_d"], g["n_rel"]), shard_pose_params(p, g, dctx)).to(dev)
    m.dmt.dist_ctx = dctx
    replicated = [v for k, v in m.named_parameters() if k not in ("gg.embedding", "gd.target_feat")]

    def finish():
        if dctx.defer_grad_reduce:
            dctx.reduce_gradients(replicated)

    def loss_of(out):
        return (dctx.loss_value if dctx.defer_grad_reduce else out[0]).detach().clone().view(())

    snaps = []
    side = torch.cuda.Stream()           # not the legacy stream: the same model is captured below
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for step in range(4):
            m.zero_grad(set_to_none=True)
            out = m(data)
            out[0].backward()
            finish()
            torch.cuda.synchronize()
            snaps.append([loss_of(out), out[1].detach().clone()] +
                         [v.grad.clone() for v in m.parameters() if v.grad is not None])
    torch.cuda.synchronize()
    used_peer = dctx.peer_gathers > 0
    assert not dctx.peer_failed(), "a peer gather timed out"
    assert torch.equal(snaps[0][1], snaps[2][1]), "peer-memory gather differs from the NCCL gather"
    for a, b in zip(snaps[0], snaps[2]):
        check("peer vs nccl", a, b, 1e-6)
    for a, b in zip(snaps[2], snaps[3]):
        assert torch.equal(a, b), "two peer-memory steps differ"
    # captured + replayed
    step = CapturedStep(lambda: m(data), m.parameters(), warmup=1, post_backward=finish)
    for _ in range(3):
        out = step.replay()
    torch.cuda.synchronize()
    assert not dctx.peer_failed()
    assert torch.equal(loss_of(out), snaps[2][0]) and torch.equal(out[1].detach(), snaps[2][1])
    for a, b in zip([v.grad for v in m.parameters() if v.grad is not None], snaps[2][2:]):
        assert torch.equal(a, b), "replayed gradients differ from the eager peer-memory step"
    return used_peer, dctx.peer_gathers, dctx.nccl_gathers, dctx.peer_reductions, dctx.nccl_reductions


def main():
    import gripnet_b200  # noqa: F401
    from gripnet_b200.parallel import DistContext
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        if "MASTER_ADDR" not in os.environ:
            os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", "29549"
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dctx = DistContext()
    w1 = run_pose(dctx, dev, with_oracle=(rank == 0))
    w2 = max(run_chain(dctx, dev), run_chain(dctx, dev, streamed=True), run_chain(dctx, dev, halo="force"),
             run_chain(dctx, dev, streamed=True, halo="off"))
    if world > 1:        # halo-packed exchange over peer memory (arena steps) and inside a captured step
        w2 = max(w2, run_chain_peer(DistContext(defer_grad_reduce=True), dev))
    w1 = max(w1, run_pose(DistContext(defer_grad_reduce=True), dev, with_oracle=False))
    peer = (False, 0, 0, 0, 0)
    if world > 1:
        run_peer(DistContext(), dev)
        peer = run_peer(DistContext(defer_grad_reduce=True), dev)
    t = torch.tensor([w1, w2], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dist_worker: world={world} OK  worst rel err pose {float(t[0]):.2e}  chain {float(t[1]):.2e}  "
              f"peer-memory exchange {'ON' if peer[0] else 'off'} ({peer[1]} peer / {peer[2]} nccl gathers, "
              f"{peer[3]} peer / {peer[4]} nccl reductions)", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)          # captured graphs hold NCCL work: skip the communicator teardown (see bench.py)


if __name__ == "__main__":
    main()
