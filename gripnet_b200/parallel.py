"""Destination-partitioned multi-GPU execution — one process per GPU.

The reference has no distributed path (single process, single device); this is new design.

* Every supervertex's nodes are cut into ``world`` contiguous blocks of ``B = ceil(n / world)`` rows
  (the last block may be short).  Rank p owns block p: those rows of every activation / gradient, the
  matching rows of the row-partitioned parameters (``embedding``, ``target_feat``) and rows
  ``[p*B, p*B + n_p)`` of the dst-sorted CSR (forward) and of the src-sorted transpose CSR (backward).
  The local CSRs are SLICES of the global arrays (``graph.py``): bit-identical to slicing.
* Exchange step: the operand of each SpMM (``Y = X W`` forward, ``dZ`` backward) is produced straight
  into this rank's slot of a ``[world*B, F]`` buffer and completed by an in-place NCCL all-gather over
  NVLink — because the blocks are uniform, gathered row index == global node id, so CSR column indices
  need no remapping.  Weight gradients are partial sums: one small all-reduce each.
* The decoder partitions the EDGE lists; ``z`` is all-gathered (differentiable: the backward is a
  reduce-scatter of the partial ``dz``).

A global edge list is marked as partitioned by ``distribute_edges`` (a registry keyed by tensor
identity), after which the ordinary drop-in modules (``homoGraph``, ``interGraph`` ...) — constructed
with LOCAL node counts — run the partitioned path.  ``torch.distributed`` does the plumbing; all
arithmetic stays in the library's kernels.
"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

PEER_MODE = os.environ.get("GRIPNET_B200_PEER", "auto")     # "auto" | "off"
MULTICAST = os.environ.get("GRIPNET_B200_MULTICAST", "1") != "0"


class PeerArena:
    """Symmetric device arena mapped by every rank of the node (torch symmetric memory: CUDA VMM
    allocation + peer mapping over NVLink) from which the SpMM gather buffers are carved, so that
    ``gn_peer_allgather`` can push slots straight into the peers' buffers.

    Layout (identical on every rank): ``[flag block | data]``.  ``alloc`` is a bump allocator reset at
    every step; all ranks run the same sequence of allocations, so a buffer has the same offset — and
    the same flag index — everywhere."""
    MAX_BUFFERS = 1024

    def __init__(self, dctx, data_bytes):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        self.max_world = int(lib.gn_peer_max_world())
        if dctx.world > self.max_world:
            raise RuntimeError("peer all-gather supports at most %d ranks" % self.max_world)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.flag_bytes = self.MAX_BUFFERS * self.max_world * 8
        self.capacity = int(data_bytes)
        self.t = symm.empty(self.flag_bytes + self.capacity, dtype=torch.uint8, device=dev)
        group = dctx.group if dctx.group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.t, group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(ptrs) != dctx.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer pointers")
        # last entry: the multicast (NVLS) mapping of the arena, when the fabric offers one — one store then
        # reaches every rank (GRIPNET_B200_MULTICAST=0 keeps the unicast pushes)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if MULTICAST else 0
        self.multicast = mc != 0
        self.bases = (ctypes.c_uint64 * (dctx.world + 1))(*ptrs, mc)
        self.t[: self.flag_bytes].zero_()
        self.seq = torch.zeros(self.MAX_BUFFERS, dtype=torch.int64, device=dev)
        self.done = torch.zeros(self.MAX_BUFFERS, dtype=torch.int32, device=dev)
        self.abort = torch.zeros(1, dtype=torch.int32, device=dev)      # raised by a timed-out wait
        self.off = 0
        self.index = 0
        torch.cuda.synchronize()
        dist.barrier(group=dctx.group)

    def reset(self):
        self.off = 0
        self.index = 0

    def alloc(self, nbytes):
        """-> (uint8 view of the local arena, byte offset in the arena, buffer index) or None when full."""
        nbytes = int(nbytes)
        padded = (nbytes + 255) // 256 * 256
        if self.off + padded > self.capacity or self.index >= self.MAX_BUFFERS:
            return None
        start = self.flag_bytes + self.off
        view = self.t[start:start + nbytes]
        tok = (view, start, self.index)
        self.off += padded
        self.index += 1
        return tok


class DistContext:
    """Rank / world / process group of the partitioned run."""

    def __init__(self, group=None, defer_grad_reduce=False):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        # False: every replicated-parameter gradient is all-reduced where it is produced (simple, always
        # correct).  True: the layer stacks leave partial sums in ``.grad`` and the caller runs ONE bucketed
        # all-reduce after ``backward()`` with ``reduce_gradients`` (fewer, larger collectives).
        self.defer_grad_reduce = bool(defer_grad_reduce)
        self._bucket = None
        # peer-memory exchange (PeerArena): sized from the first step, which still runs on NCCL
        self.peer_mode = PEER_MODE if self.world > 1 else "off"
        self.arena = None
        self._tokens = {}
        self._step_bytes = 0
        self._step_gathers = 0
        self.peer_gathers = 0            # statistics: gathers served by gn_peer_allgather / by NCCL
        self.nccl_gathers = 0
        self.peer_reductions = 0         # reductions served by gn_peer_push + gn_slot_sum / by NCCL
        self.nccl_reductions = 0
        # deferred mode: this rank's share of the global mean loss, summed over ranks (in rank order, together
        # with the gradient bucket) by ``reduce_gradients`` into ``loss_value``
        self._loss_share = None
        self.loss_value = None

    # ---- block partition -------------------------------------------------
    def block(self, n):
        """Uniform block size for ``n`` rows."""
        return block_size(n, self.world)

    def bounds(self, n, rank=None):
        return block_bounds(n, self.world, self.rank if rank is None else rank)

    def local_count(self, n, rank=None):
        r0, r1 = self.bounds(n, rank)
        return r1 - r0

    def shard_rows(self, t, n=None):
        """This rank's rows of a global row-major tensor."""
        n = t.size(0) if n is None else n
        r0, r1 = self.bounds(n)
        return t[r0:r1]

    def edge_slice(self, e, rank=None):
        """Contiguous slice of an edge list of length ``e`` scored by this rank."""
        return block_bounds(e, self.world, self.rank if rank is None else rank)

    # ---- peer-memory gather buffers -----------------------------------------
    def begin_step(self):
        """Call once at the start of every training step (the model containers in ``pipelines.py`` do).
        Rewinds the arena's bump allocator; after the first step — whose gather buffers came from the
        ordinary allocator and were exchanged with NCCL while their sizes were recorded — it creates the
        symmetric arena (a collective: every rank takes the same decision from the same shapes)."""
        if self.peer_mode != "auto":
            return
        if self.arena is None and self._step_bytes > 0:
            if self._step_gathers >= 2:           # buffer-reuse argument of peer.cu needs >= 2 gathers per step
                self._create_arena(self._step_bytes)
            else:
                self.peer_mode = "off"
        if self.arena is not None:
            self.arena.reset()
        self._tokens.clear()
        self._step_bytes = 0
        self._step_gathers = 0

    def _create_arena(self, data_bytes):
        ok = torch.ones(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        arena = None
        try:
            arena = PeerArena(self, data_bytes)
        except Exception as e:                    # no symmetric memory on this system: keep NCCL
            sys.stderr.write(f"gripnet_b200: peer all-gather disabled on rank {self.rank} ({type(e).__name__}: {e})\n")
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            self.arena = arena
        else:
            self.peer_mode = "off"

    def peer_failed(self):
        """True if a peer-memory gather timed out (one host read; call outside captured regions)."""
        return self.arena is not None and int(self.arena.abort.item()) != 0

    def slot_buffer(self, rows, f, like, uniform=True):
        """A ``[rows, f]`` fp32 gather buffer (``rows = world * B``, or the packed row count of a halo plan with
        ``uniform=False``): from the symmetric arena when it exists (then ``all_gather_slots`` /
        ``halo_gather`` exchange it over peer memory), else from the allocator."""
        nbytes = int(rows) * int(f) * 4
        if self.peer_mode == "auto":
            self._step_bytes += (nbytes + 255) // 256 * 256
            self._step_gathers += 1
            ok = (nbytes // self.world) % 16 == 0 and nbytes % self.world == 0 if uniform else (int(f) % 4 == 0)
            if self.arena is not None and ok:
                tok = self.arena.alloc(nbytes)
                if tok is not None:
                    t = tok[0].view(torch.float32).view(int(rows), int(f))
                    self._tokens[t.data_ptr()] = tok
                    return t
        return torch.empty((rows, f), dtype=torch.float32, device=like.device)

    def exchange_buffer(self, slot_bytes):
        """A ``[world][slot_bytes]`` buffer of the symmetric arena for ``gn_peer_push`` / ``gn_slot_sum``
        -> ``(uint8 view, arena offset, flag index)``, or None while the arena does not exist (first step,
        peer mode off): the caller then uses NCCL.  Sizes are recorded like those of the gather buffers."""
        slot_bytes = (int(slot_bytes) + 15) // 16 * 16
        nbytes = slot_bytes * self.world
        if self.peer_mode != "auto" or self.world == 1:
            return None
        self._step_bytes += (nbytes + 255) // 256 * 256
        self._step_gathers += 1
        if self.arena is None:
            return None
        tok = self.arena.alloc(nbytes)
        return None if tok is None else tok + (slot_bytes,)

    def _peer_push(self, tok, segments):
        """``segments``: (tensor-or-pointer, bytes, slot_offset, peer or -1) -> one ``gn_peer_push``."""
        from . import _lib
        from .graph import _stream
        a = self.arena
        _, offset, index, slot_bytes = tok
        table = (_lib.GnPeerSegment * len(segments))()
        for i, (ptr, nbytes, off, peer) in enumerate(segments):
            table[i] = _lib.GnPeerSegment(ptr, nbytes, off, peer)
        _lib.check(_lib.load().gn_peer_push(a.bases, self.world, self.rank, offset, slot_bytes,
                                            ctypes.cast(table, ctypes.c_void_p), len(segments), 0, index,
                                            a.seq[index:].data_ptr(), a.done[index:].data_ptr(), a.abort.data_ptr(),
                                            _stream()), "gn_peer_push")

    def _slot_sum(self, tok, segments):
        """``segments``: (dst tensor, n floats, slot_offset) -> one ``gn_slot_sum`` over this rank's buffer."""
        from . import _lib
        from .graph import _stream
        view, _, _, slot_bytes = tok
        table = (_lib.GnSumSegment * len(segments))()
        for i, (dst, n, off) in enumerate(segments):
            table[i] = _lib.GnSumSegment(dst.data_ptr(), n, off)
        _lib.check(_lib.load().gn_slot_sum(view.data_ptr(), self.world, slot_bytes, ctypes.cast(table, ctypes.c_void_p),
                                           len(segments), _stream()), "gn_slot_sum")

    # ---- collectives -----------------------------------------------------
    def all_gather_slots(self, full):
        """``full`` is ``[world*B, F]`` with this rank's slot already written: complete it in place —
        over NVLink peer memory (``gn_peer_allgather``) for arena buffers, with NCCL otherwise."""
        if self.world == 1:
            return full
        tok = self._tokens.get(full.data_ptr())
        if tok is not None:
            from . import _lib
            from .graph import _stream
            a = self.arena
            _, offset, index = tok
            slot_bytes = full.numel() * 4 // self.world
            _lib.check(_lib.load().gn_peer_allgather(a.bases, self.world, self.rank, offset, slot_bytes, 0, index,
                                                     a.seq[index:].data_ptr(), a.done[index:].data_ptr(),
                                                     a.abort.data_ptr(), _stream()), "gn_peer_allgather")
            self.peer_gathers += 1
            return full
        b = full.size(0) // self.world
        dist.all_gather_into_tensor(full, full[self.rank * b:(self.rank + 1) * b], group=self.group)
        self.nccl_gathers += 1
        return full

    def halo_gather(self, full, plan):
        """Complete a halo-packed operand (``graph.HaloPlan``): this rank's rows sit at the head of ``full``; every
        peer's referenced rows are stored behind them — by ``gn_peer_halo_push`` over NVLink peer memory for arena
        buffers, by an NCCL all-to-all of the packed rows otherwise (first step / peer mode off)."""
        tok = self._tokens.get(full.data_ptr())
        if tok is not None:
            from . import _lib
            from .graph import _stream
            a = self.arena
            _, offset, index = tok
            table = (_lib.GnHaloPeer * self.world)()
            for p in range(self.world):
                cnt = plan.send_counts[p] if p != self.rank else 0
                table[p] = _lib.GnHaloPeer(plan.send_idx[p].data_ptr() if cnt else None, cnt, plan.dst_row[p])
            _lib.check(_lib.load().gn_peer_halo_push(a.bases, self.world, self.rank, offset, full.numel() * 4,
                                                     full.data_ptr(), full.size(1), ctypes.cast(table, ctypes.c_void_p),
                                                     0, index, a.seq[index:].data_ptr(), a.done[index:].data_ptr(),
                                                     a.abort.data_ptr(), _stream()), "gn_peer_halo_push")
            self.peer_gathers += 1
            return full
        send = full[: plan.block].index_select(0, plan.send_flat.long()) if plan.send_flat.numel() else \
            full.new_empty((0, full.size(1)))
        recv = full[plan.block: plan.block + sum(plan.recv_counts)]
        dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts,
                               group=self.group)
        self.nccl_gathers += 1
        return full

    def all_reduce_(self, t):
        if self.world > 1 and t is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def reduce_param_grad_(self, t):
        """Gradient of a replicated parameter (a partial sum on every rank)."""
        if not self.defer_grad_reduce:
            self.all_reduce_(t)
        return t

    def reduce_gradients(self, params):
        """All-reduce of the ``.grad`` of replicated parameters left as partial sums by a
        ``defer_grad_reduce=True`` step, together with this rank's share of the loss (``global_mean_loss``):
        every rank stores its gradients into slot ``rank`` of every rank's exchange buffer over NVLink peer
        memory (``gn_peer_push``), then sums the slots in rank order straight back into the ``.grad`` tensors
        (``gn_slot_sum``) — two launches, no NCCL, bit-identical results on every rank.  While the symmetric
        arena does not exist (first step, or peer mode off) one bucketed NCCL all-reduce does the same.
        ``params`` must not contain row-partitioned parameters."""
        if self.world == 1:
            if self._loss_share is not None:
                self.loss_value, self._loss_share = self._loss_share, None
            return
        grads = [p.grad for p in params if p.grad is not None]
        share, self._loss_share = self._loss_share, None
        if not grads and share is None:
            return
        if any(not g.is_contiguous() for g in grads):
            raise RuntimeError("gripnet_b200: gradients of replicated parameters must be contiguous")
        pieces = list(grads)
        if share is not None:
            if self.loss_value is None or self.loss_value.device != share.device:
                self.loss_value = torch.zeros(1, dtype=torch.float32, device=share.device)
            pieces.append(share.view(1))
        cuda = pieces[0].is_cuda
        offs, off = [], 0
        for t in pieces:
            offs.append(off)
            off += (t.numel() * 4 + 15) // 16 * 16
        tok = self.exchange_buffer(off) if cuda and len(pieces) <= self._max_segments() else None
        if tok is not None:
            self._peer_push(tok, [(t.data_ptr(), t.numel() * 4, o, -1) for t, o in zip(pieces, offs)])
            dsts = list(grads) + ([self.loss_value] if share is not None else [])
            self._slot_sum(tok, [(d, d.numel(), o) for d, o in zip(dsts, offs)])
            self.peer_reductions += 1
            return
        sizes = [t.numel() for t in pieces]
        if self._bucket is None or self._bucket.numel() != sum(sizes) or self._bucket.device != pieces[0].device:
            self._bucket = torch.empty(sum(sizes), dtype=pieces[0].dtype, device=pieces[0].device)
        torch.cat([t.reshape(-1) for t in pieces], out=self._bucket)
        dist.all_reduce(self._bucket, op=dist.ReduceOp.SUM, group=self.group)
        parts = list(self._bucket.split(sizes))
        torch._foreach_copy_([g.view(-1) for g in grads], parts[: len(grads)])
        if share is not None:
            self.loss_value.copy_(parts[-1])
        self.nccl_reductions += 1

    def _max_segments(self):
        from . import _lib
        return int(_lib.load().gn_peer_max_segments())

    def reduce_scatter_rows(self, full):
        """Sum ``full`` ``[world*B, F]`` over ranks and return this rank's ``[B, F]`` block: block q of every
        rank is stored into rank q's exchange buffer over peer memory (an all-to-all, ``gn_peer_push``) and the
        ``world`` contributions are summed in rank order (``gn_slot_sum``); NCCL reduce-scatter as fallback."""
        b = full.size(0) // self.world
        if self.world == 1:
            return full[:b]
        full = full.contiguous()
        out = torch.empty((b,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        blk_bytes = out.numel() * 4
        tok = self.exchange_buffer(blk_bytes) if (full.is_cuda and full.dtype == torch.float32 and
                                                  self.world <= self._max_segments()) else None
        if tok is not None:
            self._peer_push(tok, [(full.data_ptr() + q * blk_bytes, blk_bytes, 0, q) for q in range(self.world)])
            self._slot_sum(tok, [(out, out.numel(), 0)])
            self.peer_reductions += 1
            return out
        dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=self.group)
        self.nccl_reductions += 1
        return out


def block_size(n, world):
    return max(1, -(-int(n) // int(world)))


def block_bounds(n, world, rank):
    b = block_size(n, world)
    r0 = min(int(n), rank * b)
    return r0, min(int(n), r0 + b)


# --------------------------------------------------------------------------
# registry of partitioned edge lists (keyed by tensor identity; the entry keeps
# the tensor alive so its storage cannot be recycled under a stale key)
# --------------------------------------------------------------------------
class EdgeSpec:
    __slots__ = ("tensor", "ctx", "n_src", "n_dst", "shards", "n_edges_global")

    def __init__(self, tensor, ctx, n_src, n_dst, shards=None, n_edges_global=None):
        self.tensor, self.ctx, self.n_src, self.n_dst = tensor, ctx, int(n_src), int(n_dst)
        self.shards, self.n_edges_global = shards, n_edges_global


_registry = {}


def _key(t):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), str(t.dtype), str(t.device))


def distribute_edges(edge_index, ctx, n_src, n_dst=None):
    """Mark a GLOBAL ``edge_index`` (global node ids, identical on every rank) as destination-partitioned
    over ``ctx``.  ``n_src`` / ``n_dst`` are the GLOBAL node counts (``n_dst`` defaults to ``n_src``:
    a square, intra-supervertex graph).  Returns ``edge_index``; modules that receive it run the
    partitioned path and expect / return LOCAL rows."""
    _registry[_key(edge_index)] = EdgeSpec(edge_index, ctx, n_src, n_src if n_dst is None else n_dst)
    return edge_index


def distribute_edge_shards(by_dst, by_src, ctx, n_src, n_dst=None, n_edges_global=None, w_dst=None, w_src=None):
    """Register THIS RANK'S shard of a destination-partitioned edge list, for supergraphs whose global edge
    list no rank should hold: ``by_dst`` = the edges (global ids, original relative order) whose destination
    lies in this rank's block, ``by_src`` = those whose source does (``graph.filter_edges`` of a streamed
    generator or loader gives both).  Returns the key tensor to pass as ``edge_index`` to the modules."""
    n_dst = n_src if n_dst is None else n_dst
    total = n_edges_global if n_edges_global is not None else by_dst.size(1)
    _registry[_key(by_dst)] = EdgeSpec(by_dst, ctx, n_src, n_dst, shards=((by_dst, w_dst), (by_src, w_src)),
                                       n_edges_global=total)
    return by_dst


def lookup(edge_index):
    return _registry.get(_key(edge_index))


def clear_registry():
    """Forget every partitioned edge list (the registry keeps the key tensors alive: call this when a partitioned
    graph is dropped, as ``bench.py`` does between workloads)."""
    _registry.clear()


def forget_edges(edge_index):
    """Drop one partitioned edge list from the registry (and let its tensor be freed)."""
    _registry.pop(_key(edge_index), None)


# --------------------------------------------------------------------------
# differentiable row all-gather (decoder input)
# --------------------------------------------------------------------------
class AllGatherRows(torch.autograd.Function):
    """``z_local [n_p, D]`` -> ``z_full [world*B, D]`` (rows past ``n`` are zero padding).
    Backward: reduce-scatter of the partial ``dz_full`` — each rank keeps the sum for its own rows."""

    @staticmethod
    def forward(ctx, z_local, dctx, n_global):
        b = dctx.block(n_global)
        if z_local.is_cuda:
            # gather buffer from the symmetric arena; the local rows are placed by a library copy kernel.
            # Rows past n (padding of the last block) are never indexed by an edge list.
            from . import _lib, ops
            full = dctx.slot_buffer(dctx.world * b, z_local.size(1), z_local)
            if z_local.size(0) > 0:
                ops.map2d(_lib.EW_COPY, ops.M(ops._as_rows(z_local, "z")),
                          ops.M(full[dctx.rank * b: dctx.rank * b + z_local.size(0)]))
        else:                                           # gloo host tests (CPU tensors): plumbing only
            full = torch.zeros((dctx.world * b, z_local.size(1)), dtype=z_local.dtype, device=z_local.device)
            full[dctx.rank * b: dctx.rank * b + z_local.size(0)].copy_(z_local)
        dctx.all_gather_slots(full)
        ctx.dctx, ctx.n_local = dctx, z_local.size(0)
        return full

    @staticmethod
    def backward(ctx, grad_full):
        out = ctx.dctx.reduce_scatter_rows(grad_full)
        return out[: ctx.n_local], None, None


def all_gather_rows(z_local, dctx, n_global):
    return AllGatherRows.apply(z_local, dctx, n_global)


class _ReplicatedParam(torch.autograd.Function):
    """Identity on a replicated parameter whose gradient is a partial sum on every rank:
    the backward all-reduces it."""

    @staticmethod
    def forward(ctx, p, dctx):
        ctx.dctx = dctx
        return p.view_as(p)

    @staticmethod
    def backward(ctx, g):
        if ctx.dctx.defer_grad_reduce:
            return g, None
        g = g.contiguous().clone()
        ctx.dctx.all_reduce_(g)
        return g, None


def replicated(p, dctx):
    return p if dctx is None or dctx.world == 1 else _ReplicatedParam.apply(p, dctx)


class _ScaleAllReduceLoss(torch.autograd.Function):
    """Global mean loss from per-rank means: ``sum_p w_p * loss_p`` (value all-reduced for reporting;
    the gradient of this rank's term is just ``w_p``)."""

    @staticmethod
    def forward(ctx, loss_local, weight, dctx):
        ctx.weight = float(weight)
        out = loss_local.detach() * ctx.weight
        dctx.all_reduce_(out)
        return out

    @staticmethod
    def backward(ctx, g):
        return g * ctx.weight, None, None


class _LossShare(torch.autograd.Function):
    """``w_p * loss_p`` by a library kernel (deferred mode: the sum over ranks is formed later, with the
    gradient bucket, by ``DistContext.reduce_gradients``)."""

    @staticmethod
    def forward(ctx, loss_local, weight):
        from . import ops
        ctx.weight = float(weight)
        out = torch.empty(1, dtype=torch.float32, device=loss_local.device)
        ops.axpby(ops.M(loss_local.detach().view(1, 1)), ctx.weight, None, 0.0, ops.M(out.view(1, 1)))
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        from . import ops
        out = torch.empty(1, dtype=torch.float32, device=g.device)
        ops.axpby(ops.M(g.contiguous().view(1, 1)), ctx.weight, None, 0.0, ops.M(out.view(1, 1)))
        return out.view(()), None


def global_mean_loss(loss_local, n_local, n_global, dctx):
    """Global mean of a per-sample loss from this rank's mean over its ``n_local`` samples.

    ``defer_grad_reduce=False``: the value is all-reduced here (NCCL) and returned.
    ``defer_grad_reduce=True`` (CUDA): returns THIS RANK'S SHARE ``n_local / n_global * loss_local`` — its
    backward is exactly the global loss's — and parks it for ``reduce_gradients``, which sums the shares in
    rank order into ``dctx.loss_value`` with the same exchange that reduces the weight gradients."""
    w = float(n_local) / float(n_global)
    if dctx.defer_grad_reduce and loss_local.is_cuda:
        share = _LossShare.apply(loss_local, w)
        dctx._loss_share = share.detach()
        return share
    return _ScaleAllReduceLoss.apply(loss_local, w, dctx)
