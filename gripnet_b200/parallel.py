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
import torch
import torch.distributed as dist


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

    # ---- collectives -----------------------------------------------------
    def all_gather_slots(self, full):
        """``full`` is ``[world*B, F]`` with this rank's slot already written: complete it in place."""
        if self.world == 1:
            return full
        b = full.size(0) // self.world
        dist.all_gather_into_tensor(full, full[self.rank * b:(self.rank + 1) * b], group=self.group)
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
        """Bucketed all-reduce of the ``.grad`` of replicated parameters (``defer_grad_reduce=True``):
        pack -> one all-reduce -> unpack.  ``params`` must not contain row-partitioned parameters."""
        if self.world == 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        sizes = [g.numel() for g in grads]
        if self._bucket is None or self._bucket.numel() != sum(sizes) or self._bucket.device != grads[0].device:
            self._bucket = torch.empty(sum(sizes), dtype=grads[0].dtype, device=grads[0].device)
        torch.cat([g.reshape(-1) for g in grads], out=self._bucket)
        dist.all_reduce(self._bucket, op=dist.ReduceOp.SUM, group=self.group)
        torch._foreach_copy_([g.view(-1) for g in grads], list(self._bucket.split(sizes)))

    def reduce_scatter_rows(self, full):
        """Sum ``full`` ``[world*B, F]`` over ranks and return this rank's ``[B, F]`` block."""
        b = full.size(0) // self.world
        if self.world == 1:
            return full[:b]
        out = torch.empty((b,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=self.group)
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
    __slots__ = ("tensor", "ctx", "n_src", "n_dst")

    def __init__(self, tensor, ctx, n_src, n_dst):
        self.tensor, self.ctx, self.n_src, self.n_dst = tensor, ctx, int(n_src), int(n_dst)


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


def lookup(edge_index):
    return _registry.get(_key(edge_index))


def clear_registry():
    _registry.clear()


# --------------------------------------------------------------------------
# differentiable row all-gather (decoder input)
# --------------------------------------------------------------------------
class AllGatherRows(torch.autograd.Function):
    """``z_local [n_p, D]`` -> ``z_full [world*B, D]`` (rows past ``n`` are zero padding).
    Backward: reduce-scatter of the partial ``dz_full`` — each rank keeps the sum for its own rows."""

    @staticmethod
    def forward(ctx, z_local, dctx, n_global):
        b = dctx.block(n_global)
        full = torch.zeros((dctx.world * b, z_local.size(1)), dtype=z_local.dtype, device=z_local.device) \
            if z_local.size(0) != b else \
            torch.empty((dctx.world * b, z_local.size(1)), dtype=z_local.dtype, device=z_local.device)
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


def global_mean_loss(loss_local, n_local, n_global, dctx):
    return _ScaleAllReduceLoss.apply(loss_local, float(n_local) / float(n_global), dctx)
