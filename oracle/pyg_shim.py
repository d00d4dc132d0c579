"""Stand-ins for the three third-party symbols ``gripnet/layers.py:3-5`` imports.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``torch_geometric<2.0`` and ``torch_scatter`` are not installable in the build
image, so the unmodified reference cannot be imported without them.  These are
restatements of their published PyG-1.x behaviour, kept
as small as the reference's three call patterns need:

* ``torch_scatter.scatter_add(src, index, dim=0, dim_size=N)``
  (used at ``gripnet/layers.py:65`` with a 1-D ``src``)
* ``torch_geometric.utils.add_remaining_self_loops``
  (used at ``gripnet/layers.py:60``)
* ``torch_geometric.nn.conv.MessagePassing`` with ``aggr`` in {"add","mean"},
  ``flow="source_to_target"`` (base class at ``gripnet/layers.py:15,108``;
  ``propagate`` call sites ``:92`` and ``:167``)

``install()`` registers them in ``sys.modules`` under the names the reference
imports.
"""
import inspect
import sys
import types

import torch


def scatter_add(src, index, dim=0, out=None, dim_size=None, fill_value=0):
    if dim != 0:
        raise NotImplementedError("shim covers dim=0 only (the reference's use)")
    if out is None:
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        out = src.new_full((dim_size,) + tuple(src.shape[1:]), fill_value)
    return out.index_add_(0, index, src)


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1, num_nodes=None):
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    new_index = torch.cat([edge_index[:, keep], loops.unsqueeze(0).repeat(2, 1)], dim=1)
    if edge_weight is None:
        return new_index, None
    loop_w = edge_weight.new_full((num_nodes,), fill_value)
    drop = ~keep
    # sequential assignment: with several self-loops on one node the last one wins
    for node, w in zip(src[drop].tolist(), edge_weight[drop].tolist()):
        loop_w[node] = w
    return new_index, torch.cat([edge_weight[keep], loop_w], dim=0)


class MessagePassing(torch.nn.Module):
    """PyG-1.x style base class: gather `*_j`/`*_i` kwargs, message, scatter, update."""

    def __init__(self, aggr="add", flow="source_to_target", **kwargs):
        super().__init__()
        if aggr not in ("add", "mean"):
            raise NotImplementedError(aggr)
        if flow != "source_to_target":
            raise NotImplementedError(flow)
        self.aggr = aggr
        self._msg_params = list(inspect.signature(self.message).parameters)
        self._upd_params = list(inspect.signature(self.update).parameters)[1:]

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        n_out = None
        args = []
        for name in self._msg_params:
            if name.endswith("_j"):
                t = kwargs[name[:-2]]
                n_out = t.size(0) if n_out is None else n_out
                args.append(t.index_select(0, src))
            elif name.endswith("_i"):
                t = kwargs[name[:-2]]
                n_out = t.size(0) if n_out is None else n_out
                args.append(t.index_select(0, dst))
            elif name == "edge_index":
                args.append(edge_index)
            else:
                args.append(kwargs[name])
        msg = self.message(*args)
        out = msg.new_zeros((n_out,) + tuple(msg.shape[1:])).index_add_(0, dst, msg)
        if self.aggr == "mean":
            cnt = torch.zeros(n_out, dtype=msg.dtype, device=msg.device)
            cnt.index_add_(0, dst, torch.ones_like(dst, dtype=msg.dtype))
            out = out / cnt.clamp(min=1).unsqueeze(-1)
        return self.update(out, **{k: kwargs[k] for k in self._upd_params})

    def message(self, x_j):  # pragma: no cover - overridden
        return x_j

    def update(self, aggr_out):  # pragma: no cover - overridden
        return aggr_out


def install():
    """Register the shim modules under the names ``gripnet/layers.py`` imports."""
    if "torch_scatter" not in sys.modules:
        m = types.ModuleType("torch_scatter")
        m.scatter_add = scatter_add
        sys.modules["torch_scatter"] = m
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg_utils = types.ModuleType("torch_geometric.utils")
        tg_utils.add_remaining_self_loops = add_remaining_self_loops
        tg_nn = types.ModuleType("torch_geometric.nn")
        tg_conv = types.ModuleType("torch_geometric.nn.conv")
        tg_conv.MessagePassing = MessagePassing
        tg_nn.conv = tg_conv
        tg.utils, tg.nn = tg_utils, tg_nn
        sys.modules.update({
            "torch_geometric": tg,
            "torch_geometric.utils": tg_utils,
            "torch_geometric.nn": tg_nn,
            "torch_geometric.nn.conv": tg_conv,
        })
