"""Drop-in ``gripnet.layers``: myGCN, myRGCN, homoGraph, interGraph on the CUDA path.

Same constructor arguments, ``forward`` signatures, parameter names / shapes
(``state_dict`` keys) and initial distributions as the reference
(``/root/reference/gripnet/layers.py``; per-class line references below), but no
PyG / torch_scatter: every forward/backward runs hand-written sm_100a kernels
through ``gripnet_b200.ops``.  Inputs must be CUDA fp32 tensors — there is no CPU
fallback (a CPU tensor raises ``RuntimeError``).
"""
import math

import torch
from torch.nn import Module, ModuleList, Parameter

from . import graph as G
from . import ops


def _glorot_uniform_(t):
    bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-bound, bound)


class myGCN(Module):
    """GCN convolution ``A_hat (x W) + b`` with ``A_hat = D^-1/2 (A + I) D^-1/2``.

    Reference: ``layers.py:15-105``.  ``D`` is the weighted in-degree by TARGET;
    existing self-loops are replaced, one loop per node is appended
    (``norm``, ``layers.py:52-69``).  With ``cached=True`` the preprocessed graph
    of the first call is reused and only the edge COUNT of later calls is
    checked (``layers.py:75-90``), exactly like the reference.
    """

    def __init__(self, in_channels, out_channels, improved=False, cached=False, bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.weight = Parameter(torch.empty(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self._graph = None
        self._aug = None
        self.cached_num_edges = None
        self.reset_parameters()

    # cached graphs hold native handles (ctypes structs over device arrays): they are rebuilt on demand, never
    # copied or pickled — copy.deepcopy(model) / torch.save(model) keep working after the first forward
    _TRANSIENT = ("_graph", "_aug", "_graph_source")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._TRANSIENT:
            if k in state:
                state[k] = None
        state["cached_num_edges"] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            setattr(new, k, copy.deepcopy(v, memo))
        return new

    def reset_parameters(self):
        _glorot_uniform_(self.weight)                 # layers.py:42-44
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()
        self._graph, self._aug, self.cached_num_edges = None, None, None

    # -- ``cached_result`` mirrors the reference attribute: (edge_index', norm) in the
    #    reference's edge order, materialised lazily; assigning None drops the cache.
    @property
    def cached_result(self):
        if self._graph is None:
            return None
        if self._aug is None and not self._graph.bipartite:
            src = self._graph_source
            self._aug = myGCN.norm(src[0], src[1], src[2], self.improved)
        return self._aug

    @cached_result.setter
    def cached_result(self, value):
        if value is not None:
            raise AttributeError("cached_result can only be reset to None")
        self._graph, self._aug, self.cached_num_edges = None, None, None

    @staticmethod
    def norm(edge_index, num_nodes, edge_weight, improved=False, dtype=None):
        """(edge_index', norm) exactly as ``layers.py:52-69`` orders them (K1 kernels)."""
        g = G.GcnGraph(edge_index, num_nodes, num_nodes, edge_weight, improved, bipartite=False, want_aug=True)
        nrm = g.aug_norm
        if dtype is not None and dtype != torch.float32:
            nrm = nrm.to(dtype)
        return g.aug_edge_index, nrm

    def _resolve_graph(self, edge_index, n_src, n_dst, edge_weight, bipartite):
        G.require_cuda(edge_index, "edge_index", torch.int64)
        if self.cached and self._graph is not None:
            if edge_index.size(1) != self.cached_num_edges:
                raise RuntimeError("Cached {} number of edges, but found {}".format(
                    self.cached_num_edges, edge_index.size(1)))
            return self._graph
        g = G.gcn_graph(edge_index, n_src, n_dst, edge_weight, self.improved, bipartite)
        self.cached_num_edges = edge_index.size(1)
        self._graph_source = (edge_index, n_dst, edge_weight)
        self._graph, self._aug = g, None
        return g

    def forward(self, x, edge_index, edge_weight=None):
        g = self._resolve_graph(edge_index, x.size(0), x.size(0), edge_weight, False)
        return ops.GcnStack.apply(x, g, (False,), False, self.weight, self.bias)

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)


class myRGCN(Module):
    """Basis-decomposed relational convolution with ONE joint mean over all in-edges.

    Reference: ``layers.py:108-205``.  ``W_r = sum_b att[r,b] basis[b]``; the relation
    of an edge is defined by ``range_list`` only (``edge_type`` is ignored there,
    ``layers.py:178-186``); ``out_i = mean_{e->i} x[src_e] W_{r(e)} + x_i root (+ bias)``.
    """

    def __init__(self, in_channels, out_channels, num_relations, num_bases, after_relu, bias=False, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases, self.after_relu = num_relations, num_bases, after_relu
        self.basis = Parameter(torch.empty(num_bases, in_channels, out_channels))
        self.att = Parameter(torch.empty(num_relations, num_bases))
        self.root = Parameter(torch.empty(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            self.att.normal_(std=1.0 / math.sqrt(self.num_bases))            # layers.py:152
            std = 2.0 / self.in_channels if self.after_relu else 1.0 / math.sqrt(self.in_channels)  # :154-160
            self.root.normal_(std=std)
            self.basis.normal_(std=std)
            if self.bias is not None:
                self.bias.zero_()

    def forward(self, x, edge_index, edge_type, range_list):
        g = G.rgcn_graph(edge_index, range_list, x.size(0), self.num_relations)
        return ops.RgcnStack.apply(x, g, (False,), False, self.basis, self.att, self.root, self.bias)

    def __repr__(self):
        return "{}({}, {}, num_relations={})".format(self.__class__.__name__, self.in_channels, self.out_channels,
                                                     self.num_relations)


class homoGraph(Module):
    """Internal module of one supervertex: optional learned embedding + GCN/RGCN stack.

    Reference: ``layers.py:208-318``.  ReLU follows every layer, the last included;
    ``if_catout`` returns ``cat([x0, h1, ..., hL], dim=1)``.  The whole stack runs as
    one autograd node: each layer writes straight into its column slice of the
    output buffer and the backward fuses ReLU masks and slice gradients.
    """

    def __init__(self, nhid_list, requires_grad=True, start_graph=False, in_dim=None, multi_relational=False,
                 n_rela=None, n_base=32):
        super().__init__()
        self.multi_relational, self.start_graph = multi_relational, start_graph
        self.out_dim = nhid_list[-1]
        self.n_cov = len(nhid_list) - 1
        if start_graph:
            self.embedding = Parameter(torch.empty(in_dim, nhid_list[0]))
            self.embedding.requires_grad = requires_grad
            self.reset_parameters()
        pairs = list(zip(nhid_list[:-1], nhid_list[1:]))
        if multi_relational:
            assert n_rela is not None
            self.conv_list = ModuleList(myRGCN(i, o, n_rela, n_base, after_relu=(l > 0))
                                        for l, (i, o) in enumerate(pairs))
        else:
            self.conv_list = ModuleList(myGCN(i, o, cached=True) for i, o in pairs)

    def reset_parameters(self):
        with torch.no_grad():
            self.embedding.normal_()                                  # layers.py:249-250

    def prologue(self, n_rows):
        """Start everything of the next ``forward`` that depends on the parameters only (the relational layers'
        ``W[r]`` and their tensor-core operand images for an ``n_rows``-row input) on a background stream.  Optional:
        a training step calls it before its first kernel; ``forward`` picks the results up (``ops.RelPrologue``)."""
        convs = list(self.conv_list)
        if self.multi_relational and all(type(c) is myRGCN for c in convs):
            return ops.RelPrologue([(c.basis, c.att) for c in convs], int(n_rows))
        return None

    def forward(self, x, homo_edge_index, edge_weight=None, edge_type=None, range_list=None, if_catout=False):
        if self.start_graph:
            x = self.embedding                                        # the input x is ignored, layers.py:261-262
        convs = list(self.conv_list)
        relu = (True,) * len(convs)
        if self.multi_relational:
            assert edge_type is not None
            assert range_list is not None
            if all(type(c) is myRGCN for c in convs):
                g = G.rgcn_graph(homo_edge_index, range_list, x.size(0), convs[0].num_relations)
                params = [p for c in convs for p in (c.basis, c.att, c.root, c.bias)]
                return ops.RgcnStack.apply(x, g, relu, bool(if_catout), *params)
        elif all(type(c) is myGCN for c in convs):
            # every layer shares one preprocessed graph (the reference caches an identical copy per layer)
            g = convs[0]._resolve_graph(homo_edge_index, x.size(0), x.size(0), edge_weight, False)
            for c in convs[1:]:
                if c.cached and c._graph is not None and homo_edge_index.size(1) != c.cached_num_edges:
                    raise RuntimeError("Cached {} number of edges, but found {}".format(
                        c.cached_num_edges, homo_edge_index.size(1)))
                c._graph, c._aug, c.cached_num_edges = g, None, homo_edge_index.size(1)
                c._graph_source = convs[0]._graph_source
            params = [p for c in convs for p in (c.weight, c.bias)]
            return ops.GcnStack.apply(x, g, relu, bool(if_catout), *params)
        # conv_list is an ordinary ModuleList, so a caller could swap a layer for a foreign module; this package
        # has no eager PyTorch path to run it with
        raise RuntimeError("gripnet_b200.homoGraph: conv_list must hold only this package's myGCN (or, when "
                           "multi_relational, myRGCN) layers")


class interGraph(Module):
    """External module: bipartite propagation parent supervertex -> child supervertex.

    Reference: ``layers.py:322-387``.  The reference stacks both node sets, zero-pads
    the child rows and runs myGCN over ``n_source + n_target`` nodes; here the same
    result comes from a rectangular CSR (closed form):
    ``h_t = (1 + sum_{s->t} w)^-1/2 * sum_{s->t} w (x_s W) + b`` — no padding, no
    dead self-loop work.
    """

    def __init__(self, source_dim, target_dim, n_target, target_feat_dim=32, requires_grad=True,
                 if_one_external=True):
        super().__init__()
        self.source_dim, self.target_dim = source_dim, target_dim
        self.target_feat_dim, self.n_target = target_feat_dim, n_target
        self.if_one_external = if_one_external
        if if_one_external:
            self.target_feat = Parameter(torch.empty(n_target, target_feat_dim))
            self.target_feat.requires_grad = requires_grad
            if target_dim != target_feat_dim:
                self.target_feat_down = Parameter(torch.empty(target_feat_dim, target_dim))
                self.target_feat_down.requires_grad = requires_grad
                with torch.no_grad():
                    self.target_feat_down.normal_()                   # layers.py:348-353
        self.conv = myGCN(source_dim, target_dim, cached=True)
        if if_one_external:
            self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            self.target_feat.normal_()                                # layers.py:359-360

    def forward(self, x, inter_edge_index, edge_weight=None, if_relu=True, mod="cat"):
        g = self.conv._resolve_graph(inter_edge_index, x.size(0), self.n_target, edge_weight, True)
        h = ops.GcnStack.apply(x, g, (bool(if_relu),), False, self.conv.weight, self.conv.bias)
        if not self.if_one_external:
            return h                                                  # layers.py:372-373
        down = getattr(self, "target_feat_down", None)
        return ops.InterTail.apply(h, self.target_feat, down, "cat" if mod == "cat" else "add",
                                   getattr(g, "ctx", None))
