"""Host-side helpers of the reference's ``gripnet/utils.py`` that the training
scripts call around the hot path (constants, range lists, negative sampling).

These are data-preparation utilities: kept so a script
written against ``gripnet.utils`` imports cleanly; they are not kernels.
"""
import itertools
from collections import OrderedDict

import numpy as np
import torch

EPS = 1e-13                                    # utils.py:10


def auprc_auroc_ap(target_tensor, score_tensor):
    """utils.py:28-35 — on the device for CUDA tensors (``metrics.lp_metrics``: one sort, no sklearn)."""
    from . import metrics
    return metrics.auprc_auroc_ap(target_tensor, score_tensor)


def micro_macro(target_tensor, score_tensor):
    """utils.py:38-46 — micro / macro F1 of class predictions, on the device."""
    from . import metrics
    return metrics.micro_macro(target_tensor, score_tensor)


def acc(target_tensor, score_tensor):
    """utils.py:49-52 — accuracy of class predictions, on the device."""
    from . import metrics
    return metrics.acc(target_tensor, score_tensor)


def get_range_list(edge_list, is_node=False):
    """Half-open ``[start, end)`` index ranges of consecutive blocks (utils.py:141-148)."""
    axis = 0 if is_node else 1
    bounds = np.cumsum([0] + [int(e.shape[axis]) for e in edge_list])
    return torch.tensor(np.stack([bounds[:-1], bounds[1:]], axis=1), dtype=torch.int64)


def to_bidirection(edge_index, edge_type=None):
    """Append the reversed edges (utils.py:132-138)."""
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    return both if edge_type is None else (both, torch.cat([edge_type, edge_type]))


def remove_bidirection(edge_index, edge_type=None):
    """Keep one direction (source > target) of every pair (utils.py:122-129)."""
    keep = (edge_index[0] > edge_index[1]).nonzero().view(-1)
    return edge_index[:, keep] if edge_type is None else (edge_index[:, keep], edge_type[keep])


class NegativeSampler:
    """One uniformly random NON-positive node pair per positive edge, drawn on the device
    (``gn_negsample_build`` / ``gn_negsample_draw``; reference ``gripnet/utils.py:98-119``).

    The positives are hashed once; every ``sample()`` is one kernel (counter-based Philox4x32-10, so the
    result depends on ``(seed, epoch, edge)`` only) and advances the epoch counter, which lives in device
    memory: a ``sample()`` captured in a CUDA graph yields new negatives on every replay.  With
    ``range_list`` the typed rule applies (reject only the positives of the edge's own relation slice).
    Same distribution as the reference; its own RNG stream (bit-exact restatement: ``oracle/negsample.py``).
    """

    def __init__(self, pos_edge_index, num_nodes, range_list=None, seed=None):
        from . import _lib
        from .graph import _ptr, _stream, require_cuda
        lib = _lib.load()
        require_cuda(pos_edge_index, "pos_edge_index", torch.int64)
        if pos_edge_index.dim() != 2 or pos_edge_index.size(0) != 2:
            raise RuntimeError("pos_edge_index must have shape [2, E]")
        dev = pos_edge_index.device
        self.n_edges, self.n_nodes = int(pos_edge_index.size(1)), int(num_nodes)
        ei = pos_edge_index.contiguous()
        self.range_list, self.n_rel = None, 0
        if range_list is not None:
            rl = torch.as_tensor(range_list).to(torch.int64).cpu()
            flat = rl.flatten().tolist()
            n_rel = rl.size(0)
            ok = rl.dim() == 2 and rl.size(1) == 2 and flat[0] == 0 and flat[-1] == self.n_edges and \
                all(flat[2 * r + 1] == flat[2 * r + 2] for r in range(n_rel - 1)) and \
                all(flat[2 * r] <= flat[2 * r + 1] for r in range(n_rel))
            if not ok:
                raise RuntimeError("range_list must partition [0, E) into ascending contiguous [start, end) slices")
            self.range_list, self.n_rel = rl.to(dev).contiguous(), int(n_rel)
        self.seed = int(torch.initial_seed() if seed is None else seed) & (2 ** 64 - 1)
        self.table_bytes = int(lib.gn_negsample_table_bytes(self.n_edges))
        self.table = torch.empty(max(self.table_bytes, 8), dtype=torch.uint8, device=dev)
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)          # [epoch, scratch]
        _lib.check(lib.gn_negsample_build(_ptr(ei[0]) if self.n_edges else None, _ptr(ei[1]) if self.n_edges else None,
                                          self.n_edges, self.n_nodes, _ptr(self.range_list), self.n_rel,
                                          _ptr(self.table), self.table_bytes, _stream()), "gn_negsample_build")

    def sample(self, out=None):
        """int64 ``[2, E]`` negatives of the next epoch (into ``out`` when given)."""
        from . import _lib
        from .graph import _ptr, _stream
        if out is None:
            out = torch.empty((2, self.n_edges), dtype=torch.int64, device=self.table.device)
        elif out.shape != (2, self.n_edges) or out.dtype != torch.int64 or not out.is_contiguous():
            raise RuntimeError("out must be a contiguous int64 [2, E] tensor")
        _lib.check(_lib.load().gn_negsample_draw(_ptr(self.table), self.table_bytes, self.n_edges, self.n_nodes,
                                                 _ptr(self.range_list), self.n_rel, self.seed, _ptr(self.state),
                                                 _ptr(out[0]) if self.n_edges else None,
                                                 _ptr(out[1]) if self.n_edges else None, _stream()),
                   "gn_negsample_draw")
        torch.autograd.graph.increment_version(out)      # written behind torch's back: structures cached per
        from .graph import mark_valid                    # (tensor, version) must be rebuilt for the new draw
        mark_valid(out, self.n_nodes)                    # in range by construction: no host check downstream
        return out

    @property
    def epoch(self):
        return int(self.state[0].item())


_samplers = OrderedDict()
_SAMPLER_CAPACITY = 64
_sampler_births = itertools.count()


def _fresh_seed():
    """Seed of a sampler created by the functional API.  The reference draws fresh numpy randoms on every
    call (utils.py:104-110), so two samplers must never share a stream: the seed mixes ``torch.initial_seed()``
    (reproducible under ``torch.manual_seed``) with a process-wide birth counter through SplitMix64."""
    x = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * (next(_sampler_births) + 1)) & (2 ** 64 - 1)
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
    return x ^ (x >> 31)


def _sampler_for(pos_edge_index, num_nodes, range_list):
    """Sampler cached per (edge tensor identity + version, num_nodes, range_list), LRU eviction.  A cache hit
    continues that sampler's epoch sequence; a miss (a fresh clone or slice of the positives every call, or more
    than ``_SAMPLER_CAPACITY`` live edge tensors) builds a new hash table AND a new, distinct random stream —
    pass the SAME tensor every epoch to avoid the rebuild."""
    from .graph import _Cache
    key = (_Cache.tkey(pos_edge_index), int(num_nodes),
           None if range_list is None else tuple(torch.as_tensor(range_list).flatten().tolist()))
    hit = _samplers.get(key)
    if hit is None:
        while len(_samplers) >= _SAMPLER_CAPACITY:
            _samplers.popitem(last=False)
        hit = _samplers[key] = (NegativeSampler(pos_edge_index, num_nodes, range_list, seed=_fresh_seed()),
                                pos_edge_index)
    else:
        _samplers.move_to_end(key)
    return hit[0]


def negative_sampling(pos_edge_index, num_nodes):
    """Uniform node pairs that are not positive edges, one per positive (utils.py:98-112), sampled on the
    device (``NegativeSampler``, cached per edge tensor; successive calls give successive epochs) — no
    device<->host round trip.  Same distribution as the reference; the RNG stream is this package's own
    (Philox4x32-10, seeded from ``torch.initial_seed()``).  CUDA tensors only: there is no CPU path."""
    from .graph import require_cuda
    require_cuda(pos_edge_index, "pos_edge_index", torch.int64)
    return _sampler_for(pos_edge_index, num_nodes, None).sample()


def typed_negative_sampling(pos_edge_index, num_nodes, range_list):
    """Per-relation negative sampling (utils.py:115-119): a draw is rejected only if it is a positive
    pair of the SAME ``range_list`` slice.  CUDA tensors only."""
    from .graph import require_cuda
    require_cuda(pos_edge_index, "pos_edge_index", torch.int64)
    return _sampler_for(pos_edge_index, num_nodes, range_list).sample()


# ---- host-side data preparation (gripnet/utils.py:13-25, :55-95, :151-272): not kernels; present so a script
#      written against ``gripnet.utils`` imports and prepares its splits unchanged ---------------------------
def normalize(input):
    """Rows scaled to unit L2 norm (utils.py:13-15)."""
    return input / input.pow(2).sum(dim=1, keepdim=True).sqrt()


def sparse_id(n):
    """``n x n`` sparse identity (utils.py:18-25; the scripts' one-hot node features, which ``homoGraph``
    ignores when ``start_graph=True``, layers.py:261-262)."""
    i = torch.arange(n, dtype=torch.int64)
    return torch.sparse_coo_tensor(torch.stack([i, i]), torch.ones(n), (n, n), check_invariants=False)


def load_graph(pt_file_path="./sample_graph.pt"):
    """utils.py:55-79 — opens the pickled ``torch_geometric.data.Data`` without PyG (``data.load``)."""
    from . import data
    return data.load(pt_file_path)


def load_node_idx_to_id_dict(pkl_file_path="./data/pose-1/map.pkl"):
    """utils.py:83-95."""
    import pickle
    with open(pkl_file_path, "rb") as f:
        return pickle.load(f)


def _bernoulli_split(n, p):
    keep = np.random.binomial(1, p, n).astype(bool)
    return np.flatnonzero(keep), np.flatnonzero(~keep)


def process_edge(raw_edges):
    """90/10 edge split on one direction, both halves mirrored back (utils.py:151-165)."""
    one_way = remove_bidirection(raw_edges, None)
    tr, te = _bernoulli_split(one_way.shape[1], 0.9)
    return to_bidirection(one_way[:, tr], None), to_bidirection(one_way[:, te], None)


def process_edge_multirelational(raw_edge_list, p=0.9):
    """Per-relation Bernoulli(p) split, mirrored, concatenated relation-major (utils.py:168-198): returns
    ``train_idx, train_et, train_range, test_idx, test_et, test_range``."""
    parts = {"train": [], "test": []}
    for idx in raw_edge_list:
        tr, te = _bernoulli_split(idx.shape[1], p)
        parts["train"].append(to_bidirection(idx[:, tr]))
        parts["test"].append(to_bidirection(idx[:, te]))
    out = []
    for name in ("train", "test"):
        lists = parts[name]
        et = torch.cat([torch.full((e.shape[1],), r, dtype=torch.long) for r, e in enumerate(lists)])
        out += [torch.cat(lists, dim=1), et, get_range_list(lists)]
    return tuple(out)


def process_node(raw_nodes, p=0.9):
    """Bernoulli(0.9) node split (utils.py:201-209; ``p`` is accepted and ignored there too)."""
    tr, te = _bernoulli_split(len(raw_nodes), 0.9)
    return raw_nodes[tr], raw_nodes[te]


def process_node_multilabel(raw_nodes_list):
    """Per-class node split (utils.py:212-247): ``train_idx, train_class, train_range, test_idx, test_class,
    test_range``."""
    splits = [process_node(idx) for idx in raw_nodes_list]
    out = []
    for k in (0, 1):
        lists = [s[k] for s in splits]
        cls = torch.cat([torch.full((len(v),), c, dtype=torch.long) for c, v in enumerate(lists)])
        out += [torch.cat(lists), cls, get_range_list(lists, is_node=True)]
    return tuple(out)


def process_data_multiclass(torch_tensor, n_class):
    """``[node ids; labels]`` regrouped class-major (utils.py:250-262): ``node_idx, node_class, range`` with
    ``range`` a list of ``[start, end]`` pairs."""
    nodes, labels = torch_tensor[0], torch_tensor[1]
    groups = [nodes[labels == c] for c in range(n_class)]
    sizes = [int(g.shape[0]) for g in groups]
    bounds = np.cumsum([0] + sizes).tolist()
    cls = torch.cat([torch.full((m,), c, dtype=torch.int64) for c, m in enumerate(sizes)])
    return torch.cat(groups), cls, [[bounds[c], bounds[c + 1]] for c in range(n_class)]
