"""Host-side helpers of the reference's ``gripnet/utils.py`` that the training
scripts call around the hot path (constants, range lists, negative sampling).

These are data-preparation utilities: kept so a script
written against ``gripnet.utils`` imports cleanly; they are not kernels.
"""
import numpy as np
import torch

EPS = 1e-13                                    # utils.py:10


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


def negative_sampling(pos_edge_index, num_nodes, generator=None):
    """Uniform node pairs that are not positive edges, one per positive (utils.py:98-112).

    Same distribution as the reference (rejection sampling over ``num_nodes**2``
    pair codes); the RNG stream is this function's own.  Returns int64 ``[2, E]`` on
    the device of ``pos_edge_index``.
    """
    rs = generator if generator is not None else np.random
    pos = pos_edge_index.detach().cpu().numpy().astype(np.int64)
    taken = np.unique(pos[0] * num_nodes + pos[1])
    code = rs.randint(0, num_nodes * num_nodes, size=pos.shape[1]).astype(np.int64)
    bad = np.isin(code, taken)
    while bad.any():
        code[bad] = rs.randint(0, num_nodes * num_nodes, size=int(bad.sum()))
        bad = np.isin(code, taken)
    out = torch.from_numpy(np.stack([code // num_nodes, code % num_nodes]))
    return out.to(pos_edge_index.device)


def typed_negative_sampling(pos_edge_index, num_nodes, range_list, generator=None):
    """Per-relation negative sampling (utils.py:115-119)."""
    parts = [negative_sampling(pos_edge_index[:, int(s):int(e)], num_nodes, generator) for s, e in range_list]
    return torch.cat(parts, dim=1)
