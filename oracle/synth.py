"""Parameter sets + re-exported synthetic supergraphs for the oracle / tests.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The graph generators live in
the neutral top-level ``synthdata`` module (the benchmark's own arm needs them and may not import
``oracle/``; the reference arm and this package may not import the product); this module adds parameter dictionaries drawn from the reference's
initial distributions under the reference's ``state_dict`` names.
"""
import numpy as np
import torch

from synthdata import (  # noqa: F401
    aminer_full, aminer_small, freebase_d_full, freebase_d_small, nc_graph, pose2_graph, pose2_rel_sizes,
    pose_edges_per_epoch, pose_graph, pose_medium, pose_small)


# ---------------------------------------------------------------------------
# parameters: same distributions as the reference initialisers,
# drawn from a private generator (stream parity with the reference is NOT a goal)
# ---------------------------------------------------------------------------
def _gcn(gen, i, o, prefix, p):
    b = float(np.sqrt(6.0 / (i + o)))                               # layers.py:42-44
    p[prefix + "weight"] = (torch.rand(i, o, generator=gen) * 2 - 1) * b
    p[prefix + "bias"] = torch.zeros(o)                              # layers.py:46-47


def _rgcn(gen, i, o, n_rel, n_base, after_relu, prefix, p):
    std = 2.0 / i if after_relu else 1.0 / np.sqrt(i)                # layers.py:154-160
    p[prefix + "basis"] = torch.randn(n_base, i, o, generator=gen) * std
    p[prefix + "att"] = torch.randn(n_rel, n_base, generator=gen) / np.sqrt(n_base)  # :152
    p[prefix + "root"] = torch.randn(i, o, generator=gen) * std


def homo_params(gen, nhids, prefix, p, in_dim=None, n_rel=None, n_base=32):
    if in_dim is not None:
        p[prefix + "embedding"] = torch.randn(in_dim, nhids[0], generator=gen)   # layers.py:249-250
    for l in range(len(nhids) - 1):
        k = prefix + f"conv_list.{l}."
        if n_rel is None:
            _gcn(gen, nhids[l], nhids[l + 1], k, p)
        else:
            _rgcn(gen, nhids[l], nhids[l + 1], n_rel, n_base, l > 0, k, p)    # layers.py:232


def inter_params(gen, src_dim, tgt_dim, n_target, tf_dim, prefix, p, if_one_external=True):
    if if_one_external:
        p[prefix + "target_feat"] = torch.randn(n_target, tf_dim, generator=gen)  # layers.py:343,359-360
        if tgt_dim != tf_dim:
            p[prefix + "target_feat_down"] = torch.randn(tf_dim, tgt_dim, generator=gen)  # :348-353
    _gcn(gen, src_dim, tgt_dim, prefix + "conv.", p)


def pose_params(g, seed=7, gg=(32, 16, 16), gd=(16, 32), dd_out=32, n_base=32, bias_jitter=True):
    """Parameters of the pose model (GripNet-pose.py:86-99).  ``bias_jitter`` replaces the
    all-zero GCN biases by small random values so bias handling is actually tested."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    homo_params(gen, list(gg), "gg.", p, in_dim=g["n_g"])
    inter_params(gen, sum(gg), gd[0], g["n_d"], gd[1], "gd.", p)
    dd = [sum(gd), dd_out]
    homo_params(gen, dd, "dd.", p, n_rel=g["n_rel"], n_base=n_base)
    d = sum(dd)
    p["dmt.weight"] = torch.randn(g["n_rel"], d, generator=gen) / np.sqrt(d)     # decoder.py:25-26
    if bias_jitter:
        for k in p:
            if k.endswith(".bias"):
                p[k] = (torch.rand(p[k].shape, generator=gen) - 0.5) * 0.2
    return p


def aminer_params(g, seed=7, pp=(128, 64, 64), pa=(64, 64), aa_hid=(128, 32), bias_jitter=True):
    """GripNet-aminer.py:96-108."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    homo_params(gen, list(pp), "pp.", p, in_dim=g["n_p"])
    inter_params(gen, sum(pp), pa[0], g["n_a"], pa[1], "pa.", p)
    aa = [sum(pa)] + list(aa_hid)
    homo_params(gen, aa, "aa.", p)
    d = sum(aa)
    b = float(np.sqrt(6.0 / (d + g["n_class"])))                                 # decoder.py:47-49
    p["mcip.weight"] = (torch.rand(d, g["n_class"], generator=gen) * 2 - 1) * b
    if bias_jitter:
        for k in p:
            if k.endswith(".bias"):
                p[k] = (torch.rand(p[k].shape, generator=gen) - 0.5) * 0.2
    return p


def freebase_d_params(g, seed=7, pp=(256, 128, 128), pa=(128, 128), aa_out=32, bias_jitter=True):
    """GripNet-freebase-d.py:103-137."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    homo_params(gen, list(pp), "pp.", p, in_dim=g["n_p"])
    inter_params(gen, sum(pp), pa[0], g["n_a"], pa[1], "pa.", p, if_one_external=False)
    homo_params(gen, list(pp), "qq.", p, in_dim=g["n_q"])
    inter_params(gen, sum(pp), pa[0], g["n_a"], pa[1], "qa.", p, if_one_external=False)
    p["aa_embeddings"] = torch.randn(g["n_a"], pa[1], generator=gen)
    homo_params(gen, [pa[1], aa_out], "aa.", p)
    b = float(np.sqrt(6.0 / (aa_out + g["n_class"])))
    p["mcip.weight"] = (torch.rand(aa_out, g["n_class"], generator=gen) * 2 - 1) * b
    if bias_jitter:
        for k in p:
            if k.endswith(".bias"):
                p[k] = (torch.rand(p[k].shape, generator=gen) - 0.5) * 0.2
    return p
