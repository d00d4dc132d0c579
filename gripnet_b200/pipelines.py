"""The reference training scripts' model wiring, as reusable containers.

``PoseModel`` = ``GripNet-pose.py:73-99`` (``Model(gg, gd, dd, dmt)``) + the
forward/loss part of its ``train()`` (``:117-142``); ``AminerModel`` =
``GripNet-aminer.py:83-108,124-133``; ``FreebaseDModel`` =
``GripNet-freebase-d.py:78-137,151-166``.  Submodule names match the scripts, so
a ``state_dict`` saved by the reference scripts loads unchanged.
Used by ``bench.py``, ``__graft_entry__.smoke()`` and the parity tests.
"""
import torch
from torch.nn import Module, Parameter

from .decoder import multiClassInnerProductDecoder, multiRelaInnerProductDecoder
from .layers import homoGraph, interGraph
from .losses import link_prediction_loss, node_classification_loss


class PoseModel(Module):
    def __init__(self, n_g, n_d, n_rel, gg=(32, 16, 16), gd=(16, 32), dd_out=32, n_base=32):
        super().__init__()
        dd = [sum(gd), dd_out]
        self.gg = homoGraph(list(gg), start_graph=True, in_dim=n_g)
        self.gd = interGraph(sum(gg), gd[0], n_d, target_feat_dim=gd[1])
        self.dd = homoGraph(dd, multi_relational=True, n_rela=n_rel, n_base=n_base)
        self.dmt = multiRelaInnerProductDecoder(sum(dd), n_rel)

    def embed(self, data):
        z = self.gg(None, data["gg_edge_index"], edge_weight=data.get("gg_edge_weight"), if_catout=True)
        z = self.gd(z, data["gd_edge_index"], mod="cat", if_relu=True)
        return self.dd(z, data["dd_edge_index"], edge_type=data["dd_edge_type"],
                       range_list=data["dd_range_list"], if_catout=True)

    def forward(self, data, neg_edge_index=None):
        """Returns (loss, z, pos_score, neg_score) for one training step's forward."""
        z = self.embed(data)
        neg = data["neg_edge_index"] if neg_edge_index is None else neg_edge_index
        pos_score = self.dmt(z, data["dd_edge_index"], data["dd_edge_type"])
        neg_score = self.dmt(z, neg, data["dd_edge_type"])
        return link_prediction_loss(pos_score, neg_score), z, pos_score, neg_score


class AminerModel(Module):
    def __init__(self, n_p, n_a, n_class, pp=(128, 64, 64), pa=(64, 64), aa_hid=(128, 32)):
        super().__init__()
        aa = [sum(pa)] + list(aa_hid)
        self.pp = homoGraph(list(pp), start_graph=True, in_dim=n_p)
        self.pa = interGraph(sum(pp), pa[0], n_a, target_feat_dim=pa[1])
        self.aa = homoGraph(aa)
        self.mcip = multiClassInnerProductDecoder(sum(aa), n_class)

    def forward(self, data):
        z = self.pp(None, data["pp_edge_index"], edge_weight=data.get("pp_edge_weight"), if_catout=True)
        z = self.pa(z, data["pa_edge_index"], if_relu=True, mod="cat")
        z = self.aa(z, data["aa_edge_index"], edge_weight=data.get("aa_edge_weight"), if_catout=True)
        score = self.mcip(z, data["train_node_idx"])
        return node_classification_loss(score, data["train_node_class"]), z, score


class FreebaseDModel(Module):
    def __init__(self, n_p, n_q, n_a, n_class, pp=(256, 128, 128), pa=(128, 128), aa_out=32):
        super().__init__()
        self.pp = homoGraph(list(pp), start_graph=True, in_dim=n_p)
        self.pa = interGraph(sum(pp), pa[0], n_a, target_feat_dim=pa[1], if_one_external=False)
        self.qq = homoGraph(list(pp), start_graph=True, in_dim=n_q)
        self.qa = interGraph(sum(pp), pa[0], n_a, target_feat_dim=pa[1], if_one_external=False)
        self.aa_embeddings = Parameter(torch.randn(n_a, pa[1]))
        self.aa = homoGraph([pa[1], aa_out])
        self.mcip = multiClassInnerProductDecoder(aa_out, n_class)

    def forward(self, data):
        z = self.pa(self.pp(None, data["pp_edge_index"], if_catout=True), data["pa_edge_index"], mod="add",
                    if_relu=True)
        z1 = self.qa(self.qq(None, data["qq_edge_index"], if_catout=True), data["qa_edge_index"], mod="add",
                     if_relu=True)
        z = self.aa((z + z1 + self.aa_embeddings) / 3, data["aa_edge_index"])       # freebase-d.py:160-164
        score = self.mcip(z, data["train_node_idx"])
        return node_classification_loss(score, data["train_node_class"]), z, score


def to_device(data, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in data.items()}


def load_flat_params(model, flat):
    """Load a flat {"gg.embedding": ..., "dmt.weight": ...} dict (oracle / reference naming)."""
    sd = model.state_dict()
    missing = set(sd.keys()) ^ set(flat.keys())
    if missing:
        raise KeyError(f"parameter name mismatch: {sorted(missing)}")
    model.load_state_dict({k: torch.as_tensor(v) for k, v in flat.items()})
    return model
