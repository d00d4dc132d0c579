"""Drop-in ``gripnet.encoder`` (reference ``gripnet/encoder.py:6-25``).

The reference class cannot run: its ``forward`` reads ``self.embed`` while the
constructor defines ``self.embedding`` (``encoder.py:11,21``).  The class, its
constructor signature and ``state_dict`` keys are kept; ``forward`` does what the
reference evidently intends (project, then two myRGCN layers, no activation
between them) so the module is usable.
"""
import torch
from torch.nn import Module, Parameter

from . import ops
from .layers import myRGCN


class RGCN(Module):
    def __init__(self, feat_dim, r1_in_dim, r1_out_dim, r2_out_dim, n_relations, n_bases):
        super().__init__()
        self.embedding = Parameter(torch.empty(feat_dim, r1_in_dim))
        with torch.no_grad():
            self.embedding.normal_()
        self.rgcn1 = myRGCN(r1_in_dim, r1_out_dim, n_relations, n_bases, after_relu=False)
        self.rgcn2 = myRGCN(r1_out_dim, r2_out_dim, n_relations, n_bases, after_relu=True)

    def forward(self, x, edge_index, edge_et, edge_range):
        x = ops.matmul(x, self.embedding)
        x = self.rgcn1(x, edge_index, edge_et, edge_range)
        return self.rgcn2(x, edge_index, edge_et, edge_range)
