"""Helpers to read the committed golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _t(a):
    a = np.asarray(a)
    return torch.from_numpy(a.copy()) if a.ndim else a.item()


def load_case(name):
    """Full-model fixture -> dict(in=..., p=..., out=..., grad=...) of torch tensors / scalars."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = {"in": {}, "p": {}, "out": {}, "grad": {}}
    for k in z.files:
        sec, key = k.split(".", 1)
        case[sec][key] = _t(z[k])
    return case


def load_grouped(name):
    """Fixture whose keys are '<case>.<field>' -> {case: {field: tensor}}."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        case, field = k.split(".", 1)
        out.setdefault(case, {})[field] = _t(z[k])
    return out


def rel_err(a, b):
    """max|a-b| / max|b| per tensor (SURVEY.md §7 'tolerance definition')."""
    a = torch.as_tensor(a).detach().to("cpu", torch.float64)
    b = torch.as_tensor(b).detach().to("cpu", torch.float64)
    if a.shape != b.shape:
        return float("inf")
    if b.numel() == 0:
        return 0.0
    d = float((a - b).abs().max())
    return d / max(float(b.abs().max()), 1e-30)
