"""Dataset container and loader without PyG.

The reference's datasets (``datasets/pose/pose-{0,1,2}.pt``, ``datasets/aminer.pt`` …, ``README.md:37-51``) are
pickled ``torch_geometric.data.Data`` objects: ``torch.load`` needs the ``torch_geometric`` package only to
resolve that one class name.  ``load`` unpickles them with every ``torch_geometric.*`` class mapped to the
plain attribute bag ``Data`` below, so the files open on a machine without PyG.  ``Data`` keeps the handful of
methods the training scripts call (``GripNet-pose.py:39-71``: attribute access, ``.to(device)``,
``Data.from_dict``).  ``pose_inputs`` / ``nc_inputs`` turn a loaded dataset into the ``data`` dict of
``pipelines.PoseModel`` / ``AminerModel`` (field names: ``GripNet-pose.py:50-55,117-131``;
``GripNet-aminer.py:47-65``).
"""
import pickle
import types

import torch


class Data:
    """Attribute bag standing in for ``torch_geometric.data.Data`` (PyG 1.x: a plain object whose
    ``__dict__`` holds the tensors; PyG 2.x: a ``_store`` mapping)."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_dict(cls, dictionary):
        return cls(**dictionary)

    # -- unpickling: PyG 1.x pickles the instance __dict__; PyG 2.x nests it under "_store" ---------------
    def __setstate__(self, state):
        if isinstance(state, tuple):                       # (dict, slots-dict)
            state = {**(state[0] or {}), **(state[1] or {})}
        store = state.get("_store") if isinstance(state, dict) else None
        if store is not None:
            inner = getattr(store, "__dict__", {})
            state = dict(inner.get("_mapping", inner))
        self.__dict__.update({k: v for k, v in state.items() if not (k.startswith("__") and k.endswith("__"))})

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys

    def __iter__(self):
        for k in sorted(self.keys):
            yield k, getattr(self, k)

    def to(self, device, *keys):
        """Move every tensor attribute (or the named ones) — ``data = data.to(device)``, GripNet-pose.py:63."""
        for k in (keys or self.keys):
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
            elif isinstance(v, (list, tuple)) and v and all(torch.is_tensor(t) for t in v):
                setattr(self, k, type(v)(t.to(device) for t in v))
        return self

    def to_dict(self):
        return {k: getattr(self, k) for k in self.keys}

    def __repr__(self):
        def d(v):
            return list(v.shape) if torch.is_tensor(v) else (f"[{len(v)}]" if isinstance(v, (list, tuple, dict)) else v)
        return "Data(" + ", ".join(f"{k}={d(v)}" for k, v in self) + ")"


class _Stub:
    """Any other torch_geometric class met inside a pickle (storages, batch helpers): state kept, no code."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch_geometric" or module.startswith("torch_geometric."):
            return Data if name in ("Data", "HeteroData", "Batch") else type(name, (_Stub,), {})
        return super().find_class(module, name)


_pickle_module = types.ModuleType("gripnet_b200._pickle")
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **kw).load()
_pickle_module.loads = pickle.loads
_pickle_module.dump, _pickle_module.dumps, _pickle_module.Pickler = pickle.dump, pickle.dumps, pickle.Pickler
_pickle_module.__name__ = "pickle"


def load(path, map_location="cpu"):
    """``torch.load`` of a reference dataset (``gripnet/utils.py:55-79`` ``load_graph``; the scripts call
    ``torch.load("datasets/pose/pose-0.pt")``) without ``torch_geometric`` installed.  Plain dict files written
    by ``save`` come back as ``Data`` too.  Datasets are trusted local files: this unpickles arbitrary objects,
    like the reference's ``torch.load`` does."""
    obj = torch.load(path, map_location=map_location, pickle_module=_pickle_module, weights_only=False)
    if isinstance(obj, dict):
        obj = Data.from_dict(obj)
    return obj


def save(data, path):
    """Write a dataset as a plain ``{name: value}`` dict — readable by ``load`` and by bare ``torch.load``."""
    torch.save(data.to_dict() if isinstance(data, Data) else dict(data), path)


def _long(t):
    return t.to(torch.int64).contiguous()


def pose_inputs(data, split="train"):
    """``pipelines.PoseModel`` inputs from a pose dataset (fields used by ``GripNet-pose.py:50-55,95-98,117-131``:
    ``n_g_node, n_d_node, n_dd_edge_type, gg_edge_index, gd_edge_index, {train,test}_{idx,et,range}``)."""
    d = {
        "n_g": int(data.n_g_node), "n_d": int(data.n_d_node), "n_rel": int(data.n_dd_edge_type),
        "gg_edge_index": _long(data.gg_edge_index), "gd_edge_index": _long(data.gd_edge_index),
        "dd_edge_index": _long(getattr(data, split + "_idx")), "dd_edge_type": _long(getattr(data, split + "_et")),
        "dd_range_list": _long(torch.as_tensor(getattr(data, split + "_range"))),
    }
    w = getattr(data, "edge_weight", None)
    if w is not None:
        d["gg_edge_weight"] = w.to(torch.float32)
    return d


def nc_inputs(data, split="train"):
    """``pipelines.AminerModel`` / ``FreebaseDModel`` inputs (``GripNet-aminer.py:47-65``,
    ``GripNet-freebase-d.py:60-66``: ``n_{a,p,q}_node, n_a_type, {pp,pa,aa,qq,qa}_edge_idx,
    {train,test}_node_{idx,class}``)."""
    d = {"n_class": int(data.n_a_type), "train_node_idx": _long(getattr(data, split + "_node_idx")),
         "train_node_class": _long(getattr(data, split + "_node_class"))}
    for v in ("a", "p", "q"):
        n = getattr(data, f"n_{v}_node", None)
        if n is not None:
            d["n_" + v] = int(n)
    for e in ("pp", "pa", "aa", "qq", "qa"):
        idx = getattr(data, e + "_edge_idx", None)
        if idx is not None:
            d[e + "_edge_index"] = _long(idx)
        w = getattr(data, e + "_edge_weight", None)
        if w is not None:
            d[e + "_edge_weight"] = w.to(torch.float32)
    return d
