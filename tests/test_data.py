"""Dataset loader without PyG: a file pickled as ``torch_geometric.data.data.Data`` (as the reference's
datasets are) opens with ``gripnet_b200.data.load`` while ``torch_geometric`` is not importable."""
import sys
import types

import pytest
import torch


def _write_fake_pyg_file(path, style):
    """Pickle an object whose class lives in a module NAMED torch_geometric.data.data, then forget the module."""
    pkg, sub, mod = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.data"), \
        types.ModuleType("torch_geometric.data.data")

    class Data:                                   # PyG 1.x layout: attributes in the instance __dict__
        pass
    Data.__module__, Data.__qualname__ = "torch_geometric.data.data", "Data"
    mod.Data = Data
    sys.modules.update({"torch_geometric": pkg, "torch_geometric.data": sub, "torch_geometric.data.data": mod})
    try:
        d = Data()
        fields = dict(n_g_node=30, n_d_node=8, n_dd_edge_type=2, n_gg_edge=40,
                      gg_edge_index=torch.randint(0, 30, (2, 40), dtype=torch.int32),
                      gd_edge_index=torch.stack([torch.randint(0, 30, (12,)), torch.randint(0, 8, (12,))]),
                      train_idx=torch.randint(0, 8, (2, 20)), train_et=torch.arange(20) // 10,
                      train_range=torch.tensor([[0, 10], [10, 20]]),
                      test_idx=torch.randint(0, 8, (2, 6)), test_et=torch.arange(6) // 3,
                      test_range=torch.tensor([[0, 3], [3, 6]]), x=None, edge_index=None)
        if style == "v1":
            d.__dict__.update(fields)
        else:                                     # PyG 2.x layout: a storage object under _store
            class GlobalStorage:
                pass
            GlobalStorage.__module__, GlobalStorage.__qualname__ = "torch_geometric.data.storage", "GlobalStorage"
            st_mod = types.ModuleType("torch_geometric.data.storage")
            st_mod.GlobalStorage = GlobalStorage
            sys.modules["torch_geometric.data.storage"] = st_mod
            st = GlobalStorage()
            st.__dict__["_mapping"] = fields
            d.__dict__["_store"] = st
        torch.save(d, path)
        return fields
    finally:
        for k in [k for k in sys.modules if k == "torch_geometric" or k.startswith("torch_geometric.")]:
            del sys.modules[k]


@pytest.mark.parametrize("style", ["v1", "v2"])
def test_load_reference_style_dataset_without_pyg(tmp_path, style):
    from gripnet_b200 import data as gd
    path = str(tmp_path / "pose-x.pt")
    fields = _write_fake_pyg_file(path, style)
    with pytest.raises(ModuleNotFoundError):
        import torch_geometric  # noqa: F401
    with pytest.raises(Exception):
        torch.load(path, weights_only=False)              # the stock loader needs the package
    d = gd.load(path)
    assert isinstance(d, gd.Data) and d.n_g_node == 30 and "x" not in d.keys
    assert torch.equal(d.gg_edge_index, fields["gg_edge_index"]) and torch.equal(d["train_range"], fields["train_range"])
    inp = gd.pose_inputs(d)
    assert inp["gg_edge_index"].dtype == torch.int64 and inp["n_rel"] == 2 and inp["dd_edge_index"].shape == (2, 20)
    assert gd.pose_inputs(d, "test")["dd_edge_type"].shape == (6,)
    # plain-dict round trip
    out = str(tmp_path / "plain.pt")
    gd.save(d, out)
    back = gd.load(out)
    assert sorted(back.keys) == sorted(d.keys) and torch.equal(back.train_idx, d.train_idx)
    assert isinstance(torch.load(out, weights_only=True), dict)


def test_data_container_surface():
    from gripnet_b200.data import Data, nc_inputs
    d = Data.from_dict({"a": torch.zeros(2), "b": [torch.ones(1), torch.ones(2)], "c": 3, "none": None})
    assert sorted(d.keys) == ["a", "b", "c"] and "a" in d and "none" not in d
    d.to("cpu")
    d["e"] = 5
    assert d.e == 5 and dict(iter(d))["c"] == 3 and "Data(" in repr(d)
    nc = Data(n_a_node=4, n_p_node=6, n_a_type=3, pp_edge_idx=torch.zeros(2, 3, dtype=torch.int32),
              pa_edge_idx=torch.zeros(2, 2), aa_edge_idx=torch.zeros(2, 2), train_node_idx=torch.arange(2),
              train_node_class=torch.zeros(2), pp_edge_weight=torch.ones(3, dtype=torch.float64))
    inp = nc_inputs(nc)
    assert inp["n_a"] == 4 and inp["n_p"] == 6 and "n_q" not in inp and inp["pp_edge_index"].dtype == torch.int64
    assert inp["pp_edge_weight"].dtype == torch.float32 and inp["train_node_class"].dtype == torch.int64


def test_host_utils_of_the_reference_surface():
    import numpy as np
    from gripnet_b200 import utils
    x = torch.tensor([[3.0, 4.0], [0.0, 2.0]])
    assert torch.allclose(utils.normalize(x), torch.tensor([[0.6, 0.8], [0.0, 1.0]]))
    assert torch.equal(utils.sparse_id(3).to_dense(), torch.eye(3))
    idx, cls, rng = utils.process_data_multiclass(torch.tensor([[10, 11, 12, 13, 14], [1, 0, 1, 2, 0]]), 3)
    assert idx.tolist() == [11, 14, 10, 12, 13] and cls.tolist() == [0, 0, 1, 1, 2] and rng == [[0, 2], [2, 4], [4, 5]]
    np.random.seed(0)
    raw = [torch.stack([torch.arange(50) + 1, torch.zeros(50, dtype=torch.long)]) for _ in range(3)]
    tr, tr_et, tr_rng, te, te_et, te_rng = utils.process_edge_multirelational(raw)
    assert tr.shape[1] + te.shape[1] == 2 * 150 and tr_rng.shape == (3, 2) and int(tr_rng[-1, 1]) == tr.shape[1]
    assert tr_et.shape[0] == tr.shape[1] and int(te_rng[-1, 1]) == te_et.shape[0]
    for r, (s, e) in enumerate(tr_rng.tolist()):
        assert (tr_et[s:e] == r).all() and (e - s) % 2 == 0
        half = (e - s) // 2
        assert torch.equal(tr[:, s:s + half], tr[:, s + half:e].flip(0))       # mirrored halves
    both = utils.to_bidirection(torch.tensor([[5, 7, 9], [1, 2, 3]]))
    a, b = utils.process_edge(both)
    assert a.shape[1] + b.shape[1] == 6
    out = utils.process_node_multilabel([torch.arange(40), torch.arange(40, 70)])
    assert out[0].shape[0] + out[3].shape[0] == 70 and int(out[2][-1, 1]) == out[0].shape[0]
