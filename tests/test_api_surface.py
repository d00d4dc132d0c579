"""Drop-in boundary: the module API of ``gripnet_b200`` against the public surface of the
UNMODIFIED reference, recorded in ``tests/golden/api_surface.json`` by ``tests/golden/make_api_surface.py``:
same classes, same constructor / ``forward`` parameter names, order and defaults, same ``state_dict`` keys and
shapes, same ``gripnet.utils`` functions.  Extra trailing OPTIONAL parameters are allowed on this side."""
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SURFACE = json.load(open(os.path.join(HERE, "golden", "api_surface.json")))


def _params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        if p.name == "self" or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
            continue
        out.append([p.name, None if p.default is p.empty else repr(p.default)])
    return out


def _assert_compatible(ours, ref, what):
    assert len(ours) >= len(ref), (what, ours, ref)
    for (on, od), (rn, rd) in zip(ours, ref):
        assert on == rn, (what, on, rn)
        assert od == rd, (what, on, od, rd)
    for name, default in ours[len(ref):]:
        assert default is not None, f"{what}: extra parameter {name} must be optional"


@pytest.mark.parametrize("name", sorted(SURFACE["classes"]))
def test_class_signatures_match_the_reference(name):
    import gripnet_b200 as gb
    ref = SURFACE["classes"][name]
    mod = getattr(gb, ref["module"])
    cls = getattr(mod, name)
    assert getattr(gb, name) is cls                       # also exported at package level
    _assert_compatible(_params(cls.__init__), ref["init"], f"{name}.__init__")
    _assert_compatible(_params(cls.forward), ref["forward"], f"{name}.forward")
    if "norm" in ref:
        _assert_compatible(_params(cls.norm), ref["norm"], f"{name}.norm")


def test_utils_functions_match_the_reference():
    from gripnet_b200 import utils
    assert utils.EPS == SURFACE["utils_constants"]["EPS"]
    for name, ref in SURFACE["utils"].items():
        assert hasattr(utils, name), f"gripnet.utils.{name} is missing"
        _assert_compatible(_params(getattr(utils, name)), ref, f"utils.{name}")


def test_state_dict_keys_and_shapes_match_the_reference():
    from gripnet_b200 import decoder, layers
    ours = {
        "myGCN": layers.myGCN(4, 3), "myRGCN": layers.myRGCN(4, 3, 2, 2, False),
        "homoGraph": layers.homoGraph([4, 3, 2], start_graph=True, in_dim=5),
        "homoGraph_rel": layers.homoGraph([4, 3], multi_relational=True, n_rela=2, n_base=2),
        "interGraph": layers.interGraph(4, 3, 6, target_feat_dim=5),
        "interGraph_down": layers.interGraph(4, 3, 6, target_feat_dim=5),
        "multiRelaInnerProductDecoder": decoder.multiRelaInnerProductDecoder(4, 3),
        "multiClassInnerProductDecoder": decoder.multiClassInnerProductDecoder(4, 3),
    }
    for k, ref in SURFACE["state_dict"].items():
        got = {n: list(t.shape) for n, t in ours[k].state_dict().items()}
        assert got == ref, (k, got, ref)


def test_install_as_gripnet_aliases_the_reference_import_names():
    import importlib
    import sys
    import gripnet_b200 as gb
    saved = {k: v for k, v in sys.modules.items() if k == "gripnet" or k.startswith("gripnet.")}
    try:
        gb.install_as_gripnet()
        from gripnet.layers import homoGraph, interGraph          # the scripts' own import lines
        from gripnet.decoder import multiClassInnerProductDecoder, multiRelaInnerProductDecoder
        from gripnet.utils import EPS, micro_macro, process_data_multiclass, sparse_id
        assert homoGraph is gb.layers.homoGraph and interGraph is gb.layers.interGraph
        assert multiRelaInnerProductDecoder is gb.decoder.multiRelaInnerProductDecoder
        assert multiClassInnerProductDecoder is gb.decoder.multiClassInnerProductDecoder
        assert EPS == 1e-13 and callable(micro_macro) and callable(process_data_multiclass) and callable(sparse_id)
        assert importlib.import_module("gripnet.encoder").RGCN is gb.encoder.RGCN
    finally:
        for k in [k for k in sys.modules if k == "gripnet" or k.startswith("gripnet.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_modules_copy_and_pickle_without_their_native_graph_handles():
    """cached graphs hold ctypes handles; deepcopy / pickle must drop them, not choke on them."""
    import copy
    import io
    import torch
    import gripnet_b200 as gb
    conv = gb.myGCN(8, 4, cached=True)
    conv._graph, conv._aug, conv._graph_source, conv.cached_num_edges = object(), None, (None, 3, None), 5   # stand-ins
    c2 = copy.deepcopy(conv)
    assert c2._graph is None and c2.cached_num_edges is None and torch.equal(c2.weight, conv.weight)
    assert c2.weight is not conv.weight
    buf = io.BytesIO()
    torch.save(conv, buf)
    buf.seek(0)
    c3 = torch.load(buf, weights_only=False)
    assert c3._graph is None and torch.equal(c3.weight, conv.weight)
    h = gb.homoGraph([8, 4, 4], start_graph=True, in_dim=10)
    h.conv_list[0]._graph = object()
    h2 = copy.deepcopy(h)
    assert h2.conv_list[0]._graph is None and sorted(h2.state_dict()) == sorted(h.state_dict())


def test_every_environment_switch_is_documented():
    """Each ``GRIPNET_B200_*`` / ``GRIPNET_BENCH_*`` variable read anywhere in the package, the kernels or bench.py is
    listed in README.md's table of A/B switches (the table uses ``_SUFFIX`` shorthand for variables that share a
    prefix)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    found = set()
    for base, _, files in os.walk(os.path.join(root, "gripnet_b200")):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                found |= set(re.findall(r'"(GRIPNET_B(?:200|ENCH)_[A-Z0-9_]+)"', open(os.path.join(base, f)).read()))
    found |= set(re.findall(r'"(GRIPNET_B(?:200|ENCH)_[A-Z0-9_]+)"', open(os.path.join(root, "bench.py")).read()))
    readme = open(os.path.join(root, "README.md")).read()
    assert found, "no switches found: the scan is broken"
    for name in sorted(found):
        suffix = name.replace("GRIPNET_B200", "")
        assert name in readme or ("`" + suffix + "`") in readme, f"{name} is not documented in README.md"
