"""Record the public surface of the UNMODIFIED reference ``gripnet`` package (class constructors, ``forward``
signatures, ``state_dict`` keys, ``gripnet.utils`` functions) as ``api_surface.json``.

Run in the build container only (needs ``/root/reference``):  python tests/golden/make_api_surface.py
"""
import inspect
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

L, D, U = ref_loader.load()


def params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        if p.name == "self" or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
            continue
        out.append([p.name, None if p.default is p.empty else repr(p.default)])
    return out


surface = {"classes": {}, "utils": {}, "state_dict": {}}
for mod, names in ((L, ["myGCN", "myRGCN", "homoGraph", "interGraph"]),
                   (D, ["multiRelaInnerProductDecoder", "multiClassInnerProductDecoder"])):
    for n in names:
        cls = getattr(mod, n)
        surface["classes"][n] = {"module": mod.__name__.split(".")[-1], "init": params(cls.__init__),
                                 "forward": params(cls.forward)}
surface["classes"]["myGCN"]["norm"] = params(L.myGCN.norm)
for n, fn in inspect.getmembers(U, inspect.isfunction):
    if fn.__module__ == U.__name__:
        surface["utils"][n] = params(fn)
surface["utils_constants"] = {"EPS": U.EPS}
instances = {
    "myGCN": L.myGCN(4, 3), "myRGCN": L.myRGCN(4, 3, 2, 2, False),
    "homoGraph": L.homoGraph([4, 3, 2], start_graph=True, in_dim=5),
    "homoGraph_rel": L.homoGraph([4, 3], multi_relational=True, n_rela=2, n_base=2),
    "interGraph": L.interGraph(4, 3, 6, target_feat_dim=5), "interGraph_down": L.interGraph(4, 3, 6, target_feat_dim=5),
    "multiRelaInnerProductDecoder": D.multiRelaInnerProductDecoder(4, 3),
    "multiClassInnerProductDecoder": D.multiClassInnerProductDecoder(4, 3),
}
for k, m in instances.items():
    surface["state_dict"][k] = {name: list(t.shape) for name, t in m.state_dict().items()}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_surface.json"), "w") as f:
    json.dump(surface, f, indent=1, sort_keys=True)
print(json.dumps(surface, indent=1, sort_keys=True)[:3000])
