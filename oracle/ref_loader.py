"""Import the UNMODIFIED reference ``gripnet`` modules from ``/root/reference``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works only where the
read-only reference checkout exists (the build container); on the GPU box
``available()`` is False and everything falls back to the committed golden
fixtures + ``oracle.port``.  No reference source is copied: the files are
executed from where they lie, under private module names so they never shadow
the product's ``gripnet`` alias.
"""
import importlib.util
import os
import sys
import types

from . import pyg_shim

REF_ROOT = os.environ.get("GRIPNET_REFERENCE_ROOT", "/root/reference")
_PKG = "_gripnet_reference"


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "gripnet", "layers.py"))


def _load(name):
    full = f"{_PKG}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [os.path.join(REF_ROOT, "gripnet")]
        sys.modules[_PKG] = pkg
    path = os.path.join(REF_ROOT, "gripnet", name + ".py")
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    setattr(sys.modules[_PKG], name, mod)
    return mod


def load():
    """Return (layers, decoder, utils) modules of the reference package."""
    if not available():
        raise RuntimeError(f"reference checkout not found under {REF_ROOT}")
    pyg_shim.install()
    import torch
    state = torch.random.get_rng_state()  # the reference reseeds at import (layers.py:11-12)
    try:
        return _load("layers"), _load("decoder"), _load("utils")
    finally:
        torch.random.set_rng_state(state)
