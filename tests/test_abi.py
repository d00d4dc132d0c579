"""The C-ABI library loads on a machine without a GPU and exports exactly what include/gripnet_b200.h declares
(no compute calls here: those are the -m gpu tests)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gripnet_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from gripnet_b200 import _lib
    names = _declared()
    assert len(names) >= 45 and "gn_spmm" in names and "gn_adam_step" in names and "gn_lp_metrics" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported by the library: {missing}"
    assert sorted(_lib.SIGNATURES) == names, (sorted(set(_lib.SIGNATURES) ^ set(names)))


def test_host_only_entry_points():
    from gripnet_b200 import _lib
    lib = _lib.load()
    assert lib.gn_version() >= 100
    assert lib.gn_error_string(0) == b"ok"
    assert b"workspace" in lib.gn_error_string(-3)
    assert lib.gn_launch_count() >= 0
    assert lib.gn_peer_max_world() >= 8
    assert lib.gn_adam_max_tensors_per_launch() >= 8
    # workspace queries are pure host arithmetic and grow with the problem
    small = lib.gn_gcn_prep_workspace_bytes(1000, 100, 100)
    big = lib.gn_gcn_prep_workspace_bytes(1000000, 100000, 100000)
    assert 0 < small < big
    assert 0 < lib.gn_lp_metrics_workspace_bytes(10, 10, 2) < lib.gn_lp_metrics_workspace_bytes(100000, 100000, 16)
    assert lib.gn_nc_metrics_workspace_bytes(8) >= 8 * 8 * 4 + 4
    assert lib.gn_negsample_table_bytes(1000) > 0


def test_struct_mirrors_match_the_header_layout():
    from gripnet_b200 import _lib
    assert ctypes.sizeof(_lib.GnCsr) == 6 * 4 + 7 * 8
    assert ctypes.sizeof(_lib.GnAdamTensor) == 4 * 8 + 8
