"""CPU: libcgnn.so loads and exports every function include/cgnn.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "cgnn.h")
LIB = os.path.join(ROOT, "cosmology_gnn_simulation_b200", "libcgnn.so")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgnn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    names = declared_functions()
    for must in ["cgnn_knn_periodic", "cgnn_edge_features", "cgnn_csr_transpose", "cgnn_mlp_rows_fwd",
                 "cgnn_mlp_rows_bwd", "cgnn_mp_edge_fwd", "cgnn_mp_node_fwd", "cgnn_mp_edge_bwd",
                 "cgnn_mp_node_bwd", "cgnn_loss_fwd_bwd", "cgnn_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(LIB)
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in cgnn.h but not exported"
    handle.cgnn_version.restype = ctypes.c_char_p
    assert b"sm_100a" in handle.cgnn_version()


def test_python_binding_matches_header():
    from cosmology_gnn_simulation_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    _lib.lib()      # sets argtypes on every symbol: raises if one is missing


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cosmology_gnn_simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "from oracle" not in src and "import oracle" not in src, f


def test_cpu_inputs_fail_loudly():
    import torch
    from cosmology_gnn_simulation_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.knn_periodic(torch.zeros(8, 3), 1.0, 4)
