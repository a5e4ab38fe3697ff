"""The C-ABI library loads and exports every symbol include/spotv2_gat.h declares (no compute)."""
import ctypes
import os
import re

import pytest
import torch

import spotv2net_b200
from spotv2net_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "spotv2_gat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spotv2_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    lib = spotv2net_b200.load_library()
    assert lib.spotv2_abi_version() == 5


def test_every_declared_symbol_is_exported_and_bound():
    lib = spotv2net_b200.load_library()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in spotv2_gat.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in spotv2net_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names)


def test_descriptor_layout_matches_header():
    assert ctypes.sizeof(_lib.GatDesc) == 17 * 4
    assert [f[0] for f in _lib.GatDesc._fields_] == ["B", "N", "F", "Fe", "H", "C", "R", "concat",
                                                     "negative_slope", "ldp", "gemm_algo", "attn_bwd_algo",
                                                     "dropout_p", "edge_mode", "dropout_seed_lo", "dropout_seed_hi", "p_format"]


def test_ldp_query_and_argument_validation_without_a_gpu():
    lib = spotv2net_b200.load_library()
    assert lib.spotv2_gat_ldp(6, 500) == 3012 and lib.spotv2_gat_ldp(8, 256) == 2064 and lib.spotv2_gat_ldp(1, 1) == 32
    bad = _lib.GatDesc(0, 30, 1260, 126, 6, 500, 870, 0, 0.2, 3012, 0, 0)
    a = ctypes.c_size_t()
    rc = lib.spotv2_gat_workspace_bytes(ctypes.byref(bad), ctypes.byref(a), None, None)
    assert rc == 1 and b"non-positive" in lib.spotv2_last_error()
    bad = _lib.GatDesc(4, 30, 1260, 126, 6, 500, 870, 0, 0.2, 3000, 0, 0)
    assert lib.spotv2_gat_workspace_bytes(ctypes.byref(bad), ctypes.byref(a), None, None) == 1
    assert b"ldp" in lib.spotv2_last_error()


def test_product_path_refuses_cpu_tensors():
    layer = spotv2net_b200.GATConv(8, 4, heads=2, edge_dim=3)
    ei = spotv2net_b200.complete_graph_edge_index(5)
    with pytest.raises(spotv2net_b200.SpotV2Error, match="no CPU fallback"):
        layer(torch.zeros(5, 8), ei, torch.zeros(20, 3))


def test_missing_library_is_an_error_not_a_fallback(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("SPOTV2_GAT_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(spotv2net_b200.SpotV2Error, match="not found"):
        _lib.load()
    monkeypatch.delenv("SPOTV2_GAT_LIB")
    monkeypatch.setattr(_lib, "_lib", None)
    _lib.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spotv2net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_host_side_sizing_rules_without_a_gpu():
    """fp16-pair row pitch and the weight-gradient split-K factor are host-side rules the workspace query shares."""
    lib = spotv2net_b200.load_library()
    # rows of fp16 operand pairs start on 32-byte sectors
    assert [lib.spotv2_gat_ld16(c) for c in (1, 16, 17, 1260, 3012)] == [16, 16, 32, 1264, 3024]
    # config A: dW_aug [3012, 1260] over 122 880 rows -> 24 x 5 tiles; 16 splits fill 12.97 waves of 148 CTAs (15: 12.16)
    assert lib.spotv2_diag_weight_grad_splits(4096 * 30, 3012, 1260) == 16
    # small problems are not split; the factor never exceeds 32 and never drops below ceil(rows / 8192)
    assert lib.spotv2_diag_weight_grad_splits(30, 3012, 1260) == 1
    assert lib.spotv2_diag_weight_grad_splits(8192, 64, 64) == 1
    for rows, m, n in [(20000, 3012, 1260), (122880, 2064, 2048), (10 ** 6, 512, 512), (65536, 100, 100)]:
        s = lib.spotv2_diag_weight_grad_splits(rows, m, n)
        base = min(32, max(1, -(-rows // 8192)))
        assert base <= s <= min(32, base + base // 4 + 1)
        tiles = -(-m // 128) * -(-n // 256)
        eff = lambda k: tiles * k / (-(-(tiles * k) // 148) * 148)
        assert eff(s) >= eff(base) - 1e-12
