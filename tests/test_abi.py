"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/taxidispatch.h declares; without a GPU every compute entry point refuses loudly
(no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "taxidispatch.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tdh?_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from taxidispatcher_b200 import _lib
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(_lib._SIGNATURES) == names, "ctypes table and header disagree"
    assert b"sm_100a" in lib.td_version()
    assert lib.td_strerror(_lib.TD_ERR_NO_DEVICE).decode().startswith("no CUDA device")


def test_struct_layouts_match_header():
    from taxidispatcher_b200 import _lib
    assert ctypes.sizeof(_lib.LcmParams) == 24
    assert ctypes.sizeof(_lib.AssignStats) == 40
    assert ctypes.sizeof(_lib.PoolStats) == 32


def test_workspace_queries_are_pure():
    from taxidispatcher_b200 import _lib
    lib = _lib.lib()
    assert lib.td_lcm_workspace_bytes(2000) >= 2000 * 2000 * 4
    assert lib.td_pool_workspace_bytes(722, 50, 4, 1 << 20) > (1 << 20) * 32
    assert lib.td_pool_merge_workspace_bytes(700, 722) > 0
    assert lib.td_assign_workspace_bytes(200) > 0


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from taxidispatcher_b200 import _lib, dispatch
    lib = _lib.lib()
    assert lib.td_device_count() == _lib.TD_ERR_NO_DEVICE
    c = np.ones((2, 2), np.int32)
    out = np.zeros(2, np.int32)
    obj = ctypes.c_int64()
    rc = lib.tdh_assign_exact(c.ctypes.data_as(ctypes.c_void_p), 2, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(obj), None, None)
    assert rc == _lib.TD_ERR_NO_DEVICE
    with pytest.raises(dispatch.TaxiDispatchError):
        dispatch.calculate_cost([[0, 1], [1, 0]], [(0, 0, 1)], [(0, 1, 0)])
    with pytest.raises(dispatch.TaxiDispatchError):
        dispatch.LCM(2, c)
    # degenerate sizes keep the reference's return shapes without touching the device
    assert dispatch.solve(0, []) == (0, [])
    assert dispatch.calculate_cost([[0]], [], []) == (0, 0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "taxidispatcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f
