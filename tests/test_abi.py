"""The C-ABI library loads, exports every symbol include/odesat_b200.h declares, and refuses
to compute without a CUDA device (no CPU fallback).  No compute happens in these tests."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from odesat_b200 import _lib as L

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "odesat_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(odesat_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    decl = declared_functions()
    assert len(decl) >= 30
    assert sorted(L.exported_symbols()) == decl


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(str(L.SO_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported by the .so"
    assert L.lib().odesat_abi_version() == 2


def test_params_struct_layout_matches_header():
    assert C.sizeof(L.Params) == 56
    p = L.make_params(step_size=0.01, steps=7)
    assert p.n_gpus == 1 and p.sub_batches == 0
    assert p.steps == 7 and p.step_size == 0.01 and np.isnan(p.tolerance) and np.isnan(p.learning_rate)


def test_invalid_formula_is_rejected_before_touching_the_device():
    off = np.array([0, 2], np.int64)
    lits = np.array([1, 5], np.int32)                      # 5 > varnum: the reference would panic (system.rs:48)
    h = C.c_void_p()
    rc = L.lib().odesat_formula_create(3, 1, off.ctypes.data_as(C.c_void_p), lits.ctypes.data_as(C.c_void_p), C.byref(h))
    assert rc == L.EINVAL and not h.value
    assert b"varnum" in L.lib().odesat_last_error()


def test_no_cpu_fallback_without_a_device():
    if L.lib().odesat_device_count() > 0:
        pytest.skip("a CUDA device is present")
    off = np.array([0, 2], np.int64)
    lits = np.array([1, -2], np.int32)
    h = C.c_void_p()
    rc = L.lib().odesat_formula_create(2, 1, off.ctypes.data_as(C.c_void_p), lits.ctypes.data_as(C.c_void_p), C.byref(h))
    assert rc == L.ECUDA and not h.value
    with pytest.raises(L.OdesatError) as e:
        L.check(rc)
    assert "no CPU fallback" in str(e.value)
