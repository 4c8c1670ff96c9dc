"""ctypes front-end of the CPU oracle (oracle/dmm_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` leg; the product package (odesat_b200/) must never import it.
PARITY UNPINNED BY THE REFERENCE (no golden vectors upstream, no Rust toolchain here) — pinned
against SURVEY.md §8c's hand-derived KATs and oracle/pyref.py instead.

All state arrays use the reference's host layout: one contiguous vector per replica
(``v[R][N]``, ``xs[R][M]``, ``xl[R][M]``), dtype float64 or float32.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libdmm_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _HERE / "dmm_oracle.cpp"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _declare(_lib)
    return _lib


_P = C.c_void_p
_D = C.c_double
_I64 = C.c_int64


def _declare(L):
    L.dmm_formula_new.restype = _P
    L.dmm_formula_new.argtypes = [_I64, _I64, _P, _P]
    L.dmm_formula_free.argtypes = [_P]
    L.dmm_default_zeta.restype = _D
    L.dmm_default_zeta.argtypes = [_P]
    L.dmm_evaluate_cnf.restype = C.c_int
    L.dmm_evaluate_cnf.argtypes = [_P, _P]
    L.dmm_stoch_step.restype = C.c_int
    L.dmm_stoch_step.argtypes = [_P, _P, _P, C.c_uint64, _I64, _I64]
    L.dmm_stoch_batch.restype = None
    L.dmm_stoch_batch.argtypes = [_P, _I64, _P, _P, C.c_uint64, _I64, _I64, _P]
    for s in ("f64", "f32"):
        f = getattr(L, f"dmm_compute_derivatives_{s}")
        f.restype, f.argtypes = C.c_int, [_P, _P, _P, _P, _D, _P, _P, _P]
        f = getattr(L, f"dmm_update_state_{s}")
        f.restype, f.argtypes = None, [_P, _P, _P, _P, _P, _P, _P, _D]
        f = getattr(L, f"dmm_max_error_{s}")
        f.restype, f.argtypes = _D, [_I64, _I64, _P, _P, _P, _P, _P, _P]
        f = getattr(L, f"dmm_euler_step_fixed_{s}")
        f.restype, f.argtypes = C.c_int, [_P, _P, _P, _P, _D, _D]
        f = getattr(L, f"dmm_euler_step_{s}")
        f.restype, f.argtypes = C.c_int, [_P, _P, _P, _P, _D, C.POINTER(_D), _D]
        f = getattr(L, f"dmm_init_short_term_memory_{s}")
        f.restype, f.argtypes = None, [_P, _P]
        f = getattr(L, f"dmm_simulate_{s}")
        f.restype, f.argtypes = C.c_int, [_P, _P, _P, _P, _D, _D, _I64, _D, _P, C.POINTER(_I64),
                                          C.POINTER(_D)]
        f = getattr(L, f"dmm_simulate_inter_{s}")
        f.restype, f.argtypes = _I64, [_P, _I64, _P, _P, _P, _D, _D, _I64, _D, _P, C.POINTER(_I64)]
        f = getattr(L, f"dmm_init_v0_{s}")
        f.restype, f.argtypes = None, [C.c_uint64, _I64, _I64, _P]
        f = getattr(L, f"dmm_batch_fixed_{s}")
        f.restype, f.argtypes = None, [_P, _I64, _P, _P, _P, _D, _D, _I64, C.c_int, _P, C.c_int]
        f = getattr(L, f"dmm_batch_adaptive_{s}")
        f.restype, f.argtypes = None, [_P, _I64, _P, _P, _P, _D, _D, _I64, _P, _P, C.c_int]


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


NAN = float("nan")


class OracleFormula:
    """CSR view of a normalised CNF formula (cnf.rs:53-57 flattened)."""

    def __init__(self, varnum: int, clause_off, lits):
        self.off = np.ascontiguousarray(clause_off, dtype=np.int64)
        self.lits = np.ascontiguousarray(lits, dtype=np.int32)
        self.N = int(varnum)
        self.M = int(len(self.off) - 1)
        self._h = lib().dmm_formula_new(self.N, self.M, _ptr(self.off), _ptr(self.lits))
        if not self._h:
            raise ValueError("literal index outside 1..varnum")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.dmm_formula_free(self._h)
            self._h = None

    # --- system.rs mirrors -------------------------------------------------------------
    def default_zeta(self) -> float:
        return lib().dmm_default_zeta(self._h)

    def init_short_term_memory(self, dtype=np.float64) -> np.ndarray:
        xs = np.empty(self.M, dtype=dtype)
        getattr(lib(), f"dmm_init_short_term_memory_{_sfx(dtype)}")(self._h, _ptr(xs))
        return xs

    def compute_derivatives(self, v, xs, xl, zeta):
        s = _sfx(v.dtype)
        dv, dxs, dxl = np.empty_like(v), np.empty_like(xs), np.empty_like(xl)
        a = getattr(lib(), f"dmm_compute_derivatives_{s}")(self._h, _ptr(v), _ptr(xs), _ptr(xl),
                                                           zeta, _ptr(dv), _ptr(dxs), _ptr(dxl))
        return dv, dxs, dxl, bool(a)

    def update_state(self, v, xs, xl, dv, dxs, dxl, dt):
        getattr(lib(), f"dmm_update_state_{_sfx(v.dtype)}")(self._h, _ptr(v), _ptr(xs), _ptr(xl),
                                                            _ptr(dv), _ptr(dxs), _ptr(dxl), dt)

    def euler_step_fixed(self, v, xs, xl, dt, zeta) -> bool:
        return bool(getattr(lib(), f"dmm_euler_step_fixed_{_sfx(v.dtype)}")(
            self._h, _ptr(v), _ptr(xs), _ptr(xl), dt, zeta))

    def euler_step(self, v, xs, xl, tol, dt, zeta):
        d = _D(dt)
        a = getattr(lib(), f"dmm_euler_step_{_sfx(v.dtype)}")(self._h, _ptr(v), _ptr(xs), _ptr(xl),
                                                              tol, C.byref(d), zeta)
        return bool(a), d.value

    def simulate(self, v, xs, xl, tol=NAN, step_size=NAN, steps=-1, zeta=NAN):
        """Returns (assignment uint8[N], flagged, steps_taken, final_dt); state mutated in place."""
        assign = np.empty(self.N, dtype=np.uint8)
        st, dt = _I64(0), _D(0)
        a = getattr(lib(), f"dmm_simulate_{_sfx(v.dtype)}")(
            self._h, _ptr(v), _ptr(xs), _ptr(xl), tol, step_size, steps, zeta, _ptr(assign),
            C.byref(st), C.byref(dt))
        return assign, bool(a), st.value, dt.value

    def simulate_inter(self, v, xs, xl, tol=NAN, step_size=NAN, steps=-1, zeta=NAN):
        """v[R][N], xs/xl[R][M].  Returns (assignment, winner (-1: none → replica 0), outer steps)."""
        R = v.shape[0]
        assign = np.empty(self.N, dtype=np.uint8)
        st = _I64(0)
        w = getattr(lib(), f"dmm_simulate_inter_{_sfx(v.dtype)}")(
            self._h, R, _ptr(v), _ptr(xs), _ptr(xl), tol, step_size, steps, zeta, _ptr(assign),
            C.byref(st))
        return assign, int(w), st.value

    def evaluate_cnf(self, assign) -> bool:
        a = np.ascontiguousarray(assign, dtype=np.uint8)
        return bool(lib().dmm_evaluate_cnf(self._h, _ptr(a)))

    # --- src/stoch.rs ---------------------------------------------------------------------
    def stoch_step(self, v, xl, seed: int, replica: int, step: int) -> bool:
        """stoch.rs:26-78 on v uint8[N] / xl uint64[M] in place; the flip draws are the counter-based stand-in."""
        assert v.dtype == np.uint8 and xl.dtype == np.uint64
        return bool(lib().dmm_stoch_step(self._h, _ptr(v), _ptr(xl), seed, replica, step))

    def stoch_batch(self, v, xl, seed: int, steps: int, replica_offset: int = 0) -> np.ndarray:
        """stoch.rs:80-110 for R replicas (v uint8[R][N], xl uint64[R][M], in place) → first flagged step or -1."""
        assert v.dtype == np.uint8 and xl.dtype == np.uint64
        R = v.shape[0]
        solved = np.full(R, -1, dtype=np.int64)
        lib().dmm_stoch_batch(self._h, R, _ptr(v), _ptr(xl), seed, replica_offset, steps, _ptr(solved))
        return solved

    # --- batch helpers (CPU baseline) -----------------------------------------------------
    def init_v0(self, seed: int, replica: int, dtype=np.float64) -> np.ndarray:
        v = np.empty(self.N, dtype=dtype)
        getattr(lib(), f"dmm_init_v0_{_sfx(dtype)}")(seed, replica, self.N, _ptr(v))
        return v

    def init_batch(self, seed: int, R: int, dtype=np.float64, replica_offset: int = 0):
        v = np.stack([self.init_v0(seed, replica_offset + r, dtype) for r in range(R)]) if R else \
            np.empty((0, self.N), dtype=dtype)
        xs = np.tile(self.init_short_term_memory(dtype), (R, 1))
        xl = np.ones((R, self.M), dtype=dtype)
        return np.ascontiguousarray(v), np.ascontiguousarray(xs), xl

    def batch_fixed(self, v, xs, xl, dt, zeta, steps, freeze=True, nthreads=1) -> np.ndarray:
        R = v.shape[0]
        solved = np.full(R, -1, dtype=np.int64)
        getattr(lib(), f"dmm_batch_fixed_{_sfx(v.dtype)}")(
            self._h, R, _ptr(v), _ptr(xs), _ptr(xl), dt, zeta, steps, int(freeze), _ptr(solved),
            nthreads)
        return solved

    def batch_adaptive(self, v, xs, xl, tol, zeta, steps, nthreads=1):
        R = v.shape[0]
        solved = np.full(R, -1, dtype=np.int64)
        dts = np.empty(R, dtype=v.dtype)
        getattr(lib(), f"dmm_batch_adaptive_{_sfx(v.dtype)}")(
            self._h, R, _ptr(v), _ptr(xs), _ptr(xl), tol, zeta, steps, _ptr(solved), _ptr(dts),
            nthreads)
        return solved, dts


def max_error(a, b) -> float:
    """system.rs:101-109 on two (v, xs, xl) triples."""
    av, axs, axl = a
    bv, bxs, bxl = b
    return getattr(lib(), f"dmm_max_error_{_sfx(av.dtype)}")(
        av.shape[0], axs.shape[0], _ptr(av), _ptr(axs), _ptr(axl), _ptr(bv), _ptr(bxs), _ptr(bxl))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
