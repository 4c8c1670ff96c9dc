"""Host-side mirror of the reference's `odesat::system` module (src/system.rs) over the C ABI.

Same names, argument order and meaning as the Rust functions, so parity tests read like the
reference's call sites (main.rs:176, 292, 360; benches/benchmarks.rs:41, 69).  Everything here
runs on the GPU through libodesat_b200.so; there is no CPU fallback.

The reference's `&mut SlabState` scratch argument has no counterpart (the kernels keep the
per-clause min / second-min in registers) and is dropped from the signatures.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .cnf import Formula


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("state arrays must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return ""
    if dtype == np.float32:
        return "_f32"
    raise TypeError(f"state dtype must be float64 or float32, got {dtype}")


@dataclass
class State:
    """system.rs:6-11 — also used as the derivative container, as in the reference."""
    v: np.ndarray    # variable values      [N]
    xs: np.ndarray   # short-term memory    [M]
    xl: np.ndarray   # long-term memory     [M]

    def clone(self) -> "State":
        return State(self.v.copy(), self.xs.copy(), self.xl.copy())

    @staticmethod
    def zeros(formula: "DeviceFormula", dtype=np.float64) -> "State":
        return State(np.zeros(formula.varnum, dtype), np.zeros(formula.n_clauses, dtype),
                     np.zeros(formula.n_clauses, dtype))


class DeviceFormula:
    """`CNFFormula` (cnf.rs:53-57) resident on the current CUDA device as CSR + transpose."""

    def __init__(self, formula: Formula):
        self.host = formula
        self.varnum = int(formula.varnum)
        self.n_clauses = int(formula.n_clauses)
        self._off = np.ascontiguousarray(formula.clause_off, dtype=np.int64)
        self._lits = np.ascontiguousarray(formula.lits, dtype=np.int32)
        h = C.c_void_p()
        L.check(L.lib().odesat_formula_create(self.varnum, self.n_clauses, _ptr(self._off), _ptr(self._lits),
                                              C.byref(h)))
        self._h = h

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            L.lib().odesat_formula_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_zeta(self) -> float:
        z = C.c_double()
        L.check(L.lib().odesat_formula_default_zeta(self._h, C.byref(z)))
        return z.value

    def uniform_k(self) -> int:
        k = C.c_int32()
        L.check(L.lib().odesat_formula_info(self._h, None, None, None, C.byref(k)))
        return k.value


def init_short_term_memory(formula: DeviceFormula, dtype=np.float64) -> np.ndarray:
    """system.rs:362-372."""
    xs = np.empty(formula.n_clauses, dtype=dtype)
    L.check(getattr(L.lib(), "odesat_init_short_term_memory" + _sfx(dtype))(formula.handle, _ptr(xs)))
    return xs


def compute_derivatives(y: State, dy: State, formula: DeviceFormula, zeta: float) -> bool:
    """system.rs:25-31: fills dy, returns the all-clauses-satisfied flag."""
    a = C.c_int()
    L.check(getattr(L.lib(), "odesat_compute_derivatives" + _sfx(y.v.dtype))(
        formula.handle, _ptr(y.v), _ptr(y.xs), _ptr(y.xl), zeta, _ptr(dy.v), _ptr(dy.xs), _ptr(dy.xl), C.byref(a)))
    return bool(a.value)


def update_state(state: State, derivatives: State, dt: float, formula: DeviceFormula) -> None:
    """system.rs:93 (clause_nums is taken from the formula handle)."""
    L.check(getattr(L.lib(), "odesat_update_state" + _sfx(state.v.dtype))(
        formula.handle, _ptr(state.v), _ptr(state.xs), _ptr(state.xl), _ptr(derivatives.v), _ptr(derivatives.xs),
        _ptr(derivatives.xl), dt))


def max_error(a: State, b: State, formula: DeviceFormula) -> float:
    """system.rs:101."""
    e = C.c_double()
    L.check(getattr(L.lib(), "odesat_max_error" + _sfx(a.v.dtype))(
        formula.handle, _ptr(a.v), _ptr(a.xs), _ptr(a.xl), _ptr(b.v), _ptr(b.xs), _ptr(b.xl), C.byref(e)))
    return e.value


def euler_step_fixed(state: State, formula: DeviceFormula, dt: float, zeta: float) -> bool:
    """system.rs:141-148."""
    a = C.c_int()
    L.check(getattr(L.lib(), "odesat_euler_step_fixed" + _sfx(state.v.dtype))(
        formula.handle, _ptr(state.v), _ptr(state.xs), _ptr(state.xl), dt, zeta, C.byref(a)))
    return bool(a.value)


def euler_step(state: State, formula: DeviceFormula, tolerance: float, dt: float, zeta: float) -> Tuple[bool, float]:
    """system.rs:111-119; the `&mut dt` comes back as the second result."""
    a, d = C.c_int(), C.c_double(dt)
    L.check(getattr(L.lib(), "odesat_euler_step" + _sfx(state.v.dtype))(
        formula.handle, _ptr(state.v), _ptr(state.xs), _ptr(state.xl), tolerance, C.byref(d), zeta, C.byref(a)))
    return bool(a.value), d.value


@dataclass
class SimInfo:
    steps_taken: int
    allsat: bool
    final_dt: float


def simulate(state: State, formula: DeviceFormula, tolerance: Optional[float] = None,
             step_size: Optional[float] = None, steps: Optional[int] = None,
             learning_rate: Optional[float] = None, *, precision: Optional[int] = None, chunk: int = 0,
             info: Optional[list] = None) -> List[bool]:
    """system.rs:156-163 → Vec<bool>; `state` is integrated in place."""
    prec = (L.F32 if state.v.dtype == np.float32 else L.F64) if precision is None else precision
    p = L.make_params(tolerance, step_size, steps, learning_rate, prec, L.ENGINE_AUTO, L.SCHED_EXACT, chunk)
    assign = np.empty(formula.varnum, dtype=np.uint8)
    st, a, d = C.c_int64(), C.c_int(), C.c_double()
    L.check(getattr(L.lib(), "odesat_simulate" + _sfx(state.v.dtype))(
        formula.handle, _ptr(state.v), _ptr(state.xs), _ptr(state.xl), C.byref(p), _ptr(assign), C.byref(st),
        C.byref(a), C.byref(d)))
    if info is not None:
        info.append(SimInfo(st.value, bool(a.value), d.value))
    return [bool(x) for x in assign]


def simulate_inter(states: Sequence[State], formula: DeviceFormula, tolerance: Optional[float] = None,
                   step_size: Optional[float] = None, steps: Optional[int] = None,
                   learning_rate: Optional[float] = None, *, chunk: int = 0, info: Optional[list] = None) -> List[bool]:
    """system.rs:241-248 → Vec<bool>.  Without `step_size` the replicas share one adaptive dt and step
    one after the other, exactly as the reference does (quirk Q7)."""
    R = len(states)
    v = np.ascontiguousarray(np.stack([s.v for s in states]), dtype=np.float64)
    xs = np.ascontiguousarray(np.stack([s.xs for s in states]), dtype=np.float64)
    xl = np.ascontiguousarray(np.stack([s.xl for s in states]), dtype=np.float64)
    p = L.make_params(tolerance, step_size, steps, learning_rate, L.F64, L.ENGINE_AUTO, L.SCHED_EXACT, chunk)
    assign = np.empty(formula.varnum, dtype=np.uint8)
    win, st = C.c_int64(), C.c_int64()
    L.check(L.lib().odesat_simulate_inter(formula.handle, R, _ptr(v), _ptr(xs), _ptr(xl), C.byref(p), _ptr(assign),
                                          C.byref(win), C.byref(st)))
    for r, s in enumerate(states):
        s.v[:], s.xs[:], s.xl[:] = v[r], xs[r], xl[r]
    if info is not None:
        info.append((win.value, st.value))
    return [bool(x) for x in assign]
