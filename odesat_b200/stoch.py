"""Host-side mirror of the reference's `odesat::stoch` module (src/stoch.rs) over the C ABI: the weighted
random-flip local search, run on the GPU for a batch of independent replicas.

Same names and meaning as the Rust items (`State`, `step`, `search`); the `&mut SlabState` scratch and the
`&mut ThreadRng` have no counterpart — the per-variable weights live in registers, and the flips are drawn from a
counter-based generator keyed by (seed, replica, step, variable) because the reference's OS-seeded RNG cannot be
reproduced.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _lib as L
from .system import DeviceFormula, _ptr


@dataclass
class State:
    """stoch.rs:8-12"""
    v: np.ndarray    # bool as uint8 [N]
    xl: np.ndarray   # uint64 [M]

    @staticmethod
    def initial(formula: DeviceFormula) -> "State":
        """stoch.rs:84-87: every variable false, every weight 1."""
        return State(np.zeros(formula.varnum, np.uint8), np.ones(formula.n_clauses, np.uint64))


def step(y: State, formula: DeviceFormula, seed: int, step_index: int, replica: int = 0) -> bool:
    """stoch.rs:26-78 → all clauses satisfied; `y` is updated in place."""
    a = C.c_int()
    L.check(L.lib().odesat_stoch_step(formula.handle, _ptr(y.v), _ptr(y.xl), seed, replica, step_index, C.byref(a)))
    return bool(a.value)


@dataclass
class SearchResult:
    solved_step: np.ndarray
    verified: np.ndarray
    winner: int
    assignment: np.ndarray
    steps_run: int


def search_batch(formula: DeviceFormula, R: int, steps: Optional[int] = None, *, seed: int = 1, replica_offset: int = 0,
                 v: Optional[np.ndarray] = None, xl: Optional[np.ndarray] = None, chunk: int = 0,
                 write_back: bool = False) -> SearchResult:
    """R independent searches (stoch.rs:80-110) in one call of `odesat_stoch_search`."""
    solved = np.full(R, -1, np.int64)
    ver = np.zeros(R, np.uint8)
    assign = np.zeros(formula.varnum, np.uint8)
    win, run = C.c_int64(-1), C.c_int64(0)
    if v is not None and (v.dtype != np.uint8 or v.shape != (R, formula.varnum)):
        raise ValueError("v must be uint8 of shape [R][N]")
    if xl is not None and (xl.dtype != np.uint64 or xl.shape != (R, formula.n_clauses)):
        raise ValueError("xl must be uint64 of shape [R][M]")
    L.check(L.lib().odesat_stoch_search(formula.handle, R, _ptr(v), _ptr(xl), seed, replica_offset,
                                        -1 if steps is None else int(steps), chunk, int(write_back), _ptr(solved), _ptr(ver),
                                        C.byref(win), _ptr(assign), C.byref(run)))
    return SearchResult(solved, ver, win.value, assign, run.value)


def search(formula: DeviceFormula, steps: Optional[int], *, seed: int = 1) -> List[bool]:
    """stoch.rs:80 search(&CNFFormula, Option<usize>) -> Vec<bool> (one trajectory)."""
    return [bool(x) for x in search_batch(formula, 1, steps, seed=seed).assignment]
