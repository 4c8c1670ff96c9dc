"""Independent pure-Python (float64) restatement of odesat's `src/system.rs`, small cases only.

TEST INFRASTRUCTURE ONLY.  Written separately from dmm_oracle.cpp (nested clause lists instead
of CSR, Python floats instead of templates) so that the two restatements can be checked
against each other and against the hand-derived KATs of SURVEY.md §8c.  Python floats are
IEEE-754 binary64 with no FMA contraction, i.e. the arithmetic Rust performs.

A formula is ``(varnum, clauses)`` with ``clauses = [[(var, is_negated), ...], ...]``.
"""
from __future__ import annotations

import math

ALPHA, BETA, GAMMA, DELTA, EPSILON = 5.0, 20.0, 0.25, 0.05, 0.001      # system.rs:19-23
INF = float("inf")


def _rmax(a, b):   # Rust f64::max: a NaN operand is ignored
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def _rmin(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def compute_derivatives(formula, v, xs, xl, zeta):
    """system.rs:25-91 → (dv, dxs, dxl, allsat)."""
    varnum, clauses = formula
    dv = [0.0] * varnum
    dxs, dxl = [0.0] * len(clauses), [0.0] * len(clauses)
    sat = []
    for m, clause in enumerate(clauses):
        mn = sm = INF
        slab = []
        for var, neg in clause:
            q = -1.0 if neg else 1.0
            value = 1.0 - q * v[var]
            if value < mn:
                sm, mn = mn, value
            elif value < sm:
                sm = value
            slab.append((var, value, q))
        c_m = 0.5 * mn
        for i, val, q in slab:
            g = 0.5 * q * (mn if val != mn else sm)
            r = 0.5 * (q - v[i]) if c_m == (1.0 - q * v[i]) else 0.0
            dv[i] += xl[m] * xs[m] * g + (1.0 + zeta * xl[m]) * (1.0 - xs[m]) * r
        dxs[m] = BETA * (xs[m] + EPSILON) * (c_m - GAMMA)
        dxl[m] = ALPHA * (c_m - DELTA)
        sat.append(c_m < GAMMA)
    return dv, dxs, dxl, all(sat)


def update_state(state, d, dt, clause_nums):
    """system.rs:93-97, in place on state=(v, xs, xl)."""
    v, xs, xl = state
    dv, dxs, dxl = d
    for m in range(len(xs)):
        xs[m] = _rmin(_rmax(xs[m] + dt * dxs[m], EPSILON), 1.0 - EPSILON)
    for m in range(len(xl)):
        xl[m] = _rmin(_rmax(xl[m] + dt * dxl[m], 1.0), 1e4 * float(clause_nums))
    for i in range(len(v)):
        v[i] = _rmin(_rmax(v[i] + dt * dv[i], -1.0), 1.0)


def max_error(a, b):
    """system.rs:101-109."""
    out = []
    for x, y in zip(a, b):
        e = float("nan")
        for p, q in zip(x, y):
            e = _rmax(e, abs(p - q))
        out.append(e)
    return _rmax(out[0], _rmax(out[1], out[2]))


def euler_step_fixed(formula, state, dt, zeta):
    """system.rs:141-154."""
    dv, dxs, dxl, allsat = compute_derivatives(formula, *state, zeta)
    update_state(state, (dv, dxs, dxl), dt, len(formula[1]))
    return allsat


def euler_step(formula, state, tol, dt, zeta):
    """system.rs:111-139 → (allsat, new_dt)."""
    M = len(formula[1])
    dv, dxs, dxl, allsat = compute_derivatives(formula, *state, zeta)
    if not allsat:
        t1 = [list(a) for a in state]
        update_state(t1, (dv, dxs, dxl), dt, M)
        update_state(state, (dv, dxs, dxl), 0.5 * dt, M)
        dv, dxs, dxl, _ = compute_derivatives(formula, *state, zeta)
        update_state(state, (dv, dxs, dxl), 0.5 * dt, M)
        err = max_error(t1, state)
        ratio = tol / err if err != 0.0 else (INF if tol > 0 else float("nan"))
        dt = _rmax(_rmin(dt * math.sqrt(ratio), 1e3), 2.0 ** -7)
    return allsat, dt


def default_zeta(formula):
    """system.rs:164-173."""
    d = len(formula[1]) / formula[0]
    return 0.1 if d >= 6.0 else (0.01 if d >= 4.9 else 0.001)


def init_short_term_memory(formula):
    """system.rs:362-372."""
    return [1.0 if any(neg for _, neg in c) else -1.0 for c in formula[1]]


def simulate(formula, state, tol=None, step_size=None, steps=None, zeta=None):
    """system.rs:156-239 → (assignment, flagged, steps_taken)."""
    zeta = default_zeta(formula) if zeta is None else zeta
    tol = 1e-3 if tol is None else tol
    it, flag, dt = 0, False, 0.01
    while steps is None or it < steps:
        it += 1
        if step_size is not None:
            flag = euler_step_fixed(formula, state, step_size, zeta)
        else:
            flag, dt = euler_step(formula, state, tol, dt, zeta)
        if flag:
            break
    return [x > 0.0 for x in state[0]], flag, it


def from_csr(varnum, off, lits):
    clauses = []
    for m in range(len(off) - 1):
        clauses.append([(abs(int(l)) - 1, int(l) < 0) for l in lits[off[m]:off[m + 1]]])
    return varnum, clauses
