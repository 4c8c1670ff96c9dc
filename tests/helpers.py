"""Shared test helpers: formulas with the edge cases the reference can produce."""
import numpy as np

from odesat_b200 import cnf


def ragged_formula(seed: int, n_vars: int = 40, n_clauses: int = 120) -> cnf.Formula:
    """Mixed clause lengths 1..6, repeated variables inside a clause, one empty clause, unused
    variables (header varnum > distinct) — what `solve`'s preprocessing and sloppy DIMACS give."""
    rng = np.random.default_rng(seed)
    off, lits = [0], []
    for m in range(n_clauses):
        k = int(rng.integers(1, 7))
        if m == 7:
            k = 0                                  # blank line → empty clause (quirk Q9)
        vs = rng.integers(1, n_vars - 3, size=k)   # last 3 variables unused; repeats allowed
        sg = rng.integers(0, 2, size=k) * 2 - 1
        lits.extend(int(a * b) for a, b in zip(vs, sg))
        off.append(len(lits))
    return cnf.Formula(n_vars, np.asarray(off, np.int64), np.asarray(lits, np.int32), {})


def random_state(rng, N, M, dtype=np.float64, R=None):
    shape_v = (N,) if R is None else (R, N)
    shape_m = (M,) if R is None else (R, M)
    v = rng.uniform(-1, 1, size=shape_v).astype(dtype)
    # sprinkle saturated values so the rigidity branch (system.rs:73) and ties are exercised
    mask = rng.random(shape_v) < 0.15
    v[mask] = np.sign(v[mask]) + (v[mask] == 0)
    xs = rng.uniform(0.001, 0.999, size=shape_m).astype(dtype)
    xl = rng.uniform(1.0, 50.0, size=shape_m).astype(dtype)
    return np.ascontiguousarray(v), np.ascontiguousarray(xs), np.ascontiguousarray(xl)


def repeated_var_formula(seed: int, n_vars: int = 60, n_clauses: int = 300) -> cnf.Formula:
    """Uniform 3-literal clauses in which a variable may repeat inside a clause, also with both signs
    (x ∨ ¬x ∨ y): uniform length sends it to the streaming clause kernel, the repeats keep it off the
    tile engine, and the summation order inside a clause matters for dv."""
    rng = np.random.default_rng(seed)
    var = rng.integers(1, n_vars + 1, size=(n_clauses, 3))
    var[::5, 1] = var[::5, 0]                       # force repeats
    sign = rng.integers(0, 2, size=(n_clauses, 3)) * 2 - 1
    lits = (var * sign).astype(np.int32).reshape(-1)
    off = np.arange(n_clauses + 1, dtype=np.int64) * 3
    return cnf.Formula(n_vars, off, lits, {})
