"""Formula plumbing for the DMM hot path: DIMACS reader, variable normaliser, CSR flattening and
the synthetic random k-SAT generator used by the benchmarks.

This is the host-side boundary type only: the reference's
`CNFFormula { clauses: Array1<CNFClause{literals: Array1<Literal>}>, varnum }` (cnf.rs:5-57)
flattened to the CSR the C ABI takes — ``clause_off[M+1]`` (int64) and ``lits[L]`` (int32,
signed ``±(index+1)`` over normalised variables ``0..varnum-1``).  Preprocessing (cnf.rs:317-840)
and the stochastic search are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class CNF:
    """A parsed formula over the file's own variable names (cnf.rs:53-57)."""
    clauses: List[List[int]]            # signed DIMACS literals, names as in the file
    varnum: int                         # header value, or the distinct-variable count


def parse_dimacs_format(text: str) -> CNF:
    """cnf.rs:138-172, line for line: lines starting with 'c' are skipped, a line starting
    with "p cnf" sets varnum, EVERY other line is a clause (tokens up to the first "0") —
    including blank lines, which become empty clauses."""
    lines = text.split("\n")
    if lines and lines[-1] == "":       # str::lines() yields no trailing empty line
        lines.pop()
    clauses: List[List[int]] = []
    varnum: Optional[int] = None
    for line in lines:
        if line.endswith("\r"):
            line = line[:-1]
        if line.startswith("c"):
            continue
        if line.startswith("p cnf"):
            varnum = int(line.split()[2])
            continue
        lits: List[int] = []
        for tok in line.split():
            if tok == "0":
                break
            lits.append(int(tok))       # ValueError ↔ the reference's unwrap() panic
        clauses.append(lits)
    if varnum is None:                  # CNFFormula::new(.., None): distinct-variable count
        varnum = len({abs(l) for c in clauses for l in c})
    return CNF(clauses, varnum)


@dataclass
class Formula:
    """Normalised formula in the C-ABI layout (variables are indices 0..varnum-1)."""
    varnum: int
    clause_off: np.ndarray              # int64 [M+1]
    lits: np.ndarray                    # int32 [L], ±(index+1)
    name_map: Dict[int, int] = field(default_factory=dict)   # file name → index

    @property
    def n_clauses(self) -> int:
        return int(len(self.clause_off) - 1)

    @property
    def n_literals(self) -> int:
        return int(self.clause_off[-1])

    def default_zeta(self) -> float:
        """system.rs:164-173."""
        d = self.n_clauses / self.varnum if self.varnum else float("inf")
        return 0.1 if d >= 6.0 else (0.01 if d >= 4.9 else 0.001)

    def evaluate(self, assignment: Sequence[int]) -> bool:
        """cnf.rs:246-264 on a dense 0/1 assignment over the normalised variables (host-side,
        numpy; the device-side check is the library's verify kernel)."""
        a = np.asarray(assignment).astype(bool)
        if self.n_literals == 0:
            return self.n_clauses == 0
        var = np.abs(self.lits) - 1
        litval = a[var] ^ (self.lits < 0)
        lens = np.diff(self.clause_off)
        if (lens == 0).any():
            return False
        sat = np.logical_or.reduceat(litval, self.clause_off[:-1])
        return bool(sat.all())

    def map_values_by_indices(self, assignment: Sequence[int]) -> Dict[int, bool]:
        """cnf.rs:301-315: file variable name → value."""
        return {name: bool(assignment[idx]) for name, idx in self.name_map.items()}


def normalize_cnf_variables(cnf: CNF) -> Formula:
    """cnf.rs:206-219.  The reference numbers variables in `HashSet` iteration order (random
    per process); any bijection is equivalent, this one uses ascending file name.  `varnum`
    stays the header value (cnf.rs:198), so unused tail indices simply receive dv = 0."""
    names = sorted({abs(l) for c in cnf.clauses for l in c})
    if len(names) > cnf.varnum:
        raise ValueError(f"{len(names)} distinct variables exceed header varnum {cnf.varnum} "
                         "(the reference would panic on an out-of-bounds index)")
    name_map = {n: i for i, n in enumerate(names)}
    off = np.zeros(len(cnf.clauses) + 1, dtype=np.int64)
    flat: List[int] = []
    for m, c in enumerate(cnf.clauses):
        for l in c:
            idx = name_map[abs(l)] + 1
            flat.append(-idx if l < 0 else idx)
        off[m + 1] = len(flat)
    return Formula(cnf.varnum, off, np.asarray(flat, dtype=np.int32), name_map)


def load_dimacs(path: str) -> Formula:
    with open(path, "r") as fh:
        return normalize_cnf_variables(parse_dimacs_format(fh.read()))


def render_variable_map(values: Dict[int, bool]) -> str:
    """cnf.rs:289-298 ("<var> <0|1>\\n"; the reference emits HashMap order, here ascending)."""
    return "".join(f"{k} {1 if values[k] else 0}\n" for k in sorted(values))


def random_ksat(n_vars: int, alpha: float, seed: int, k: int = 3) -> Formula:
    """Uniform random k-SAT: M = round(alpha·N) clauses, each with k DISTINCT
    variables drawn uniformly and independent fair signs."""
    m = int(round(alpha * n_vars))
    rng = np.random.default_rng(seed)
    var = rng.integers(0, n_vars, size=(m, k), dtype=np.int64)
    while True:
        s = np.sort(var, axis=1)
        bad = (s[:, 1:] == s[:, :-1]).any(axis=1)
        nb = int(bad.sum())
        if nb == 0:
            break
        var[bad] = rng.integers(0, n_vars, size=(nb, k), dtype=np.int64)
    sign = rng.integers(0, 2, size=(m, k), dtype=np.int64) * 2 - 1
    lits = ((var + 1) * sign).astype(np.int32).reshape(-1)
    off = (np.arange(m + 1, dtype=np.int64) * k)
    return Formula(n_vars, off, lits, {i + 1: i for i in range(n_vars)})


def to_dimacs(f: Formula) -> str:
    """cnf.rs:221-244 shape ("p cnf N M" then "l1 l2 .. 0")."""
    out = [f"p cnf {f.varnum} {f.n_clauses}"]
    for m in range(f.n_clauses):
        c = f.lits[f.clause_off[m]:f.clause_off[m + 1]]
        out.append(" ".join(str(int(x)) for x in c) + " 0")
    return "\n".join(out) + "\n"
