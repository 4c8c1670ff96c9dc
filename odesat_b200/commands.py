"""The `solve` / `batch` / `inter` command bodies of the reference's CLI (main.rs:143-204, 254-323,
326-386) over the GPU integrator: read → parse → (preprocess) → normalise → init → simulate →
map → (trace replay) → verify against the ORIGINAL formula → render.

Same flags (-f -o -t -n -s -l -r -b as keyword arguments), same console lines, same output
format ("<var> <0|1>\\n", cnf.rs:289-298).  The state initialisation draws v0 ~ U[-1,1) from a
seeded numpy generator (the reference's thread_rng is OS-seeded and unseedable, main.rs:169).
Everything numeric runs through libodesat_b200 on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import numpy as np

from . import _lib as L
from . import batch as B
from . import cnf, preprocess
from . import stoch as ST
from .system import DeviceFormula, State, init_short_term_memory, simulate


@dataclass
class CommandResult:
    is_satisfiable: bool               # evaluate_cnf on the original formula (main.rs:190, 303, 371)
    values: Dict[int, bool]            # file variable name → value
    rendered: str                      # render_variable_map
    steps: int = 0
    winner: int = -1
    n_vars: int = 0                    # size of the integrated (preprocessed, normalised) formula
    n_clauses: int = 0
    seconds_preprocess: float = 0.0    # host-side ratio preprocessing (solve only)
    seconds_integrate: float = 0.0     # formula upload + GPU integration


def evaluate_cnf(values: Dict[int, bool], clauses: List[List[int]]) -> bool:
    """cnf.rs:246-264 on the original clauses; missing variables are inserted as false."""
    for c in clauses:
        ok = False
        for l in c:
            val = values.setdefault(abs(l), False)
            ok = ok or (not val if l < 0 else val)
        if not ok:
            return False
    return True


def _finish(values, original, output, log, res: CommandResult, newline="") -> CommandResult:
    res.is_satisfiable = evaluate_cnf(values, original.clauses)
    log(f"{newline}Checking if solution vector satisfies formula: {'true' if res.is_satisfiable else 'false'}")
    log("Rendering variable assignments...")
    res.values = values
    res.rendered = cnf.render_variable_map(values)
    if output is not None:
        log("Writing results to file...")
        with open(output, "w") as fh:
            fh.write(res.rendered)
    else:
        log(f"Variable assignments:\n{res.rendered}")
    return res


def _read(path, log) -> cnf.CNF:
    log("Reading CNF formula from file...")
    with open(path, "r") as fh:
        text = fh.read()
    log("Parsing CNF formula...")
    return cnf.parse_dimacs_format(text)


def solve(input: str, output: Optional[str] = None, tolerance: Optional[float] = None,
          step_number: Optional[int] = None, step_size: Optional[float] = None,
          learning_rate: Optional[float] = None, ctv_ratio: Optional[float] = None, *, seed: int = 0,
          precision: int = L.F64, log: Callable[[str], None] = print) -> CommandResult:
    """main.rs:143-204 (`odesat solve -f F [-o O] [-t tol] [-n steps] [-s dt] [-l zeta] [-r ratio]`)."""
    ratio = 7.0 if ctv_ratio is None else ctv_ratio                              # main.rs:150-154
    original = _read(input, log)
    log("Preprocessing CNF formula...")
    t0 = time.perf_counter()
    clauses, varnum, trace = preprocess.repeatedly_resolve_and_update(
        preprocess.to_clause_set(original.clauses), original.varnum, ratio, log=log)
    t_pre = time.perf_counter() - t0
    # convert_to_cnf_formula (cnf.rs:397-416): clause and literal order = BTreeSet order
    reduced = cnf.CNF([preprocess.sorted_literals(c) for c in preprocess.sorted_clauses(clauses)], varnum)
    formula = cnf.normalize_cnf_variables(reduced)
    log("Simulating...")
    t0 = time.perf_counter()
    F = DeviceFormula(formula)
    dtype = np.float32 if precision == L.F32 else np.float64
    rng = np.random.default_rng(seed)
    state = State((rng.random(formula.varnum) * 2.0 - 1.0).astype(dtype),       # main.rs:170-174
                  init_short_term_memory(F, dtype), np.ones(formula.n_clauses, dtype))
    info: list = []
    result = simulate(state, F, tolerance, step_size, step_number, learning_rate, info=info)
    F.close()
    t_int = time.perf_counter() - t0
    log("Mapping values...")
    values = formula.map_values_by_indices(result)
    preprocess.calculate_trace(values, trace)                                    # main.rs:186-187
    log("Evaluating CNF formula...")
    res = CommandResult(False, {}, "", steps=info[0].steps_taken, n_vars=formula.varnum, n_clauses=formula.n_clauses,
                        seconds_preprocess=t_pre, seconds_integrate=t_int)
    return _finish(values, original, output, log, res)


def stoch(input: str, output: Optional[str] = None, step_number: Optional[int] = None, ctv_ratio: Optional[float] = None, *,
          batch_size: int = 1, seed: int = 0, log: Callable[[str], None] = print) -> CommandResult:
    """main.rs:206-252 (`odesat stoch -f F [-o O] [-n steps] [-r ratio]`): ratio preprocessing, then the weighted
    random-flip search of src/stoch.rs on the GPU (`batch_size` independent replicas; the reference runs one), trace
    replay, exact verification."""
    ratio = 7.0 if ctv_ratio is None else ctv_ratio                              # main.rs:209-213
    original = _read(input, log)
    log("Preprocessing CNF formula...")
    clauses, varnum, trace = preprocess.repeatedly_resolve_and_update(
        preprocess.to_clause_set(original.clauses), original.varnum, ratio, log=log)
    reduced = cnf.CNF([preprocess.sorted_literals(c) for c in preprocess.sorted_clauses(clauses)], varnum)
    formula = cnf.normalize_cnf_variables(reduced)
    log("Simulating...")
    F = DeviceFormula(formula)
    r = ST.search_batch(F, batch_size, step_number, seed=seed)
    F.close()
    log("Mapping values...")
    values = formula.map_values_by_indices(r.assignment)
    preprocess.calculate_trace(values, trace)                                    # main.rs:237-238
    log("Evaluating CNF formula...")
    res = CommandResult(False, {}, "", steps=r.steps_run, winner=r.winner, n_vars=formula.varnum, n_clauses=formula.n_clauses)
    return _finish(values, original, output, log, res)


def batch(input: str, step_number: int, batch_size: int, output: Optional[str] = None,
          tolerance: Optional[float] = None, step_size: Optional[float] = None,
          learning_rate: Optional[float] = None, *, seed: int = 0, precision: int = L.F64,
          log: Callable[[str], None] = print) -> CommandResult:
    """main.rs:254-323: B independent replicas, all at once on the GPU; the winner is the lowest
    replica whose thresholded final state verifies (the sequential loop's `break`, main.rs:305-307)."""
    original = _read(input, log)
    log("Normalizing CNF formula...")
    formula = cnf.normalize_cnf_variables(original)
    log("Simulating...")
    if batch_size == 0:   # main.rs:276-308: the loop does not run — empty map, is_satisfiable stays false
        res = CommandResult(False, {}, "", n_vars=formula.varnum, n_clauses=formula.n_clauses)
        log("\nChecking if solution vector satisfies formula: false")
        log("Rendering variable assignments...")
        if output is not None:
            log("Writing results to file...")
            open(output, "w").close()
        else:
            log("Variable assignments:\n")
        return res
    F = DeviceFormula(formula)
    r = B.simulate_batch(F, batch_size, seed=seed, tolerance=tolerance, step_size=step_size, steps=step_number,
                         learning_rate=learning_rate, precision=precision, mode=L.MODE_BATCH)
    F.close()
    values = formula.map_values_by_indices(r.assignment)
    res = CommandResult(False, {}, "", steps=r.steps_run, winner=r.winner, n_vars=formula.varnum, n_clauses=formula.n_clauses)
    return _finish(values, original, output, log, res, newline="\n")


def inter(input: str, batch_size: int, output: Optional[str] = None, tolerance: Optional[float] = None,
          step_number: Optional[int] = None, step_size: Optional[float] = None,
          learning_rate: Optional[float] = None, *, seed: int = 0, precision: int = L.F64,
          log: Callable[[str], None] = print) -> CommandResult:
    """main.rs:326-386: B replicas in lock-step, stop at the first flag (system.rs:241-359).  Without -s
    the reference shares one adaptive dt across the replicas (quirk Q7); the library runs that loop
    literally (sequential over replicas)."""
    original = _read(input, log)
    log("Normalizing CNF formula...")
    formula = cnf.normalize_cnf_variables(original)
    log("Simulating...")
    if batch_size == 0:
        raise ValueError("inter needs at least one replica (the reference indexes states[0], system.rs:357)")
    F = DeviceFormula(formula)
    r = B.simulate_batch(F, batch_size, seed=seed, tolerance=tolerance, step_size=step_size,
                         steps=step_number, learning_rate=learning_rate, precision=precision, mode=L.MODE_INTER)
    F.close()
    values = formula.map_values_by_indices(r.assignment)
    res = CommandResult(False, {}, "", steps=r.steps_run, winner=r.winner, n_vars=formula.varnum, n_clauses=formula.n_clauses)
    return _finish(values, original, output, log, res, newline="\n")
