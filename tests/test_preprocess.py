"""Ratio preprocessing of `solve` (cnf.rs:317-840) restated on the host: BVE + BCE + subsumption
and the trace replay.  The reference's tie-breaking follows HashSet order (arbitrary), so the
tests pin properties every valid run of the reference has, plus hand-checked small cases."""
import itertools

import numpy as np
import pytest

from odesat_b200 import cnf
from odesat_b200 import preprocess as P


def fs(*lits):
    return frozenset(lits)


def brute_force(clauses, variables):
    variables = sorted(variables)
    for bits in itertools.product([False, True], repeat=len(variables)):
        a = dict(zip(variables, bits))
        if all(any(a[abs(l)] != (l < 0) for l in c) for c in clauses):
            return a
    return None


def test_resolvents_drop_tautologies_and_empties():
    cl = {fs(1, 2), fs(-1, 3), fs(-1, -2), fs(-1)}
    idx = P.calculate_variable_indices(cl)
    # (1 ∨ 2) on variable 1: with (¬1 ∨ 3) → (2 ∨ 3); with (¬1 ∨ ¬2) → tautology, dropped (cleared);
    # with (¬1) → (2)
    assert sorted(map(sorted, P.calculate_resolvents(idx, fs(1, 2), 1))) == [[2], [2, 3]]
    # the unit clauses (1) and (¬1) resolve to the empty clause, which the reference drops (cnf.rs:473)
    idx2 = P.calculate_variable_indices({fs(1), fs(-1)})
    assert P.calculate_resolvents(idx2, fs(1), 1) == []


def test_subsumption_and_blocked_clause():
    s = {fs(1, 2), fs(1, 2, 3), fs(-1, 4), fs(4)}
    P.subsume_clauses(s)
    assert s == {fs(1, 2), fs(4)}
    # (1 ∨ 2) is blocked on 1 when every clause with ¬1 also holds ¬2
    cl = {fs(1, 2), fs(-1, -2, 3), fs(-3, 2)}
    idx = P.calculate_variable_indices(cl)
    assert P.is_blocked(fs(1, 2), idx) == 1
    # a pure literal blocks its clause (no resolvents at all → all([]) is true)
    assert P.is_blocked(fs(5, 6), P.calculate_variable_indices({fs(5, 6), fs(-6, 7)})) == 5


def test_literal_and_clause_order_is_btreeset_order():
    # derive(Ord) on Literal { variable, is_negated }: by variable, positive before negated
    assert P.sorted_literals(fs(-3, 2, -2, 1)) == [1, 2, -2, -3]
    assert P.sorted_clauses([fs(2), fs(1, 3), fs(1, -2), fs(1, 2)]) == [fs(1, 2), fs(1, -2), fs(1, 3), fs(2)]


@pytest.mark.parametrize("seed", range(12))
def test_trace_replay_extends_any_model_of_the_reduced_formula(seed):
    rng = np.random.default_rng(seed)
    n, m = 12, int(rng.integers(20, 44))
    raw = []
    for _ in range(m):
        k = int(rng.integers(2, 4))
        vs = rng.choice(np.arange(1, n + 1), size=k, replace=False)
        raw.append([int(v) * int(rng.choice([-1, 1])) for v in vs])
    original = P.to_clause_set(raw)
    model0 = brute_force(original, range(1, n + 1))
    ratio = float(rng.choice([3.0, 5.0, 7.0]))
    reduced, varnum, trace = P.repeatedly_resolve_and_update(original, n, ratio)
    # satisfiability is preserved (the converse does not hold: the reference drops empty resolvents,
    # cnf.rs:473, so an UNSAT input can reduce to a satisfiable formula — `solve` then prints `false`)
    red_vars = {abs(l) for c in reduced for l in c}
    model = brute_force(reduced, red_vars)
    if model0 is not None:
        assert model is not None
    assert varnum == n - sum(1 for s in trace.steps if s[0] == "ve")           # cnf.rs:685
    if model is not None and model0 is not None:
        # any model of the reduced formula, replayed through the trace, satisfies the original
        P.calculate_trace(model, trace)
        assert all(any(model.get(abs(l), False) != (l < 0) for l in c) for c in original)
    # no clause of the result is subsumed by another (cnf.rs:805)
    assert not any(a != b and a >= b for a in reduced for b in reduced)


def test_ratio_is_reached_or_no_candidate_remains(golden_dir):
    f = cnf.parse_dimacs_format((golden_dir / "aim100_sat.cnf").read_text())
    lines = []
    reduced, varnum, trace = P.repeatedly_resolve_and_update(P.to_clause_set(f.clauses), f.varnum, 7.0, log=lines.append)
    assert lines == [f"Clauses: {len(reduced)} | Vars: {varnum}"]               # cnf.rs:822-826
    assert varnum < f.varnum and len(trace.steps) > 0
    # the loop stops when every remaining candidate would push the ratio past the target
    assert np.float32(len(reduced)) / np.float32(max(varnum, 1)) <= np.float32(7.0) + np.float32(1.0)
    # the preprocessed formula has ragged clause lengths — the general engine's job
    assert len({len(c) for c in reduced}) > 1


@pytest.mark.parametrize("fixture,ratio", [("aim100_sat.cnf", 7.0), ("aim100_unsat.cnf", 3.0), ("toy_mixed.cnf", 7.0)])
def test_cpp_host_preprocessing_matches_the_python_restatement(golden_dir, fixture, ratio):
    """odesat_b200/host/preprocess.hpp (used by the C++ CLI's `solve -r`) and preprocess.py make the same
    choices (ascending variable order where the reference follows HashSet order): identical reduced
    formula, identical trace.  `odesat_b200_cli preprocess` is host-only (no GPU)."""
    import subprocess
    from pathlib import Path
    cli = Path(__file__).resolve().parent.parent / "odesat_b200" / "csrc" / "odesat_b200_cli"
    out = subprocess.run([str(cli), "preprocess", "-f", str(golden_dir / fixture), "-r", str(ratio)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    f = cnf.parse_dimacs_format((golden_dir / fixture).read_text())
    reduced, varnum, trace = P.repeatedly_resolve_and_update(P.to_clause_set(f.clauses), f.varnum, ratio)
    assert lines[0] == f"Clauses: {len(reduced)} | Vars: {varnum}"
    assert lines[1] == f"p cnf {varnum} {len(reduced)}"
    want = [" ".join(str(l) for l in P.sorted_literals(c)) + " 0" for c in P.sorted_clauses(reduced)]
    got = [l.strip() for l in lines[2:2 + len(reduced)]]
    assert got == want
    tl = lines[2 + len(reduced):]
    heads = [l.split() for l in tl if not l.startswith("t  ")]
    assert [(h[1], int(h[2])) for h in heads] == [(s[0], s[1]) for s in trace.steps]
    # the clauses recorded with each step agree too
    k = 0
    for s in trace.steps:
        assert tl[k].split()[1] == s[0]
        cl = s[2] if s[0] == "ve" else [s[2]]
        want_c = [" ".join(str(l) for l in P.sorted_literals(c)) + " 0" for c in P.sorted_clauses(cl)]
        got_c = [x[3:].strip() for x in tl[k + 1:k + 1 + len(want_c)]]
        assert got_c == want_c
        k += 1 + len(want_c)
