"""Pins the CPU oracle: hand-derived KATs (derived from src/system.rs by hand),
an independent pure-Python restatement, the committed golden trajectories, and the
behavioural facts the reference's fixtures imply (easy.cnf SAT, hard.cnf never flags)."""
import math

import numpy as np
import pytest

from odesat_b200 import cnf
from oracle import oracle as O
from oracle import pyref

from helpers import ragged_formula, random_state


def toy(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "toy_mixed.cnf"))
    return f, O.OracleFormula(f.varnum, f.clause_off, f.lits)


# ---- KATs on small.cnf, normalised {1→0, 3→1, 4→2, 5→3}, N = 5 (index 4 unused) -------------
KATS = [
    dict(v=[0, 0, 0, 0, 0], dv=[0, 0, 0.5, 0, 0], dxs=[5.004999999999999] * 3, dxl=[2.25] * 3, allsat=False,
         v1=[0, 0, 0.005, 0, 0], xs1=[0.999] * 3, xl1=[1.0225] * 3),
    dict(v=[0.5, -0.25, 0.75, -1.0, 0], dv=[-0.125, -0.75, 0.25, 0, 0],
         dxs=[-5.004999999999999, -2.5024999999999995, 2.5024999999999995], dxl=[-0.25, 0.375, 1.625], allsat=False,
         v1=[0.49875, -0.2575, 0.7525, -1.0, 0], xs1=[0.94995, 0.974975, 0.999], xl1=[1.0, 1.00375, 1.01625]),
    dict(v=[1, -1, -1, 1, 0], dv=[1, 0, 0, 1, 0], dxs=[-5.004999999999999] * 3, dxl=[-0.25] * 3, allsat=True,
         v1=[1, -1, -1, 1, 0], xs1=[0.94995] * 3, xl1=[1.0] * 3),
]


def test_toy_normalisation(golden_dir):
    f, _ = toy(golden_dir)
    assert f.varnum == 5 and f.n_clauses == 3
    assert f.name_map == {1: 0, 3: 1, 4: 2, 5: 3}
    assert list(f.clause_off) == [0, 3, 7, 9]
    assert list(f.lits) == [1, -4, 3, -1, 4, 2, 3, -2, -3]
    assert f.default_zeta() == 0.001


@pytest.mark.parametrize("kat", KATS)
def test_kat_rhs_and_fixed_step(golden_dir, kat):
    f, F = toy(golden_dir)
    v = np.array(kat["v"], np.float64)
    xs = F.init_short_term_memory()
    assert list(xs) == [1.0, 1.0, 1.0]                      # every clause has a negated literal
    xl = np.ones(3)
    dv, dxs, dxl, allsat = F.compute_derivatives(v, xs, xl, 0.001)
    np.testing.assert_allclose(dv, kat["dv"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(dxs, kat["dxs"], rtol=1e-15)
    np.testing.assert_allclose(dxl, kat["dxl"], rtol=1e-15)
    assert allsat is kat["allsat"]
    flag = F.euler_step_fixed(v, xs, xl, 0.01, 0.001)
    assert flag is kat["allsat"]                            # flag of the PRE-update state (Q4)
    np.testing.assert_allclose(v, kat["v1"], rtol=1e-15, atol=1e-18)
    np.testing.assert_allclose(xs, kat["xs1"], rtol=1e-15)
    np.testing.assert_allclose(xl, kat["xl1"], rtol=1e-15)


def test_rigidity_term_is_identically_zero(golden_dir):
    """Quirk Q1: zeta never changes the result for states inside [-1, 1]."""
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    rng = np.random.default_rng(0)
    for _ in range(5):
        v, xs, xl = random_state(rng, F.N, F.M)
        a = F.compute_derivatives(v, xs, xl, 0.001)
        b = F.compute_derivatives(v, xs, xl, 123.0)
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_independent_python_restatement(seed):
    f = ragged_formula(seed)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    P = pyref.from_csr(f.varnum, f.clause_off, f.lits)
    rng = np.random.default_rng(seed)
    v, xs, xl = random_state(rng, F.N, F.M)
    zeta = 0.01
    dv, dxs, dxl, allsat = F.compute_derivatives(v, xs, xl, zeta)
    pdv, pdxs, pdxl, pall = pyref.compute_derivatives(P, list(v), list(xs), list(xl), zeta)
    assert np.array_equal(dv, np.array(pdv), equal_nan=True)
    assert np.array_equal(dxs, np.array(pdxs), equal_nan=True)
    assert np.array_equal(dxl, np.array(pdxl), equal_nan=True)
    assert allsat == pall
    # 25 fixed + 25 adaptive steps, bit for bit
    st = [list(v), list(xs), list(xl)]
    for _ in range(25):
        a = F.euler_step_fixed(v, xs, xl, 0.01, zeta)
        b = pyref.euler_step_fixed(P, st, 0.01, zeta)
        assert a == b
    dt = pdt = 0.01
    for _ in range(25):
        a, dt = F.euler_step(v, xs, xl, 1e-3, dt, zeta)
        b, pdt = pyref.euler_step(P, st, 1e-3, pdt, zeta)
        assert a == b and dt == pdt
    assert np.array_equal(v, np.array(st[0]), equal_nan=True)
    assert np.array_equal(xs, np.array(st[1]), equal_nan=True)
    assert np.array_equal(xl, np.array(st[2]), equal_nan=True)


def test_max_error_nan_semantics():
    a = (np.array([0.0, np.nan]), np.array([1.0]), np.array([2.0]))
    b = (np.array([0.5, 1.0]), np.array([1.0]), np.array([5.0]))
    assert O.max_error(a, b) == 3.0                        # NaN element ignored (f64::max)
    e = (np.zeros(0), np.zeros(0), np.zeros(0))
    assert math.isnan(O.max_error(e, e))                   # folds start at NaN (system.rs:103)
    pa = [list(x) for x in a]
    pb = [list(x) for x in b]
    assert pyref.max_error(pa, pb) == 3.0


def test_unit_and_empty_clause_quirks():
    """Quirk Q9: a unit clause drives v to ±1 through g = ±inf; an empty clause is never satisfied."""
    f = cnf.normalize_cnf_variables(cnf.parse_dimacs_format("p cnf 2 3\n1 0\n\n-2 1 0\n"))
    assert f.n_clauses == 3 and list(np.diff(f.clause_off)) == [1, 0, 2]
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    v = np.array([-0.5, 0.3]); xs = F.init_short_term_memory(); xl = np.ones(3)
    assert list(xs) == [-1.0, -1.0, 1.0]
    dv, dxs, dxl, allsat = F.compute_derivatives(v, xs, xl, 0.001)
    assert dv[0] == -np.inf or dv[0] == np.inf or np.isnan(dv[0])
    assert np.isinf(dxl[1]) and not allsat
    F.euler_step_fixed(v, xs, xl, 0.01, 0.001)
    assert abs(v[0]) == 1.0 and np.all(np.isfinite(v))


@pytest.mark.parametrize("name", ["traj_aim100_f64", "traj_aim100_f32", "traj_rand300_f64", "traj_toy_f64"])
def test_oracle_reproduces_golden(golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    F = O.OracleFormula(int(g["varnum"]), g["clause_off"], g["lits"])
    dtype = g["v0"].dtype
    zeta = float(g["zeta"])
    assert np.array_equal(F.init_v0(int(g["seed"]), 0, dtype), g["v0"])
    dv, dxs, dxl, allsat = F.compute_derivatives(g["v0"].copy(), g["xs0"].copy(), g["xl0"].copy(), zeta)
    assert np.array_equal(dv, g["dv"]) and np.array_equal(dxs, g["dxs"]) and np.array_equal(dxl, g["dxl"])
    v, xs, xl = g["v0"].copy(), g["xs0"].copy(), g["xl0"].copy()
    flags = [F.euler_step_fixed(v, xs, xl, 0.01, zeta) for _ in range(len(g["fixed_flags"]))]
    assert flags == list(g["fixed_flags"])
    assert np.array_equal(v, g["fixed_v"]) and np.array_equal(xs, g["fixed_xs"]) and np.array_equal(xl, g["fixed_xl"])
    v, xs, xl = g["v0"].copy(), g["xs0"].copy(), g["xl0"].copy()
    dt = 0.01
    for k in range(len(g["adapt_dt"])):
        a, dt = F.euler_step(v, xs, xl, 1e-3, dt, zeta)
        assert dt == g["adapt_dt"][k] and a == g["adapt_flags"][k]
    assert np.array_equal(v, g["adapt_v"]) and np.array_equal(xs, g["adapt_xs"]) and np.array_equal(xl, g["adapt_xl"])


def test_sat_fixture_solves_and_unsat_never_flags(golden_dir):
    sat = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    F = O.OracleFormula(sat.varnum, sat.clause_off, sat.lits)
    solved = 0
    for seed in range(6):
        v = F.init_v0(seed, 0); xs = F.init_short_term_memory(); xl = np.ones(F.M)
        assign, flag, steps, dt = F.simulate(v, xs, xl, steps=20000)       # adaptive, tol 1e-3
        if flag:
            solved += 1
            # adaptive leaves the flagged state untouched (system.rs:122) ⇒ the flag implies SAT
            assert F.evaluate_cnf(assign) and sat.evaluate(assign)
    assert solved >= 5
    unsat = cnf.load_dimacs(str(golden_dir / "aim100_unsat.cnf"))
    U = O.OracleFormula(unsat.varnum, unsat.clause_off, unsat.lits)
    v, xs, xl = U.init_batch(1, 8)
    st = U.batch_fixed(v, xs, xl, 0.01, 0.001, 1000, freeze=True)
    assert (st == -1).all()


def test_simulate_inter_semantics(golden_dir):
    sat = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    F = O.OracleFormula(sat.varnum, sat.clause_off, sat.lits)
    v, xs, xl = F.init_batch(11, 16)
    # Q8: zero steps → state_res all true → replica 0 "wins"
    a, w, st = F.simulate_inter(v.copy(), xs.copy(), xl.copy(), step_size=0.01, steps=0)
    assert w == 0 and st == 0 and np.array_equal(a, (v[0] > 0).astype(np.uint8))
    a, w, st = F.simulate_inter(v, xs, xl, step_size=0.01, steps=5000)
    assert w >= 0
    # the winner is the lowest replica flagged at the stopping step; cross-check with batch_fixed
    v2, xs2, xl2 = F.init_batch(11, 16)
    first = F.batch_fixed(v2, xs2, xl2, 0.01, sat.default_zeta(), 5000, freeze=True)
    s_star = min(s for s in first if s >= 0)
    assert st == s_star + 1 and w == int(np.argmax(first == s_star))
    assert np.array_equal(a, (v2[w] > 0).astype(np.uint8))    # state one step past the flag (Q4)
    assert sat.evaluate(a)


def test_f32_oracle_tracks_f64_over_a_short_horizon(golden_dir):
    sat = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    F = O.OracleFormula(sat.varnum, sat.clause_off, sat.lits)
    v = F.init_v0(3, 0); xs = F.init_short_term_memory(); xl = np.ones(F.M)
    v32, xs32, xl32 = v.astype(np.float32), xs.astype(np.float32), xl.astype(np.float32)
    for _ in range(20):
        F.euler_step_fixed(v, xs, xl, 0.01, 0.001)
        F.euler_step_fixed(v32, xs32, xl32, 0.01, 0.001)
    assert np.max(np.abs(v - v32)) < 1e-4
