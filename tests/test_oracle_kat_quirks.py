"""More hand-derived known-answer tests for the CPU oracle, one per behaviour that the first three KATs
(tests/test_oracle_kat.py) leave unpinned: ties at the minimum (quirk Q2), unit and empty clauses (Q9), one full
ADAPTIVE step with the exact dt it returns (system.rs:111-139), max_error's NaN folds (system.rs:101-109) and
simulate_inter with zero steps (Q8).

Every expected number below was derived from /root/reference/src/system.rs statement by statement for the concrete
instance, NOT by running oracle/dmm_oracle.cpp or oracle/pyref.py; the derivation is written next to each value
(`:NN` = line of system.rs).  Decimal literals are the shortest round-trip form of the IEEE-754 double that the
statement produces, so the comparisons are exact (==), in f64.
"""
import math

import numpy as np
import pytest

from odesat_b200 import cnf
from oracle import oracle as O
from oracle import pyref

INF = float("inf")


def formula(text):
    f = cnf.normalize_cnf_variables(cnf.parse_dimacs_format(text))
    return f, O.OracleFormula(f.varnum, f.clause_off, f.lits)


def both(F, f, v, xs, xl, zeta):
    """RHS from the C++ oracle and from the independent Python restatement (they must agree bit for bit)."""
    dv, dxs, dxl, a = F.compute_derivatives(np.array(v, np.float64), np.array(xs, np.float64), np.array(xl, np.float64), zeta)
    P = pyref.from_csr(f.varnum, f.clause_off, f.lits)
    pdv, pdxs, pdxl, pa = pyref.compute_derivatives(P, list(v), list(xs), list(xl), zeta)
    assert np.array_equal(dv, np.array(pdv), equal_nan=True) and np.array_equal(dxs, np.array(pdxs), equal_nan=True)
    assert np.array_equal(dxl, np.array(pdxl), equal_nan=True) and a == pa
    return dv, dxs, dxl, a


# ---- Q2: ties at the minimum (system.rs:50-55, :66) ------------------------------------------------------------
def test_kat_tie_at_minimum_first_two_literals():
    """Clause (x0 ∨ x1 ∨ ¬x2), v = [0.5, 0.5, 0.25], xs = 0.5, xl = 2.
    :49 values 1−0.5 = 0.5, 1−0.5 = 0.5, 1+0.25 = 1.25.
    :50-55 lit 0: 0.5 < inf → min = 0.5; lit 1: 0.5 < 0.5 false, 0.5 < inf → second_min = 0.5; lit 2: neither.
    :60 c_m = 0.25.  :66 lits 0 and 1 have val == min → both take second_min = 0.5; lit 2 takes min = 0.5.
    :64 g = 0.5·q·0.5 = +0.25, +0.25, −0.25.  :73 c_m (0.25) equals none of the values → r = 0.
    :80 dv_i = xl·xs·g + (…)·0 = 2·0.5·(±0.25) = ±0.25.
    :84 dxs = 20·(0.5+0.001)·(0.25−0.25) = 0.  :85 dxl = 5·(0.25−0.05) = 5·0.2 = 1.0.  :88 0.25 < 0.25 is false."""
    f, F = formula("p cnf 3 1\n1 2 -3 0\n")
    dv, dxs, dxl, a = both(F, f, [0.5, 0.5, 0.25], [0.5], [2.0], 0.001)
    assert list(dv) == [0.25, 0.25, -0.25]
    assert list(dxs) == [0.0] and list(dxl) == [5.0 * (0.25 - 0.05)] and dxl[0] == 1.0
    assert a is False


def test_kat_tie_at_minimum_with_a_larger_value_in_between():
    """Clause (x0 ∨ ¬x1 ∨ x2), v = [0.5, 0.25, 0.5], xs = 1, xl = 1.
    :49 values 0.5, 1.25, 0.5.  :50-55 lit 0 → min 0.5; lit 1: 1.25 < inf → second_min 1.25; lit 2: 0.5 < 0.5 false,
    0.5 < 1.25 → second_min = 0.5.  So the tied pair gets second_min = 0.5 (by VALUE, :66), lit 1 gets min = 0.5.
    g = +0.25, −0.25, +0.25;  dv = 1·1·g.  dxs = 20·1.001·0 = 0, dxl = 1.0."""
    f, F = formula("p cnf 3 1\n1 -2 3 0\n")
    dv, dxs, dxl, a = both(F, f, [0.5, 0.25, 0.5], [1.0], [1.0], 0.001)
    assert list(dv) == [0.25, -0.25, 0.25] and list(dxs) == [0.0] and list(dxl) == [1.0] and a is False


# ---- Q9: unit clause and empty clause -------------------------------------------------------------------------
def test_kat_unit_and_empty_clause():
    """p cnf 2 3: clause 0 = (x0), clause 1 = empty (blank line, cnf.rs:155-167), clause 2 = (¬x1 ∨ x0).
    :362-372 xs0 = [−1 (no negation), −1 (`any` over nothing is false), +1];  xl = 1;  v = [−0.5, 0.3], zeta = 0.001.
    clause 0: value 1.5 → min 1.5, second_min inf; c = 0.75; g = 0.5·1·inf = inf (val == min → second_min);
              r = 0 (0.75 ≠ 1.5);  dv0 += 1·(−1)·inf + (1.001)·(2)·0 = −inf + 0 = −inf.
              dxs0 = 20·(−1+0.001)·(0.75−0.25) = (20·−0.999)·0.5;  dxl0 = 5·(0.75−0.05).
    clause 1: no literals → min = inf, c = inf;  dxs1 = (20·−0.999)·(inf−0.25) = −inf;  dxl1 = 5·inf = inf;  not satisfied.
    clause 2: ¬x1: 1−(−1·0.3) = 1.3 → min; x0: 1−(−0.5) = 1.5 → second_min; c = 0.65;
              g(¬x1) = 0.5·(−1)·1.5 = −0.75 → dv1 = −0.75;  g(x0) = 0.5·1·1.3 = 0.65 → dv0 = −inf + 0.65 = −inf.
              dxs2 = (20·1.001)·(0.65−0.25);  dxl2 = 5·(0.65−0.05).
    fixed step dt = 0.01 (:94-96): v0 = (−0.5 + 0.01·−inf).max(−1).min(1) = −1;  v1 = 0.3 + 0.01·−0.75;
              xs → [0.001 (−1.0999 clamped up), 0.001 (−inf clamped up), 0.999];  xl1 = inf.min(1e4·3) = 30000."""
    f, F = formula("p cnf 2 3\n1 0\n\n-2 1 0\n")
    assert list(np.diff(f.clause_off)) == [1, 0, 2]
    xs0 = F.init_short_term_memory()
    assert list(xs0) == [-1.0, -1.0, 1.0]
    v = [-0.5, 0.3]
    dv, dxs, dxl, a = both(F, f, v, list(xs0), [1.0, 1.0, 1.0], 0.001)
    assert list(dv) == [-INF, -0.75]
    assert list(dxs) == [20.0 * (-1.0 + 0.001) * (0.75 - 0.25), -INF, 20.0 * (1.0 + 0.001) * (0.65 - 0.25)]
    assert list(dxl) == [5.0 * (0.75 - 0.05), INF, 5.0 * (0.65 - 0.05)]
    assert a is False
    v = np.array(v); xs = xs0.copy(); xl = np.ones(3)
    assert F.euler_step_fixed(v, xs, xl, 0.01, 0.001) is False
    assert list(v) == [-1.0, 0.3 + 0.01 * -0.75]
    assert list(xs) == [0.001, 0.001, 0.999]
    assert list(xl) == [1.0 + 0.01 * dxl[0], 30000.0, 1.0 + 0.01 * dxl[2]]


# ---- one full adaptive step, exact dt out (system.rs:111-139) ---------------------------------------------------
def test_kat_adaptive_step_with_exact_dt(golden_dir):
    """small.cnf normalised {1→0, 3→1, 4→2, 5→3}: clauses [(0+),(3−),(2+)], [(0−),(3+),(1+),(2+)], [(1−),(2−)];
    state of KAT-2: v = [0.5, −0.25, 0.75, −1, 0], xs = xl = [1, 1, 1]; dt = 0.01, tolerance = 1e-3, zeta = 0.001.
    :120 k1 = KAT-2's derivatives: dv = [−0.125, −0.75, 0.25, 0, 0], dxs = [−5.005, −2.5025, 2.5025] (rounded as in
         KAT-2), dxl = [−0.25, 0.375, 1.625]; not all-satisfied, so the step is taken.
    :124-125 full step (dt): v = [0.49875, −0.2575, 0.7525, −1, 0], xs = [0.94995, 0.974975, 0.999], xl = [1, 1.00375, 1.01625].
    :128 half step (0.5·dt = 0.005): v½ = [0.499375, −0.25375, 0.75125, −1, 0], xs½ = [0.974975, 0.9874875, 0.999 (clamped)],
         xl½ = [1 (0.99875 clamped up), 1.001875, 1.008125].
    :129 k2 at the half state:
       clause 0: values 0.500625, 1−(−1·−1) = 0, 0.24875000000000003 → min 0 (lit 3−), second_min 0.24875…; c = 0;
                 g(0+) = 0.5·1·0 = 0; g(3−) = 0.5·−1·0.24875… ; g(2+) = 0;  :73 c == value only for lit 3− (0 == 0):
                 r = 0.5·(−1 − (−1)) = 0 — the rigidity branch is TAKEN and contributes 0 (quirk Q1).
                 dv3 = 1·0.974975·g(3−) + (1+0.001·1)·(1−0.974975)·0 = −0.1212… (see the literal below).
                 dxs0 = 20·(0.974975+0.001)·(0−0.25) = −4.879875;  dxl0 = 5·(0−0.05) = −0.25.
       clause 1: values 1.4993750000000001, 2, 1.25375, 0.24875000000000003 → min (2+), second_min 1.25375 (1+); c = 0.124375…
       clause 2: values 0.7462500000000001, 1.75125 → min (1−), second_min (2−).
       k2: dv = [−0.12304904298339844, −0.7588076706884764, 0.24441142612792965, 0.001786527358398432, 0],
           dxs = [−4.879875, −2.4835748437499996, 2.462500000000001], dxl = [−0.25, 0.37187500000000007, 1.6156250000000003].
    :130 second half step → v = [0.498759754785083, −0.25754403835344236, 0.7524720571306396, −0.999991067363208, 0],
         xs = [0.950575625, 0.97506962578125, 0.999], xl = [1, 1.003734375, 1.0162031249999999].
    :132 error = max |full − two halves| = |0.94995 − 0.950575625| = 0.000625625000000074 (xs of clause 0).
    :133-135 dt = (0.01·sqrt(1e-3 / error)).min(1e3).max(2^-7) = 0.012642790824819533."""
    f = cnf.load_dimacs(str(golden_dir / "toy_mixed.cnf"))
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    v = np.array([0.5, -0.25, 0.75, -1.0, 0.0]); xs = np.ones(3); xl = np.ones(3)
    flag, dt = F.euler_step(v, xs, xl, 1e-3, 0.01, 0.001)
    assert flag is False
    assert list(v) == [0.498759754785083, -0.25754403835344236, 0.7524720571306396, -0.999991067363208, 0.0]
    assert list(xs) == [0.950575625, 0.97506962578125, 0.999]
    assert list(xl) == [1.0, 1.003734375, 1.0162031249999999]
    assert dt == 0.012642790824819533
    assert dt == max(min(0.01 * math.sqrt(1e-3 / 0.000625625000000074), 1e3), 2.0 ** -7)
    # the independent Python restatement takes the same step
    P = pyref.from_csr(f.varnum, f.clause_off, f.lits)
    st = [[0.5, -0.25, 0.75, -1.0, 0.0], [1.0] * 3, [1.0] * 3]
    pflag, pdt = pyref.euler_step(P, st, 1e-3, 0.01, 0.001)
    assert pflag is False and pdt == dt and st[0] == list(v) and st[1] == list(xs) and st[2] == list(xl)


def test_kat_adaptive_step_leaves_a_satisfied_state_untouched(golden_dir):
    """:122 `if !allsat`: KAT-3's state is all-satisfied → state and dt unchanged, flag true (unlike the fixed step, Q4)."""
    f = cnf.load_dimacs(str(golden_dir / "toy_mixed.cnf"))
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    v = np.array([1.0, -1.0, -1.0, 1.0, 0.0]); xs = np.ones(3); xl = np.ones(3)
    flag, dt = F.euler_step(v, xs, xl, 1e-3, 0.01, 0.001)
    assert flag is True and dt == 0.01
    assert list(v) == [1.0, -1.0, -1.0, 1.0, 0.0] and list(xs) == [1.0] * 3 and list(xl) == [1.0] * 3


def test_kat_adaptive_dt_floor():
    """:133-135 the clamp order is .min(1e3) THEN .max(2^-7): a large error floors dt at 2^-7 = 0.0078125.
    One unit clause (x0), v = [−0.5], xs = xl = 1, dt = 0.01, tolerance = 1e-3:
    k1: value 1.5, c = 0.75, g = 0.5·1·inf = inf (second_min of a unit clause, Q9); dxs = 20·1.001·0.5 = 10.01; dxl = 5·0.7 = 3.5.
    full step: v = clamp(−0.5 + 0.01·inf) = 1; xs = clamp(1.1001) = 0.999; xl = 1.035.
    half step: v = 1; xs = clamp(1.05005) = 0.999; xl = 1 + 0.005·3.5 = 1.0175.
    k2 at v = 1: value 0, c = 0; dxs = 20·(0.999+0.001)·(0−0.25) = −5; dxl = 5·(0−0.05) = −0.25.
    second half: xs = 0.999 + 0.005·(−5) = 0.974; xl = 1.0175 + 0.005·(−0.25) = 1.01625; v stays 1.
    error = max(|1−1|, |0.999−0.974| = 0.025, |1.035−1.01625| = 0.01875) = 0.025…; dt = 0.01·sqrt(1e-3/0.025) = 0.002 → 2^-7."""
    f, F = formula("p cnf 1 1\n1 0\n")
    v = np.array([-0.5]); xs = np.array([1.0]); xl = np.array([1.0])
    flag, dt = F.euler_step(v, xs, xl, 1e-3, 0.01, 0.001)
    assert flag is False and dt == 2.0 ** -7
    assert list(v) == [1.0]
    assert xs[0] == 0.999 + 0.005 * (20.0 * (0.999 + 0.001) * (0.0 - 0.25)) and abs(xs[0] - 0.974) < 1e-15
    assert xl[0] == (1.0 + 0.005 * 3.5) + 0.005 * -0.25


# ---- max_error NaN folds (system.rs:101-109) -------------------------------------------------------------------
def test_kat_max_error_nan_folds():
    """Each fold starts at NaN and f64::max ignores a NaN operand (:103-107); the final f64::max does too (:108).
    v: |NaN − 0| = NaN → fold stays NaN.  xs: |0.5 − 0.25| = 0.25.  xl: empty → NaN.  max(NaN, max(0.25, NaN)) = 0.25."""
    a = (np.array([np.nan]), np.array([0.5]), np.zeros(0))
    b = (np.array([0.0]), np.array([0.25]), np.zeros(0))
    assert O.max_error(a, b) == 0.25
    assert pyref.max_error([list(x) for x in a], [list(x) for x in b]) == 0.25
    # inf − inf = NaN is ignored as well; a lone inf difference wins
    a = (np.array([INF, 1.0]), np.array([INF]), np.array([2.0]))
    b = (np.array([INF, 3.5]), np.array([0.0]), np.array([2.0]))
    assert O.max_error(a, b) == INF
    a = (np.array([INF, 1.0]), np.array([1.0]), np.array([2.0]))
    b = (np.array([INF, 3.5]), np.array([1.0]), np.array([2.0]))
    assert O.max_error(a, b) == 2.5


# ---- Q8: simulate_inter with zero steps -----------------------------------------------------------------------
@pytest.mark.parametrize("step_size", [0.01, float("nan")])
def test_kat_simulate_inter_zero_steps(golden_dir, step_size):
    """:274 state_res starts all-true, the step loop does not run, :353 position() = 0: replica 0 is returned although it
    is NOT satisfied, and no state is touched."""
    f = cnf.load_dimacs(str(golden_dir / "toy_mixed.cnf"))
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    v = np.array([[-0.5, 0.5, 0.25, -0.75, 0.0], [1.0, -1.0, -1.0, 1.0, 0.0]])
    xs = np.ones((2, 3)); xl = np.ones((2, 3))
    v0 = v.copy()
    a, w, st = F.simulate_inter(v, xs, xl, step_size=step_size, steps=0)
    assert w == 0 and st == 0
    assert list(a) == [0, 1, 1, 0, 0]                          # :238 v > 0 of replica 0
    assert np.array_equal(v, v0) and np.all(xs == 1.0) and np.all(xl == 1.0)
    assert not f.evaluate(a)                                   # clause (x1 ∨ ¬x5 ∨ x4): 0 ∨ ¬… check below
