"""GPU `stoch` (src/stoch.rs on the device, csrc/stoch.cuh) against the CPU oracle: integer state, so the bar is
bit-exact — every weight, every flip, every flag — for single steps, long batches with freezing and ragged formulas."""
import numpy as np
import pytest

from odesat_b200 import cnf
from odesat_b200 import stoch as ST
from odesat_b200 import system as S
from oracle import oracle as O

from helpers import ragged_formula, repeated_var_formula

pytestmark = pytest.mark.gpu


def both(f):
    return S.DeviceFormula(f), O.OracleFormula(f.varnum, f.clause_off, f.lits)


FORMULAS = {
    "toy": lambda g: cnf.load_dimacs(str(g / "toy_mixed.cnf")),
    "aim_sat": lambda g: cnf.load_dimacs(str(g / "aim100_sat.cnf")),
    "rand3": lambda g: cnf.random_ksat(300, 4.0, seed=4),
    "ragged": lambda g: ragged_formula(3),            # unit / empty clauses, repeated variables, unused variables
    "repeat3": lambda g: repeated_var_formula(9),
}


@pytest.mark.parametrize("name", list(FORMULAS))
def test_single_steps_bit_exact(golden_dir, name):
    f = FORMULAS[name](golden_dir)
    D, F = both(f)
    rng = np.random.default_rng(1)
    y = ST.State(rng.integers(0, 2, F.N).astype(np.uint8), rng.integers(1, 500, F.M).astype(np.uint64))
    y.xl[: min(3, F.M)] = np.uint64((1 << 64) - 7)                # saturating_add is exercised
    ov, oxl = y.v.copy(), y.xl.copy()
    for k in range(12):
        a = ST.step(y, D, seed=7, step_index=k, replica=2)
        b = F.stoch_step(ov, oxl, seed=7, replica=2, step=k)
        assert a == b and np.array_equal(y.v, ov) and np.array_equal(y.xl, oxl)


@pytest.mark.parametrize("R", [1, 33, 100])
@pytest.mark.parametrize("name", ["aim_sat", "ragged", "rand3"])
def test_batch_search_bit_exact_with_freezing(golden_dir, name, R):
    f = FORMULAS[name](golden_dir)
    D, F = both(f)
    steps = 700
    v = np.zeros((R, F.N), np.uint8); xl = np.ones((R, F.M), np.uint64)
    ov, oxl = v.copy(), xl.copy()
    first = F.stoch_batch(ov, oxl, seed=3, steps=steps, replica_offset=11)
    # chunk = steps: nobody stops early, every replica runs to its own flag (then frozen) or out of steps
    r = ST.search_batch(D, R, steps, seed=3, replica_offset=11, v=v, xl=xl, chunk=steps, write_back=True)
    assert np.array_equal(v, ov) and np.array_equal(xl, oxl)
    if (first >= 0).any():
        s_star = first[first >= 0].min()
        assert r.winner == int(np.argmax(first == s_star)) and r.steps_run == s_star + 1
        assert np.array_equal(r.solved_step, np.where(first == s_star, first, -1))
        assert f.evaluate(r.assignment) and r.verified[r.winner] == 1
    else:
        assert r.winner == -1 and r.steps_run == steps and np.array_equal(r.assignment, ov[0])
    assert np.array_equal(r.verified, np.array([f.evaluate(ov[q]) for q in range(R)], np.uint8))


def test_search_solves_the_fixture_and_stops_at_the_first_flag(golden_dir):
    f = FORMULAS["aim_sat"](golden_dir)
    D, F = both(f)
    r = ST.search_batch(D, 256, None, seed=2, chunk=32)            # unbounded: until some replica's step returns true
    assert r.winner >= 0 and f.evaluate(r.assignment)
    ov = np.zeros((256, F.N), np.uint8); oxl = np.ones((256, F.M), np.uint64)
    first = F.stoch_batch(ov, oxl, seed=2, steps=r.steps_run)
    assert first[r.winner] == r.steps_run - 1 and (first[first >= 0] >= r.steps_run - 1).all()
    assert ST.search(D, 20000, seed=2) is not None                 # the single-trajectory mirror runs
