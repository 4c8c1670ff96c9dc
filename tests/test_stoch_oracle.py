"""src/stoch.rs restated in the CPU oracle: hand-derived known answers for the deterministic part of a step (clause
satisfaction, the saturating weight updates, the per-variable weights that decide the flip probability), the flip rule
`r <= unsat` against draws recomputed here, and the flip frequencies against unsat / total."""
import numpy as np

from odesat_b200 import cnf
from oracle import oracle as O

M64 = (1 << 64) - 1


def sm64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def draw(seed, replica, step, var, total):
    """r in 1..=total: the counter-based stand-in for `rng.gen_range(1..=total)` (stoch.rs:68)."""
    h = sm64(sm64(seed ^ sm64(replica)) ^ sm64((step + 0x632BE59BD9B4E019) & M64) ^ ((var * 0xD1342543DE82EF95) & M64))
    return 1 + ((h * total) >> 64)


def toy(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "toy_mixed.cnf"))       # clauses [(0+),(3−),(2+)], [(0−),(3+),(1+),(2+)], [(1−),(2−)]
    return f, O.OracleFormula(f.varnum, f.clause_off, f.lits)


def test_kat_all_satisfied_state_returns_true_and_changes_nothing_but_the_weights(golden_dir):
    """v = all false (stoch.rs:84-87): ¬x3, ¬x0, ¬x1 make every clause true → xl ← max(xl ⊖ 1, 1): 1 → 1, 5 → 4; no
    unsatisfied weight anywhere → no flip; the step returns true (stoch.rs:77)."""
    f, F = toy(golden_dir)
    v = np.zeros(5, np.uint8); xl = np.array([1, 5, 2], np.uint64)
    assert F.stoch_step(v, xl, seed=9, replica=0, step=0) is True
    assert list(v) == [0, 0, 0, 0, 0] and list(xl) == [1, 4, 1]


def test_kat_unsatisfied_clause_weights_and_flip_rule(golden_dir):
    """v = [F, F, F, T, F]: clause 0 (x0 ∨ ¬x3 ∨ x2) is false → xl0 = 1 ⊕ 20 = 21; clauses 1 (¬x0 true) and 2 are true
    → stay 1.  Weights (UPDATED xl, stoch.rs:54-59): var0 total 21+1 = 22 / unsat 21; var3 22 / 21; var2 21+1+1 = 23 / 21;
    var1 1+1 = 2 / 0 (never flips); var4 occurs nowhere (the reference would panic on 1..=0; here: no flip).
    Flip iff r <= unsat (stoch.rs:70)."""
    f, F = toy(golden_dir)
    for seed in range(40):
        v = np.array([0, 0, 0, 1, 0], np.uint8); xl = np.ones(3, np.uint64)
        assert F.stoch_step(v, xl, seed=seed, replica=3, step=7) is False
        assert list(xl) == [21, 1, 1]
        exp = [0 ^ (draw(seed, 3, 7, 0, 22) <= 21), 0, 0 ^ (draw(seed, 3, 7, 2, 23) <= 21), 1 ^ (draw(seed, 3, 7, 3, 22) <= 21), 0]
        assert list(v) == [int(x) for x in exp]


def test_kat_saturating_weights(golden_dir):
    f, F = toy(golden_dir)
    v = np.array([0, 0, 0, 1, 0], np.uint8)
    xl = np.array([M64 - 5, 1, 7], np.uint64)
    F.stoch_step(v, xl, seed=1, replica=0, step=0)
    assert list(xl) == [M64, 1, 6]                                # saturating_add(20) hits u64::MAX; 7 ⊖ 1 = 6; max(1 ⊖ 1, 1) = 1


def test_flip_frequency_matches_unsat_over_total(golden_dir):
    f, F = toy(golden_dir)
    flips = np.zeros(5)
    n = 4000
    for seed in range(n):
        v = np.array([0, 0, 0, 1, 0], np.uint8); xl = np.ones(3, np.uint64)
        F.stoch_step(v, xl, seed=seed, replica=0, step=0)
        flips += v != np.array([0, 0, 0, 1, 0])
    p = flips / n
    assert abs(p[0] - 21 / 22) < 0.02 and abs(p[3] - 21 / 22) < 0.02 and abs(p[2] - 21 / 23) < 0.02 and p[1] == 0 and p[4] == 0


def test_search_solves_the_satisfiable_fixture_and_the_flag_implies_sat(golden_dir):
    sat = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    F = O.OracleFormula(sat.varnum, sat.clause_off, sat.lits)
    R = 8
    v = np.zeros((R, F.N), np.uint8); xl = np.ones((R, F.M), np.uint64)
    st = F.stoch_batch(v, xl, seed=5, steps=20000)
    assert (st >= 0).sum() >= 1
    for r in np.nonzero(st >= 0)[0]:
        assert sat.evaluate(v[r])                                 # the flagged state is left untouched and satisfies the CNF
