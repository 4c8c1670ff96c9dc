"""Ragged clause lengths in the TILE engine (csrc/tile_ragged.cuh): lengths 0..7, unit and empty clauses (quirk Q9: ±inf
derivatives), variables repeated inside a clause, uniform k = 4 — bit-identical to the oracle with the EXACT schedule
and to the GATHER engine, flags and freezing included."""
import numpy as np
import pytest

from odesat_b200 import _lib as L
from odesat_b200 import batch as B
from odesat_b200 import cnf
from odesat_b200 import system as S
from oracle import oracle as O

from helpers import ragged_formula, random_state, repeated_var_formula

pytestmark = pytest.mark.gpu


def eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


def both(f):
    return S.DeviceFormula(f), O.OracleFormula(f.varnum, f.clause_off, f.lits)


def mixed_23sat(n_vars, n2, n3, seed):
    """Binary and ternary clauses with distinct variables (what structured instances mostly consist of)."""
    rng = np.random.default_rng(seed)
    off, lits = [0], []
    for m in range(n2 + n3):
        k = 2 if m % (n2 + n3) < n2 else 3
        vs = rng.choice(n_vars, size=k, replace=False) + 1
        sg = rng.integers(0, 2, size=k) * 2 - 1
        lits.extend(int(a * b) for a, b in zip(vs, sg))
        off.append(len(lits))
    order = rng.permutation(n2 + n3)                          # interleave the two kinds
    o2, l2 = [0], []
    for m in order:
        l2.extend(lits[off[m]:off[m + 1]])
        o2.append(len(l2))
    return cnf.Formula(n_vars, np.asarray(o2, np.int64), np.asarray(l2, np.int32), {})


def long_clause_formula(n_vars, seed):
    """3-literal clauses mixed with clauses of 4, 7, 9, 17, 33 and 40 distinct literals (group sizes 4 / 8 / 16 / 32 and the
    serial walk beyond 32)."""
    rng = np.random.default_rng(seed)
    off, lits = [0], []
    for m in range(900):
        k = int(rng.choice([3, 3, 3, 3, 4, 7, 9, 17, 33, 40]))
        vs = rng.choice(n_vars, size=k, replace=False) + 1
        sg = rng.integers(0, 2, size=k) * 2 - 1
        lits.extend(int(a * b) for a, b in zip(vs, sg))
        off.append(len(lits))
    return cnf.Formula(n_vars, np.asarray(off, np.int64), np.asarray(lits, np.int32), {})


FORMULAS = {
    "long": lambda: long_clause_formula(500, 8),
    "ragged": lambda: ragged_formula(3, 300, 1500),           # lengths 0..6, repeats, one empty clause, unused variables
    "k4": lambda: cnf.random_ksat(400, 5.0, seed=1, k=4),
    "repeat3": lambda: repeated_var_formula(2, 200, 900),     # uniform 3 with (x, x, y) and (x, -x, y)
    "mixed23": lambda: mixed_23sat(600, 700, 1500, 4),
}


@pytest.mark.parametrize("prec", [L.F64, L.F32])
@pytest.mark.parametrize("name", sorted(FORMULAS))
@pytest.mark.parametrize("R,groups", [(1, "1"), (33, "1"), (33, "0")])
def test_tile_ragged_exact_vs_oracle(prec, name, R, groups, monkeypatch):
    """groups = 1: clauses of 4..32 distinct literals are GROUP clauses (one lane per literal, the EXACT default);
    groups = 0: one thread walks them (clause_loop8 / clause_loop, the BALANCED default).  Same bits either way."""
    monkeypatch.setenv("ODESAT_TILE_GROUPS", groups)
    f = FORMULAS[name]()
    D, F = both(f)
    dtype = B.np_dtype(prec)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    assert b.engine == L.ENGINE_TILE
    v, xs, xl = F.init_batch(4, R, dtype)
    b.upload(v, xs, xl)
    for n in (1, 70, 59):                                     # 130 steps: crosses the 64-step launch chunk
        b.run_fixed(0.01, f.default_zeta(), n, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), 130, freeze=False, nthreads=4)
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_tile_ragged_states_outside_the_unit_box_and_large_zeta(prec):
    """The literal first step (STRICT) and the loop clauses' literal arithmetic with a rigidity term that matters."""
    f = ragged_formula(5, 120, 500)
    D, F = both(f)
    dtype = B.np_dtype(prec)
    rng = np.random.default_rng(5)
    R = 16
    v, xs, xl = random_state(rng, F.N, F.M, dtype, R=R)
    v[:, ::7] = (rng.uniform(-3, 3, size=v[:, ::7].shape)).astype(dtype)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    g = B.ReplicaBatch(D, R, prec, L.ENGINE_GATHER)
    for zeta in (0.5, float("inf")):
        b.upload(v, xs, xl); g.upload(v, xs, xl)
        b.run_fixed(0.01, zeta, 5, freeze=False); g.run_fixed(0.01, zeta, 5, freeze=False)
        ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
        F.batch_fixed(ov, oxs, oxl, 0.01, zeta, 5, freeze=False)
        for q in (b, g):
            gv, gxs, gxl = q.download()
            assert eq(gv, ov) and eq(gxs, oxs) and eq(gxl, oxl)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_tile_ragged_flags_freezing_and_verification(prec):
    """An easy mixed 2/3-SAT instance: replicas flag at different steps, take their update (system.rs:149-153) and
    freeze; flags, states and the exact verification of every replica equal the oracle's / the host evaluation."""
    f = mixed_23sat(400, 250, 700, 9)
    D, F = both(f)
    dtype = B.np_dtype(prec)
    R = 41
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    v, xs, xl = F.init_batch(6, R, dtype)
    b.upload(v, xs, xl)
    for n in (300, 500, 700):
        b.run_fixed(0.01, f.default_zeta(), n, freeze=True)
    ost = F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), 1500, freeze=True, nthreads=4)
    assert (ost >= 0).sum() >= 3
    gst, _ = b.status()
    gv, gxs, gxl = b.download()
    assert eq(gst, ost) and eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    exp = np.array([f.evaluate(v[r] > 0) for r in range(R)], np.uint8)
    assert eq(b.verify(), exp)
    assert eq(b.assignment(3), (v[3] > 0).astype(np.uint8))


def test_tile_ragged_mid_size_equals_gather_and_balanced_is_close():
    """N = 6 000 with 30 % binary clauses, f32, 64 replicas, 60 steps: tile (EXACT) == gather bit for bit; BALANCED within
    rounding."""
    f = mixed_23sat(6000, 7000, 17000, 12)
    D = S.DeviceFormula(f)
    R = 64
    t = B.ReplicaBatch(D, R, L.F32, L.ENGINE_AUTO, L.SCHED_EXACT)
    assert t.engine == L.ENGINE_TILE
    g = B.ReplicaBatch(D, R, L.F32, L.ENGINE_GATHER)
    c = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
    for q in (t, g, c):
        q.init(1, 0)
        q.run_fixed(0.01, f.default_zeta(), 60, freeze=False)
    tv, txs, txl = t.download()
    gv, gxs, gxl = g.download()
    cv, cxs, cxl = c.download()
    assert eq(tv, gv) and eq(txs, gxs) and eq(txl, gxl)
    assert np.max(np.abs(cv - gv)) <= 2e-4 and np.max(np.abs(cxs - gxs)) <= 2e-4 and np.max(np.abs(cxl / gxl - 1)) <= 2e-4


def test_simulate_batch_on_a_ragged_formula_through_the_c_abi():
    """One odesat_simulate_batch call (AUTO → tile engine) on the mixed instance: flags and verification as the oracle's."""
    f = mixed_23sat(400, 250, 700, 9)
    D, F = both(f)
    R, steps = 40, 1200
    v, xs, xl = F.init_batch(6, R, np.float64)
    res = B.simulate_batch(D, R, v=v.copy(), step_size=0.01, steps=steps, precision=L.F64, mode=L.MODE_BATCH)
    ost = F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), steps, freeze=True, nthreads=4)
    assert eq(res.solved_step, ost)
    exp = np.array([f.evaluate(v[r] > 0) for r in range(R)], np.uint8)
    assert eq(res.verified, exp)


@pytest.mark.parametrize("groups", ["1", "0"])
def test_ragged_tile_kernel_soak_full_batch_equals_gather_engine(groups, monkeypatch):
    """4 096 replicas of a 10 000-variable formula with binary, ternary, 5- and 8-literal clauses, 40 steps, EXACT: the
    ragged tile kernel (group clauses / one thread per long clause) and the gather engine end in identical states."""
    monkeypatch.setenv("ODESAT_TILE_GROUPS", groups)
    rng = np.random.default_rng(3)
    ks = np.concatenate([np.full(n, k) for k, n in {2: 12_000, 3: 28_000, 5: 2_500, 8: 500}.items()])
    rng.shuffle(ks)
    off, lits = [0], []
    for k in ks:
        vs = rng.choice(10_000, size=int(k), replace=False) + 1
        sg = rng.integers(0, 2, size=int(k)) * 2 - 1
        lits.extend(int(a * b) for a, b in zip(vs, sg))
        off.append(len(lits))
    f = cnf.Formula(10_000, np.asarray(off, np.int64), np.asarray(lits, np.int32), {})
    D = S.DeviceFormula(f)
    R = 4096
    t = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_EXACT)
    g = B.ReplicaBatch(D, R, L.F32, L.ENGINE_GATHER)
    for q in (t, g):
        q.init(3, 0)
        q.run_fixed(0.01, f.default_zeta(), 40, freeze=False)
    tv, txs, txl = t.download()
    gv, gxs, gxl = g.download()
    assert eq(tv, gv) and eq(txs, gxs) and eq(txl, gxl)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
@pytest.mark.parametrize("name", sorted(FORMULAS))
def test_tile_ragged_adaptive_steps_exact_vs_oracle(prec, name, monkeypatch):
    """Adaptive steps (system.rs:111-139) on ragged formulas in the tile engine (k_tile_adaptive<…, RAGGED>): one- and
    two-literal packed clauses and loop clauses in both passes; states, per-replica dt and flags equal the oracle's.
    (Group clauses — the EXACT default for 4..32 literals — have no adaptive form: switched off here.)"""
    monkeypatch.setenv("ODESAT_TILE_GROUPS", "0")
    f = FORMULAS[name]()
    D, F = both(f)
    dtype = B.np_dtype(prec)
    R = 33
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    assert b.engine == L.ENGINE_TILE
    v, xs, xl = F.init_batch(4, R, dtype)
    b.upload(v, xs, xl)
    for n in (1, 40, 29):
        b.run_adaptive(1e-3, f.default_zeta(), n)
    ost, odt = F.batch_adaptive(v, xs, xl, 1e-3, f.default_zeta(), 70, nthreads=4)
    gv, gxs, gxl = b.download()
    assert eq(b.status()[0], ost) and eq(b.dt(), odt)
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)


def test_adaptive_batch_on_ragged_formulas_picks_an_engine_that_can():
    """AUTO for adaptive steps: 2/3-SAT mixes run on the tile engine's adaptive kernel; a formula whose long clauses are
    group clauses (EXACT) falls back to the gather engine instead of failing — same results as the oracle either way."""
    for make, steps in ((lambda: mixed_23sat(6000, 7000, 17000, 12), 12), (lambda: long_clause_formula(500, 8), 30)):
        f = make()
        D, F = both(f)
        R = 24
        v, xs, xl = F.init_batch(2, R, np.float64)
        res = B.simulate_batch(D, R, v=v.copy(), steps=steps, precision=L.F64, mode=L.MODE_BATCH, write_back=False)
        ost, _ = F.batch_adaptive(v, xs, xl, 1e-3, f.default_zeta(), steps, nthreads=4)
        assert eq(res.solved_step, ost)
        exp = np.array([f.evaluate(v[r] > 0) for r in range(R)], np.uint8)
        assert eq(res.verified, exp)
