"""GPU parity proper: every entry point of the C ABI against the CPU oracle on identical inputs.

Bar: BIT-EXACT in both precisions (the kernels are compiled without FMA contraction and add
dv contributions in the reference's order), which is stronger than the north-star tolerances
(1e-12 relative f64, 1e-5 relative f32)."""
import numpy as np
import pytest

from odesat_b200 import _lib as L
from odesat_b200 import batch as B
from odesat_b200 import cnf
from odesat_b200 import system as S
from oracle import oracle as O

from helpers import ragged_formula, random_state, repeated_var_formula

pytestmark = pytest.mark.gpu


def both(f):
    return S.DeviceFormula(f), O.OracleFormula(f.varnum, f.clause_off, f.lits)


def eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


FORMULAS = {
    "toy": lambda g: cnf.load_dimacs(str(g / "toy_mixed.cnf")),
    "aim_sat": lambda g: cnf.load_dimacs(str(g / "aim100_sat.cnf")),
    "rand3": lambda g: cnf.random_ksat(300, 4.3, seed=4),
    "rand4": lambda g: cnf.random_ksat(120, 9.0, seed=5, k=4),
    "ragged": lambda g: ragged_formula(3),
    "repeat3": lambda g: repeated_var_formula(9),
}


@pytest.mark.parametrize("name", list(FORMULAS))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rhs_update_and_steps_bit_exact(golden_dir, name, dtype):
    f = FORMULAS[name](golden_dir)
    D, F = both(f)
    rng = np.random.default_rng(1)
    zeta = 0.01
    for trial in range(3):
        v, xs, xl = random_state(rng, F.N, F.M, dtype)
        if trial == 0:
            xs = F.init_short_term_memory(dtype)            # raw ±1 first-step memories (quirk Q5)
            xl = np.ones(F.M, dtype)
        y = S.State(v.copy(), xs.copy(), xl.copy())
        dy = S.State.zeros(D, dtype)
        allsat = S.compute_derivatives(y, dy, D, zeta)
        dv, dxs, dxl, oall = F.compute_derivatives(v, xs, xl, zeta)
        assert eq(dy.v, dv) and eq(dy.xs, dxs) and eq(dy.xl, dxl) and allsat == oall
        # update_state
        S.update_state(y, dy, 0.013, D)
        ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
        F.update_state(ov, oxs, oxl, dv, dxs, dxl, 0.013)
        assert eq(y.v, ov) and eq(y.xs, oxs) and eq(y.xl, oxl)
        # max_error
        a = S.State(v, xs, xl)
        assert S.max_error(a, y, D) == O.max_error((v, xs, xl), (ov, oxs, oxl)) or \
            (np.isnan(S.max_error(a, y, D)) and np.isnan(O.max_error((v, xs, xl), (ov, oxs, oxl))))
        # euler_step_fixed / euler_step
        s1 = S.State(v.copy(), xs.copy(), xl.copy())
        g1 = S.euler_step_fixed(s1, D, 0.01, zeta)
        o = [v.copy(), xs.copy(), xl.copy()]
        assert g1 == F.euler_step_fixed(*o, 0.01, zeta)
        assert eq(s1.v, o[0]) and eq(s1.xs, o[1]) and eq(s1.xl, o[2])
        s2 = S.State(v.copy(), xs.copy(), xl.copy())
        g2, dt2 = S.euler_step(s2, D, 1e-3, 0.01, zeta)
        o = [v.copy(), xs.copy(), xl.copy()]
        oa, odt = F.euler_step(*o, 1e-3, 0.01, zeta)
        assert g2 == oa and (dt2 == odt or (np.isnan(dt2) and np.isnan(odt)))
        assert eq(s2.v, o[0]) and eq(s2.xs, o[1]) and eq(s2.xl, o[2])


@pytest.mark.parametrize("name", ["rand3", "ragged", "repeat3"])
def test_sorted_clause_view_of_single_instances_is_invisible(golden_dir, name, monkeypatch):
    """Large single instances run on a view of the formula whose clauses are STORED by smallest variable (formula.hpp,
    -12 % at N = 1 M); the variable→clause lists keep the reference's summation order and xs / xl rows are permuted on the
    way in and out.  Forced on here for small formulas: RHS, update, error norm, fixed and adaptive steps and `simulate`
    must return the caller's clause order and the oracle's bits."""
    monkeypatch.setenv("ODESAT_GATHER_SORT", "1")
    monkeypatch.setenv("ODESAT_SMALL", "0")                  # the block-wide gather kernels, not the one-CTA small-instance kernel
    test_rhs_update_and_steps_bit_exact(golden_dir, name, np.float64)
    f = FORMULAS[name](golden_dir)
    D, F = both(f)
    v, xs, xl = F.init_batch(3, 1)
    st = S.State(v[0].copy(), xs[0].copy(), xl[0].copy())
    S.simulate(st, D, None, None, 60, None)                  # 60 adaptive steps (system.rs:156-239)
    ost, odt = F.batch_adaptive(v, xs, xl, 1e-3, f.default_zeta(), 60)
    assert eq(st.v, v[0]) and eq(st.xs, xs[0]) and eq(st.xl, xl[0])


def test_kat_satisfied_state_flags(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "toy_mixed.cnf"))
    D, _ = both(f)
    y = S.State(np.array([1., -1., -1., 1., 0.]), np.ones(3), np.ones(3))
    dy = S.State.zeros(D)
    assert S.compute_derivatives(y, dy, D, 0.001) is True
    assert list(dy.v) == [1, 0, 0, 1, 0] and list(dy.xl) == [-0.25] * 3
    assert S.euler_step_fixed(y, D, 0.01, 0.001) is True          # flag of the pre-update state
    np.testing.assert_allclose(y.xs, [0.94995] * 3, rtol=1e-15)
    assert list(S.init_short_term_memory(D)) == [1.0, 1.0, 1.0]


@pytest.mark.parametrize("name", ["traj_aim100_f64", "traj_aim100_f32", "traj_rand300_f64", "traj_toy_f64"])
@pytest.mark.parametrize("engine", [L.ENGINE_GATHER, L.ENGINE_TILE])
def test_golden_trajectories(golden_dir, name, engine):
    """100 fixed steps (dt = 0.01) and 50 adaptive steps against the committed fixtures."""
    g = np.load(golden_dir / f"{name}.npz")
    f = cnf.Formula(int(g["varnum"]), g["clause_off"], g["lits"], {})
    D = S.DeviceFormula(f)
    dtype = g["v0"].dtype
    prec = L.F32 if dtype == np.float32 else L.F64
    zeta = float(g["zeta"])
    try:
        b = B.ReplicaBatch(D, 1, prec, engine)
    except L.OdesatError as e:
        assert e.code == L.EUNSUPPORTED and engine == L.ENGINE_TILE
        pytest.skip("tile engine does not cover this formula")
    b.init(int(g["seed"]), 0)
    v, xs, xl = b.download()
    assert eq(v[0], g["v0"]) and eq(xs[0], g["xs0"]) and eq(xl[0], g["xl0"])    # device RNG == oracle RNG
    n = len(g["fixed_flags"])
    b.run_fixed(0.01, zeta, n, freeze=False)
    v, xs, xl = b.download()
    assert eq(v[0], g["fixed_v"]) and eq(xs[0], g["fixed_xs"]) and eq(xl[0], g["fixed_xl"])
    first = np.flatnonzero(g["fixed_flags"])
    st, done = b.status()
    assert done == n and st[0] == (first[0] if len(first) else -1)
    if engine == L.ENGINE_GATHER:
        b.init(int(g["seed"]), 0)
        b.run_adaptive(1e-3, zeta, len(g["adapt_dt"]))
        v, xs, xl = b.download()
        assert eq(v[0], g["adapt_v"]) and eq(xs[0], g["adapt_xs"]) and eq(xl[0], g["adapt_xl"])
        assert b.dt()[0] == g["adapt_dt"][-1]              # dt stops changing once flagged


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_simulate_matches_oracle(golden_dir, dtype):
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    D, F = both(f)
    for seed, kw in [(0, dict(step_size=0.01, steps=3000)), (1, dict(steps=4000)), (2, dict(tolerance=0.01, steps=50)),
                     (3, dict(step_size=0.01, steps=0))]:
        v = F.init_v0(seed, 0, dtype); xs = F.init_short_term_memory(dtype); xl = np.ones(F.M, dtype)
        st = S.State(v.copy(), xs.copy(), xl.copy())
        info = []
        res = S.simulate(st, D, kw.get("tolerance"), kw.get("step_size"), kw.get("steps"), None, info=info, chunk=17)
        oa, oflag, osteps, odt = F.simulate(v, xs, xl, tol=kw.get("tolerance", O.NAN), step_size=kw.get("step_size", O.NAN),
                                            steps=kw.get("steps", -1))
        assert info[0].steps_taken == osteps and info[0].allsat == oflag
        assert eq(st.v, v) and eq(st.xs, xs) and eq(st.xl, xl)
        assert res == [bool(x) for x in oa]
        if "step_size" not in kw:
            assert info[0].final_dt == odt
        if oflag:
            assert f.evaluate(res)


@pytest.mark.parametrize("engine", [L.ENGINE_GATHER, L.ENGINE_TILE])
@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_batch_fixed_bit_exact_with_freeze(golden_dir, engine, prec):
    """64 replicas × 1500 steps on the satisfiable fixture: states, per-replica first-flag steps and
    the frozen final states all equal the oracle's `batch` loop (main.rs:278-308)."""
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    D, F = both(f)
    dtype = B.np_dtype(prec)
    R, steps = 64, 1500
    b = B.ReplicaBatch(D, R, prec, engine)
    b.init(21, 100)
    v, xs, xl = F.init_batch(21, R, dtype, replica_offset=100)
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    ost = F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), steps, freeze=True)
    for n in (700, 1, 799):                                   # chunked: flags must carry across calls
        b.run_fixed(0.01, f.default_zeta(), n, freeze=True)
    gst, done = b.status()
    assert done == steps and eq(gst, ost) and (ost >= 0).any()
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    ver = b.verify()
    exp = np.array([f.evaluate(v[r] > 0) for r in range(R)], np.uint8)
    assert eq(ver, exp)
    k = b.first_solved()
    s_star = ost[ost >= 0].min()
    assert B.decode_key(k) == (s_star, int(np.argmax(ost == s_star)))
    assert eq(b.assignment(5), (v[5] > 0).astype(np.uint8))


def test_batch_adaptive_per_replica_dt(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    D, F = both(f)
    R, steps = 24, 600
    b = B.ReplicaBatch(D, R, L.F64, L.ENGINE_GATHER)
    b.init(5, 0)
    v, xs, xl = F.init_batch(5, R)
    ost, odt = F.batch_adaptive(v, xs, xl, 1e-3, f.default_zeta(), steps)
    b.run_adaptive(1e-3, f.default_zeta(), steps)
    gst, _ = b.status()
    gv, gxs, gxl = b.download()
    assert eq(gst, ost) and eq(b.dt(), odt)
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    assert (ost >= 0).sum() >= 3


def test_simulate_batch_call_batch_and_inter(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    D, F = both(f)
    R = 32
    v, xs, xl = F.init_batch(8, R)
    # INTER, caller-supplied states, lock-step write-back (chunk = 1)
    states = [S.State(v[r].copy(), xs[r].copy(), xl[r].copy()) for r in range(R)]
    info = []
    res = S.simulate_inter(states, D, None, 0.01, 4000, None, chunk=1, info=info)
    oa, ow, ost = F.simulate_inter(v, xs, xl, step_size=0.01, steps=4000)
    assert info[0] == (ow, ost) and res == [bool(x) for x in oa] and f.evaluate(res)
    for r in range(R):
        assert eq(states[r].v, v[r]) and eq(states[r].xs, xs[r]) and eq(states[r].xl, xl[r])
    # same thing with the default chunk: winner / assignment / step count unchanged
    v, xs, xl = F.init_batch(8, R)
    r2 = B.simulate_batch(D, R, v, xs, xl, step_size=0.01, steps=4000, precision=L.F64, mode=L.MODE_INTER)
    assert r2.winner == ow and r2.steps_run == ost and eq(r2.assignment, oa)
    # Q8: zero steps → replica 0
    r3 = B.simulate_batch(D, R, v, xs, xl, step_size=0.01, steps=0, precision=L.F64, mode=L.MODE_INTER)
    assert r3.winner == 0 and eq(r3.assignment, (v[0] > 0).astype(np.uint8))
    # BATCH: winner = lowest replica whose thresholded final state verifies (main.rs:305-307, quirk Q6)
    r4 = B.simulate_batch(D, R, seed=8, step_size=0.01, steps=1000, precision=L.F64, mode=L.MODE_BATCH)
    v, xs, xl = F.init_batch(8, R)
    ost = F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), 1000, freeze=True)
    exp = np.array([f.evaluate(v[r] > 0) for r in range(R)], np.uint8)
    assert eq(r4.solved_step, ost) and eq(r4.verified, exp)
    assert r4.winner == (int(np.argmax(exp)) if exp.any() else -1)
    if r4.winner >= 0:
        assert f.evaluate(r4.assignment)


def test_unsat_fixture_never_flags(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "aim100_unsat.cnf"))
    D, _ = both(f)
    res = B.simulate_batch(D, 100, seed=1, step_size=0.01, steps=1000, precision=L.F32, mode=L.MODE_BATCH)
    assert (res.solved_step == -1).all() and res.winner == -1 and not res.verified.any() and res.steps_run == 1000


def test_steps_to_solution_distribution_256_seeds(golden_dir):
    """≥256 seeds on the satisfiable fixture, dt = 0.01, n = 2000.  f64 is bit-exact, so the GPU's
    steps-to-flag vector EQUALS the oracle's; for f32 the two samples are compared with a
    two-sample KS test (p > 0.01), and every flagged replica's emitted assignment is verified."""
    from scipy import stats
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    D, F = both(f)
    R, steps = 256, 2000
    v, xs, xl = F.init_batch(77, R)
    ost = F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), steps, freeze=True, nthreads=O.host_cores())
    r64 = B.simulate_batch(D, R, seed=77, step_size=0.01, steps=steps, precision=L.F64)
    assert eq(r64.solved_step, ost)
    r32 = B.simulate_batch(D, R, seed=77, step_size=0.01, steps=steps, precision=L.F32)
    a = np.where(ost >= 0, ost, steps)
    b = np.where(r32.solved_step >= 0, r32.solved_step, steps)
    assert stats.ks_2samp(a, b).pvalue > 0.01
    assert stats.mannwhitneyu(a, b).pvalue > 0.01
    assert (r32.solved_step >= 0).sum() > R // 8
    # the throughput configuration (f32, BALANCED schedule: dv summed in colour order) draws from the same distribution
    r32b = B.simulate_batch(D, R, seed=77, step_size=0.01, steps=steps, precision=L.F32, schedule=L.SCHED_BALANCED)
    c = np.where(r32b.solved_step >= 0, r32b.solved_step, steps)
    assert stats.ks_2samp(a, c).pvalue > 0.01 and stats.mannwhitneyu(a, c).pvalue > 0.01
    # every emitted winner is verified exactly (cnf.rs:246-264), on the device and here on the host
    for r in (r64, r32, r32b):
        assert r.winner >= 0 and r.verified[r.winner] == 1 and f.evaluate(r.assignment)


def test_error_paths(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "toy_mixed.cnf"))
    D, _ = both(f)
    with pytest.raises(L.OdesatError):
        B.ReplicaBatch(D, -1)
    b = B.ReplicaBatch(D, 4, L.F64)
    with pytest.raises(L.OdesatError):
        b.assignment(9)
    with pytest.raises(ValueError):
        b.upload(np.zeros((4, 5), np.float32), np.zeros((4, 3)), np.zeros((4, 3)))
    z = B.ReplicaBatch(D, 0, L.F64)                           # empty batch is legal
    z.run_fixed(0.01, 0.001, 3)
    assert z.status()[1] == 3
    e = cnf.Formula(4, np.zeros(1, np.int64), np.zeros(0, np.int32), {})    # no clauses: allsat at once
    DE = S.DeviceFormula(e)
    y = S.State(np.array([0.1, -0.2, 0.3, 0.0]), np.zeros(0), np.zeros(0))
    assert S.euler_step_fixed(y, DE, 0.01, 0.001) is True


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_large_single_instance_adaptive_and_fixed(dtype):
    """Single large instance (the shape of BASELINE configs[3], scaled to what the oracle does in
    seconds): N = 200 000, alpha = 4.2, R = 1 — adaptive and fixed steps bit-exact."""
    f = cnf.random_ksat(200_000, 4.2, seed=20240614)
    D, F = both(f)
    v = F.init_v0(1, 0, dtype); xs = F.init_short_term_memory(dtype); xl = np.ones(F.M, dtype)
    st = S.State(v.copy(), xs.copy(), xl.copy())
    info = []
    S.simulate(st, D, 1e-3, None, 6, None, info=info)
    oa, oflag, osteps, odt = F.simulate(v, xs, xl, tol=1e-3, steps=6)
    assert info[0].steps_taken == osteps == 6 and info[0].final_dt == odt
    assert eq(st.v, v) and eq(st.xs, xs) and eq(st.xl, xl)
    S.simulate(st, D, None, 0.01, 5, None)
    F.simulate(v, xs, xl, step_size=0.01, steps=5)
    assert eq(st.v, v) and eq(st.xs, xs) and eq(st.xl, xl)


@pytest.mark.parametrize("small", ["1", "0"])
@pytest.mark.parametrize("name", ["aim_sat", "ragged", "rand4", "repeat3"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_persistent_small_kernel_and_general_engine_agree_with_oracle(golden_dir, monkeypatch, small, name, dtype):
    """The persistent one-CTA-per-replica kernel (kernels_small.cuh; ODESAT_SMALL=1, default) and the
    general two-phase engine (ODESAT_SMALL=0) integrate the same trajectories as the oracle, bit for
    bit: adaptive steps with per-replica dt, fixed steps with and without freezing, ragged clauses."""
    monkeypatch.setenv("ODESAT_SMALL", small)
    f = FORMULAS[name](golden_dir)
    D, F = both(f)
    prec = L.F64 if dtype == np.float64 else L.F32
    R = 19
    zeta = f.default_zeta()
    # adaptive, 300 steps (some aim-100 replicas flag on the way and must stop untouched)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_GATHER)
    v, xs, xl = F.init_batch(7, R, dtype)
    b.upload(v, xs, xl)
    for n in (1, 120, 179):
        b.run_adaptive(1e-3, zeta, n)
    ost, odt = F.batch_adaptive(v, xs, xl, 1e-3, zeta, 300)
    gst, nsteps = b.status()
    gv, gxs, gxl = b.download()
    assert nsteps == 300 and eq(gst, ost) and eq(b.dt().astype(dtype), odt)
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    # fixed with freezing, continuing from the adaptive state (flags reset by the upload)
    for freeze in (True, False):
        b.upload(v, xs, xl)
        b.run_fixed(0.01, zeta, 150, freeze=freeze)
        ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
        ost = F.batch_fixed(ov, oxs, oxl, 0.01, zeta, 150, freeze=freeze)
        gst, _ = b.status()
        gv, gxs, gxl = b.download()
        assert eq(gst, ost)
        assert eq(gv, ov) and eq(gxs, oxs) and eq(gxl, oxl)
    b.close()


@pytest.mark.parametrize("small", ["1", "0"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_adaptive_inter_shares_one_dt_across_replicas_like_the_reference(golden_dir, monkeypatch, small, dtype):
    """system.rs:312-349 (quirk Q7): in adaptive `inter` the replicas step one after the other inside an
    outer step and share ONE dt.  Both device paths — the sequential one-CTA kernel (ODESAT_SMALL=1) and
    the general engine stepping one replica at a time (ODESAT_SMALL=0) — reproduce the oracle exactly:
    winner, outer steps, and the final state of EVERY replica (the loop stops right after the outer step
    in which the first replica flags, as the reference's does)."""
    monkeypatch.setenv("ODESAT_SMALL", small)
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    D, F = both(f)
    R = 5
    prec = L.F64 if dtype == np.float64 else L.F32
    for steps in (40, 4000):                                # a budget that runs out, and one that reaches a flag
        v, xs, xl = F.init_batch(21, R, dtype)
        gv, gxs, gxl = v.copy(), xs.copy(), xl.copy()
        oa, owin, osteps = F.simulate_inter(v, xs, xl, tol=1e-3, steps=steps)
        res = B.simulate_batch(D, R, gv, gxs, gxl, tolerance=1e-3, steps=steps, precision=prec, mode=L.MODE_INTER,
                               write_back=True)
        assert res.winner == owin and eq(res.assignment, oa)
        assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
        if owin >= 0:
            assert res.steps_run == osteps and f.evaluate(res.assignment)
    # the Rust-shaped mirror (f64 states in a list, as main.rs:348-360 builds them)
    if dtype == np.float64:
        v, xs, xl = F.init_batch(3, 3, np.float64)
        states = [S.State(v[r].copy(), xs[r].copy(), xl[r].copy()) for r in range(3)]
        info = []
        got = S.simulate_inter(states, D, 1e-3, None, 3000, None, info=info)
        oa, owin, osteps = F.simulate_inter(v, xs, xl, tol=1e-3, steps=3000)
        assert [int(x) for x in got] == list(oa) and info[0][0] == owin
        for r in range(3):
            assert eq(states[r].v, v[r]) and eq(states[r].xs, xs[r]) and eq(states[r].xl, xl[r])
