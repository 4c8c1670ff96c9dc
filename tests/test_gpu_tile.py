"""The shared-memory-resident TILE engine against the oracle and against the GATHER engine:
bit-exact with the EXACT (order-preserving) schedule, rounding-level with BALANCED; ragged replica
counts, states outside [-1, 1] (literal rigidity term), and BASELINE.json's full size."""
import numpy as np
import pytest

from odesat_b200 import _lib as L
from odesat_b200 import batch as B
from odesat_b200 import cnf
from odesat_b200 import system as S
from oracle import oracle as O

from helpers import random_state

pytestmark = pytest.mark.gpu


def eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
@pytest.mark.parametrize("R", [1, 2, 33, 257])
def test_tile_exact_vs_oracle_ragged_replica_counts(prec, R):
    f = cnf.random_ksat(500, 4.3, seed=12)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    assert b.engine == L.ENGINE_TILE
    v, xs, xl = F.init_batch(4, R, dtype)
    b.upload(v, xs, xl)
    for n in (1, 70, 59):                                     # 130 steps: crosses the 64-step launch chunk
        b.run_fixed(0.01, 0.001, n, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, 0.001, 130, freeze=False, nthreads=4)
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_tile_literal_rigidity_term_for_states_outside_unit_box(prec):
    """|v| > 1 on entry makes system.rs:73's branch reachable with r != 0; the engine then runs the
    first step with the literal term.  zeta is large so a wrong shortcut would show."""
    f = cnf.random_ksat(200, 4.3, seed=3)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    rng = np.random.default_rng(5)
    R = 16
    v, xs, xl = random_state(rng, F.N, F.M, dtype, R=R)
    # literals with value exactly half the clause minimum: v = 1 - c·… engineered via 1-q·v = 0.5·min
    v[:, ::7] = (rng.uniform(-3, 3, size=v[:, ::7].shape)).astype(dtype)
    v[:, 1] = 3.0; v[:, 2] = 2.0                              # 1-v = -2, -1 → c = 0.5·(-2) = -1 == 1-2: branch taken
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    g = B.ReplicaBatch(D, R, prec, L.ENGINE_GATHER)
    for zeta in (0.5, float("inf")):
        b.upload(v, xs, xl); g.upload(v, xs, xl)
        b.run_fixed(0.01, zeta, 5, freeze=False); g.run_fixed(0.01, zeta, 5, freeze=False)
        ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
        F.batch_fixed(ov, oxs, oxl, 0.01, zeta, 5, freeze=False)
        for got in (b.download(), g.download()):
            assert eq(got[0], ov) and eq(got[1], oxs) and eq(got[2], oxl)


def test_tile_balanced_schedule_agrees_to_rounding():
    f = cnf.random_ksat(800, 4.3, seed=2)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 32
    v, xs, xl = F.init_batch(9, R, np.float64)
    b = B.ReplicaBatch(D, R, L.F64, L.ENGINE_TILE, L.SCHED_BALANCED)
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.001, 1, freeze=False)
    g1 = b.download()
    o = [v.copy(), xs.copy(), xl.copy()]
    F.batch_fixed(*o, 0.01, 0.001, 1, freeze=False)
    # one step: only the summation order of dv differs → north-star tolerance 1e-12 relative
    np.testing.assert_allclose(g1[0], o[0], rtol=1e-12, atol=1e-15)
    assert eq(g1[1], o[1]) and eq(g1[2], o[2])                # memories do not depend on the order
    # and it is deterministic run to run
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.001, 1, freeze=False)
    g2 = b.download()
    assert eq(g1[0], g2[0])
    # 100 steps: trajectories stay together far below the f32-vs-f64 spread
    b.run_fixed(0.01, 0.001, 99, freeze=False)
    F.batch_fixed(*o, 0.01, 0.001, 99, freeze=False)
    g3 = b.download()
    assert np.max(np.abs(g3[0] - o[0])) < 1e-9


def test_fixed_trajectory_tolerance_f32_vs_f64_reference():
    """North star: 100 fixed steps vs the reference's f64 arithmetic.  The f32 kernels are bit-exact
    against the f32 oracle; against f64 the stated bound is the measured f32-oracle-vs-f64-oracle
    spread (the dynamics switch on argmin, so rounding is amplified): max |Δv| ≤ 5e-3 on this instance."""
    f = cnf.random_ksat(1000, 4.3, seed=20240611)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 8
    v, xs, xl = F.init_batch(1, R, np.float64)
    b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE)
    b.upload(v.astype(np.float32), xs.astype(np.float32), xl.astype(np.float32))
    b.run_fixed(0.01, 0.001, 100, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, 0.001, 100, freeze=False)
    gv, gxs, gxl = b.download()
    assert np.max(np.abs(gv - v)) <= 5e-3
    assert np.max(np.abs(gxs - xs)) <= 5e-3
    np.testing.assert_allclose(gxl, xl, rtol=2e-3)


@pytest.mark.parametrize("prec", [L.F32, L.F64])
def test_full_size_baseline_config_engines_agree_and_match_oracle(prec):
    """BASELINE.json configs[2]: N = 10 000, alpha = 4.3, 4096 replicas.  (a) the two engines —
    independent kernels, different layouts — produce bit-identical states for ALL replicas after
    12 steps; (b) a sample of replicas equals the oracle; (c) invariants: clamps hold, run is
    deterministic, every thresholded state is checked by the exact verifier without a false SAT."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240613)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 4096 if prec == L.F32 else 1024
    dtype = B.np_dtype(prec)
    t = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE)
    t.init(1, 0)
    t.run_fixed(0.01, 0.001, 12, freeze=False)
    tv, txs, txl = t.download()
    ver = t.verify()
    t.close()
    g = B.ReplicaBatch(D, R, prec, L.ENGINE_GATHER)
    g.init(1, 0)
    g.run_fixed(0.01, 0.001, 12, freeze=False)
    gv, gxs, gxl = g.download()
    g.close()
    assert eq(tv, gv) and eq(txs, gxs) and eq(txl, gxl)
    sample = [0, 1, R // 2 + 1, R - 1]
    for r in sample:
        v = F.init_v0(1, r, dtype); xs = F.init_short_term_memory(dtype); xl = np.ones(F.M, dtype)
        for _ in range(12):
            F.euler_step_fixed(v, xs, xl, 0.01, 0.001)
        assert eq(tv[r], v) and eq(txs[r], xs) and eq(txl[r], xl)
    assert tv.min() >= -1 and tv.max() <= 1 and txs.min() >= dtype(0.001) and txs.max() <= dtype(1) - dtype(0.001)
    assert txl.min() >= 1 and txl.max() <= 1e4 * F.M
    for r in sample:
        assert bool(ver[r]) == f.evaluate(tv[r] > 0)


def test_tile_balanced_f32_meets_north_star_tolerance():
    """Throughput configuration (f32, BALANCED schedule): one RHS + Euler step on identical state
    within 1e-5 relative of the f32 oracle (north star), memories bit-exact."""
    f = cnf.random_ksat(3000, 4.3, seed=8)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    rng = np.random.default_rng(2)
    R = 64
    v, xs, xl = random_state(rng, F.N, F.M, np.float32, R=R)
    b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.001, 1, freeze=False)
    gv, gxs, gxl = b.download()
    F.batch_fixed(v, xs, xl, 0.01, 0.001, 1, freeze=False)
    np.testing.assert_allclose(gv, v, rtol=1e-5, atol=1e-6)
    assert eq(gxs, xs) and eq(gxl, xl)


@pytest.mark.parametrize("prec", [L.F32, L.F64])
def test_long_trajectory_stays_bit_identical(prec):
    """5 000 fixed steps of a random 3-SAT instance near the threshold: variables saturate at ±1, clause
    minima tie, memories hit their clamps, some replicas flag and freeze — the EXACT tile schedule, the
    general engine and the oracle still agree bit for bit at the end (and on the flag steps)."""
    f = cnf.random_ksat(2000, 4.0, seed=77)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    R, steps = 12, 5000
    v, xs, xl = F.init_batch(5, R, dtype)
    t = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    g = B.ReplicaBatch(D, R, prec, L.ENGINE_GATHER)
    for b in (t, g):
        b.upload(v, xs, xl)
        b.run_fixed(0.05, 0.001, steps, freeze=True)
    ost = F.batch_fixed(v, xs, xl, 0.05, 0.001, steps, freeze=True, nthreads=O.host_cores())
    for b in (t, g):
        st, _ = b.status()
        gv, gxs, gxl = b.download()
        assert eq(st, ost)
        assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    assert (np.abs(v) == 1).mean() > 0.02                      # the run really touches the clamps


@pytest.mark.parametrize("prec", [L.F32, L.F64])
def test_tma_fed_kernel_equals_the_per_thread_ring_at_the_bench_configuration(monkeypatch, prec):
    """BASELINE configs[2] shape, BALANCED schedule (the bench default): the kernel whose ring is fed by
    cp.async.bulk + mbarriers (k_tile_fixed_tma, default at 768 threads) and the per-thread cp.async kernel walk
    the same schedule, so 70 steps must agree BIT FOR BIT; and one step agrees with the oracle within the
    north-star tolerance (the BALANCED order differs from the reference's only in the summation order of dv)."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240613)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    R = 96
    v, xs, xl = F.init_batch(3, R, dtype)
    out = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("ODESAT_TILE_TMA", tma)
        b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_BALANCED)
        b.upload(v, xs, xl)
        b.run_fixed(0.01, 0.001, 1, freeze=False)
        first = b.download()
        b.run_fixed(0.01, 0.001, 69, freeze=False)          # crosses the 64-step launch chunk
        out[tma] = (first, b.download())
        b.close()
    for a, c in zip(out["1"][0] + out["1"][1], out["0"][0] + out["0"][1]):
        assert eq(a, c)
    o = [v.copy(), xs.copy(), xl.copy()]
    F.batch_fixed(*o, 0.01, 0.001, 1, freeze=False, nthreads=O.host_cores())
    g = out["1"][0]
    tol = dict(rtol=1e-5, atol=1e-6) if prec == L.F32 else dict(rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(g[0], o[0], **tol)
    assert eq(g[1], o[1]) and eq(g[2], o[2])


def test_hundred_step_tolerance_of_the_benched_configuration():
    """north_star: "fixed-step trajectories must agree over 100 steps within a stated tolerance".  For the configuration
    bench.py times — f32, BALANCED schedule, the TMA-fed kernel, the N = 10 000 / alpha = 4.3 formula of BASELINE
    configs[2] — the STATED bound against the f32 oracle after 100 steps of dt = 0.01 is
        max |dv| <= 2e-5,  max |dxs| <= 2e-5,  max |xl / xl_ref - 1| <= 2e-5.
    BALANCED differs from the reference only in the order in which a variable's clause contributions are added; on the
    CPU, two f32 oracle runs that differ only in that order (clauses permuted) are 8e-7 / 1.4e-6 / 1.2e-6 apart after
    100 steps and 7e-5 after 300 (the dynamics switch on argmin, so rounding differences grow), which is where the
    bound comes from.  The EXACT schedule has no such term: bit-identical (test_long_trajectory_stays_bit_identical)."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)          # bench.py's formula
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 64
    v, xs, xl = F.init_batch(1, R, np.float32)
    b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
    b.upload(v, xs, xl)
    b.run_fixed(0.01, f.default_zeta(), 100, freeze=False)
    gv, gxs, gxl = b.download()
    b.close()
    F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), 100, freeze=False, nthreads=O.host_cores())
    dv, dxs, dxl = np.abs(gv - v).max(), np.abs(gxs - xs).max(), np.abs(gxl / xl - 1).max()
    print(f"100-step deviation of the benched configuration vs the f32 oracle: v {dv:.3g} xs {dxs:.3g} xl(rel) {dxl:.3g}")
    assert dv <= 2e-5 and dxs <= 2e-5 and dxl <= 2e-5
    assert dv > 0                                                  # it IS a different summation order


def test_tma_ring_soak_4096_replicas_1024_steps(monkeypatch):
    """Race evidence in lieu of compute-sanitizer (closed on this pool): the TMA-fed kernel (bulk copies + mbarriers +
    cross-proxy fences) and the per-thread cp.async ring integrate the full bench batch — 4 096 replicas, N = 10 000 —
    for 1 024 steps (16 launches of 64 steps, ring wrap-arounds across steps and launches) and must end BIT-IDENTICAL."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)
    D = S.DeviceFormula(f)
    R = 4096
    out = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("ODESAT_TILE_TMA", tma)
        b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
        b.init(1, 0)
        b.run_fixed(0.01, f.default_zeta(), 1024, freeze=False)
        out[tma] = b.download()
        b.close()
    for a, c in zip(out["1"], out["0"]):
        assert eq(a, c)
    assert (np.abs(out["1"][0]) == 1).mean() > 0.02


@pytest.mark.parametrize("prec", [L.F32, L.F64])
def test_warp_specialised_kernel_equals_the_per_thread_ring(monkeypatch, prec):
    """k_tile_ws (producer warp + full/empty mbarriers, barriers only between the wide BALANCED levels, persistent CTAs
    with a (sub-chunk, tile) work queue) walks the same schedule as the per-thread cp.async kernel: bit-identical
    states, whatever the sub-chunk size (1 step: every step of a tile may run on another SM; 3: ragged last
    sub-chunk; 0: automatic), across the 64-step launch chunk."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240613)
    D = S.DeviceFormula(f)
    dtype = B.np_dtype(prec)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 331                                                       # odd: the last tile is half empty (f32)
    v, xs, xl = F.init_batch(3, R, dtype)
    out = {}
    for name, ws, ksub in (("ring", "0", "0"), ("ws", "1", "0"), ("ws1", "1", "1"), ("ws3", "1", "3")):
        monkeypatch.setenv("ODESAT_TILE_WS", ws)
        monkeypatch.setenv("ODESAT_TILE_TMA", "0")
        monkeypatch.setenv("ODESAT_TILE_NT", "704")              # same CTA width, hence the same schedule, for both kernels
        monkeypatch.setenv("ODESAT_TILE_KSUB", ksub)
        b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_BALANCED)
        b.upload(v, xs, xl)
        b.run_fixed(0.01, 0.001, 1, freeze=False)
        first = b.download()
        b.run_fixed(0.01, 0.001, 69, freeze=False)
        out[name] = first + b.download()
        b.close()
    for name in ("ws", "ws1", "ws3"):
        for a, c in zip(out[name], out["ring"]):
            assert eq(a, c), name
    o = [v.copy(), xs.copy(), xl.copy()]
    F.batch_fixed(*o, 0.01, 0.001, 1, freeze=False, nthreads=O.host_cores())
    tol = dict(rtol=1e-5, atol=1e-6) if prec == L.F32 else dict(rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(out["ws"][0], o[0], **tol)
    assert eq(out["ws"][1], o[1]) and eq(out["ws"][2], o[2])


def test_warp_specialised_kernel_flags_and_freezes_like_the_ring(monkeypatch):
    """Replicas that flag are frozen (system.rs:149-153 then dt = 0) and their flag steps recorded — also when a tile's
    sub-chunks run on different SMs, and when a whole tile is frozen (its work items become no-ops)."""
    f = cnf.random_ksat(4000, 3.0, seed=5)                        # flags between steps ~340 and ~450 at dt = 0.1
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 301
    v, xs, xl = F.init_batch(11, R, np.float32)
    out = {}
    for name, ws, ksub in (("ring", "0", "0"), ("ws", "1", "0"), ("ws2", "1", "2")):
        monkeypatch.setenv("ODESAT_TILE_WS", ws)
        monkeypatch.setenv("ODESAT_TILE_TMA", "0")
        monkeypatch.setenv("ODESAT_TILE_NT", "512")
        monkeypatch.setenv("ODESAT_TILE_KSUB", ksub)
        b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
        b.upload(v, xs, xl)
        b.run_fixed(0.1, 0.001, 400, freeze=True)
        st, _ = b.status()
        out[name] = (st,) + b.download()
        b.close()
    nflag = int((out["ring"][0] >= 0).sum())
    assert R // 8 < nflag                                          # replicas do flag on this instance
    for name in ("ws", "ws2"):
        for a, c in zip(out[name], out["ring"]):
            assert eq(a, c), name


def test_steps_to_flag_distribution_of_the_benched_kernel(monkeypatch):
    """north_star: "the distribution of steps-to-solution over >= 256 seeds must be statistically indistinguishable from
    the reference's" — here for the kernel bench.py times (f32, BALANCED wide levels, warp-specialised persistent
    kernel with a work queue cut into sub-chunks) on a formula large enough to use it: 256 replicas of a random 3-SAT
    instance below the threshold, two-sample KS and Mann-Whitney against the f32 oracle's flag steps, and every flagged
    replica's thresholded state verified exactly."""
    from scipy import stats
    monkeypatch.setenv("ODESAT_TILE_WS", "1")
    monkeypatch.setenv("ODESAT_TILE_NT", "768")
    monkeypatch.setenv("ODESAT_TILE_KSUB", "7")
    f = cnf.random_ksat(4000, 3.0, seed=5)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R, steps = 256, 600
    v, xs, xl = F.init_batch(21, R, np.float32)
    b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
    b.upload(v, xs, xl)
    b.run_fixed(0.1, f.default_zeta(), steps, freeze=True)
    st, _ = b.status()
    gv, _, _ = b.download()
    ver = b.verify()
    b.close()
    ost = F.batch_fixed(v, xs, xl, 0.1, f.default_zeta(), steps, freeze=True, nthreads=O.host_cores())
    a = np.where(ost >= 0, ost, steps)
    c = np.where(st >= 0, st, steps)
    assert (ost >= 0).sum() > R // 2
    assert stats.ks_2samp(a, c).pvalue > 0.01 and stats.mannwhitneyu(a, c).pvalue > 0.01
    print(f"flag steps: oracle median {np.median(a)}, kernel median {np.median(c)}, identical for {(a == c).mean():.0%} of the replicas")
    for r in np.flatnonzero(st >= 0)[:16]:
        assert bool(ver[r]) == f.evaluate(gv[r] > 0)


@pytest.mark.parametrize("prec", [L.F32, L.F64])
@pytest.mark.parametrize("ksub", ["1", "3"])
def test_exact_kernel_on_the_work_queue_is_bit_identical(monkeypatch, prec, ksub):
    """The per-thread-ring kernel (EXACT schedule, the library default) as persistent CTAs taking (sub-chunk, tile) work
    items: every sub-chunk of a tile may run on another SM; states, flag steps and frozen replicas equal the oracle's."""
    monkeypatch.setenv("ODESAT_TILE_KSUB", ksub)
    monkeypatch.setenv("ODESAT_TILE_QUEUE", "1")
    f = cnf.random_ksat(4000, 3.0, seed=5)                        # flags between steps ~340 and ~450 at dt = 0.1
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    R = 203
    v, xs, xl = F.init_batch(11, R, dtype)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    b.upload(v, xs, xl)
    for n in (1, 330, 69):
        b.run_fixed(0.1, 0.001, n, freeze=True)
    st, _ = b.status()
    gv, gxs, gxl = b.download()
    b.close()
    ost = F.batch_fixed(v, xs, xl, 0.1, 0.001, 400, freeze=True, nthreads=O.host_cores())
    assert 0 < (ost >= 0).sum() < R or (ost >= 0).all()
    assert eq(st, ost) and eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)


def test_warp_specialised_kernel_soak_4096_replicas_1024_steps(monkeypatch):
    """Race evidence in lieu of compute-sanitizer (closed on this pool) for the kernel bench.py times: the full bench batch
    — 4 096 replicas, N = 10 000, the benched width and ring (704 threads, four stages) — for 1 024 steps (16 launches;
    producer warp, full / empty mbarriers, cross-proxy fences, work queue) must end BIT-IDENTICAL to the per-thread
    cp.async ring walking the same schedule; and a 512-replica shard cut into sub-chunks of 5 steps (every tile hops
    between SMs 13 times per launch) must equal the first 512 replicas of that run."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)
    D = S.DeviceFormula(f)
    monkeypatch.setenv("ODESAT_TILE_NT", "704")
    monkeypatch.setenv("ODESAT_TILE_TMA", "0")
    out = {}
    for name, ws, R in (("ws", "1", 4096), ("ring", "0", 4096), ("shard", "1", 512)):
        monkeypatch.setenv("ODESAT_TILE_WS", ws)
        b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
        b.init(1, 0)
        b.run_fixed(0.01, f.default_zeta(), 1024, freeze=False)
        out[name] = b.download()
        b.close()
    for a, c in zip(out["ws"], out["ring"]):
        assert eq(a, c)
    for a, c in zip(out["shard"], out["ws"]):
        assert eq(a, c[:512])
    assert (np.abs(out["ws"][0]) == 1).mean() > 0.02
