"""Adaptive Euler steps (system.rs:111-139) in the TILE engine (csrc/tile_adaptive.cuh) against the oracle and against the
GATHER engine: with the EXACT schedule every replica's state, dt sequence and flag step are bit-identical; flagged
replicas stop untouched (system.rs:122); states outside the fast-arithmetic domain take the literal first step."""
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


def both(f):
    return S.DeviceFormula(f), O.OracleFormula(f.varnum, f.clause_off, f.lits)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
@pytest.mark.parametrize("R", [1, 2, 33, 130])
def test_tile_adaptive_exact_vs_oracle_ragged_replica_counts(prec, R, monkeypatch):
    monkeypatch.setenv("ODESAT_TILE_SMALL", "0")             # the block-wide kernels, not the one-warp-per-tile kernel
    f = cnf.random_ksat(500, 4.3, seed=12)
    D, F = both(f)
    dtype = B.np_dtype(prec)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    assert b.engine == L.ENGINE_TILE
    v, xs, xl = F.init_batch(4, R, dtype)
    b.upload(v, xs, xl)
    for n in (1, 70, 59):                                     # 130 steps: crosses the 64-step launch chunk
        b.run_adaptive(1e-3, 0.001, n)
    ost, odt = F.batch_adaptive(v, xs, xl, 1e-3, 0.001, 130, nthreads=4)
    gv, gxs, gxl = b.download()
    gst, _ = b.status()
    assert eq(gst, ost) and eq(b.dt(), odt)
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    assert len(np.unique(odt)) > 1 or R == 1                  # the controller really moved the step sizes apart


@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_tile_adaptive_flagged_replicas_stop_untouched(prec, monkeypatch):
    """An easy instance: the replicas flag at different steps (231 … 400); each must keep the state that satisfied the
    check (system.rs:122) and its dt while its tile mate goes on."""
    monkeypatch.setenv("ODESAT_TILE_SMALL", "0")
    f = cnf.random_ksat(300, 2.0, seed=5)
    D, F = both(f)
    dtype = B.np_dtype(prec)
    R = 41
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    g = B.ReplicaBatch(D, R, prec, L.ENGINE_GATHER)
    v, xs, xl = F.init_batch(9, R, dtype)
    b.upload(v, xs, xl)
    g.upload(v, xs, xl)
    for n in (200, 100, 50):
        b.run_adaptive(1e-3, f.default_zeta(), n)
        g.run_adaptive(1e-3, f.default_zeta(), n)
    ost, odt = F.batch_adaptive(v, xs, xl, 1e-3, f.default_zeta(), 350, nthreads=4)
    assert 3 <= (ost >= 0).sum() < R                          # some flagged on the way, some still run
    gst, _ = b.status()
    gv, gxs, gxl = b.download()
    assert eq(gst, ost) and eq(b.dt(), odt)
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    hst, _ = g.status()
    hv, hxs, hxl = g.download()
    assert eq(hst, ost) and eq(hv, v) and eq(hxs, xs) and eq(hxl, xl) and eq(g.dt(), odt)
    # exact verification of the flagged replicas' thresholded states (cnf.rs:246-264)
    ver = b.verify()
    assert all(ver[r] == 1 for r in range(R) if ost[r] >= 0)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_tile_adaptive_first_step_outside_the_fast_domain(prec, monkeypatch):
    """|v| > 1, raw memories and a large zeta: the first step runs the reference's statements literally (the device
    decides), the rest the fast forms — bit-identical to the oracle throughout."""
    monkeypatch.setenv("ODESAT_TILE_SMALL", "0")
    f = cnf.random_ksat(200, 4.3, seed=3)
    D, F = both(f)
    dtype = B.np_dtype(prec)
    rng = np.random.default_rng(5)
    R = 16
    v, xs, xl = random_state(rng, F.N, F.M, dtype, R=R)
    v[:, ::7] = (rng.uniform(-3, 3, size=v[:, ::7].shape)).astype(dtype)
    v[:, 1] = 3.0; v[:, 2] = 2.0
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    for zeta in (0.5, float("inf")):
        b.upload(v, xs, xl)
        b.run_adaptive(1e-3, zeta, 6)
        ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
        ost, odt = F.batch_adaptive(ov, oxs, oxl, 1e-3, zeta, 6)
        gv, gxs, gxl = b.download()
        assert eq(gv, ov) and eq(gxs, oxs) and eq(gxl, oxl) and eq(b.dt(), odt)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
def test_tile_adaptive_mid_size_auto_engine_and_balanced(prec):
    """N = 3 000 (block-wide kernels by default): AUTO takes the tile engine for an adaptive batch; EXACT is bit-identical
    to the oracle; BALANCED (dv summed in colour order) agrees within rounding over a few steps."""
    f = cnf.random_ksat(3000, 4.3, seed=21)
    D, F = both(f)
    dtype = B.np_dtype(prec)
    R = 24
    v, xs, xl = F.init_batch(2, R, dtype)
    ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
    ost, odt = F.batch_adaptive(ov, oxs, oxl, 1e-3, f.default_zeta(), 12, nthreads=4)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_AUTO, L.SCHED_EXACT)
    assert b.engine == L.ENGINE_TILE
    b.upload(v, xs, xl)
    b.run_adaptive(1e-3, f.default_zeta(), 12)
    gv, gxs, gxl = b.download()
    assert eq(gv, ov) and eq(gxs, oxs) and eq(gxl, oxl) and eq(b.dt(), odt)
    c = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_BALANCED)
    c.upload(v, xs, xl)
    c.run_adaptive(1e-3, f.default_zeta(), 12)
    cv, cxs, cxl = c.download()
    tol = 1e-9 if prec == L.F64 else 2e-4
    assert np.max(np.abs(cv - ov)) <= tol and np.max(np.abs(cxs - oxs)) <= tol
    assert np.max(np.abs(cxl / oxl - 1)) <= tol and np.max(np.abs(c.dt() / odt - 1)) <= (1e-6 if prec == L.F64 else 1e-2)


def test_tile_adaptive_headline_size_equals_gather_engine():
    """The formula of BASELINE configs[2] (N = 10 000, alpha = 4.3), 64 replicas, f32, 40 adaptive steps: the tile kernel
    (EXACT) and the gather engine — different kernels, different layouts — end in identical states and step sizes."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)
    D = S.DeviceFormula(f)
    R = 64
    t = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_EXACT)
    g = B.ReplicaBatch(D, R, L.F32, L.ENGINE_GATHER)
    t.init(1, 0)
    g.init(1, 0)
    t.run_adaptive(1e-3, f.default_zeta(), 40)
    g.run_adaptive(1e-3, f.default_zeta(), 40)
    tv, txs, txl = t.download()
    gv, gxs, gxl = g.download()
    assert eq(tv, gv) and eq(txs, gxs) and eq(txl, gxl) and eq(t.dt(), g.dt())


def test_simulate_batch_adaptive_takes_the_tile_engine_and_matches_the_oracle(monkeypatch):
    """`batch` without -s through the one-call C ABI (main.rs:278-308): per-replica flags and verification equal the
    oracle's, whichever engine AUTO picks (the tile engine for this size)."""
    monkeypatch.setenv("ODESAT_TILE_SMALL", "0")
    f = cnf.random_ksat(300, 2.0, seed=5)
    D, F = both(f)
    R, steps = 40, 380
    v, xs, xl = F.init_batch(9, R, np.float64)
    res = B.simulate_batch(D, R, v=v.copy(), steps=steps, precision=L.F64, mode=L.MODE_BATCH)
    ost, _ = F.batch_adaptive(v, xs, xl, 1e-3, f.default_zeta(), steps, nthreads=4)
    assert eq(res.solved_step, ost)
    exp = np.array([f.evaluate(v[r] > 0) for r in range(R)], np.uint8)
    assert eq(res.verified, exp)


def test_adaptive_batch_sharded_over_the_devices_of_one_process():
    """odesat_params::n_gpus with adaptive steps: every shard runs the tile engine's adaptive kernel on its own device;
    flags and verification equal the single-device call's."""
    if L.lib().odesat_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    f = cnf.random_ksat(5000, 2.0, seed=5)                    # M = 10 000: the block-wide kernels (not one warp per tile)
    D = S.DeviceFormula(f)
    R, steps = 48, 650
    one = B.simulate_batch(D, R, seed=4, steps=steps, precision=L.F32, mode=L.MODE_BATCH, n_gpus=1)
    two = B.simulate_batch(D, R, seed=4, steps=steps, precision=L.F32, mode=L.MODE_BATCH, n_gpus=2)
    assert eq(one.solved_step, two.solved_step) and eq(one.verified, two.verified) and one.winner == two.winner
    assert (one.solved_step >= 0).sum() >= 1


def test_adaptive_tile_kernel_soak_full_batch_equals_gather_engine():
    """Race evidence in lieu of the closed compute-sanitizer: the FULL bench batch (4 096 replicas of the configs[2] formula,
    every SM busy, C_m and v_full scratch written and re-read under load), 48 adaptive steps in launches of 1 + 31 + 16,
    EXACT schedule — bit-identical to the gather engine (which is pinned to the oracle above): states, step sizes, flags."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)
    D = S.DeviceFormula(f)
    R = 4096
    t = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_EXACT)
    g = B.ReplicaBatch(D, R, L.F32, L.ENGINE_GATHER)
    for q in (t, g):
        q.init(3, 0)
        for n in (1, 31, 16):
            q.run_adaptive(1e-3, f.default_zeta(), n)
    assert eq(t.dt(), g.dt()) and eq(t.status()[0], g.status()[0])
    tv, txs, txl = t.download()
    gv, gxs, gxl = g.download()
    assert eq(tv, gv) and eq(txs, gxs) and eq(txl, gxl)
    assert len(np.unique(t.dt())) > 100


@pytest.mark.parametrize("R", [4096, 512, 37])
def test_warp_specialised_adaptive_kernel_equals_the_per_thread_ring(R, monkeypatch):
    """k_tile_adaptive_ws (producer warp, bulk copies, C_m through the async proxy, work queue with sub-chunks when the
    shard has few tiles) against k_tile_adaptive on the SAME BALANCED schedule: identical states, step sizes and flags —
    race evidence for the ring / proxy-fence protocol at the full batch and on a shard whose tiles hop between SMs."""
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)
    D = S.DeviceFormula(f)
    out = []
    for ws in ("1", "0"):
        monkeypatch.setenv("ODESAT_TILE_ADWS", ws)
        b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
        b.init(5, 0)
        for n in (1, 17, 6):
            b.run_adaptive(1e-3, f.default_zeta(), n)
        out.append((b.download(), b.dt(), b.status()[0]))
        b.close()
    (s1, d1, f1), (s0, d0, f0) = out
    assert eq(d1, d0) and eq(f1, f0)
    assert all(eq(x, y) for x, y in zip(s1, s0))
