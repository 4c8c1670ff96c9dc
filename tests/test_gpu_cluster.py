"""The thread-block-CLUSTER tile engine (tile_cluster.cuh): variables distributed over the shared
memories of 1 / 2 / 4 CTAs, gathered through distributed shared memory.  Forced on small formulas
with ODESAT_TILE_CLUSTER so the oracle finishes in seconds; selected automatically at N = 50 000."""
import os

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


@pytest.fixture
def force_cluster():
    saved = {k: os.environ.get(k) for k in ("ODESAT_TILE_CLUSTER", "ODESAT_TILE_NT", "ODESAT_TILE_D")}

    def setter(cl, nt=None, depth=None):
        os.environ["ODESAT_TILE_CLUSTER"] = str(cl)
        for key, val in (("ODESAT_TILE_NT", nt), ("ODESAT_TILE_D", depth)):
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = str(val)
    yield setter
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("prec", [L.F64, L.F32])
@pytest.mark.parametrize("cl,nt,depth", [(1, 1024, 2), (2, 1024, 3), (2, 512, 4), (4, 512, 2), (4, 1024, 4)])
def test_cluster_exact_schedule_is_bit_identical_to_the_oracle(force_cluster, prec, cl, nt, depth):
    force_cluster(cl, nt, depth)
    f = cnf.random_ksat(3000, 4.3, seed=31 + cl)              # levels wider than one item at nt = 512
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    R = 37
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    v, xs, xl = F.init_batch(4, R, dtype)
    b.upload(v, xs, xl)
    for n in (1, 70, 29):                                     # 100 steps, crossing the 64-step launch chunk
        b.run_fixed(0.01, 0.001, n, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, 0.001, 100, freeze=False, nthreads=8)
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)


@pytest.mark.parametrize("cl", [2, 4])
def test_cluster_strict_first_step_and_freeze(force_cluster, cl):
    """|v| > 1 on entry (literal rigidity term in the first step) and per-replica freezing after the flag."""
    force_cluster(cl)
    f = cnf.random_ksat(200, 4.3, seed=3)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    rng = np.random.default_rng(5)
    R = 16
    v, xs, xl = random_state(rng, F.N, F.M, np.float64, R=R)
    v[:, ::7] = rng.uniform(-3, 3, size=v[:, ::7].shape)
    v[:, 1] = 3.0; v[:, 2] = 2.0
    b = B.ReplicaBatch(D, R, L.F64, L.ENGINE_TILE, L.SCHED_EXACT)
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.5, 5, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, 0.5, 5, freeze=False)
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    # freeze: an easy satisfiable instance, every replica flags and stops one update past its flag
    sat = cnf.random_ksat(300, 2.0, seed=11)
    DS = S.DeviceFormula(sat)
    FS = O.OracleFormula(sat.varnum, sat.clause_off, sat.lits)
    R = 24
    v, xs, xl = FS.init_batch(2, R, np.float64)
    b = B.ReplicaBatch(DS, R, L.F64, L.ENGINE_TILE, L.SCHED_EXACT)
    b.upload(v, xs, xl)
    for n in (100, 300):
        b.run_fixed(0.01, 0.001, n, freeze=True)
    st, _ = b.status()
    ost = FS.batch_fixed(v, xs, xl, 0.01, 0.001, 400, freeze=True, nthreads=8)
    gv, gxs, gxl = b.download()
    assert np.array_equal(st, ost) and (st >= 0).any()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    assert b.verify()[st >= 0].all()


def test_cluster_balanced_agrees_to_rounding_and_is_deterministic(force_cluster):
    force_cluster(2)
    f = cnf.random_ksat(3000, 4.3, seed=8)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    rng = np.random.default_rng(2)
    R = 64
    v, xs, xl = random_state(rng, F.N, F.M, np.float32, R=R)
    b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_TILE, L.SCHED_BALANCED)
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.001, 1, freeze=False)
    g1 = b.download()
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.001, 1, freeze=False)
    g2 = b.download()
    F.batch_fixed(v, xs, xl, 0.01, 0.001, 1, freeze=False)
    np.testing.assert_allclose(g1[0], v, rtol=1e-5, atol=1e-6)      # north-star tolerance, f32
    assert eq(g1[1], xs) and eq(g1[2], xl)
    assert eq(g1[0], g2[0])


@pytest.mark.parametrize("prec", [L.F32, L.F64])
def test_config4_size_n50k_cluster_engine_vs_general_engine_vs_oracle(prec):
    """BASELINE.json configs[4] shape (N = 50 000, alpha = 4.25).  The cluster tile engine (2 CTAs per
    replica in f32, 4 in f64; explicit request — AUTO keeps the general engine at this size, which is
    faster) and the slab-scheduled general engine produce bit-identical states for all
    replicas; a sample equals the oracle."""
    for k in ("ODESAT_TILE_CLUSTER", "ODESAT_TILE_NT", "ODESAT_TILE_D"):
        os.environ.pop(k, None)
    f = cnf.random_ksat(50_000, 4.25, seed=20240615)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    R = 96
    a = B.ReplicaBatch(D, R, prec, L.ENGINE_AUTO, L.SCHED_EXACT)
    assert a.engine == L.ENGINE_GATHER
    a.close()
    t = B.ReplicaBatch(D, R, prec, L.ENGINE_TILE, L.SCHED_EXACT)
    assert t.engine == L.ENGINE_TILE
    t.init(1, 0)
    t.run_fixed(0.01, 0.001, 10, freeze=False)
    tv, txs, txl = t.download()
    t.close()
    g = B.ReplicaBatch(D, R, prec, L.ENGINE_GATHER)
    g.init(1, 0)
    g.run_fixed(0.01, 0.001, 10, freeze=False)
    gv, gxs, gxl = g.download()
    g.close()
    assert eq(tv, gv) and eq(txs, gxs) and eq(txl, gxl)
    for r in (0, R - 1):
        v = F.init_v0(1, r, dtype); xs = F.init_short_term_memory(dtype); xl = np.ones(F.M, dtype)
        for _ in range(10):
            F.euler_step_fixed(v, xs, xl, 0.01, 0.001)
        assert eq(tv[r], v) and eq(txs[r], xs) and eq(txl[r], xl)


def test_auto_takes_the_single_cta_cluster_engine_between_13k_and_27k_variables():
    """f32, N = 20 000: too many 16-byte rows for the two-replica tile kernel, but the 8-byte rows of
    the one-replica kernel fit in one CTA (no distributed-shared-memory traffic) — AUTO picks it."""
    for k in ("ODESAT_TILE_CLUSTER", "ODESAT_TILE_NT", "ODESAT_TILE_D"):
        os.environ.pop(k, None)
    f = cnf.random_ksat(20_000, 4.3, seed=5)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 64
    t = B.ReplicaBatch(D, R, L.F32, L.ENGINE_AUTO, L.SCHED_EXACT)
    assert t.engine == L.ENGINE_TILE
    v, xs, xl = F.init_batch(4, R, np.float32)
    t.upload(v, xs, xl)
    t.run_fixed(0.01, 0.001, 20, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, 0.001, 20, freeze=False, nthreads=8)
    gv, gxs, gxl = t.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
