"""The SLAB engine (csrc/slab_engine.cuh: one persistent kernel per chunk of steps over 32-byte replica slabs, the
gather engine's deterministic two-phase RHS with the contributions consumed out of L2) against the oracle and the
gather engine: bit for bit in both precisions, ragged replica counts, chunk boundaries, flags and freezing, states
outside the fast-arithmetic domain (literal first step), BASELINE configs[4]'s formula size, engine selection."""
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


@pytest.mark.parametrize("prec", [L.F32, L.F64])
@pytest.mark.parametrize("R", [1, 7, 100, 515])
def test_slab_engine_is_bit_identical_to_the_oracle(prec, R):
    f = cnf.random_ksat(3000, 4.3, seed=14)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_SLAB)
    assert b.engine == L.ENGINE_SLAB
    v, xs, xl = F.init_batch(4, R, dtype)
    b.upload(v, xs, xl)
    for n in (1, 40, 29):                                     # 70 steps: crosses the 32-step launch chunk twice
        b.run_fixed(0.01, 0.001, n, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, 0.001, 70, freeze=False, nthreads=O.host_cores())
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)


@pytest.mark.parametrize("cta", ["0", "1"])
def test_slab_engine_flags_and_freezes_like_the_gather_engine(monkeypatch, cta):
    """Replicas flag at different steps, freeze (system.rs:149-153, then dt = 0) and keep their flag step; units handed
    out per warp (default) and per CTA."""
    monkeypatch.setenv("ODESAT_SLAB_CTA", cta)
    f = cnf.random_ksat(4000, 3.0, seed=5)                        # flags between steps ~340 and ~450 at dt = 0.1
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = 301
    v, xs, xl = F.init_batch(11, R, np.float32)
    out = {}
    for name, eng in (("gather", L.ENGINE_GATHER), ("slab", L.ENGINE_SLAB)):
        b = B.ReplicaBatch(D, R, L.F32, eng)
        b.upload(v, xs, xl)
        b.run_fixed(0.1, 0.001, 400, freeze=True)
        st, _ = b.status()
        out[name] = (st,) + b.download()
        ver = b.verify()
        b.close()
        for r in np.flatnonzero(st >= 0)[:5]:
            assert bool(ver[r]) == f.evaluate(out[name][1][r] > 0)
    nflag = int((out["gather"][0] >= 0).sum())
    assert R // 8 < nflag
    for a, c in zip(out["slab"], out["gather"]):
        assert eq(a, c)


@pytest.mark.parametrize("prec", [L.F32, L.F64])
def test_slab_literal_first_step_for_states_outside_the_fast_domain(prec):
    """|v| > 1 on entry makes system.rs:73's branch reachable with r != 0, an inf memory makes inf·0 = NaN: the first
    step after such an import runs the reference's statements literally (decided on the device); a non-finite zeta
    keeps every step literal."""
    f = cnf.random_ksat(500, 4.3, seed=3)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    rng = np.random.default_rng(5)
    R = 19
    v, xs, xl = random_state(rng, F.N, F.M, dtype, R=R)
    v[:, ::7] = (rng.uniform(-3, 3, size=v[:, ::7].shape)).astype(dtype)
    v[:, 1] = 3.0; v[:, 2] = 2.0
    xl[3, 5] = np.inf
    b = B.ReplicaBatch(D, R, prec, L.ENGINE_SLAB)
    for zeta in (0.5, float("inf")):
        b.upload(v, xs, xl)
        b.run_fixed(0.01, zeta, 5, freeze=False)
        ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
        F.batch_fixed(ov, oxs, oxl, 0.01, zeta, 5, freeze=False)
        got = b.download()
        assert eq(got[0], ov) and eq(got[1], oxs) and eq(got[2], oxl)


def test_slab_engine_at_config4_size_and_engine_selection(monkeypatch):
    """BASELINE configs[4]'s formula (N = 50 000, alpha = 4.25): AUTO keeps the (faster) gather engine unless ODESAT_SLAB=1;
    12 steps on the slab engine equal the gather engine's for ALL replicas and the oracle's for a sample."""
    f = cnf.random_ksat(50_000, 4.25, seed=20240615)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    small = B.ReplicaBatch(D, 96, L.F32, L.ENGINE_AUTO)
    assert small.engine == L.ENGINE_GATHER
    small.close()
    R = 300
    a = B.ReplicaBatch(D, R, L.F32, L.ENGINE_AUTO)
    assert a.engine == L.ENGINE_GATHER
    a.close()
    monkeypatch.setenv("ODESAT_SLAB", "1")
    t = B.ReplicaBatch(D, R, L.F32, L.ENGINE_AUTO)
    assert t.engine == L.ENGINE_SLAB
    t.init(1, 0)
    t.run_fixed(0.01, 0.001, 12, freeze=False)
    tv, txs, txl = t.download()
    t.close()
    g = B.ReplicaBatch(D, R, L.F32, L.ENGINE_GATHER)
    g.init(1, 0)
    g.run_fixed(0.01, 0.001, 12, freeze=False)
    gv, gxs, gxl = g.download()
    g.close()
    assert eq(tv, gv) and eq(txs, gxs) and eq(txl, gxl)
    for r in (0, R - 1):
        v = F.init_v0(1, r, np.float32); xs = F.init_short_term_memory(np.float32); xl = np.ones(F.M, np.float32)
        for _ in range(12):
            F.euler_step_fixed(v, xs, xl, 0.01, 0.001)
        assert eq(tv[r], v) and eq(txs[r], xs) and eq(txl[r], xl)


def test_slab_engine_rejects_what_it_cannot_run():
    g = cnf.random_ksat(100, 5.0, seed=1, k=4)
    with pytest.raises(L.OdesatError) as e:
        B.ReplicaBatch(S.DeviceFormula(g), 8, L.F32, L.ENGINE_SLAB)
    assert e.value.code == L.EUNSUPPORTED
