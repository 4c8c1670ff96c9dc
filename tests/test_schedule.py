"""Host-side schedule compiler of the TILE engine (no GPU needed): the level structure must make
the shared-memory read-modify-write race-free, and the EXACT schedule must preserve the
reference's per-variable summation order (system.rs:35-80)."""
import ctypes as C

import numpy as np
import pytest

from odesat_b200 import _lib as L
from odesat_b200 import cnf


def compile_schedule(f, sched, threads, depth):
    out = np.zeros(4, np.int64)
    perm = np.full(f.n_clauses * 8 + f.n_literals * 2 + 4096, -7, np.int32)   # group clauses take a slot per literal (rounded up to 4 / 8 / 16 / 32)
    items = np.zeros(f.n_clauses * 2 + 4096, np.uint32)
    P = C.c_void_p
    L.check(L.lib().odesat_tile_schedule_stats(
        f.varnum, f.n_clauses, f.clause_off.ctypes.data_as(P), f.lits.ctypes.data_as(P), sched, threads, depth,
        perm.ctypes.data_as(P), len(perm), items.ctypes.data_as(P), len(items), out.ctypes.data_as(P)))
    nlev, n_items, slots, wf = (int(x) for x in out)
    return nlev, items[:n_items], perm[:slots], wf / 1000.0


def levels_of(items, perm):
    """→ list of levels, each the list of clause indices it processes (holes dropped)."""
    levels, cur = [], []
    for it in items:
        base, cnt, last = int(it) & 0xFFFFF, (int(it) >> 20) & 0x7FF, bool(int(it) >> 31)
        cur.extend(int(m) for m in perm[base:base + cnt] if m >= 0)
        if last:
            levels.append(cur)
            cur = []
    assert not cur
    return levels


@pytest.mark.parametrize("sched", [L.SCHED_EXACT, L.SCHED_BALANCED])
@pytest.mark.parametrize("threads,depth", [(128, 4), (512, 5), (1024, 4), (512, 3)])
def test_schedule_invariants(sched, threads, depth):
    f = cnf.random_ksat(2000, 4.3, seed=5)
    var = (np.abs(f.lits) - 1).reshape(-1, 3)
    nlev, items, perm, wf = compile_schedule(f, sched, threads, depth)
    # every clause exactly once; padding only as holes
    assert sorted(perm[perm >= 0]) == list(range(f.n_clauses))
    # ring invariants of the kernel: item i lives in ring slot i % depth in every step, and a slot
    # is stored before it is prefetched again; the strict first-step kernel uses a ring of 2
    assert len(items) % depth == 0 and len(items) % 2 == 0 and len(items) > depth
    # items tile the slot array in order, never wider than the CTA
    pos = 0
    for it in items:
        base, cnt = int(it) & 0xFFFFF, (int(it) >> 20) & 0x7FF
        if cnt == 0:
            continue
        assert 0 <= base - pos < 8 and cnt <= threads and base % 8 == 0    # levels start 128-byte aligned
        assert (perm[pos:base] == -1).all() and (perm[base:base + cnt] >= 0).all()
        pos = base + cnt
    assert 0 <= len(perm) - pos < 8
    levels = levels_of(items, perm)
    assert len(levels) == nlev
    # race freedom: no variable twice inside a level
    level_of = np.empty(f.n_clauses, np.int64)
    for li, lv in enumerate(levels):
        vs = var[lv].reshape(-1)
        assert len(np.unique(vs)) == len(vs)
        level_of[lv] = li
    if sched == L.SCHED_EXACT:
        # order preservation: each variable meets its clauses in ascending clause index
        for i in range(f.varnum):
            occ = np.flatnonzero((var == i).any(axis=1))
            assert (np.diff(level_of[occ]) > 0).all()
    else:
        sizes = np.array([len(lv) for lv in levels])
        assert sizes.max() - np.median(sizes) <= 64            # balanced classes
    assert 1.0 <= wf < 1.6                                       # bank-conflict packing quality


def test_schedule_is_deterministic_and_rejects_unsupported():
    f = cnf.random_ksat(500, 4.3, seed=1)
    a = compile_schedule(f, L.SCHED_BALANCED, 512, 4)
    b = compile_schedule(f, L.SCHED_BALANCED, 512, 4)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    empty = cnf.Formula(3, np.array([0], np.int64), np.array([], np.int32), {})
    with pytest.raises(L.OdesatError):
        compile_schedule(empty, L.SCHED_EXACT, 128, 2)


@pytest.mark.parametrize("sched", [L.SCHED_EXACT, L.SCHED_BALANCED])
@pytest.mark.parametrize("which", ["ragged", "k4", "repeat"])
def test_ragged_schedules_keep_the_level_invariants(sched, which):
    """Clause lengths other than 3, unit / empty clauses and variables repeated inside a clause compile into LOOP
    clauses (tile_ragged.cuh): still every clause exactly once, no variable shared by two CLAUSES of a level, and — EXACT —
    every variable meets its clauses in ascending clause index."""
    from helpers import ragged_formula, repeated_var_formula
    f = {"ragged": lambda: ragged_formula(3, 300, 1500), "k4": lambda: cnf.random_ksat(400, 5.0, seed=1, k=4),
         "repeat": lambda: repeated_var_formula(2, 200, 900)}[which]()
    nlev, items, perm, wf = compile_schedule(f, sched, 512, 4)
    assert sorted(perm[perm >= 0]) == list(range(f.n_clauses))
    assert len(items) % 4 == 0 and len(items) > 4
    levels = levels_of(items, perm)
    assert len(levels) == nlev
    vars_of = [set(int(abs(l)) - 1 for l in f.lits[f.clause_off[m]:f.clause_off[m + 1]]) for m in range(f.n_clauses)]
    level_of = np.empty(f.n_clauses, np.int64)
    for li, lv in enumerate(levels):
        seen = set()
        for m in lv:
            assert not (seen & vars_of[m])
            seen |= vars_of[m]
        level_of[lv] = li
    if sched == L.SCHED_EXACT:
        last = {}
        for m in range(f.n_clauses):
            for v in vars_of[m]:
                assert last.get(v, -1) < level_of[m]
                last[v] = level_of[m]


def test_aim_fixture_schedule(golden_dir):
    f = cnf.load_dimacs(str(golden_dir / "aim100_sat.cnf"))
    nlev, items, perm, wf = compile_schedule(f, L.SCHED_EXACT, 128, 6)
    assert nlev >= 8 and sorted(perm[perm >= 0]) == list(range(160))


@pytest.mark.parametrize("threads", [128, 512, 1024])
def test_exact_levels_are_list_scheduled_to_the_cta_width(threads):
    """EXACT levels are capped at the CTA width (one item per level) without lengthening the schedule
    beyond the critical path of the per-variable clause chains, and stay order-preserving."""
    f = cnf.random_ksat(4000, 4.3, seed=11)
    nlev, items, perm, _ = compile_schedule(f, L.SCHED_EXACT, threads, 2)
    levels = levels_of(items, perm)
    assert max(len(lv) for lv in levels) <= threads
    assert sum(1 for it in items if ((int(it) >> 20) & 0x7FF) > 0) == len(levels)      # one item per level
    # critical path: longest chain of clauses linked through a shared variable, in clause order
    var = (np.abs(f.lits) - 1).reshape(-1, 3)
    depth = np.zeros(f.n_clauses, int)
    last = {}
    for m in range(f.n_clauses):
        d = 0
        for v in var[m]:
            if v in last:
                d = max(d, depth[last[v]] + 1)
            last[v] = m
        depth[m] = d
    critical = int(depth.max()) + 1
    assert nlev >= critical
    if threads >= 512:
        assert nlev == critical                      # enough lanes: the cap costs no extra level
    # order preservation: every variable meets its clauses in ascending clause index, one level apart at least
    level_of = {m: k for k, lv in enumerate(levels) for m in lv}
    last_level = {}
    for m in range(f.n_clauses):
        for v in var[m]:
            assert last_level.get(v, -1) < level_of[m]
            last_level[v] = level_of[m]


@pytest.mark.parametrize("ipl", ["1", "0"])
def test_wide_balanced_levels_at_the_bench_size(monkeypatch, ipl):
    """BALANCED levels of the benched formula for 768-thread CTAs: one item per level (56 levels, the round-1 form) or
    WIDE levels — as many colours as the maximum variable degree (28), two full 768-clause items each, so half the
    level barriers for the same 56 items.  Either way no variable occurs twice inside a level."""
    monkeypatch.setenv("ODESAT_TILE_IPL", ipl)
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)
    var = (np.abs(f.lits) - 1).reshape(-1, 3)
    nlev, items, perm, wf = compile_schedule(f, L.SCHED_BALANCED, 768, 3)
    levels = levels_of(items, perm)
    assert sorted(perm[perm >= 0]) == list(range(f.n_clauses))
    for lv in levels:
        vs = var[lv].reshape(-1)
        assert len(np.unique(vs)) == len(vs)
    real = sum(1 for it in items if ((int(it) >> 20) & 0x7FF) > 0)
    max_degree = int(np.bincount(var.reshape(-1)).max())
    if ipl == "1":
        assert nlev == real == 56
    else:
        assert nlev == max_degree == 28 and real == 56
    assert len(items) % 6 == 0 and 1.0 <= wf < 1.6
