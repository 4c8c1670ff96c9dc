"""The step-loop driver of the C ABI (csrc/odesat_b200.cu::drive): lock-step `inter` with any chunk, speculative
chunks with the device-side stop key, sub-batches, multi-GPU shards inside one process, the device-side choice of the
literal first step, and the asynchronous entry points a multi-process host layer builds its collective on.
All against the CPU oracle or against the single-shard run, bit for bit."""
import ctypes as C

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


def devices():
    return L.lib().odesat_device_count()


def easy_instance():
    """Random 3-SAT far below the threshold: flags within a few hundred fixed steps, at different steps per replica."""
    f = cnf.random_ksat(300, 3.0, seed=21)
    return f, S.DeviceFormula(f), O.OracleFormula(f.varnum, f.clause_off, f.lits)


@pytest.mark.parametrize("chunk", [0, 7, 1])
@pytest.mark.parametrize("engine", [L.ENGINE_TILE, L.ENGINE_GATHER, L.ENGINE_SLAB])
def test_inter_write_back_is_lock_step_for_every_chunk(engine, chunk):
    """system.rs:279-293: when the loop stops, EVERY replica has taken exactly `steps` Euler steps — also when the
    device polls only every `chunk` steps (the overshoot of the non-winners)."""
    f, D, F = easy_instance()
    R = 24
    v, xs, xl = F.init_batch(5, R)
    ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
    oa, ow, ost = F.simulate_inter(ov, oxs, oxl, step_size=0.01, steps=5000)
    assert ow >= 0 and ost > 40                                   # the winning step lies inside a later chunk
    r = B.simulate_batch(D, R, v, xs, xl, step_size=0.01, steps=5000, precision=L.F64, engine=engine, chunk=chunk,
                         mode=L.MODE_INTER, write_back=True)
    assert (r.winner, r.steps_run) == (ow, ost) and eq(r.assignment, oa) and f.evaluate(r.assignment)
    assert eq(v, ov) and eq(xs, oxs) and eq(xl, oxl)              # all replicas, not only the winner
    # flags: exactly the replicas that flagged on the winning step
    v2, xs2, xl2 = F.init_batch(5, R)
    first = F.batch_fixed(v2, xs2, xl2, 0.01, f.default_zeta(), ost, freeze=False)
    assert eq(r.solved_step, np.where(first == ost - 1, first, -1))


def test_simulate_inter_entry_point_default_chunk_is_lock_step():
    f, D, F = easy_instance()
    R = 9
    v, xs, xl = F.init_batch(6, R)
    states = [S.State(v[r].copy(), xs[r].copy(), xl[r].copy()) for r in range(R)]
    info = []
    res = S.simulate_inter(states, D, None, 0.01, 5000, None, info=info)          # chunk = 0 → library default (32)
    oa, ow, ost = F.simulate_inter(v, xs, xl, step_size=0.01, steps=5000)
    assert info[0] == (ow, ost) and res == [bool(x) for x in oa]
    for r in range(R):
        assert eq(states[r].v, v[r]) and eq(states[r].xs, xs[r]) and eq(states[r].xl, xl[r])


@pytest.mark.parametrize("engine", [L.ENGINE_TILE, L.ENGINE_GATHER, L.ENGINE_SLAB])
@pytest.mark.parametrize("chunk", [3, 32])
def test_speculative_chunks_keep_winner_steps_and_flags(engine, chunk):
    """Without write-back the next chunk is already enqueued when the host reads a chunk's key; its kernels see the key
    on the device and do nothing.  Winner, step count, assignment and the flags up to the winning step are unchanged."""
    f, D, F = easy_instance()
    R = 40
    v, xs, xl = F.init_batch(9, R, np.float32)
    oa, ow, ost = F.simulate_inter(v.copy(), xs.copy(), xl.copy(), step_size=0.01, steps=5000)
    r = B.simulate_batch(D, R, v, xs, xl, step_size=0.01, steps=5000, precision=L.F32, engine=engine, chunk=chunk,
                         mode=L.MODE_INTER)
    assert (r.winner, r.steps_run) == (ow, ost) and eq(r.assignment, oa) and f.evaluate(r.assignment)
    assert r.verified[ow] == 1
    assert (r.solved_step[r.solved_step >= 0] == ost - 1).all() and r.solved_step[ow] == ost - 1
    # BATCH: every replica runs to its own flag, whatever the chunk
    v, xs, xl = F.init_batch(9, R, np.float32)
    first = F.batch_fixed(v.copy(), xs.copy(), xl.copy(), 0.01, f.default_zeta(), 1100, freeze=True)
    assert 0 < (first >= 0).sum() < R                              # some replicas flag and freeze, others run out of steps
    rb = B.simulate_batch(D, R, v, xs, xl, step_size=0.01, steps=1100, precision=L.F32, engine=engine, chunk=chunk,
                          mode=L.MODE_BATCH, write_back=True)
    assert eq(rb.solved_step, first) and rb.steps_run == 1100
    ov, oxs, oxl = F.init_batch(9, R, np.float32)
    F.batch_fixed(ov, oxs, oxl, 0.01, f.default_zeta(), 1100, freeze=True)
    assert eq(v, ov) and eq(xs, oxs) and eq(xl, oxl)


def test_sub_batches_give_identical_results():
    """odesat_params::sub_batches: the shards of one device have their own streams (upload / compute overlap) and are
    otherwise invisible."""
    f, D, F = easy_instance()
    R = 70
    v0 = np.ascontiguousarray(np.stack([F.init_v0(3, r, np.float32) for r in range(R)]))
    out = []
    for sub in (1, 4, 7):
        v = v0.copy()
        out.append(B.simulate_batch(D, R, v, None, None, step_size=0.01, steps=1200, precision=L.F32, mode=L.MODE_BATCH,
                                    sub_batches=sub))
    for r in out[1:]:
        assert eq(r.solved_step, out[0].solved_step) and eq(r.verified, out[0].verified)
        assert r.winner == out[0].winner and eq(r.assignment, out[0].assignment) and r.steps_run == out[0].steps_run
    ov, oxs, oxl = F.init_batch(3, R, np.float32)
    first = F.batch_fixed(ov, oxs, oxl, 0.01, f.default_zeta(), 1200, freeze=True)
    assert eq(out[0].solved_step, first) and (first >= 0).sum() > 3


@pytest.mark.parametrize("mode", [L.MODE_BATCH, L.MODE_INTER])
def test_multi_gpu_shards_inside_one_process(mode):
    """odesat_params::n_gpus: one handle, one process, G devices; results equal the single-device call."""
    G = devices()
    if G < 2:
        pytest.skip("needs at least 2 CUDA devices in this process")
    f, D, F = easy_instance()
    R = 50
    res = []
    for g in (1, 2, min(G, 4)):
        res.append(B.simulate_batch(D, R, seed=12, step_size=0.01, steps=3000, precision=L.F64, mode=mode, n_gpus=g, chunk=16))
    for r in res[1:]:
        assert r.winner == res[0].winner and r.steps_run == res[0].steps_run and eq(r.assignment, res[0].assignment)
        assert eq(r.solved_step, res[0].solved_step) and r.verified[r.winner] == 1
        if mode == L.MODE_BATCH:                      # (inter without write-back defines only the winner's final state)
            assert eq(r.verified, res[0].verified)
    v, xs, xl = F.init_batch(12, R)
    if mode == L.MODE_INTER:
        oa, ow, ost = F.simulate_inter(v, xs, xl, step_size=0.01, steps=3000)
        assert (res[0].winner, res[0].steps_run) == (ow, ost) and eq(res[0].assignment, oa)
        # lock-step write-back across devices
        v, xs, xl = F.init_batch(12, R)
        gv, gxs, gxl = v.copy(), xs.copy(), xl.copy()
        F.simulate_inter(v, xs, xl, step_size=0.01, steps=3000)
        r = B.simulate_batch(D, R, gv, gxs, gxl, step_size=0.01, steps=3000, precision=L.F64, mode=mode, n_gpus=2,
                             write_back=True)
        assert (r.winner, r.steps_run) == (ow, ost) and eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)
    else:
        first = F.batch_fixed(v, xs, xl, 0.01, f.default_zeta(), 3000, freeze=True)
        assert eq(res[0].solved_step, first)
    with pytest.raises(L.OdesatError):
        B.simulate_batch(D, R, seed=12, step_size=0.01, steps=10, precision=L.F64, mode=mode, n_gpus=G + 1)


@pytest.mark.parametrize("prec", [L.F64, L.F32])
@pytest.mark.parametrize("engine", [L.ENGINE_TILE, L.ENGINE_GATHER])
def test_uploaded_memories_outside_the_fast_domain_take_the_literal_first_step(engine, prec):
    """the short arithmetic is bit-identical to the reference only for finite, normal memories.  inf, NaN
    and denormal xs / xl on upload are detected ON THE DEVICE and the first step runs the literal statements."""
    f = cnf.random_ksat(150, 4.3, seed=8)
    D = S.DeviceFormula(f)
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    dtype = B.np_dtype(prec)
    rng = np.random.default_rng(2)
    R = 12
    v, xs, xl = random_state(rng, F.N, F.M, dtype, R=R)
    tiny = np.finfo(dtype).tiny
    xs[0, 3] = np.inf; xl[1, 5] = np.nan; xs[2, 7] = tiny / 8; xl[3, 11] = tiny / 2; xs[4, 0] = -np.inf
    xs[5, 9] = 1e-30 if prec == L.F32 else 1e-200
    b = B.ReplicaBatch(D, R, prec, engine, L.SCHED_EXACT)
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.25, 6, freeze=False)
    ov, oxs, oxl = v.copy(), xs.copy(), xl.copy()
    F.batch_fixed(ov, oxs, oxl, 0.01, 0.25, 6, freeze=False)
    gv, gxs, gxl = b.download()
    assert eq(gv, ov) and eq(gxs, oxs) and eq(gxl, oxl)
    # and a clean state right after it goes back to the fast path with the same result as the oracle
    v, xs, xl = random_state(rng, F.N, F.M, dtype, R=R)
    b.upload(v, xs, xl)
    b.run_fixed(0.01, 0.25, 6, freeze=False)
    F.batch_fixed(v, xs, xl, 0.01, 0.25, 6, freeze=False)
    gv, gxs, gxl = b.download()
    assert eq(gv, v) and eq(gxs, xs) and eq(gxl, xl)


def test_async_entry_points_and_device_stop_key():
    """What a multi-process host layer uses: enqueue a chunk, reduce the flags into DEVICE words, hand the key of the
    previous chunk to the next one.  A chunk that starts with a set key must leave the state untouched."""
    import torch
    f, D, F = easy_instance()
    R = 16
    b = B.ReplicaBatch(D, R, L.F32, L.ENGINE_AUTO, L.SCHED_EXACT)
    b.init(4, 0)
    keys = torch.full((2, 2), L.INT64_MAX, dtype=torch.int64, device="cuda")
    stream = torch.cuda.ExternalStream(b.stream)
    zeta = f.default_zeta()
    done, c = 0, 0
    while done < 5000:
        b.run_fixed_async(0.01, zeta, 16, True, keys[(c - 1) & 1].data_ptr() if c else 0)
        b.post_key(100, keys[c & 1].data_ptr())
        done += 16
        with torch.cuda.stream(stream):
            k = keys[c & 1].cpu()
        if int(k[0]) != L.INT64_MAX:
            break
        c += 1
    key, unflagged = int(k[0]), int(k[1])
    v, xs, xl = F.init_batch(4, R, np.float32)
    oa, ow, ost = F.simulate_inter(v, xs, xl, step_size=0.01, steps=5000)
    assert key >> 32 == ost - 1 and (key & 0xFFFFFFFF) == ow + 100 and 0 <= unflagged < R
    before = b.download()
    b.run_fixed_async(0.01, zeta, 16, True, keys[c & 1].data_ptr())           # key is set: a no-op
    b.sync()
    after = b.download()
    assert all(eq(x, y) for x, y in zip(before, after))
    b.run_fixed_async(0.01, zeta, 16, True, 0)                                 # no key: the unflagged replicas move
    b.sync()
    assert not eq(b.download()[0], before[0])
    b.close()
