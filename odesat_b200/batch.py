"""Replica batches on the GPU: the `batch` (main.rs:254-323) and `inter` (system.rs:241-359)
workloads, single-GPU objects plus the one-process-per-GPU sharding layer.

Replicas are independent, so a job of R replicas shards as contiguous ranges
``[rank·R/W, (rank+1)·R/W)`` with NO data-path collective.  The only exchange is the early
exit: after every chunk of Euler steps each rank contributes one int64 key
``(first_flag_step << 32) | global_replica`` (INT64_MAX = nothing flagged) to a MIN
all-reduce — NCCL over NVLink on GPUs, gloo in the CPU tests.  The minimum names the
earliest flagged step and, within it, the lowest replica index, which is exactly
`state_res.iter().position(|&x| x)` at system.rs:353.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from .system import DeviceFormula, _ptr

NO_KEY = L.INT64_MAX


def np_dtype(precision: int):
    return np.float32 if precision == L.F32 else np.float64


class ReplicaBatch:
    """R replicas of one formula, state resident in HBM (odesat_batch of the C ABI)."""

    def __init__(self, formula: DeviceFormula, replicas: int, precision: int = L.F32,
                 engine: int = L.ENGINE_AUTO, schedule: int = L.SCHED_EXACT):
        self.formula = formula
        self.R = int(replicas)
        self.precision = precision
        self.dtype = np_dtype(precision)
        h = C.c_void_p()
        L.check(L.lib().odesat_batch_create(formula.handle, self.R, precision, engine, schedule, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            L.lib().odesat_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- info ------------------------------------------------------------------------------
    def info(self) -> Tuple[int, int, int]:
        """(engine, kernel launches so far, device bytes held)."""
        e, n, b = C.c_int32(), C.c_int64(), C.c_int64()
        L.check(L.lib().odesat_batch_info(self._h, C.byref(e), C.byref(n), C.byref(b)))
        return e.value, n.value, b.value

    @property
    def engine(self) -> int:
        return self.info()[0]

    @property
    def launches(self) -> int:
        return self.info()[1]

    # -- state in / out ----------------------------------------------------------------------
    def init(self, seed: int, replica_offset: int = 0) -> None:
        L.check(L.lib().odesat_batch_init(self._h, seed, replica_offset))

    def upload(self, v, xs, xl) -> None:
        """Host ``[R][N]``, ``[R][M]``, ``[R][M]`` (numpy arrays or raw pointers of the batch dtype)."""
        L.check(L.lib().odesat_batch_upload(self._h, self._as_ptr(v, self.formula.varnum),
                                            self._as_ptr(xs, self.formula.n_clauses),
                                            self._as_ptr(xl, self.formula.n_clauses)))

    def download(self):
        N, M = self.formula.varnum, self.formula.n_clauses
        v = np.empty((self.R, N), self.dtype)
        xs = np.empty((self.R, M), self.dtype)
        xl = np.empty((self.R, M), self.dtype)
        L.check(L.lib().odesat_batch_download(self._h, _ptr(v), _ptr(xs), _ptr(xl)))
        return v, xs, xl

    def _as_ptr(self, a, width):
        if isinstance(a, np.ndarray):
            if a.dtype != self.dtype or a.shape != (self.R, width):
                raise ValueError(f"expected {self.dtype} array of shape {(self.R, width)}, got {a.dtype} {a.shape}")
            return _ptr(a)
        return C.c_void_p(int(a))   # raw host pointer (e.g. torch pinned tensor .data_ptr())

    # -- stepping ------------------------------------------------------------------------------
    def run_fixed(self, dt: float, zeta: float, n: int, freeze: bool = True, timed: bool = False) -> Optional[float]:
        ms = C.c_float()
        L.check(L.lib().odesat_batch_run_fixed(self._h, dt, zeta, n, int(freeze), C.byref(ms) if timed else None))
        return ms.value if timed else None

    def run_adaptive(self, tolerance: float, zeta: float, n: int, timed: bool = False) -> Optional[float]:
        ms = C.c_float()
        L.check(L.lib().odesat_batch_run_adaptive(self._h, tolerance, zeta, n, C.byref(ms) if timed else None))
        return ms.value if timed else None

    # -- asynchronous pieces (multi-process host layer) --------------------------------------------
    @property
    def stream(self) -> int:
        """cudaStream_t of the batch as an integer (wrap with torch.cuda.ExternalStream)."""
        s = C.c_void_p()
        L.check(L.lib().odesat_batch_stream(self._h, C.byref(s)))
        return int(s.value or 0)

    def run_fixed_async(self, dt: float, zeta: float, n: int, freeze: bool = True, stop_key_ptr: int = 0) -> None:
        """Enqueue n fixed steps; `stop_key_ptr` is a DEVICE pointer to one int64 key (0: none)."""
        L.check(L.lib().odesat_batch_run_fixed_async(self._h, dt, zeta, n, int(freeze),
                                                     C.c_void_p(stop_key_ptr) if stop_key_ptr else None))

    def post_key(self, replica_offset: int, key_out_ptr: int) -> None:
        """Enqueue the flag reduction into two DEVICE int64 words {min key, unflagged replicas}."""
        L.check(L.lib().odesat_batch_post_key(self._h, replica_offset, C.c_void_p(key_out_ptr)))

    def sync(self) -> None:
        L.check(L.lib().odesat_batch_sync(self._h))

    # -- results ---------------------------------------------------------------------------------
    def status(self) -> Tuple[np.ndarray, int]:
        s = np.empty(self.R, np.int64)
        n = C.c_int64()
        L.check(L.lib().odesat_batch_status(self._h, _ptr(s), C.byref(n)))
        return s, n.value

    def first_solved(self) -> int:
        k = C.c_int64()
        L.check(L.lib().odesat_batch_first_solved(self._h, C.byref(k)))
        return k.value

    def verify(self) -> np.ndarray:
        out = np.zeros(self.R, np.uint8)
        L.check(L.lib().odesat_batch_verify(self._h, _ptr(out)))
        return out

    def assignment(self, replica: int) -> np.ndarray:
        out = np.empty(self.formula.varnum, np.uint8)
        L.check(L.lib().odesat_batch_assignment(self._h, replica, _ptr(out)))
        return out

    def dt(self) -> np.ndarray:
        out = np.empty(self.R, np.float64)
        L.check(L.lib().odesat_batch_dt(self._h, _ptr(out)))
        return out


@dataclass
class BatchResult:
    solved_step: np.ndarray
    verified: np.ndarray
    winner: int
    assignment: np.ndarray
    steps_run: int


def simulate_batch(formula: DeviceFormula, R: int, v=None, xs=None, xl=None, *, seed: int = 1,
                   replica_offset: int = 0, tolerance=None, step_size=None, steps=None, learning_rate=None,
                   precision: int = L.F32, engine: int = L.ENGINE_AUTO, schedule: int = L.SCHED_EXACT,
                   chunk: int = 0, mode: int = L.MODE_BATCH, write_back: bool = False, n_gpus: int = 1,
                   sub_batches: int = 0) -> BatchResult:
    """One call of `odesat_simulate_batch[_f32]`: HOST buffers in, results out (the e2e path).
    v/xs/xl: numpy arrays of shape [R][N]/[R][M]/[R][M] or raw host pointers (ints), or None to
    generate on the device.  Host element type: float32 when precision is F32, else float64.
    n_gpus: devices of THIS process the library shards the replicas over (odesat_params::n_gpus)."""
    p = L.make_params(tolerance, step_size, steps, learning_rate, precision, engine, schedule, chunk, n_gpus, sub_batches)
    fn = L.lib().odesat_simulate_batch_f32 if precision == L.F32 else L.lib().odesat_simulate_batch

    def ptr(a):
        if a is None:
            return None
        return _ptr(a) if isinstance(a, np.ndarray) else C.c_void_p(int(a))

    solved = np.full(R, -1, np.int64)
    ver = np.zeros(R, np.uint8)
    assign = np.zeros(formula.varnum, np.uint8)
    win, run = C.c_int64(-1), C.c_int64(0)
    L.check(fn(formula.handle, R, ptr(v), ptr(xs), ptr(xl), seed, replica_offset, C.byref(p), mode, int(write_back),
               _ptr(solved), _ptr(ver), C.byref(win), _ptr(assign), C.byref(run)))
    return BatchResult(solved, ver, win.value, assign, run.value)


# ---------------------------------------------------------------------------------------------
# one-process-per-GPU sharding
# ---------------------------------------------------------------------------------------------

def shard_range(R: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous replica range owned by `rank`."""
    return R * rank // world, R * (rank + 1) // world


def encode_key(step: int, replica: int) -> int:
    return (int(step) << 32) | int(replica)


def decode_key(key: int) -> Tuple[int, int]:
    return int(key) >> 32, int(key) & 0xFFFFFFFF


def globalize_key(local_key: int, replica_offset: int) -> int:
    """Local (step, replica) key → global replica numbering; NO_KEY stays NO_KEY."""
    if local_key == NO_KEY:
        return NO_KEY
    s, r = decode_key(local_key)
    return encode_key(s, r + replica_offset)


def allreduce_min_key(key: int, device=None) -> int:
    """The early-exit collective: MIN all-reduce of one int64 (8 bytes) over the default
    process group (NCCL on GPUs, gloo on CPU).  Without an initialised group it is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(key)
    t = torch.tensor([int(key)], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


@dataclass
class ShardedResult:
    key: int                 # global early-exit key (NO_KEY: nothing flagged)
    steps_run: int           # Euler steps every rank executed
    winner: int              # global replica index or -1
    winner_rank: int         # rank that owns the winner or -1


def run_sharded_inter(batch, dt: float, zeta: float, max_steps: int, chunk: int, replica_offset: int,
                      shard_sizes, device=None, reduce_fn=None) -> ShardedResult:
    """`inter` across ranks: every rank steps its shard `chunk` Euler steps at a time, then the
    8-byte MIN all-reduce decides whether anybody flagged.  `batch` needs `.run_fixed(dt, zeta, n,
    freeze)` and `.first_solved()` (ReplicaBatch, or a stub in the CPU tests).  max_steps < 0 runs
    until some replica flags (system.rs:296-311).  `reduce_fn` replaces the collective (identity for a
    single-process reference run inside an initialised group)."""
    done = 0
    key = NO_KEY
    if max_steps == 0:
        # system.rs:274, 353 (quirk Q8): `state_res` starts all-true, so with zero steps replica 0 is returned
        return ShardedResult(NO_KEY, 0, 0, 0)
    while max_steps < 0 or done < max_steps:
        n = chunk if max_steps < 0 else min(chunk, max_steps - done)
        batch.run_fixed(dt, zeta, n, True)
        done += n
        local_key = globalize_key(batch.first_solved(), replica_offset)
        key = reduce_fn(local_key) if reduce_fn is not None else allreduce_min_key(local_key, device)
        if key != NO_KEY:
            break
    winner, wrank = -1, -1
    if key != NO_KEY:
        _, winner = decode_key(key)
        lo = 0
        for rk, sz in enumerate(shard_sizes):
            if lo <= winner < lo + sz:
                wrank = rk
                break
            lo += sz
    return ShardedResult(key, done, winner, wrank)
