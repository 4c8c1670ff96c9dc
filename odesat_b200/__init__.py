"""odesat_b200 — B200-native digital-memcomputing ODE integrator (the hot path of AHartNtkn/odesat).

Layout: `csrc/` hand-written sm_100a CUDA kernels + the C ABI (include/odesat_b200.h);
`system.py` mirrors the reference's `odesat::system` functions over that ABI; `batch.py` holds the
device-resident replica batches and the one-process-per-GPU sharding layer; `cnf.py` is the
formula plumbing (DIMACS reader, normaliser, random k-SAT generator).
"""
from . import _lib  # noqa: F401
from ._lib import (ENGINE_AUTO, ENGINE_GATHER, ENGINE_SLAB, ENGINE_TILE, F32, F64, MODE_BATCH, MODE_INTER, SCHED_BALANCED,  # noqa: F401
                   SCHED_EXACT, OdesatError)

__all__ = ["_lib", "cnf", "system", "batch", "OdesatError"]
