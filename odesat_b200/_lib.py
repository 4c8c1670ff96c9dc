"""ctypes binding of libodesat_b200.so (the C ABI of include/odesat_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a compute call
runs without a CUDA device, it fails loudly (OdesatError).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
SO_PATH = Path(os.environ.get("ODESAT_B200_SO", str(_HERE / "csrc" / "libodesat_b200.so")))   # override: tuning builds only

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, 1, 2, 3, 4
F64, F32 = 0, 1
ENGINE_AUTO, ENGINE_GATHER, ENGINE_TILE, ENGINE_SLAB = 0, 1, 2, 3
SCHED_EXACT, SCHED_BALANCED = 0, 1
MODE_BATCH, MODE_INTER = 0, 1
INT64_MAX = (1 << 63) - 1


class OdesatError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"odesat_b200 status {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """odesat_params (include/odesat_b200.h)."""
    _fields_ = [("tolerance", C.c_double), ("step_size", C.c_double), ("steps", C.c_int64),
                ("learning_rate", C.c_double), ("precision", C.c_int32), ("engine", C.c_int32),
                ("schedule", C.c_int32), ("chunk", C.c_int32), ("n_gpus", C.c_int32), ("sub_batches", C.c_int32)]


def make_params(tolerance=None, step_size=None, steps=None, learning_rate=None, precision=F64,
                engine=ENGINE_AUTO, schedule=SCHED_EXACT, chunk=0, n_gpus=1, sub_batches=0) -> Params:
    nan = float("nan")
    return Params(nan if tolerance is None else float(tolerance),
                  nan if step_size is None else float(step_size),
                  -1 if steps is None else int(steps),
                  nan if learning_rate is None else float(learning_rate),
                  int(precision), int(engine), int(schedule), int(chunk), int(n_gpus), int(sub_batches))


def build(force: bool = False) -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = list((_HERE / "csrc").glob("*.cu")) + list((_HERE / "csrc").glob("*.cuh")) + \
        list((_HERE / "csrc").glob("*.hpp")) + [(_HERE.parent / "include" / "odesat_b200.h")]
    stale = not SO_PATH.exists() or any(s.stat().st_mtime > SO_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if not Path(nvcc).exists():
            raise OdesatError(ECUDA, f"{SO_PATH} is missing/stale and nvcc was not found at {nvcc}")
        subprocess.run(["make", "-C", str(_HERE / "csrc"), "-s", f"NVCC={nvcc}"], check=True)
    return SO_PATH


_P, _D, _I64, _I32 = C.c_void_p, C.c_double, C.c_int64, C.c_int32
_PI = C.POINTER(C.c_int)
_PD = C.POINTER(C.c_double)
_PI64 = C.POINTER(C.c_int64)

# name -> argtypes (restype is int unless listed in _RESTYPE); doubles as the export list the
# CPU-side test checks against include/odesat_b200.h.
_SIGS = {
    "odesat_last_error": [],
    "odesat_abi_version": [],
    "odesat_device_count": [],
    "odesat_formula_create": [_I64, _I64, _P, _P, C.POINTER(_P)],
    "odesat_formula_destroy": [_P],
    "odesat_formula_info": [_P, _PI64, _PI64, _PI64, C.POINTER(_I32)],
    "odesat_formula_default_zeta": [_P, _PD],
    "odesat_init_short_term_memory": [_P, _P],
    "odesat_init_short_term_memory_f32": [_P, _P],
    "odesat_compute_derivatives": [_P, _P, _P, _P, _D, _P, _P, _P, _PI],
    "odesat_compute_derivatives_f32": [_P, _P, _P, _P, _D, _P, _P, _P, _PI],
    "odesat_update_state": [_P, _P, _P, _P, _P, _P, _P, _D],
    "odesat_update_state_f32": [_P, _P, _P, _P, _P, _P, _P, _D],
    "odesat_max_error": [_P, _P, _P, _P, _P, _P, _P, _PD],
    "odesat_max_error_f32": [_P, _P, _P, _P, _P, _P, _P, _PD],
    "odesat_euler_step_fixed": [_P, _P, _P, _P, _D, _D, _PI],
    "odesat_euler_step_fixed_f32": [_P, _P, _P, _P, _D, _D, _PI],
    "odesat_euler_step": [_P, _P, _P, _P, _D, _PD, _D, _PI],
    "odesat_euler_step_f32": [_P, _P, _P, _P, _D, _PD, _D, _PI],
    "odesat_simulate": [_P, _P, _P, _P, C.POINTER(Params), _P, _PI64, _PI, _PD],
    "odesat_simulate_f32": [_P, _P, _P, _P, C.POINTER(Params), _P, _PI64, _PI, _PD],
    "odesat_simulate_batch": [_P, _I64, _P, _P, _P, C.c_uint64, _I64, C.POINTER(Params), _I32, _I32,
                              _P, _P, _PI64, _P, _PI64],
    "odesat_simulate_batch_f32": [_P, _I64, _P, _P, _P, C.c_uint64, _I64, C.POINTER(Params), _I32,
                                  _I32, _P, _P, _PI64, _P, _PI64],
    "odesat_simulate_inter": [_P, _I64, _P, _P, _P, C.POINTER(Params), _P, _PI64, _PI64],
    "odesat_stoch_step": [_P, _P, _P, C.c_uint64, _I64, _I64, _PI],
    "odesat_stoch_search": [_P, _I64, _P, _P, C.c_uint64, _I64, _I64, _I32, _I32, _P, _P, _PI64, _P, _PI64],
    "odesat_tile_schedule_stats": [_I64, _I64, _P, _P, _I32, _I32, _I32, _P, _I64, _P, _I64, _P],
    "odesat_batch_create": [_P, _I64, _I32, _I32, _I32, C.POINTER(_P)],
    "odesat_batch_destroy": [_P],
    "odesat_batch_info": [_P, C.POINTER(_I32), _PI64, _PI64],
    "odesat_batch_init": [_P, C.c_uint64, _I64],
    "odesat_batch_upload": [_P, _P, _P, _P],
    "odesat_batch_download": [_P, _P, _P, _P],
    "odesat_batch_run_fixed": [_P, _D, _D, _I64, _I32, C.POINTER(C.c_float)],
    "odesat_batch_run_adaptive": [_P, _D, _D, _I64, C.POINTER(C.c_float)],
    "odesat_batch_stream": [_P, C.POINTER(_P)],
    "odesat_batch_run_fixed_async": [_P, _D, _D, _I64, _I32, _P],
    "odesat_batch_post_key": [_P, _I64, _P],
    "odesat_batch_sync": [_P],
    "odesat_batch_status": [_P, _P, _PI64],
    "odesat_batch_first_solved": [_P, _PI64],
    "odesat_batch_verify": [_P, _P],
    "odesat_batch_assignment": [_P, _I64, _P],
    "odesat_batch_dt": [_P, _P],
}
_RESTYPE = {"odesat_last_error": C.c_char_p, "odesat_formula_destroy": None, "odesat_batch_destroy": None}

_lib = None


def lib():
    """Load the shared library (never builds implicitly on a GPU box: the .so ships in-tree)."""
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise OdesatError(ECUDA, f"{SO_PATH} not found — run `python -c 'import __graft_entry__ as g; "
                                     "g.build()'` (there is no CPU fallback)")
        L = C.CDLL(str(SO_PATH))
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = L
    return _lib


def check(code: int):
    if code != OK:
        raise OdesatError(code, lib().odesat_last_error().decode("utf-8", "replace"))


def exported_symbols():
    return list(_SIGS)
