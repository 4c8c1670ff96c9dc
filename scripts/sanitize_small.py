"""Small workload for compute-sanitizer (racecheck / memcheck): both engines, both schedules,
f32 and f64, ragged replica count, a few fixed and adaptive steps.  Run on a GPU box:
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from odesat_b200 import _lib as L, batch as B, cnf  # noqa: E402
from odesat_b200.system import DeviceFormula  # noqa: E402

f = cnf.random_ksat(300, 4.3, seed=1)
F = DeviceFormula(f)
for prec in (L.F32, L.F64):
    for engine, sched in ((L.ENGINE_TILE, L.SCHED_EXACT), (L.ENGINE_TILE, L.SCHED_BALANCED), (L.ENGINE_GATHER, L.SCHED_EXACT)):
        b = B.ReplicaBatch(F, 37, prec, engine, sched)
        b.init(3, 0)
        b.run_fixed(0.01, 0.001, 6, freeze=True)
        if engine == L.ENGINE_GATHER:
            b.run_adaptive(1e-3, 0.001, 3)
        v, xs, xl = b.download()
        assert np.isfinite(v).all()
        b.verify()
        b.close()
print("sanitize_small: ok")
