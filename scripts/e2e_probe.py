#!/usr/bin/env python
"""Tuning aid: wall time of ONE odesat_simulate_batch call (host v0 in, flags / verification / assignment out) at the
bench configuration, for several sub-batch counts.  Prints one JSON line per setting."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from odesat_b200 import _lib as L
    from odesat_b200 import batch as B
    from odesat_b200 import cnf
    from odesat_b200.system import DeviceFormula
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    f = cnf.random_ksat(10_000, 4.3, seed=20240611 + 2)
    F = DeviceFormula(f)
    hv = torch.empty((R, f.varnum), dtype=torch.float32).pin_memory()
    hv.uniform_(-1.0, 1.0, generator=torch.Generator().manual_seed(1))
    for sub in (1, 2, 4, 8, 16):
        def call():
            return B.simulate_batch(F, R, hv.data_ptr(), None, None, step_size=0.01, steps=steps, precision=L.F32,
                                    schedule=L.SCHED_BALANCED, mode=L.MODE_BATCH, chunk=max(32, steps), sub_batches=sub)
        call(); call()
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"sub_batches": sub, "steps": steps, "replicas": R, "ms_best": min(ts) * 1e3, "ms_median": float(np.median(ts)) * 1e3,
                          "clause_evals_per_s": steps * f.n_clauses * R / min(ts)}), flush=True)


if __name__ == "__main__":
    main()
