#!/usr/bin/env python
"""All five BASELINE.json configs on ONE B200 (multi-GPU configs: the per-GPU share).  Prints one JSON
line per config; the committed copy is profiles/rNN_configs.jsonl.  (bench.py is the contract benchmark
and the only place the CPU oracle is timed — `bench.py --workload hard|rand10k|… ` prints `cpu_baseline`;
this script runs product code only.)

    python scripts/run_configs.py [--quick]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bench import algorithmic_bytes_per_step, measured_peak   # noqa: E402
from odesat_b200 import _lib as L                             # noqa: E402
from odesat_b200 import batch as B                            # noqa: E402
from odesat_b200 import cnf, commands                         # noqa: E402
from odesat_b200.system import DeviceFormula                  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    peak, _ = measured_peak()

    # ---- configs[0]: tests/easy.cnf via `solve` (adaptive, -r 7) ---------------------------------
    times, steps, ok, tpre, tint = [], [], [], [], []
    for seed in range(8):
        t0 = time.perf_counter()
        r = commands.solve(str(GOLD / "aim100_sat.cnf"), output="/dev/null", step_number=1_000_000, seed=seed, log=lambda s: None)
        times.append(time.perf_counter() - t0)
        steps.append(r.steps)
        ok.append(r.is_satisfiable)
        tpre.append(r.seconds_preprocess)
        tint.append(r.seconds_integrate)
    emit(config=0, what="easy.cnf `solve` adaptive -r 7 (8 seeds): preprocessing + GPU integration + trace replay, end to end",
         all_verified_sat=all(ok), steps_median=float(np.median(steps)), seconds_median_end_to_end=float(np.median(times)),
         seconds_median_host_preprocessing_python=float(np.median(tpre)), seconds_median_upload_plus_gpu_integration=float(np.median(tint)))

    # ---- configs[1]: tests/hard.cnf `batch -b 100 -n 1000 -s 0.01` --------------------------------
    fh = cnf.load_dimacs(str(GOLD / "aim100_unsat.cnf"))
    FH = DeviceFormula(fh)
    for prec, name in ((L.F64, "f64"), (L.F32, "f32")):
        B.simulate_batch(FH, 100, seed=1, step_size=0.01, steps=1000, precision=prec)     # warm
        t0 = time.perf_counter()
        r = B.simulate_batch(FH, 100, seed=1, step_size=0.01, steps=1000, precision=prec)
        sec = time.perf_counter() - t0
        emit(config=1, what=f"hard.cnf batch -b 100 -n 1000 -s 0.01, {name}, one odesat_simulate_batch call (host in/out)",
             seconds=sec, clause_evals_per_s=1000 * 160 * 100 / sec, any_verified=bool(r.verified.any()), steps_run=int(r.steps_run))

    # ---- configs[2..4]: throughput of the step loop, device-resident state -------------------------
    def throughput(cfg, what, f, R, prec, adaptive, steps, warm, sched=L.SCHED_BALANCED):
        F = DeviceFormula(f)
        b = B.ReplicaBatch(F, R, prec, L.ENGINE_AUTO, sched)
        b.init(1, 0)
        zeta = f.default_zeta()
        run = (lambda n, t=False: b.run_adaptive(1e-3, zeta, n, timed=t)) if adaptive else \
              (lambda n, t=False: b.run_fixed(0.01, zeta, n, freeze=False, timed=t))
        run(warm)
        ms = run(steps, True)
        P = 4 if prec == L.F32 else 8
        by = algorithmic_bytes_per_step(f.varnum, f.n_clauses, f.n_literals, R, P, adaptive)
        eng = {L.ENGINE_GATHER: "gather", L.ENGINE_TILE: "tile", L.ENGINE_SLAB: "slab"}[b.engine]
        b.close()
        F.close()
        emit(config=cfg, what=what, engine=eng, precision="f32" if prec == L.F32 else "f64", ms_per_step=ms / steps,
             clause_evals_per_s=steps * f.n_clauses * R / (ms * 1e-3), algorithmic_GBps=by * steps / (ms * 1e-3) / 1e9,
             frac_of_measured_hbm_peak=by * steps / (ms * 1e-3) / 1e9 / peak, frac_of_8TBps=by * steps / (ms * 1e-3) / 8e12)

    k = 1 if args.quick else 4
    f2 = cnf.random_ksat(10_000, 4.3, seed=20240613)
    for prec, R in ((L.F32, 4096), (L.F64, 2048)):
        for sched, sn_ in ((L.SCHED_BALANCED, "balanced"), (L.SCHED_EXACT, "exact")):
            throughput(2, f"random 3-SAT N=10k alpha=4.3, {R} replicas, fixed step, {sn_} schedule", f2, R, prec, False, 32 * k, 8, sched)
    f3 = cnf.random_ksat(1_000_000, 4.2, seed=20240614)
    for prec in (L.F32, L.F64):
        throughput(3, "random 3-SAT N=1M alpha=4.2 single instance, adaptive step (2 RHS evaluations per step)", f3, 1, prec, True, 16 * k, 4)
    f4 = cnf.random_ksat(50_000, 4.25, seed=20240615)
    for prec in (L.F32, L.F64):
        throughput(4, "random 3-SAT N=50k alpha=4.25, 2048 replicas (per-GPU share of 16384 over 8 GPUs), fixed step", f4, 2048, prec, False, 8 * k, 3)

if __name__ == "__main__":
    main()
