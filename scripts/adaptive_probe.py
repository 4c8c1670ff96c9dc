#!/usr/bin/env python
"""Adaptive batch steps (system.rs:111-139), tile engine against gather engine, device-resident state.
One JSON line per (engine, schedule, precision).   python scripts/adaptive_probe.py [--replicas R] [--steps K]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))

from bench import algorithmic_bytes_per_step, measured_peak   # noqa: E402
from odesat_b200 import _lib as L                             # noqa: E402
from odesat_b200 import batch as B                            # noqa: E402
from odesat_b200 import cnf                                   # noqa: E402
from odesat_b200.system import DeviceFormula                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--n", type=int, default=10_000)
    ap.add_argument("--alpha", type=float, default=4.3)
    ap.add_argument("--f64", action="store_true")
    ap.add_argument("--mix", default="", help="ragged formula instead: length:count,... (e.g. 2:13000,3:30000)")
    args = ap.parse_args()
    peak, _ = measured_peak()
    if args.mix:
        from ragged_probe import ragged
        f = ragged(args.n, {int(a.split(":")[0]): int(a.split(":")[1]) for a in args.mix.split(",")}, seed=3)
    else:
        f = cnf.random_ksat(args.n, args.alpha, seed=20240611 + 2)
    F = DeviceFormula(f)
    zeta = f.default_zeta()
    precs = [(L.F32, "f32", 4)] + ([(L.F64, "f64", 8)] if args.f64 else [])
    for prec, pname, P in precs:
        R = args.replicas if prec == L.F32 else args.replicas // 2
        for eng, ename, sched, sname in ((L.ENGINE_TILE, "tile", L.SCHED_EXACT, "exact"), (L.ENGINE_TILE, "tile", L.SCHED_BALANCED, "balanced"),
                                         (L.ENGINE_GATHER, "gather", L.SCHED_EXACT, "-")):
            b = B.ReplicaBatch(F, R, prec, eng, sched)
            b.init(1, 0)
            b.run_adaptive(1e-3, zeta, args.warmup)
            ms = b.run_adaptive(1e-3, zeta, args.steps, timed=True)
            by = algorithmic_bytes_per_step(f.varnum, f.n_clauses, f.n_literals, R, P, True)
            launches = b.launches
            b.close()
            print(json.dumps(dict(what="adaptive batch steps" + (", ragged formula (length:count " + args.mix + ")" if args.mix else ""), engine=ename, schedule=sname, precision=pname, N=f.varnum, M=f.n_clauses,
                                  replicas=R, steps=args.steps, ms_per_step=ms / args.steps,
                                  clause_evals_per_s=args.steps * f.n_clauses * R / (ms * 1e-3),
                                  rhs_evals_per_s=2 * args.steps * f.n_clauses * R / (ms * 1e-3),
                                  algorithmic_GBps=by * args.steps / (ms * 1e-3) / 1e9,
                                  roofline_frac=by * args.steps / (ms * 1e-3) / 1e9 / peak, launches=launches)), flush=True)


if __name__ == "__main__":
    main()
