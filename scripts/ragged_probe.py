#!/usr/bin/env python
"""Fixed steps on a RAGGED formula (binary + ternary + a few long clauses), tile engine (k_tile_ragged) against the
gather engine, device-resident state.  One JSON line per (engine, schedule).   python scripts/ragged_probe.py"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from odesat_b200 import _lib as L                             # noqa: E402
from odesat_b200 import batch as B                            # noqa: E402
from odesat_b200 import cnf                                   # noqa: E402
from odesat_b200.system import DeviceFormula                  # noqa: E402


def ragged(n_vars, counts, seed):
    """counts: {length: number of clauses}; distinct variables per clause, interleaved."""
    rng = np.random.default_rng(seed)
    ks = np.concatenate([np.full(n, k) for k, n in counts.items()])
    rng.shuffle(ks)
    off, lits = [0], []
    for k in ks:
        vs = rng.choice(n_vars, size=int(k), replace=False) + 1
        sg = rng.integers(0, 2, size=int(k)) * 2 - 1
        lits.extend(int(a * b) for a, b in zip(vs, sg))
        off.append(len(lits))
    return cnf.Formula(n_vars, np.asarray(off, np.int64), np.asarray(lits, np.int32), {})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mix", default="2:12000,3:28000,5:2500,8:500", help="length:count,...")
    args = ap.parse_args()
    counts = {int(a.split(":")[0]): int(a.split(":")[1]) for a in args.mix.split(",")}
    f = ragged(10_000, counts, seed=3)
    F = DeviceFormula(f)
    zeta = f.default_zeta()
    for eng, ename, sched, sname in ((L.ENGINE_TILE, "tile", L.SCHED_EXACT, "exact"), (L.ENGINE_TILE, "tile", L.SCHED_BALANCED, "balanced"),
                                     (L.ENGINE_GATHER, "gather", L.SCHED_EXACT, "-")):
        b = B.ReplicaBatch(F, args.replicas, L.F32, eng, sched)
        b.init(1, 0)
        b.run_fixed(0.01, zeta, args.warmup, freeze=False)
        ms = b.run_fixed(0.01, zeta, args.steps, freeze=False, timed=True)
        b.close()
        print(json.dumps(dict(what="fixed steps, ragged formula (length:count " + args.mix + ")", engine=ename, schedule=sname,
                              precision="f32", N=f.varnum, M=f.n_clauses, L=f.n_literals, replicas=args.replicas, steps=args.steps,
                              ms_per_step=ms / args.steps, clause_evals_per_s=args.steps * f.n_clauses * args.replicas / (ms * 1e-3))), flush=True)


if __name__ == "__main__":
    main()
