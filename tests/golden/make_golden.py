"""Generates tests/golden/*.npz from the CPU oracle (oracle/dmm_oracle.cpp).

The reference itself cannot run here (Rust, no toolchain) and holds no golden vectors, so these
fixtures pin the ORACLE's behaviour (regression) and give the GPU tests size-stable targets; the
oracle in turn is pinned by the hand-derived KATs in tests/test_oracle_kat.py.
Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from odesat_b200 import cnf  # noqa: E402
from oracle import oracle as O  # noqa: E402

G = Path(__file__).resolve().parent


def traj(name, f, seed, dtype, fixed_steps=100, adaptive_steps=50):
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    zeta = f.default_zeta()
    v0 = F.init_v0(seed, 0, dtype)
    xs0 = F.init_short_term_memory(dtype)
    xl0 = np.ones(F.M, dtype)
    dv, dxs, dxl, allsat = F.compute_derivatives(v0, xs0, xl0, zeta)
    v, xs, xl = v0.copy(), xs0.copy(), xl0.copy()
    flags = []
    for _ in range(fixed_steps):
        flags.append(F.euler_step_fixed(v, xs, xl, 0.01, zeta))
    av, axs, axl = v0.copy(), xs0.copy(), xl0.copy()
    dt, dts, aflags = 0.01, [], []
    for _ in range(adaptive_steps):
        a, dt = F.euler_step(av, axs, axl, 1e-3, dt, zeta)
        dts.append(dt)
        aflags.append(a)
    np.savez_compressed(G / f"{name}.npz", varnum=f.varnum, clause_off=f.clause_off, lits=f.lits, zeta=zeta, seed=seed,
                        v0=v0, xs0=xs0, xl0=xl0, dv=dv, dxs=dxs, dxl=dxl, allsat=allsat,
                        fixed_v=v, fixed_xs=xs, fixed_xl=xl, fixed_flags=np.array(flags),
                        adapt_v=av, adapt_xs=axs, adapt_xl=axl, adapt_dt=np.array(dts), adapt_flags=np.array(aflags))
    print(name, "ok")


if __name__ == "__main__":
    sat = cnf.load_dimacs(str(G / "aim100_sat.cnf"))
    traj("traj_aim100_f64", sat, 5, np.float64)
    traj("traj_aim100_f32", sat, 5, np.float32)
    rnd = cnf.random_ksat(300, 4.3, seed=20240611)
    traj("traj_rand300_f64", rnd, 9, np.float64)
    toy = cnf.load_dimacs(str(G / "toy_mixed.cnf"))
    traj("traj_toy_f64", toy, 2, np.float64, fixed_steps=30, adaptive_steps=30)
