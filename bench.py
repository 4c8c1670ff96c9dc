#!/usr/bin/env python
"""bench.py — throughput of the fused DMM Euler step (the hot path of odesat's src/system.rs).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[2]): synthetic uniform random 3-SAT, N = 10 000 variables,
alpha = 4.3 (M = 43 000 clauses), 4096 replicas PER GPU, fixed step dt = 0.01, f32 — one "step" is
one fused Euler step (RHS + update + clamps + all-satisfied check) of every replica.
Metric: clause-evals/s = steps x M x replicas / seconds.  Replicas are independent, so ranks
shard them with no data-path collective ("scaling": "weak": per-GPU work is fixed).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FORMULA_SEED = 20240611 + 2          # formula seed 20240611 + config index
RUN_SEED = 1
DT = 0.01


def workload(args):
    from odesat_b200 import cnf
    if args.workload == "rand10k":
        f = cnf.random_ksat(10_000, 4.3, seed=FORMULA_SEED)
        name = f"random 3-SAT N=10000 alpha=4.3 M=43000, {args.replicas} replicas/GPU, fixed step dt=0.01"
    elif args.workload == "rand20k":
        f = cnf.random_ksat(20_000, 4.3, seed=20240611 + 5)
        name = f"random 3-SAT N=20000 alpha=4.3, {args.replicas} replicas/GPU, fixed step dt=0.01"
    elif args.workload == "rand50k":
        f = cnf.random_ksat(50_000, 4.25, seed=20240611 + 4)
        name = f"random 3-SAT N=50000 alpha=4.25, {args.replicas} replicas/GPU, fixed step dt=0.01"
    elif args.workload == "rand1m":
        f = cnf.random_ksat(1_000_000, 4.2, seed=20240611 + 3)
        args.replicas = 1
        name = "random 3-SAT N=1000000 alpha=4.2 M=4200000, single instance, adaptive step tol=1e-3 (2 RHS evaluations per step)"
    elif args.workload == "hard":
        f = cnf.load_dimacs(str(ROOT / "tests" / "golden" / "aim100_unsat.cnf"))
        name = f"tests/hard.cnf (aim-100 UNSAT), {args.replicas} replicas/GPU, fixed step dt=0.01"
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    return f, name


def algorithmic_bytes_per_step(N, M, Lits, R, P, adaptive=False):
    """every state element read once and written once, formula CSR read once
    (adaptive: pass A reads y, writes y_half and y_full; pass B reads both, writes y_new)."""
    if adaptive:
        return R * 6 * P * (N + 2 * M) + 2 * (4 * Lits + 4 * (M + 1))
    return R * 2 * P * (N + 2 * M) + 4 * Lits + 4 * (M + 1)


class ClockSampler:
    """Samples the SM clock and the throttle reasons of one GPU WHILE the timed region runs: an NVML
    polling thread (the timed call is a ctypes call, which releases the GIL), every 2 ms, started
    before the region so that no sample is lost to start-up.  Falls back to `nvidia-smi -lms`."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, uuid: str | None = None):
        self.index, self.uuid, self.rows, self.proc, self.t = index, uuid, [], None, None
        self.stop = threading.Event()
        self.nvml = None

    def _nvml_loop(self):
        n, h = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        while not self.stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                r = get_reasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b, _ in bits])
            except Exception:
                pass
            time.sleep(0.001)

    def __enter__(self):
        try:
            import pynvml as n
            n.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = n.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
                except Exception:
                    h = None
            if h is None:
                h = n.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = (n, h)
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 2.0:      # NVML is warm before the region starts
                time.sleep(0.001)
            self.rows.clear()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "10"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:      # nvidia-smi is up before the region starts
                time.sleep(0.01)
            self.rows.clear()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        self.stop.set()
        if self.nvml is not None:
            self.t.join(timeout=5)
            return
        if self.proc:
            time.sleep(0.03)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.t.join(timeout=5)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, val in zip(self.NAMES, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_for(engine_name, precision, schedule, steps, launches):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu
    capture (profiles/traffic.json, bytes per Euler step), scaled to one launch of this run."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            per_step = json.loads(p.read_text()).get(f"{engine_name}_{precision}_{schedule}")
            return None if per_step is None else per_step * steps / max(launches, 1)
        except Exception:
            return None
    return None


def cpu_sample_adaptive(f, cores):
    """Single-instance adaptive workload: the oracle's `simulate` on ONE core (the reference is single-threaded)."""
    from oracle import oracle as O
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    v = F.init_v0(RUN_SEED, 0); xs = F.init_short_term_memory(); xl = np.ones(F.M)
    F.simulate(v, xs, xl, tol=1e-3, steps=2)                                                 # warm
    steps = 8
    t0 = time.perf_counter()
    F.simulate(v, xs, xl, tol=1e-3, steps=steps)
    sec = time.perf_counter() - t0
    return steps * f.n_clauses / sec, f"1 instance x {steps} adaptive steps (2 RHS evaluations each), f64, 1 thread, {sec:.2f} s"


def cpu_sample(f, args, cores, steps_hint=None):
    """Times the oracle (a C++ restatement of system.rs, f64, one replica per thread-slot — the way
    main.rs:278-308 runs `batch`, spread over the host cores) on a bounded sample of the workload."""
    from oracle import oracle as O
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = cores * 4
    v, xs, xl = F.init_batch(RUN_SEED, R, np.float64)
    F.batch_fixed(v, xs, xl, DT, f.default_zeta(), 2, freeze=False, nthreads=cores)        # warm
    t0 = time.perf_counter()
    F.batch_fixed(v, xs, xl, DT, f.default_zeta(), 5, freeze=False, nthreads=cores)
    per_step = (time.perf_counter() - t0) / 5
    steps = steps_hint or int(max(10, min(2000, 4.0 / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    F.batch_fixed(v, xs, xl, DT, f.default_zeta(), steps, freeze=False, nthreads=cores)
    sec = time.perf_counter() - t0
    return steps * f.n_clauses * R / sec, f"{R} replicas x {steps} fixed steps, f64, {cores} threads, {sec:.2f} s"


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port; the Rust crate cannot be built in
    this image) on the host cores.  Each step = one Euler step of a bounded replica sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    f, name = workload(args)
    cores = O.host_cores()
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = cores * 4
    v, xs, xl = F.init_batch(RUN_SEED, R, np.float64)
    zeta = f.default_zeta()
    for _ in range(args.warmup):
        F.batch_fixed(v, xs, xl, DT, zeta, 1, freeze=False, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        F.batch_fixed(v, xs, xl, DT, zeta, 1, freeze=False, nthreads=cores)
    sec = time.perf_counter() - t0
    value = args.steps * f.n_clauses * R / sec
    sample = f"{R} replicas per step (bounded sample of the {args.replicas}-replica workload), f64, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "clause-evals/sec", "value": value, "unit": "clause-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "note": "CPU port of src/system.rs (oracle/dmm_oracle.cpp); the Rust crate cannot be "
                   "compiled here (no cargo/rustc)"},
        "cpu_baseline": {"value": value, "unit": "clause-evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clause-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from odesat_b200 import _lib as L
    from odesat_b200 import batch as B
    from odesat_b200.system import DeviceFormula

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    f, name = workload(args)
    prec = L.F32 if args.precision == "f32" else L.F64
    P = 4 if prec == L.F32 else 8
    engine = {"auto": L.ENGINE_AUTO, "gather": L.ENGINE_GATHER, "tile": L.ENGINE_TILE}[args.engine]
    sched = L.SCHED_EXACT if args.schedule == "exact" else L.SCHED_BALANCED
    F = DeviceFormula(f)
    R = args.replicas
    zeta = f.default_zeta()
    b = B.ReplicaBatch(F, R, prec, engine, sched)
    b.init(RUN_SEED, rank * R)
    eng_name = {L.ENGINE_GATHER: "gather", L.ENGINE_TILE: "tile"}[b.engine]

    adaptive = args.workload == "rand1m"

    def run(n, timed=False):
        if adaptive:
            return b.run_adaptive(1e-3, zeta, n, timed=timed)
        return b.run_fixed(DT, zeta, n, freeze=False, timed=timed)

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps --------------
    run(args.warmup)
    launches0 = b.launches
    barrier()
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_uuid = None
    with ClockSampler(local, gpu_uuid) as clk:
        ms = run(args.steps, timed=True)
    barrier()
    n_launch = b.launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    evals = args.steps * f.n_clauses * R * world
    value = evals / (ms_max * 1e-3)
    bytes_step = algorithmic_bytes_per_step(f.varnum, f.n_clauses, f.n_literals, R, P, adaptive)
    achieved = bytes_step * args.steps / (ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    st, _ = b.status()
    flagged = int((st >= 0).sum())
    b.close()
    other = None
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "workload": args.workload, "engine": eng_name, "schedule": args.schedule,
                              "precision": args.precision, "ms_per_step": ms_max / args.steps, "value": value,
                              "frac": achieved / peak, "frac_of_8TBs": achieved / 8000.0, "launches": n_launch,
                              "flagged": flagged, "env": {k: v for k, v in os.environ.items() if k.startswith("ODESAT_")}}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    if not args.quick and eng_name == "tile":
        # the same K steps with the other clause schedule, for the record (not the headline value)
        osched = L.SCHED_EXACT if sched == L.SCHED_BALANCED else L.SCHED_BALANCED
        ob = B.ReplicaBatch(F, R, prec, engine, osched)
        ob.init(RUN_SEED, rank * R)
        ob.run_fixed(DT, zeta, args.warmup, freeze=False)
        oms = ob.run_fixed(DT, zeta, args.steps, freeze=False, timed=True)
        ob.close()
        other = {"schedule": "exact" if osched == L.SCHED_EXACT else "balanced", "ms_per_step": oms / args.steps,
                 "roofline_frac": bytes_step * args.steps / (oms * 1e-3) / 1e9 / peak}

    # ---- end to end through the C ABI with HOST (pinned) buffers ------------------------------
    dt_t = torch.float32 if prec == L.F32 else torch.float64
    hv = torch.empty((R, f.varnum), dtype=dt_t).pin_memory()
    gen = torch.Generator().manual_seed(RUN_SEED + rank)
    hv.uniform_(-1.0, 1.0, generator=gen)
    e2e_steps = args.steps

    def e2e_call():
        # inputs = the random v0 of every replica, from pinned host memory (main.rs:283-289 draws them
        # on the host); xs0 / xl0 are functions of the formula (init_short_term_memory, ones) and are
        # built on the device, as `batch` builds them from the formula
        return B.simulate_batch(F, R, hv.data_ptr(), None, None, step_size=None if adaptive else DT,
                                tolerance=1e-3 if adaptive else None, steps=e2e_steps,
                                precision=prec, engine=engine, schedule=sched, mode=L.MODE_BATCH, write_back=False,
                                chunk=max(32, e2e_steps))
    e2e_call()                                                    # warm (allocator, schedule cache)
    barrier()
    t0 = time.perf_counter()
    res = e2e_call()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    te = torch.tensor([sec], dtype=torch.float64, device=dev)
    work = torch.tensor([float(np.where(res.solved_step >= 0, res.solved_step + 1, e2e_steps).sum()) * f.n_clauses],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    e2e_value = float(work.item()) / float(te.item())
    h2d = R * f.varnum * P
    d2h = R * 8 + R + f.varnum

    out = None
    if rank == 0:
        from oracle import oracle as O
        cores = O.host_cores()
        if world > 1:
            cpu_val, cpu_sample_desc = None, None
        elif adaptive:
            cpu_val, cpu_sample_desc = cpu_sample_adaptive(f, cores)
            cores = 1
        else:
            cpu_val, cpu_sample_desc = cpu_sample(f, args, cores)
        out = {
            "metric": "clause-evals/sec", "value": value, "unit": "clause-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": {"workload": name, "replicas_per_gpu": R, "replicas_total": R * world, "N": f.varnum,
                       "M": f.n_clauses, "engine": eng_name, "schedule": args.schedule, "formula_seed": FORMULA_SEED,
                       "parallelism": f"replica-sharded x{world}, no data-path collective",
                       "l2": (f"state {bytes_step / 2 / 1e6:.0f} MB per GPU is larger than L2 (126 MB); no flush needed" if bytes_step / 2 > 126e6 else
                              f"state {bytes_step / 2 / 1e6:.1f} MB per GPU fits in L2 (126 MB) and is NOT flushed between steps: a step's input is the previous step's output by construction"),
                       "flagged_replicas": flagged, "other_schedule_same_run": other,
                       "e2e_call": f"one odesat_simulate_batch call of {e2e_steps} steps per GPU: v0 of every replica "
                                   "from pinned host memory in; per-replica flags, exact verification and the "
                                   "winner's assignment out"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "clause-evals/s", "h2d_bytes_per_step": h2d / e2e_steps,
                    "d2h_bytes_per_step": d2h / e2e_steps, "seconds_per_call": float(te.item())},
            "gpu_launches": n_launch,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_for(eng_name, args.precision, args.schedule, args.steps, n_launch),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_step": bytes_step,
                         "algorithmic_bytes_per_launch": bytes_step * args.steps / max(n_launch, 1), "launches": n_launch,
                         "avg_launch_ms": ms / max(n_launch, 1), "frac_of_8TBs": achieved / 8000.0},
        }
        if cpu_val is not None:
            out["cpu_baseline"] = {"value": cpu_val, "unit": "clause-evals/s", "cores": cores, "kind": "port",
                                   "sample": cpu_sample_desc}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rand10k", choices=["rand10k", "rand20k", "rand50k", "rand1m", "hard"])
    ap.add_argument("--replicas", type=int, default=4096, help="replicas per GPU")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--engine", default="auto", choices=["auto", "gather", "tile"])
    ap.add_argument("--schedule", default="balanced", choices=["exact", "balanced"],
                    help="tile-engine clause schedule: balanced = throughput mode (dv summed in colour order, "
                         "agrees with the reference to rounding); exact = the reference's summation order, bit-identical")
    ap.add_argument("--quick", action="store_true", help="tuning aid: device-resident timing only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    import __graft_entry__ as g
    if int(os.environ.get("LOCAL_RANK", "0")) == 0 and os.environ.get("ODESAT_SKIP_BUILD") != "1":
        g.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
