#!/usr/bin/env python
"""bench.py — throughput of the fused DMM Euler step (the hot path of odesat's src/system.rs).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[2]): synthetic uniform random 3-SAT, N = 10 000 variables,
alpha = 4.3 (M = 43 000 clauses), 4096 replicas IN TOTAL sharded over the GPUs ("scaling": "strong", the way
configs[2] states it: "4096 replicas ... sharded 1/2/4/8"), fixed step dt = 0.01, f32 — one "step" is one fused
Euler step (RHS + update + clamps + all-satisfied check) of every replica.
Metric: clause-evals/s = steps x M x replicas / seconds.  Replicas are independent, so ranks shard them with no
data-path collective.  With N > 1 the same line also carries the weak-scaling figure (4096 replicas PER GPU,
`weak_scaling_same_run`), and — every N — `inter`: BASELINE configs[4] (`inter`, N = 50 000, alpha = 4.25, 16 384
replicas in total, chunks of 32 steps, early exit armed) with the per-chunk MIN all-reduce of the early-exit key
INSIDE the CUDA-event region, next to the same chunks without the collective.
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
        name = f"random 3-SAT N=10000 alpha=4.3 M=43000, {args.replicas} replicas, fixed step dt=0.01"
    elif args.workload == "rand20k":
        f = cnf.random_ksat(20_000, 4.3, seed=20240611 + 5)
        name = f"random 3-SAT N=20000 alpha=4.3, {args.replicas} replicas, fixed step dt=0.01"
    elif args.workload == "rand50k":
        f = cnf.random_ksat(50_000, 4.25, seed=20240611 + 4)
        name = f"random 3-SAT N=50000 alpha=4.25, {args.replicas} replicas, fixed step dt=0.01"
    elif args.workload == "rand1m":
        f = cnf.random_ksat(1_000_000, 4.2, seed=20240611 + 3)
        args.replicas = 1
        name = "random 3-SAT N=1000000 alpha=4.2 M=4200000, single instance, adaptive step tol=1e-3 (2 RHS evaluations per step)"
    elif args.workload == "hard":
        f = cnf.load_dimacs(str(ROOT / "tests" / "golden" / "aim100_unsat.cnf"))
        name = f"tests/hard.cnf (aim-100 UNSAT), {args.replicas} replicas, fixed step dt=0.01"
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    return f, name


def algorithmic_bytes_per_step(N, M, Lits, R, P, adaptive=False):
    """every state element read once and written once, formula CSR read once
    (adaptive: pass A reads y, writes y_half and y_full; pass B reads both, writes y_new)."""
    if adaptive:
        return R * 6 * P * (N + 2 * M) + 2 * (4 * Lits + 4 * (M + 1))
    return R * 2 * P * (N + 2 * M) + 4 * Lits + 4 * (M + 1)


def shard_range(R, rank, world):
    """Contiguous replica range of a rank; same rule as odesat_b200.batch.shard_range."""
    return R * rank // world, R * (rank + 1) // world


def make_config(args, f, name, world):
    """The workload's identity — the SAME dict in both arms (`--impl ours` / `--impl reference`), so that the
    driver can tell that they measured the same thing; everything arm-specific goes to `detail`."""
    P = 4 if args.precision == "f32" else 8
    per_gpu = -(-args.replicas // world)
    state_mb = per_gpu * P * (f.varnum + 2 * f.n_clauses) / 1e6
    return {"workload": name, "N": f.varnum, "M": f.n_clauses, "replicas_total": args.replicas,
            "replicas_per_gpu": per_gpu, "gpus": world, "scaling": args.scaling, "step": "fixed dt=0.01" if args.workload != "rand1m" else "adaptive tol=1e-3",
            "formula_seed": FORMULA_SEED,
            "l2": (f"GPU arm: state {state_mb:.0f} MB per GPU is larger than L2 (126 MB); no flush needed" if state_mb > 126 else
                   f"GPU arm: state {state_mb:.1f} MB per GPU fits in L2 (126 MB) and is NOT flushed between steps: a step's "
                   "input is the previous step's output by construction")}


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


def cpu_sample(f, cores, budget_s=4.0):
    """Times the oracle (a C++ restatement of system.rs, f64, one replica per thread-slot — the way
    main.rs:278-308 runs `batch`, spread over `cores` host threads) on a bounded sample of the workload."""
    from oracle import oracle as O
    F = O.OracleFormula(f.varnum, f.clause_off, f.lits)
    R = cores * 4
    v, xs, xl = F.init_batch(RUN_SEED, R, np.float64)
    F.batch_fixed(v, xs, xl, DT, f.default_zeta(), 2, freeze=False, nthreads=cores)        # warm
    t0 = time.perf_counter()
    F.batch_fixed(v, xs, xl, DT, f.default_zeta(), 5, freeze=False, nthreads=cores)
    per_step = (time.perf_counter() - t0) / 5
    steps = int(max(10, min(2000, budget_s / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    F.batch_fixed(v, xs, xl, DT, f.default_zeta(), steps, freeze=False, nthreads=cores)
    sec = time.perf_counter() - t0
    return steps * f.n_clauses * R / sec, f"{R} replicas x {steps} fixed steps, f64, {cores} thread{'s' if cores > 1 else ''}, {sec:.2f} s"


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port; the Rust crate cannot be built in
    this image) on the host cores.  Each step = one Euler step of a bounded replica sample.  Loads nothing of the
    GPU library."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "-s"], check=True)
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
    single, single_desc = cpu_sample(f, 1, budget_s=3.0)
    sample = f"{R} replicas per step (bounded sample of the {args.replicas}-replica workload), f64, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "clause-evals/sec", "value": value, "unit": "clause-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, f, name, world),
        "detail": {"note": "CPU port of src/system.rs (oracle/dmm_oracle.cpp); the Rust crate cannot be compiled here "
                           "(no cargo/rustc).  The real binary is single-threaded (see cpu_baseline.single_core)."},
        "cpu_baseline": {"value": value, "unit": "clause-evals/s", "cores": cores, "kind": "port", "sample": sample,
                         "single_core": {"value": single, "unit": "clause-evals/s", "cores": 1, "sample": single_desc}},
        "e2e": {"value": value, "unit": "clause-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def measure_inter(args, torch, dist, rank, world, local, dev, barrier):
    """BASELINE configs[4]: `inter`, random 3-SAT N = 50 000, alpha = 4.25, 16 384 replicas IN TOTAL over the ranks,
    fixed step, chunks of 32 steps with the early exit armed (freeze = 1; system.rs:279-293).  After every chunk each
    rank reduces its flags to one int64 key on its GPU, the keys are MIN all-reduced over NCCL on the same stream, the
    result goes to pinned host memory, and the host reads it ONE CHUNK LATE while the next chunk — whose kernels check
    the previous key on the device — is already running.  The CUDA events bracket the whole loop, collective included;
    the same chunks are then timed without the collective."""
    from odesat_b200 import _lib as L
    from odesat_b200 import batch as B
    from odesat_b200 import cnf
    from odesat_b200.system import DeviceFormula

    total = args.inter_replicas
    lo, hi = shard_range(total, rank, world)
    R = hi - lo
    f = cnf.random_ksat(50_000, 4.25, seed=20240611 + 4)
    F = DeviceFormula(f)
    zeta = f.default_zeta()
    b = B.ReplicaBatch(F, R, L.F32, L.ENGINE_AUTO, L.SCHED_EXACT)
    b.init(RUN_SEED, lo)
    stream = torch.cuda.ExternalStream(b.stream, device=dev)
    keys = torch.full((2, 2), L.INT64_MAX, dtype=torch.int64, device=dev)
    host = torch.full((2, 2), L.INT64_MAX, dtype=torch.int64).pin_memory()
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    chunk, warm_chunks, chunks = 32, 1, args.inter_chunks

    def loop(n_chunks, collective):
        """→ (device ms, chunk index at which a key was seen or -1)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hit = -1
        with torch.cuda.stream(stream):
            e0.record()
            for c in range(n_chunks):
                b.run_fixed_async(DT, zeta, chunk, True, keys[(c - 1) & 1].data_ptr() if c else 0)
                b.post_key(lo, keys[c & 1].data_ptr())
                if collective and world > 1:
                    dist.all_reduce(keys[c & 1][0:1], op=dist.ReduceOp.MIN)
                host[c & 1].copy_(keys[c & 1], non_blocking=True)
                evs[c & 1].record()
                if c >= 1:                                   # one chunk late: chunk c is already enqueued
                    evs[(c - 1) & 1].synchronize()
                    if int(host[(c - 1) & 1][0]) != L.INT64_MAX:
                        hit = c - 1
                        break
            e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1), hit

    def exchange_only(n):
        """The per-chunk exchange by itself — flag reduction, MIN all-reduce, copy to pinned host, host poll one round
        late — with no integration in between: its cost per chunk in isolation (a chunk of 32 steps takes 10^4 times
        longer, so the difference of the two loops above is below their run-to-run noise)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for c in range(n):
                b.post_key(lo, keys[c & 1].data_ptr())
                if world > 1:
                    dist.all_reduce(keys[c & 1][0:1], op=dist.ReduceOp.MIN)
                host[c & 1].copy_(keys[c & 1], non_blocking=True)
                evs[c & 1].record()
                if c >= 1:
                    evs[(c - 1) & 1].synchronize()
            e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    loop(warm_chunks, True)
    barrier()
    ms_c, hit = loop(chunks, True)
    barrier()
    ms_n, _ = loop(chunks, False)
    barrier()
    exchange_only(20)
    barrier()
    us_x = exchange_only(200)
    barrier()
    t = torch.tensor([ms_c, ms_n, us_x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_c, ms_n, us_x = float(t[0]), float(t[1]), float(t[2])
    eng = {L.ENGINE_GATHER: "gather", L.ENGINE_TILE: "tile", L.ENGINE_SLAB: "slab"}[b.engine]
    # the pinned block and the events were used on the batch's stream: release them while that stream still exists
    del host, keys, evs, t
    torch.cuda.synchronize()
    b.close()
    steps = chunk * chunks
    bytes_step = algorithmic_bytes_per_step(f.varnum, f.n_clauses, f.n_literals, R, 4)
    peak, _ = measured_peak()
    return {
        "workload": f"inter: random 3-SAT N=50000 alpha=4.25 M={f.n_clauses}, {total} replicas in total ({R} on rank 0), "
                    f"fixed step dt=0.01, {chunks} chunks of {chunk} steps, early exit armed",
        "value": steps * f.n_clauses * total / (ms_c * 1e-3), "unit": "clause-evals/s", "ms_per_step": ms_c / steps,
        "engine": eng, "early_exit_hit_chunk": hit,
        "collective": {"op": "MIN all-reduce of one int64 early-exit key per chunk (NCCL, on the compute stream)" if world > 1
                             else "none at 1 GPU: the key goes device -> pinned host",
                       "in_timed_region": True, "per_run": chunks, "bytes": 8,
                       "ms_per_step_without": ms_n / steps,
                       "difference_us_per_chunk": (ms_c - ms_n) / chunks * 1e3,
                       "exchange_alone_us_per_chunk": us_x,
                       "cost_fraction": us_x * 1e-3 * chunks / ms_c,
                       "note": "difference = the same chunks with and without the collective (below run-to-run noise); "
                               "exchange_alone = 200 back-to-back rounds of flag reduction + all-reduce + pinned copy + "
                               "one-round-late host poll, max over ranks; cost_fraction = exchange_alone x chunks / timed region"},
        "roofline_frac": bytes_step * steps / (ms_c * 1e-3) / 1e9 / peak,
    }


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
    engine = {"auto": L.ENGINE_AUTO, "gather": L.ENGINE_GATHER, "tile": L.ENGINE_TILE, "slab": L.ENGINE_SLAB}[args.engine]
    sched = L.SCHED_EXACT if args.schedule == "exact" else L.SCHED_BALANCED
    F = DeviceFormula(f)
    zeta = f.default_zeta()
    adaptive = args.workload == "rand1m"
    strong = args.scaling == "strong"
    if strong:
        lo, hi = shard_range(args.replicas, rank, world)      # configs[2]: the replicas of ONE job, sharded
    else:
        lo, hi = rank * args.replicas, (rank + 1) * args.replicas
    R = hi - lo
    total = args.replicas if strong else args.replicas * world

    def device_resident(R_local, offset, schedule, timed_clock=False):
        """W warm-up steps, then exactly K timed steps (CUDA events on the library's stream) on state resident in HBM."""
        b = B.ReplicaBatch(F, R_local, prec, engine, schedule)
        b.init(RUN_SEED, offset)

        def run(n, timed=False):
            if adaptive:
                return b.run_adaptive(1e-3, zeta, n, timed=timed)
            return b.run_fixed(DT, zeta, n, freeze=False, timed=timed)
        run(args.warmup)
        l0 = b.launches
        barrier()
        clk = None
        if timed_clock:
            try:
                gpu_uuid = str(torch.cuda.get_device_properties(local).uuid)
            except Exception:
                gpu_uuid = None
            with ClockSampler(local, gpu_uuid) as clk:
                ms = run(args.steps, timed=True)
        else:
            ms = run(args.steps, timed=True)
        barrier()
        n_launch = b.launches - l0
        st, _ = b.status()
        eng = {L.ENGINE_GATHER: "gather", L.ENGINE_TILE: "tile", L.ENGINE_SLAB: "slab"}[b.engine]
        b.close()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return ms, float(t.item()), n_launch, int((st >= 0).sum()), eng, clk

    # ---- device-resident throughput: the headline value ------------------------------------------
    ms, ms_max, n_launch, flagged, eng_name, clk = device_resident(R, lo, sched, timed_clock=True)
    value = args.steps * f.n_clauses * total / (ms_max * 1e-3)
    bytes_step = algorithmic_bytes_per_step(f.varnum, f.n_clauses, f.n_literals, R, P, adaptive)
    achieved = bytes_step * args.steps / (ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "workload": args.workload, "engine": eng_name, "schedule": args.schedule,
                              "precision": args.precision, "replicas_per_gpu": R, "ms_per_step": ms_max / args.steps, "value": value,
                              "frac": achieved / peak, "frac_of_8TBs": achieved / 8000.0, "launches": n_launch,
                              "flagged": flagged, "env": {k: v for k, v in os.environ.items() if k.startswith("ODESAT_")}}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    other = None
    if eng_name == "tile":
        # the same K steps with the other clause schedule, for the record (not the headline value)
        osched = L.SCHED_EXACT if sched == L.SCHED_BALANCED else L.SCHED_BALANCED
        oms, oms_max, _, _, _, _ = device_resident(R, lo, osched)
        other = {"schedule": "exact" if osched == L.SCHED_EXACT else "balanced", "ms_per_step": oms_max / args.steps,
                 "value": args.steps * f.n_clauses * total / (oms_max * 1e-3),
                 "roofline_frac": bytes_step * args.steps / (oms * 1e-3) / 1e9 / peak}
    adapt = None
    if eng_name == "tile" and not adaptive:
        # the same shard integrated with ADAPTIVE steps (system.rs:111-139, `batch` without -s) on the tile engine's
        # adaptive kernel, for the record: one step = two RHS evaluations; bytes = SURVEY §8(d)'s adaptive formula
        ab = B.ReplicaBatch(F, R, prec, engine, sched)
        ab.init(RUN_SEED, lo)
        ab.run_adaptive(1e-3, zeta, args.warmup)
        barrier()
        ams = ab.run_adaptive(1e-3, zeta, args.steps, timed=True)
        barrier()
        aeng = {L.ENGINE_GATHER: "gather", L.ENGINE_TILE: "tile", L.ENGINE_SLAB: "slab"}[ab.engine]
        ab.close()
        t = torch.tensor([ams], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ams_max = float(t.item())
        abytes = algorithmic_bytes_per_step(f.varnum, f.n_clauses, f.n_literals, R, P, True)
        adapt = {"engine": aeng, "schedule": args.schedule, "ms_per_adaptive_step": ams_max / args.steps,
                 "rhs_evals_per_s": 2 * args.steps * f.n_clauses * total / (ams_max * 1e-3),
                 "roofline_frac_on_adaptive_algorithmic_bytes": abytes * args.steps / (ams * 1e-3) / 1e9 / peak}
    weak = None
    if strong and world > 1 and not adaptive:
        # the weak-scaling figure of the same run: args.replicas replicas on EVERY GPU
        _, wms_max, _, _, _, _ = device_resident(args.replicas, rank * args.replicas, sched)
        weak = {"replicas_per_gpu": args.replicas, "replicas_total": args.replicas * world, "ms_per_step": wms_max / args.steps,
                "value": args.steps * f.n_clauses * args.replicas * world / (wms_max * 1e-3), "unit": "clause-evals/s"}

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

    inter = None
    if args.inter_chunks > 0 and args.workload == "rand10k":
        inter = measure_inter(args, torch, dist, rank, world, local, dev, barrier)

    if rank == 0:
        from oracle import oracle as O
        cores = O.host_cores()
        cpu = None
        if world == 1:
            if adaptive:
                cpu_val, cpu_desc = cpu_sample_adaptive(f, cores)
                cpu = {"value": cpu_val, "unit": "clause-evals/s", "cores": 1, "kind": "port", "sample": cpu_desc}
            else:
                cpu_val, cpu_desc = cpu_sample(f, cores)
                one_val, one_desc = cpu_sample(f, 1, budget_s=3.0)
                cpu = {"value": cpu_val, "unit": "clause-evals/s", "cores": cores, "kind": "port", "sample": cpu_desc,
                       "single_core": {"value": one_val, "unit": "clause-evals/s", "cores": 1, "sample": one_desc,
                                       "note": "the reference binary is single-threaded (no par_* call in src/)"}}
        out = {
            "metric": "clause-evals/sec", "value": value, "unit": "clause-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": make_config(args, f, name, world),
            "detail": {"engine": eng_name, "schedule": args.schedule,
                       "parallelism": f"replica-sharded x{world}, no data-path collective",
                       "flagged_replicas": flagged, "other_schedule_same_run": other, "adaptive_steps_same_run": adapt,
                       "e2e_call": f"one odesat_simulate_batch call of {e2e_steps} steps per GPU: v0 of every replica "
                                   "from pinned host memory in (uploaded sub-batch by sub-batch, overlapping the integration); "
                                   "per-replica flags, exact verification and the winner's assignment out"},
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
        if weak is not None:
            out["weak_scaling_same_run"] = weak
        if inter is not None:
            out["inter"] = inter
            out["collective"] = inter["collective"]
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
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
    ap.add_argument("--replicas", type=int, default=4096, help="replicas of the job (strong scaling: sharded over the GPUs; weak: per GPU)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE configs[2]): --replicas in total, sharded 1/2/4/8; weak: --replicas per GPU")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--engine", default="auto", choices=["auto", "gather", "tile", "slab"])
    ap.add_argument("--schedule", default="balanced", choices=["exact", "balanced"],
                    help="tile-engine clause schedule: balanced = throughput mode (dv summed in colour order, "
                         "agrees with the reference to rounding); exact = the reference's summation order, bit-identical")
    ap.add_argument("--inter-chunks", type=int, default=3, help="chunks of 32 steps of the configs[4] `inter` measurement (0 = skip)")
    ap.add_argument("--inter-replicas", type=int, default=16384, help="replicas of the `inter` measurement, in total")
    ap.add_argument("--quick", action="store_true", help="tuning aid: device-resident timing only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload == "rand1m":
        args.scaling = "weak"                     # a single instance does not shard: replicas only
    if args.impl == "reference":
        run_reference(args)                       # CPU only: never builds or loads the GPU library
        return
    import __graft_entry__ as g
    if int(os.environ.get("LOCAL_RANK", "0")) == 0 and os.environ.get("ODESAT_SKIP_BUILD") != "1":
        g.build()
    run_gpu(args)


if __name__ == "__main__":
    main()
