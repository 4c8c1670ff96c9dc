#!/usr/bin/env python
"""`inter` across GPUs (BASELINE configs[4] shape): one process per GPU, contiguous replica
shards, no data-path collective; after every chunk of Euler steps ONE 8-byte MIN all-reduce of the
early-exit key (first_flag_step << 32 | global replica) over NCCL decides whether anybody flagged.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        scripts/inter_multi_gpu.py --vars 50000 --alpha 3.6 --replicas 4096

Prints one JSON line on rank 0: winner, the step it flagged at, exact verification of the winner's
assignment (on the device and again on the host), and — with --check — that a single-GPU run of the
whole batch names the same winner at the same step."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", type=int, default=50_000)
    ap.add_argument("--alpha", type=float, default=3.6)
    ap.add_argument("--replicas", type=int, default=4096, help="total over all ranks")
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--max-steps", type=int, default=20_000)
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--dt", type=float, default=0.01)
    ap.add_argument("--check", action="store_true", help="rank 0 re-runs the whole batch alone and compares the winner")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from odesat_b200 import _lib as L
    from odesat_b200 import batch as B
    from odesat_b200 import cnf
    from odesat_b200.system import DeviceFormula

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prec = L.F32 if args.precision == "f32" else L.F64
    f = cnf.random_ksat(args.vars, args.alpha, seed=20240611 + 4)
    F = DeviceFormula(f)
    lo, hi = B.shard_range(args.replicas, rank, world)
    sizes = [B.shard_range(args.replicas, r, world)[1] - B.shard_range(args.replicas, r, world)[0] for r in range(world)]
    b = B.ReplicaBatch(F, hi - lo, prec)
    b.init(args.seed, lo)                                   # replica r gets the stream (seed, global r) on any sharding
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = B.run_sharded_inter(b, args.dt, f.default_zeta(), args.max_steps, args.chunk, lo, sizes, device=dev)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    ok_dev = ok_host = None
    assign = None
    if res.winner >= 0 and res.winner_rank == rank:
        local_r = res.winner - lo
        ok_dev = bool(b.verify()[local_r])                  # cnf.rs:246-264 on the device
        assign = b.assignment(local_r)
        ok_host = bool(f.evaluate(assign))                  # and again on the host
    flag = torch.tensor([-1 if ok_dev is None else int(ok_dev and ok_host)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        step = B.decode_key(res.key)[0] if res.key != B.NO_KEY else -1
        out = {"mode": "inter", "n_gpus": world, "N": f.varnum, "M": f.n_clauses, "replicas": args.replicas,
               "chunk": args.chunk, "dt": args.dt, "alpha": args.alpha, "engine": {L.ENGINE_GATHER: "gather", L.ENGINE_TILE: "tile"}[b.engine],
               "winner": res.winner, "winner_rank": res.winner_rank, "flag_step": step, "steps_run": res.steps_run,
               "winner_verified_sat": bool(flag.item() == 1) if res.winner >= 0 else None,
               "seconds": sec, "clause_evals_per_s": res.steps_run * f.n_clauses * args.replicas / sec,
               "allreduces": (res.steps_run + args.chunk - 1) // args.chunk, "allreduce_bytes": 8}
        if args.check:
            b.close()
            one = B.ReplicaBatch(F, args.replicas, prec)
            one.init(args.seed, 0)
            r1 = B.run_sharded_inter(one, args.dt, f.default_zeta(), args.max_steps, args.chunk, 0, [args.replicas], reduce_fn=int)
            out["single_gpu_same_winner"] = bool(r1.key == res.key)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
