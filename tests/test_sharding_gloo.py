"""The N>1 host path on CPU: replica sharding and the 8-byte MIN all-reduce early exit, run as
two gloo processes with stub batches standing in for the GPU batches."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from odesat_b200 import batch as B


def test_shard_ranges_partition_the_replicas():
    for R in (0, 1, 7, 100, 4096, 16384):
        for W in (1, 2, 3, 8):
            edges = [B.shard_range(R, r, W) for r in range(W)]
            assert edges[0][0] == 0 and edges[-1][1] == R
            assert all(edges[i][1] == edges[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1


def test_key_encoding_orders_by_step_then_replica():
    assert B.decode_key(B.encode_key(7, 123)) == (7, 123)
    assert B.encode_key(3, 4000) < B.encode_key(4, 0) < B.NO_KEY
    assert B.globalize_key(B.NO_KEY, 50) == B.NO_KEY
    assert B.globalize_key(B.encode_key(9, 2), 50) == B.encode_key(9, 52)


class StubBatch:
    """Flags replica `r` at step `s` for the given {local replica: step} table."""

    def __init__(self, table):
        self.table, self.steps = table, 0

    def run_fixed(self, dt, zeta, n, freeze):
        self.steps += n

    def first_solved(self):
        keys = [B.encode_key(s, r) for r, s in self.table.items() if s < self.steps]
        return min(keys) if keys else B.NO_KEY


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, R, tables, expect, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [B.shard_range(R, r, world)[1] - B.shard_range(R, r, world)[0] for r in range(world)]
        lo, _ = B.shard_range(R, rank, world)
        res = B.run_sharded_inter(StubBatch(tables[rank]), 0.01, 0.001, expect["max_steps"], 32, lo, sizes)
        q.put((rank, res.key, res.steps_run, res.winner, res.winner_rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [
    # rank 1 flags first (step 40, local replica 3 → global 53); rank 0 flags later
    dict(tables=[{5: 70}, {3: 40, 9: 40}], max_steps=1000, key=B.encode_key(40, 53), steps=64, winner=53, wrank=1),
    # same step on both ranks → lowest global index wins (system.rs:353)
    dict(tables=[{7: 10}, {0: 10}], max_steps=-1, key=B.encode_key(10, 7), steps=32, winner=7, wrank=0),
    # nobody flags within the budget
    dict(tables=[{}, {}], max_steps=100, key=B.NO_KEY, steps=100, winner=-1, wrank=-1),
])
def test_two_rank_early_exit_over_gloo(case):
    world, R = 2, 100
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, case["tables"], case, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, key, steps, winner, wrank in out:
        assert key == case["key"] and steps == case["steps"] and winner == case["winner"] and wrank == case["wrank"]


def test_allreduce_is_identity_without_a_group():
    assert B.allreduce_min_key(B.encode_key(3, 1)) == B.encode_key(3, 1)
