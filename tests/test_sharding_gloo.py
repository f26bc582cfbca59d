"""Multi-rank host logic on CPU (gloo, world_size 2): shards are disjoint, cover the batch, regenerate exactly the
inputs of the unsharded batch, and the reduced counters equal the single-process result.  The per-shard "solver" here
is the oracle port (test infrastructure); the GPU path uses the identical sharding in bench.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib
from ilqg_b200 import workloads as W
from ilqg_b200.sharding import shard_range

B, T, ITERS = 10, 40, 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(B, rank, world)
    x0, u0 = W.car_batch(count, T=T, first=first)
    out = oracle_lib.OracleLib("port", "car", 0).solve_batch(x0, u0, W.CAR_PARAMS, {"max_iter": float(ITERS)}, 1)
    t = torch.tensor([float(out["n_linesearch"].sum()), float(count)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    mx = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    q.put((rank, first, count, out["cost"].tolist(), t.tolist(), mx.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    for batch in (1, 7, 8, 262144):
        for world in (1, 2, 3, 8):
            r = [shard_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and sum(c for _, c in r) == batch
            assert all(r[k][0] + r[k][1] == r[k + 1][0] for k in range(world - 1))


def test_counter_based_inputs_are_shard_invariant():
    x0, u0 = W.car_batch(B, T=T)
    for world in (2, 3):
        for k in range(world):
            f, c = shard_range(B, k, world)
            xs, us = W.car_batch(c, T=T, first=f)
            assert np.array_equal(xs, x0[f:f + c]) and np.array_equal(us, u0[f:f + c])


def test_two_ranks_reduce_to_single_process_result():
    x0, u0 = W.car_batch(B, T=T)
    whole = oracle_lib.OracleLib("port", "car", 0).solve_batch(x0, u0, W.CAR_PARAMS, {"max_iter": float(ITERS)}, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    costs = sum((r[3] for r in res), [])
    assert np.array_equal(np.array(costs), whole["cost"])
    for r in res:
        assert r[4] == [float(whole["n_linesearch"].sum()), float(B)] and r[5] == 2.0
