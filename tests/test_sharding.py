"""Host-side multi-rank logic on CPU: shard bounds and the final all-gather over gloo
(world_size 2 and 3, ragged batch)."""
import os
import socket

import numpy as np
import pytest

from pycllp_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for N in (0, 1, 7, 64, 4096, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(N, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)
    assert sharding.world_and_rank(None) == (1, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, m, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(N, world, rank)
        full = _fake_results(N, m, n)
        local = {k: v[lo:hi] for k, v in full.items()}
        out = sharding.allgather_results(local, N, group=True)
        ok = all(np.array_equal(out[k], full[k]) for k in full)
        q.put((rank, ok, sharding.world_and_rank(True)))
    finally:
        dist.destroy_process_group()


def _fake_results(N, m, n):
    rng = np.random.RandomState(3)
    return dict(x=rng.rand(N, n), y=rng.rand(N, m), z=rng.rand(N, n),
                status=rng.randint(0, 6, N).astype(np.int32), iters=rng.randint(0, 200, N).astype(np.int32))


@pytest.mark.parametrize("world,N", [(2, 10), (2, 7), (3, 8)])
def test_allgather_results_gloo(world, N):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, 3, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in results) == list(range(world))
    assert all(r[1] for r in results)
    assert all(r[2][0] == world for r in results)


def _nccl_worker(rank, world, port, q):
    """One rank = one GPU: solve this rank's slice through the plugin API and all-gather over NCCL."""
    import torch
    import torch.distributed as dist
    from pycllp_b200.lp import StandardLP
    from pycllp_b200.problems import random_problem
    from pycllp_b200.solvers import solver_registry
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        lp = StandardLP(*random_problem(50, 50, 0.1, 37)).to_equality_form()     # ragged over 2 ranks
        sharded = solver_registry["cl_dense_primal_normal"](rank, group=True)
        lp.init(sharded)
        lp.solve(sharded)
        whole = solver_registry["cl_dense_primal_normal"](rank)
        lp.init(whole)
        lp.solve(whole)
        ok = all(np.array_equal(getattr(sharded, k), getattr(whole, k)) for k in ("x", "y", "z", "status", "iterations"))
        q.put((rank, ok, int(sharded.x.shape[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_solve_gathers_over_nccl():
    """Two GPUs, one process each: the sharded solve + NCCL all-gather gives every rank exactly
    what a single GPU computes for the whole batch (skipped on a one-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results) and all(r[2] == 37 for r in results)
