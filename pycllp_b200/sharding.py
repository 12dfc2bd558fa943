"""Sharding of a batch of LPs over ranks (one process per GPU) and the final gather.

The problems are independent and A is read-only, so the batch shards naturally
(SURVEY.md section 8(e)): rank r owns a contiguous slice, there is no communication
inside the IPM loop, and one all-gather at the end collects x, y, z, status and
iteration counts (NCCL over NVLink on GPUs; the same code runs over gloo on CPU,
which is how the tests exercise it).
"""
import os

import numpy as np


def default_device():
    """LOCAL_RANK under torchrun, else 0."""
    return int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(nproblems, world, rank):
    """Contiguous, balanced split: the first ``nproblems % world`` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(int(nproblems), world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def world_and_rank(group=None):
    """(world_size, rank) of ``group``; (1, 0) when ``group`` is None/False."""
    if group is None or group is False:
        return 1, 0
    import torch.distributed as dist
    g = None if group is True else group
    return dist.get_world_size(g), dist.get_rank(g)


def allgather_results(local, nproblems, group=None, device=None):
    """All-gather per-rank result dicts (numpy arrays keyed x, y, z, status, iters).

    Every rank passes its slice; every rank gets the full arrays back.  Slices may be
    ragged (nproblems not divisible by the world size): each is padded to the largest
    slice for the collective and trimmed afterwards.  One collective per array.
    """
    import torch
    import torch.distributed as dist
    g = None if (group is True or group is None) else group
    world, rank = dist.get_world_size(g), dist.get_rank(g)
    backend = dist.get_backend(g)
    dev = torch.device("cuda", device if device is not None else default_device()) \
        if backend == "nccl" else torch.device("cpu")
    bounds = [shard_bounds(nproblems, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in bounds)
    out = {}
    for key in ("x", "y", "z", "status", "iters"):
        arr = local.get(key)
        if arr is None:
            out[key] = None
            continue
        arr = np.ascontiguousarray(arr)
        cols = arr.shape[1:] if arr.ndim > 1 else ()
        pad = np.zeros((width,) + cols, dtype=arr.dtype)
        pad[: arr.shape[0]] = arr
        src = torch.from_numpy(pad).to(dev)
        # concatenated-along-dim-0 output form: accepted by both NCCL and gloo
        dst = torch.empty((world * width,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
        dist.all_gather_into_tensor(dst, src, group=g)
        full = dst.cpu().numpy().reshape((world, width) + tuple(src.shape[1:]))
        out[key] = np.concatenate([full[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)], axis=0)
    return out
