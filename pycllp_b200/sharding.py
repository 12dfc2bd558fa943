"""Sharding of a batch of LPs over ranks (one process per GPU) and the final gather.

The problems are independent and A is read-only, so the batch shards naturally
(SURVEY.md section 8(e)): rank r owns a contiguous slice, there is no communication
inside the IPM loop, and ONE all-gather at the end collects everything.  The unit of
the exchange is a packed per-problem record of ``2n + m + 1`` float64

    [ x (n) | y (m) | z (n) | status (int32), iterations (int32) ]

which the engine writes in place (``pycllp_b200_solve_device_packed``), so on GPUs the
collective runs straight from the kernel's output buffer over NCCL / NVLink -- no
device->host->device round trip, one collective instead of five.  The same code runs
over gloo on CPU tensors, which is how the tests exercise the host logic.
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


def record_width(m, n):
    return 2 * n + m + 1


def pack_records(res, m, n):
    """dict(x, y, z, status, iters) of numpy arrays -> (N, 2n+m+1) float64 records."""
    N = res["status"].shape[0]
    rec = np.zeros((N, record_width(m, n)))
    rec[:, :n] = res["x"]
    if res.get("y") is not None:
        rec[:, n:n + m] = res["y"]
    if res.get("z") is not None:
        rec[:, n + m:2 * n + m] = res["z"]
    tail = rec[:, 2 * n + m:].view(np.int32)          # (N, 2)
    tail[:, 0] = res["status"]
    tail[:, 1] = res["iters"]
    return rec


def unpack_records(rec, m, n, copy=True):
    """(N, 2n+m+1) float64 records -> dict(x, y, z, status, iters).  ``copy=False``: x, y, z are
    strided VIEWS of ``rec`` (no second pass over 2n+m doubles per problem on the host)."""
    rec = np.ascontiguousarray(rec)
    tail = rec[:, 2 * n + m:].view(np.int32)
    cp = (lambda a: a.copy()) if copy else (lambda a: a)
    return dict(x=cp(rec[:, :n]), y=cp(rec[:, n:n + m]), z=cp(rec[:, n + m:2 * n + m]),
                status=tail[:, 0].copy(), iters=tail[:, 1].copy())


def allgather_records(rec, nproblems, group=None):
    """ONE all-gather of the per-rank record blocks.

    ``rec``: this rank's (nlocal, width) float64 block -- a torch tensor (on the rank's GPU
    for NCCL, on the CPU for gloo) or a numpy array.  Slices may be ragged (nproblems not
    divisible by the world size): each block is padded to the largest slice for the
    collective and trimmed afterwards.  Returns the full (nproblems, width) block as a torch
    tensor on the same device.
    """
    import torch
    import torch.distributed as dist
    g = None if (group is True or group is None) else group
    world = dist.get_world_size(g)
    if isinstance(rec, np.ndarray):
        rec = torch.from_numpy(np.ascontiguousarray(rec))
    bounds = [shard_bounds(nproblems, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in bounds)
    cols = rec.shape[1]
    if rec.shape[0] != width:                         # ragged: pad (on the tensor's device)
        pad = torch.zeros((width, cols), dtype=rec.dtype, device=rec.device)
        pad[: rec.shape[0]] = rec
        rec = pad
    out = torch.empty((world * width, cols), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=g)
    if world * width == nproblems:
        return out
    full = out.view(world, width, cols)
    return torch.cat([full[r, : hi - lo] for r, (lo, hi) in enumerate(bounds)], dim=0)


def allgather_results(local, nproblems, group=None, device=None):
    """All-gather per-rank result dicts (numpy arrays keyed x, y, z, status, iters): every rank
    passes its slice and gets the full arrays back.  Host-side form of ``allgather_records``
    (the solver classes use the device form directly on the engine's output)."""
    import torch
    import torch.distributed as dist
    g = None if (group is True or group is None) else group
    m = local["y"].shape[1] if local.get("y") is not None else 0
    n = local["x"].shape[1]
    rec = torch.from_numpy(pack_records(local, m, n))
    if dist.get_backend(g) == "nccl":
        rec = rec.to(torch.device("cuda", device if device is not None else default_device()))
    full = allgather_records(rec, nproblems, g)
    out = unpack_records(full.cpu().numpy(), m, n)
    if local.get("y") is None:
        out["y"] = None
    if local.get("z") is None:
        out["z"] = None
    return out
