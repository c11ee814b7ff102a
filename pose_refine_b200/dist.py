"""Hypothesis-batch sharding over the GPUs of one box.

The path shards by hypothesis: every hypothesis has its own pose, cloud and accumulators; the scene
and the mesh are read-only.  So there is NO data-path collective inside the iteration loop.  The only
exchanges are (1) a broadcast of the scene depth image from rank 0 (every rank then prepares the same
scene deterministically on its own GPU) and (2) a gather of the 72-byte results.  One process per GPU;
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.

The reference has none of this (no multi-GPU code at all: SURVEY.md section 2c).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of rank's share: the first n_items % world ranks get one more."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_plan(n_items, world):
    return [shard_range(n_items, r, world) for r in range(world)]


def broadcast_scene(depth, shape, dtype, src=0, device=None):
    """Rank `src` passes the scene depth image (numpy [H,W] int32/uint16); others pass None.
    Returns the image on every rank as a numpy array."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return np.ascontiguousarray(depth)
    dtype = np.dtype(dtype)
    n_bytes = int(np.prod(shape)) * dtype.itemsize
    if dist.get_rank() == src:
        t = torch.as_tensor(np.ascontiguousarray(depth, dtype=dtype).reshape(-1).view(np.uint8).copy())
    else:
        t = torch.empty(n_bytes, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)          # raw bytes: every backend moves uint8
    return t.cpu().numpy().view(dtype).reshape(shape)


def gather_results(local, counts):
    """all_gather of per-rank result blocks ([n_r, 18] float32 tensors, n_r = counts[r]) -> [sum, 18] on every rank.
    Blocks are padded to the largest shard so a single all_gather suffices."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    width = local.shape[1]
    pad = max(counts)
    buf = torch.zeros((pad, width), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * pad, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf) if hasattr(dist, "all_gather_into_tensor") and local.is_cuda else \
        out.copy_(torch.cat(_all_gather_list(buf, world)))
    return torch.cat([out[r * pad: r * pad + counts[r]] for r in range(world)])


def _all_gather_list(buf, world):
    lst = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(lst, buf)
    return lst


def refine_sharded(refiner_run, poses, criteria=None):
    """poses: the FULL [P,4,4] batch (identical on every rank). Each rank refines its contiguous shard with
    refiner_run(poses_shard) -> cuda/cpu tensor [n,18]; returns the gathered [P,18] results on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    plan = shard_plan(len(poses), world)
    b, e = plan[rank]
    local = refiner_run(poses[b:e])
    return gather_results(local, [pe - pb for pb, pe in plan])
