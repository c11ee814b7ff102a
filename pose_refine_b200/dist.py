"""Hypothesis-batch sharding over the GPUs of one box.

The path shards by hypothesis: every hypothesis has its own pose, cloud and accumulators; the scene
and the mesh are read-only.  So there is NO data-path collective inside the iteration loop.  The only
exchanges are (1) a broadcast of the scene depth image from rank 0 (every rank then prepares the same
scene deterministically on its own GPU) and (2) a gather of the 72-byte results.  One process per GPU;
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.

The reference has none of this (no multi-GPU code at all: SURVEY.md section 2c).
"""
import ctypes as C

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


class Comm:
    """The multi-GPU entry points of the C ABI (pr_comm_*, pr_broadcast_scene, pr_gather_results: NCCL bound at run time
    inside libpose_refine_b200.so) for one process per GPU.  torch.distributed is used for ONE thing: handing the
    128-byte ncclUniqueId from rank 0 to the others."""

    def __init__(self, handle, rank, world):
        self._h, self.rank, self.world = handle, rank, world

    @classmethod
    def create(cls):
        from ._lib import lib, check
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        if world == 1:
            return cls(None, 0, 1)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            check(lib().pr_nccl_unique_id(uid.data_ptr()), "pr_nccl_unique_id")
        if dist.get_backend() == "nccl":
            uid = uid.cuda()
        dist.broadcast(uid, src=0)
        uid = uid.cpu().contiguous()
        h = C.c_void_p()
        check(lib().pr_comm_create(C.byref(h), uid.data_ptr(), world, rank), "pr_comm_create")
        return cls(h, rank, world)

    def shard(self, n_items):
        from ._lib import lib, check
        b, c = C.c_size_t(), C.c_size_t()
        check(lib().pr_shard_plan(n_items, self.world, self.rank, C.byref(b), C.byref(c)), "pr_shard_plan")
        return b.value, b.value + c.value

    def broadcast_scene(self, buf_dev, root=0):
        """In place: a contiguous cuda tensor (the scene depth image) from `root` to every rank, on the current stream."""
        if self.world == 1:
            return buf_dev
        from ._lib import lib, check
        assert buf_dev.is_cuda and buf_dev.is_contiguous()
        check(lib().pr_broadcast_scene(self._h, buf_dev.data_ptr(), buf_dev.numel() * buf_dev.element_size(), root,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pr_broadcast_scene")
        return buf_dev

    def gather_results(self, local_dev, all_dev):
        """local_dev: [n_per_rank, 18] float32 on the device (every rank the same n_per_rank: pad the shards);
        all_dev: [world * n_per_rank, 18].  Starts after the work queued on the current stream, runs on the
        communicator's own stream; wait() before reading all_dev."""
        if self.world == 1:
            if local_dev.data_ptr() != all_dev.data_ptr():
                all_dev.copy_(local_dev)
            return
        from ._lib import lib, check
        assert local_dev.is_cuda and all_dev.is_cuda and local_dev.is_contiguous() and all_dev.is_contiguous()
        check(lib().pr_gather_results(self._h, local_dev.data_ptr(), local_dev.shape[0], all_dev.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pr_gather_results")

    def wait(self, host=False):
        if self.world == 1:
            return
        from ._lib import lib, check
        check(lib().pr_gather_wait(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream), int(host)), "pr_gather_wait")

    def close(self):
        if self._h:
            from ._lib import lib
            lib().pr_comm_destroy(self._h)
            self._h = None
