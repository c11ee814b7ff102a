"""world_size-2 gloo tests (CPU) of the sharding plumbing: shard plan, scene broadcast, result gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pose_refine_b200 import dist as prd


def test_shard_plan_covers_everything_once():
    for n in (0, 1, 7, 512, 4096, 4099):
        for w in (1, 2, 3, 8):
            plan = prd.shard_plan(n, w)
            assert plan[0][0] == 0 and plan[-1][1] == n
            assert all(plan[i][1] == plan[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in plan]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(0)
        depth = rng.randint(0, 2000, size=(48, 64)).astype(np.int32)
        got = prd.broadcast_scene(depth if rank == 0 else None, (48, 64), np.int32)
        ok = np.array_equal(got, depth)
        d16 = depth.astype(np.uint16)
        got16 = prd.broadcast_scene(d16 if rank == 0 else None, (48, 64), np.uint16)
        ok = ok and got16.dtype == np.uint16 and np.array_equal(got16, d16)
        # 7 "hypotheses": fake refiner = 18 columns derived from the pose so the gather order is checkable
        poses = rng.normal(size=(7, 4, 4)).astype(np.float32)

        def fake_run(p):
            return torch.as_tensor(np.concatenate([p.reshape(len(p), 16), p.reshape(len(p), 16)[:, :2] * 2], axis=1))

        res = prd.refine_sharded(fake_run, poses)
        want = fake_run(poses)
        ok = ok and res.shape == (7, 18) and torch.equal(res, want)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
