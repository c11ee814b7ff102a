"""NN-scene ICP timing (C3-like): P hypotheses against a kd-tree scene of ~100k points.
    python scripts/time_nn.py [hyp] [scene_points]"""
import os, sys, json, ctypes as C, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
target = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = api.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
obj_depth = api.render_cuda(mesh, scene_pose[None], 640, 480, proj)[0]
scene_depth = wl.plane_scene_depth(obj_depth, target_valid=target) if target > 30000 else obj_depth
t0 = time.perf_counter()
scene = api.SceneNN().init_cuda(scene_depth, K)
t_build = time.perf_counter() - t0
depth = api.render_cuda_keep_in_gpu(mesh, wl.hypotheses(P, seed=1234), 640, 480, proj)
pts, offsets, counts = api.depth2cloud_batch(depth, K)
crit = api.ICPConvergenceCriteria(0.0, 0.0, 30)
res = api.icp_batch(pts, offsets, counts, scene, crit); torch.cuda.synchronize()
ts = []
for _ in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); res = api.icp_batch(pts, offsets, counts, scene, crit); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
n_pts = int(counts.sum())
r = res.cpu().numpy()
print(json.dumps({"hyp": P, "scene_points": int(scene.pcd.shape[0]), "nodes": len(scene.nodes_host), "build_s": round(t_build, 3),
                  "icp_ms": round(float(np.median(ts)), 2), "all_ms": [round(t, 1) for t in ts], "queries_per_s": round(n_pts * 31 / (np.median(ts) * 1e-3) / 1e9, 3),
                  "hyp_per_s": round(P / (np.median(ts) * 1e-3), 1), "mean_fitness": float(r[:, 17].mean())}))
