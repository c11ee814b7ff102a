"""One-time scene preparation, host vs device kd-tree build (fixture scene 21,960 points; composited 100k-point scene).
    python scripts/time_scene_build.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = api.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
d0 = api.render_cuda(mesh, scene_pose[None], 640, 480, proj)[0]
out = {}
for name, d in (("fixture_22k", d0), ("plane_100k", wl.plane_scene_depth(d0, target_valid=100000))):
    dd = torch.as_tensor(d).cuda()
    for kind, fn in (("device", lambda: api.SceneNN().init_cuda(dd, K)), ("host", lambda: api.SceneNN().init_host_build(d, K))):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a = time.perf_counter(); s = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - a) * 1e3)
        out[f"{name}_{kind}_ms"] = round(float(np.median(ts)), 3)
    out[f"{name}_points"] = int(s.pcd.shape[0]); out[f"{name}_nodes"] = int(len(s.nodes_host))
print(json.dumps(out))
