"""Hash-grid / tree-walk statistics of the C3 scene for a few hypotheses, before and after refinement.
   [PR_LIB=...] python scripts/nn_stats.py"""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import _lib
if os.environ.get("PR_LIB"):
    _lib.use_library(os.environ["PR_LIB"])
from pose_refine_b200 import api, workloads as wl
P = 16
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = api.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
obj = api.render_cuda(mesh, scene_pose[None], 640, 480, proj)[0]
scene_depth = wl.plane_scene_depth(obj, target_valid=100000)
ref = api.PoseRefiner(mesh, 640, 480, K, max_hyp=P)
ref.set_scene_nn(scene_depth)
poses = torch.as_tensor(wl.hypotheses(P, seed=1234).reshape(P, 16)).cuda()
crit = api.ICPConvergenceCriteria(0.0, 0.0, 30)
res = torch.empty((P, 18), dtype=torch.float32, device="cuda")
ref.run_device(poses, crit, res); torch.cuda.synchronize()
_, pts, off, cnt = ref.buffers(P)
sn = api.SceneNN().init_cuda(scene_depth, K)
L = _lib.lib()
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
snc = sn.c()
r = res.cpu().numpy()
for h in range(4):
    o, n = int(off[h]), int(cnt[h])
    cloud = pts[o:o + n].clone()
    T = torch.as_tensor(r[h, :16].reshape(4, 4)).cuda()
    moved = (cloud @ T[:3, :3].T + T[:3, 3]).contiguous()
    for name, q in (("initial", cloud), ("refined", moved)):
        stats = torch.zeros(4, dtype=torch.int64, device="cuda")
        ws, wsb = api._icp_workspace(1, n, sn)
        _lib.check(L.pr_nn_walk_stats(q.data_ptr(), n, C.byref(snc), stats.data_ptr(), ws.data_ptr(), wsb, stream), "stats")
        st = stats.cpu().numpy().astype(float) / n
        print(json.dumps({"hyp": h, "state": name, "fitness": float(r[h, 17]), "tree_nodes": round(st[0], 1), "tree_leaf_pts": round(st[1], 1),
                          "grid_answered": round(st[2], 3), "grid_pts_tested": round(st[3], 1)}))
