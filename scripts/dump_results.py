"""Run the C2 batch ICP with the current PR_ICP_IMPL and save per-hypothesis results + timing.
    python scripts/dump_results.py out.npy [hyp]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl
out = sys.argv[1]
P = int(sys.argv[2]) if len(sys.argv) > 2 else 512
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = api.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
scene_depth = api.render_cuda(mesh, scene_pose[None], 640, 480, proj)[0]
poses = wl.hypotheses(P, seed=1234)
depth = api.render_cuda_keep_in_gpu(mesh, poses, 640, 480, proj)
pts, offsets, counts = api.depth2cloud_batch(depth, K)
scene = api.SceneProjective().init_cuda(scene_depth, K)
crit = api.ICPConvergenceCriteria(0.0, 0.0, 30)
res = api.icp_batch(pts, offsets, counts, scene, crit)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = api.icp_batch(pts, offsets, counts, scene, crit)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
np.save(out, res.cpu().numpy())
print(json.dumps({"impl": os.environ.get("PR_ICP_IMPL", "persistent"), "icp_ms_wall": [round(t * 1e3, 3) for t in ts], "cap": int(pts.shape[0])}))
