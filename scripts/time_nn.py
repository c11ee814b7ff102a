"""C3 timing: 512 hypotheses against the ~100k-point kd-tree scene through PoseRefiner (device-resident poses).
    [PR_LIB=<variant.so>] python scripts/time_nn.py [hyp] [reps] [clusters e.g. 2,8]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import _lib
if os.environ.get("PR_LIB"):
    _lib.use_library(os.environ["PR_LIB"])
from pose_refine_b200 import api, workloads as wl
P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
clusters = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
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
for c in clusters:
    if c: os.environ["PR_NN_CLUSTER"] = str(c)
    ref.run_device(poses, crit, res); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ref.run_device(poses, crit, res); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    r = res.cpu().numpy()
    print(json.dumps({"lib": os.path.basename(os.environ.get("PR_LIB", "default")), "cluster": c, "step_ms": round(float(np.median(ts)), 2),
                      "hyp_per_s": round(P / np.median(ts) * 1e3, 1), "mean_fitness": float(r[:, 17].mean()), "checksum": float(np.abs(r[:, :16]).sum())}), flush=True)
