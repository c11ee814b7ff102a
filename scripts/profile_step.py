"""One or more steps of the C2 workload through pr_refiner (device-resident), for ncu captures.
    python scripts/profile_step.py [steps] [hyp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = int(sys.argv[2]) if len(sys.argv) > 2 else 512
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = api.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
scene_depth = api.render_cuda(mesh, scene_pose[None], 640, 480, proj)[0]
ref = api.PoseRefiner(mesh, 640, 480, K, max_hyp=P)
ref.set_scene_projective(scene_depth)
poses = torch.as_tensor(wl.hypotheses(P, seed=1234).reshape(P, 16)).cuda()
crit = api.ICPConvergenceCriteria(0.0, 0.0, 30)
for _ in range(steps):
    res = ref.run_device(poses, crit)
torch.cuda.synchronize()
print("done", float(res[:, 17].mean()))
