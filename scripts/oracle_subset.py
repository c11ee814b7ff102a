"""CPU oracle results for a subset of the C2 hypotheses (test infrastructure; uses oracle/).
    python scripts/oracle_subset.py out.npy idx0,idx1,..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import binding
from pose_refine_b200 import workloads as wl
out = sys.argv[1]; idx = [int(v) for v in sys.argv[2].split(",")]
chk = binding.load("port"); chk.set_threads(1)
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = chk.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
scene_depth = chk.render(mesh, scene_pose[None], 640, 480, proj)[0]
sp = chk.scene_projective(scene_depth, K)
poses = wl.hypotheses(512, seed=1234)
res = np.zeros((len(idx), 18), np.float32)
for j, i in enumerate(idx):
    d = chk.render(mesh, poses[i:i + 1], 640, 480, proj)[0]
    res[j] = chk.icp(sp, chk.depth2cloud(d, K), 0.0, 0.0, 30)["raw"]
np.save(out, res)
