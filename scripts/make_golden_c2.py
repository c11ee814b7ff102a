"""Golden oracle results for the FULL C2 batch (512 hypotheses, seed 1234, criteria (0,0,30)) -> tests/golden/c2_oracle_512.npz.

    python scripts/make_golden_c2.py        (CPU only; ~1 minute on 8 cores)

The oracle's render -> depth2cloud -> ICP pipeline is run four times with different OpenMP summation orders inside
ICP_Point2Plane_cpu (1, 2, 5 and 8 threads over the points, icp.cpp:141-148): the reference's own result is only
defined up to that order, and the spread between the four runs is what a GPU result can be held to
(tests/test_gpu_parity.py::test_full_size_all_hypotheses).  Stored: results [4, 512, 18] float32, n_pts [512].
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding            # noqa: E402  (test infrastructure)
from pose_refine_b200 import workloads as wl   # noqa: E402

P = 512
port = binding.load("port")
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = port.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
port.set_threads(port.max_threads())
scene_depth = port.render(mesh, scene_pose[None], 640, 480, proj)[0]
scene = port.scene_projective(scene_depth, K)
poses = wl.hypotheses(P, seed=1234)
out, npts = [], None
for nt in (1, 2, 5, 8):
    if nt == 1:
        port.set_threads(port.max_threads())
        sec, res, npts = port.pipeline(scene, mesh, poses, 640, 480, proj, K, 0.0, 0.0, 30, schedule=1)   # hypotheses in parallel, serial sums
    else:
        port.set_threads(nt)
        sec, res, _ = port.pipeline(scene, mesh, poses, 640, 480, proj, K, 0.0, 0.0, 30, schedule=0)      # nt-thread sums
    print(f"threads {nt}: {sec:.1f} s", flush=True)
    out.append(res)
out = np.stack(out).astype(np.float32)
spread = np.abs(out[1:, :, :16] - out[0, :, :16]).max(axis=(0, 2))
conv = out[0, :, 17] > 0.9
print(f"converging (fitness > 0.9): {conv.sum()} of {P}; oracle's own spread over thread counts: "
      f"median {np.median(spread[conv]):.2e}, max {spread[conv].max():.2e} (converging), {(spread > 1e-4).sum()} of {P} beyond 1e-4")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c2_oracle_512.npz"), results=out, n_pts=npts.astype(np.int32),
                    threads=np.array([1, 2, 5, 8], np.int32))
