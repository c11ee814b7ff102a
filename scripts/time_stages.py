"""Stage timings of the C2 workload (CUDA events, warm): render, depth2cloud, ICP; and the whole refiner step.
    python scripts/time_stages.py [hyp] [reps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl, _lib

P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.LINEMOD_K
proj = api.compute_proj(K, 640, 480)
_, scene_pose = wl.fixture_poses()
scene_depth = api.render_cuda(mesh, scene_pose[None], 640, 480, proj)[0]
ref = api.PoseRefiner(mesh, 640, 480, K, max_hyp=P)
ref.set_scene_projective(scene_depth)
poses = torch.as_tensor(wl.hypotheses(P, seed=1234).reshape(P, 16)).cuda()
crit = api.ICPConvergenceCriteria(0.0, 0.0, 30)
res = ref.run_device(poses, crit)
torch.cuda.synchronize()
depth, pts, offsets, counts = ref.buffers(P)
pts, offsets, counts = pts.clone(), offsets.clone(), counts.clone()
n_pts = int(counts.sum())
scene = api.SceneProjective().init_cuda(scene_depth, K)
tris_dev = torch.as_tensor(mesh).cuda()
L = _lib.lib()

def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out.setdefault("raw", {})[len(out["raw"])] = [round(t, 3) for t in ts]
    return float(np.median(ts))

out = {}
out = {"impl": os.environ.get("PR_ICP_IMPL", "persistent"), "hyp": P, "model_points": n_pts}
out["step_ms"] = timed(lambda: ref.run_device(poses, crit))
out["icp_ms"] = timed(lambda: api.icp_batch(pts, offsets, counts, scene, crit))
out["render_ms"] = timed(lambda: api.render_cuda_keep_in_gpu(tris_dev, poses, 640, 480, proj))
bytes_pass = 12 * n_pts + 640 * 480 * 24 + 72 * P
out["icp_GBs_algorithmic"] = bytes_pass * 31 / (out["icp_ms"] * 1e-3) / 1e9
out["icp_frac_of_6501.5"] = out["icp_GBs_algorithmic"] / 6501.5
out["hyp_per_s"] = P / (out["step_ms"] * 1e-3)
r = api.icp_batch(pts, offsets, counts, scene, crit).cpu().numpy()
out["result_checksum"] = float(np.abs(r).sum())
out["mean_fitness"] = float(r[:, 17].mean())
print(json.dumps(out))
