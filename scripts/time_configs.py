"""Whole-path timings of the other BASELINE.json configurations on one GPU through PoseRefiner (device-resident poses):
C4 shard (512 hypotheses at 1280x720, projective) and C5 (render-only, 49,920-triangle sphere, P poses, 640x480).
    python scripts/time_configs.py [c4_hyp] [c5_poses]          -> one JSON line"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl

P4 = int(sys.argv[1]) if len(sys.argv) > 1 else 512
P5 = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
out = {}


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


# ---- C4 shard: what one of 8 GPUs does for 4096 hypotheses at 1280x720
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
K = wl.k_1280x720(); W, H = 1280, 720
proj = api.compute_proj(K, W, H)
_, scene_pose = wl.fixture_poses()
scene_depth = api.render_cuda(mesh, scene_pose[None], W, H, proj)[0]
ref = api.PoseRefiner(mesh, W, H, K, max_hyp=P4)
ref.set_scene_projective(scene_depth)
poses = torch.as_tensor(wl.hypotheses(P4, seed=4321, scene_pose=scene_pose).reshape(P4, 16)).cuda()
crit = api.ICPConvergenceCriteria(0.0, 0.0, 30)
ms = timed(lambda: ref.run_device(poses, crit))
res = ref.run_device(poses, crit); torch.cuda.synchronize()
_, _, _, counts = ref.buffers(P4)
n_pts = int(counts.sum())
out["c4_shard"] = {"hyp": P4, "size": "1280x720", "model_points": n_pts, "step_ms": round(ms, 3), "hyp_per_s": round(P4 / ms * 1e3, 1),
                   "icp_stream_GBs_if_all_icp": round((12 * n_pts + W * H * 24) * 31 / (ms * 1e-3) / 1e9, 1),
                   "mean_fitness": float(res[:, 17].mean())}
ref.close(); del ref
torch.cuda.empty_cache()

# ---- C5: render-only
tris = wl.uv_sphere()
K5 = wl.LINEMOD_K
proj5 = api.compute_proj(K5, 640, 480)
p5 = torch.as_tensor(wl.shoemake_poses(P5, seed=99).reshape(P5, 16)).cuda()
import ctypes as C
from pose_refine_b200 import _lib
L = _lib.lib()
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
tris_d = torch.as_tensor(tris).cuda()
n_tris = int(tris.shape[0])
depth = torch.empty((P5, 480, 640), dtype=torch.int32, device="cuda")
ws_bytes = L.pr_render_workspace_bytes(P5, n_tris, 640, 480)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
proj_c = np.ascontiguousarray(proj5, np.float32).reshape(16)


def render_soup():      # pr_render_batch with preallocated output and workspace, poses resident
    _lib.check(L.pr_render_batch(tris_d.data_ptr(), n_tris, p5.data_ptr(), 1, P5, 640, 480, proj_c.ctypes.data, _lib.Roi(0, 0, 0, 0),
                                 depth.data_ptr(), ws.data_ptr(), ws_bytes, stream), "pr_render_batch")


ms_soup = timed(render_soup)
# clustered: indexed mesh, Morton-ordered 64-triangle clusters (device-resident mesh, reused output / workspace)
verts, faces = api.mesh_index(tris)
cf, off, cv = api.mesh_cluster(verts, faces)
v_d, f_d = torch.as_tensor(verts).cuda(), torch.as_tensor(cf).cuda()
cl_d = (torch.as_tensor(off).cuda(), torch.as_tensor(cv).cuda())
ws_c = torch.empty(L.pr_render_cloud_workspace_bytes(P5, verts.shape[0], n_tris, 640, 480), dtype=torch.uint8, device="cuda")
depth_c = torch.empty_like(depth)
ms_cl = timed(lambda: api.render_clustered_keep_in_gpu(v_d, f_d, p5, 640, 480, proj5, cl_d, out=depth_c, ws=ws_c))
assert torch.equal(depth, depth_c)
out["c5_render"] = {"poses": P5, "tris": n_tris, "render_ms": round(ms_soup, 3), "poses_per_s": round(P5 / ms_soup * 1e3, 1),
                    "clustered_render_ms": round(ms_cl, 3), "clustered_poses_per_s": round(P5 / ms_cl * 1e3, 1),
                    "depth_write_GBs": round(P5 * 640 * 480 * 4 / (ms_soup * 1e-3) / 1e9, 1),
                    "valid_px_per_pose": round(float((depth > 0).sum()) / P5, 1)}
print(json.dumps(out))
