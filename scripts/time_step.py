"""Stage timings of the refiner's own path on C2 (CUDA events, warm, L2 flushed between repetitions):
the whole step (pr_refiner_run_device), the fused render -> cloud call and the ICP call it consists of.
    [PR_LIB=<variant.so>] python scripts/time_step.py [hyp] [reps]"""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import _lib
if os.environ.get("PR_LIB"):
    _lib.use_library(os.environ["PR_LIB"])
from pose_refine_b200 import api, workloads as wl
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
res = torch.empty((P, 18), dtype=torch.float32, device="cuda")
L = _lib.lib()
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, do_flush=True):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
out = {"lib": os.path.basename(os.environ.get("PR_LIB", "default")), "hyp": P}
out["step_ms"] = timed(lambda: ref.run_device(poses, crit, res), do_flush=False)
# the two calls of a step, on their own
verts, faces = api.mesh_index(mesh)
faces, off, cv = api.mesh_cluster(verts, faces)
verts_d, faces_d = torch.as_tensor(verts).cuda(), torch.as_tensor(faces).cuda()
off_d, cv_d = torch.as_tensor(off).cuda(), torch.as_tensor(cv).cuda()
cl = _lib.MeshClusters(off_d.shape[0] - 1, off_d.data_ptr(), cv_d.data_ptr())
cap = P * 40000
depth = torch.empty((P, 480, 640), dtype=torch.int32, device="cuda")
pts = torch.empty((cap + 8, 3), dtype=torch.float32, device="cuda")
counts = torch.empty(P, dtype=torch.int32, device="cuda"); offsets = torch.empty(P + 1, dtype=torch.int32, device="cuda")
overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
ws_bytes = L.pr_render_cloud_workspace_bytes(P, verts.shape[0], faces.shape[0], 640, 480)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
Kc = np.ascontiguousarray(K.reshape(9), np.float32); pj = np.ascontiguousarray(proj.reshape(16), np.float32)
def render_cloud(with_cloud=True):
    rc = L.pr_render_cloud_batch(verts_d.data_ptr(), verts.shape[0], faces_d.data_ptr(), faces.shape[0], poses.data_ptr(), 1, P, 640, 480,
                                 pj.ctypes.data, Kc.ctypes.data, depth.data_ptr(), pts.data_ptr() if with_cloud else None, cap, 4,
                                 counts.data_ptr(), offsets.data_ptr(), overflow.data_ptr(), C.cast(C.pointer(cl), C.c_void_p), ws.data_ptr(), ws_bytes, stream)
    assert rc == 0, rc
out["render_cloud_ms"] = timed(render_cloud)
out["render_only_ms"] = timed(lambda: render_cloud(False))
scene = api.SceneProjective().init_cuda(scene_depth, K)
sc = scene.c()
packed = torch.empty(L.pr_scene_projective_packed_bytes(640, 480), dtype=torch.uint8, device="cuda")
assert L.pr_scene_projective_pack(C.byref(sc), packed.data_ptr(), stream) == 0
iws_bytes = L.pr_icp_workspace_bytes(P, cap, 0)
iws = torch.empty(iws_bytes, dtype=torch.uint8, device="cuda")
critc = _lib.Criteria(0.0, 0.0, 30)
def icp():
    rc = L.pr_icp_projective_batch_packed(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, cap, C.byref(sc), packed.data_ptr(), critc,
                                          res.data_ptr(), 0, iws.data_ptr(), iws_bytes, stream)
    assert rc == 0, rc
out["icp_ms"] = timed(icp)
n_pts = int(counts.sum())
alg = (12 * n_pts + 640 * 480 * 24 + 72 * P) * 31
out["model_points"] = n_pts
out["icp_frac"] = alg / (out["icp_ms"] * 1e-3) / 1e9 / 6553.3
out["hyp_per_s"] = P / (out["step_ms"] * 1e-3)
out["depth_crc_sum"] = int(depth.sum().item())
print(json.dumps(out))
