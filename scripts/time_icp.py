"""ICP-only timing of the C2 batch through the C ABI with preallocated workspace (CUDA events, median).
    PR_LIB=<variant.so> python scripts/time_icp.py [hyp] [reps]"""
import os, sys, json, ctypes as C
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
depth = api.render_cuda_keep_in_gpu(mesh, wl.hypotheses(P, seed=1234), 640, 480, proj)
pts, offsets, counts = api.depth2cloud_batch(depth, K)
del depth
scene = api.SceneProjective().init_cuda(scene_depth, K)
L = _lib.lib()
cap = pts.shape[0]
ws_bytes = L.pr_icp_workspace_bytes(P, cap, 640 * 480)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
res = torch.empty((P, 18), dtype=torch.float32, device="cuda")
sc = scene.c()
crit = _lib.Criteria(0.0, 0.0, 30)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    rc = L.pr_icp_projective_batch(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, cap, C.byref(sc), crit,
                                   res.data_ptr(), 0, ws.data_ptr(), ws_bytes, stream)
    assert rc == 0, rc
run(); torch.cuda.synchronize()
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
n_pts = int(counts.sum())
ms = float(np.median(ts))
r = res.cpu().numpy()
good = r[:, 17] > 0.9
out = {"lib": os.path.basename(os.environ.get("PR_LIB", "default")), "impl": os.environ.get("PR_ICP_IMPL", "persistent"),
       "icp_ms": round(ms, 4), "min_ms": round(min(ts), 4), "GBs": round((12 * n_pts + 640 * 480 * 24 + 72 * P) * 31 / (ms * 1e-3) / 1e9, 1),
       "frac": round((12 * n_pts + 640 * 480 * 24 + 72 * P) * 31 / (ms * 1e-3) / 1e9 / 6501.5, 4),
       "checksum_converged": float(np.abs(r[good]).sum()), "n_converged": int(good.sum())}
print(json.dumps(out))
if len(sys.argv) > 3:
    np.save(sys.argv[3], r)
