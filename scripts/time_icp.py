"""ICP-only timing of the C2 batch through the C ABI with preallocated workspace (CUDA events, median).
    PR_LIB=<variant.so> python scripts/time_icp.py [hyp] [reps] [clusters e.g. 1,2,4,8]
Clouds come from the fused render -> cloud path (tile order), as in the refiner."""
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
clusters = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
BIG = os.environ.get("RES") == "720"          # C4: 1280x720, ~88k points per hypothesis
W_, H_ = (1280, 720) if BIG else (640, 480)
K = wl.k_1280x720() if BIG else wl.LINEMOD_K
proj = api.compute_proj(K, W_, H_)
_, scene_pose = wl.fixture_poses()
scene_depth = api.render_cuda(mesh, scene_pose[None], W_, H_, proj)[0]
verts, faces = api.mesh_index(mesh)
faces, off, cv = api.mesh_cluster(verts, faces)
depth, pts, offsets, counts = api.render_cloud_batch(verts, faces, wl.hypotheses(P, seed=4321 if BIG else 1234), W_, H_, proj, K,
                                                     capacity_points=P * (140000 if BIG else 40000), clusters=(off, cv))
del depth
scene = api.SceneProjective().init_cuda(scene_depth, K, W_, H_)
L = _lib.lib()
cap = int(offsets[P].item())
ws_bytes = L.pr_icp_workspace_bytes(P, cap, W_ * H_)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
res = torch.empty((P, 18), dtype=torch.float32, device="cuda")
sc = scene.c()
packed = torch.empty(L.pr_scene_projective_packed_bytes(W_, H_), dtype=torch.uint8, device="cuda")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
assert L.pr_scene_projective_pack(C.byref(sc), packed.data_ptr(), stream) == 0
crit = _lib.Criteria(0.0, 0.0, 30)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    rc = L.pr_icp_projective_batch_packed(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, cap, C.byref(sc), packed.data_ptr(), crit,
                                          res.data_ptr(), 0, ws.data_ptr(), ws_bytes, stream)
    assert rc == 0, rc
n_pts = int(counts.sum())
ref = None
for c in clusters:
    if c: os.environ["PR_HYP_CLUSTER"] = str(c)
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()          # the clouds (135 MB) start in HBM, as they do after the rasteriser wrote 629 MB of depth
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    r = res.cpu().numpy()
    if ref is None: ref = r.copy()
    good = ref[:, 17] > 0.9
    alg = (12 * n_pts + W_ * H_ * 24 + 72 * P) * 31
    print(json.dumps({"lib": os.path.basename(os.environ.get("PR_LIB", "default")), "cluster": c, "icp_ms": round(ms, 4), "min_ms": round(min(ts), 4),
                      "GBs": round(alg / (ms * 1e-3) / 1e9, 1), "frac": round(alg / (ms * 1e-3) / 1e9 / 6553.3, 4),
                      "max_dev_vs_first": float(np.abs(r[good, :16] - ref[good, :16]).max()), "n_converged": int(good.sum())}), flush=True)
if len(sys.argv) > 4:
    np.save(sys.argv[4], r)
