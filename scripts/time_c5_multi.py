"""C5 render-only: k back-to-back calls between one pair of events (separates a fixed start-up latency from GPU time)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
W, H = 640, 480
proj = api.compute_proj(wl.LINEMOD_K, W, H)
tris5 = wl.uv_sphere()
verts5, faces5 = api.mesh_index(tris5)
faces5, off5, cv5 = api.mesh_cluster(verts5, faces5)
p5 = torch.as_tensor(wl.shoemake_poses(8192, seed=99)[:n].reshape(-1, 16)).cuda()
v5, f5 = torch.as_tensor(verts5).cuda(), torch.as_tensor(faces5).cuda()
cl5 = (torch.as_tensor(off5).cuda(), torch.as_tensor(cv5).cuda())
depth5 = torch.empty((n, H, W), dtype=torch.int32, device="cuda")
L = _lib.lib()
ws5 = torch.empty(L.pr_render_cloud_workspace_bytes(n, verts5.shape[0], faces5.shape[0], W, H), dtype=torch.uint8, device="cuda")
for k in (1, 1, 2, 4, 8, 1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(k):
        api.render_clustered_keep_in_gpu(v5, f5, p5, W, H, proj, cl5, out=depth5, ws=ws5)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(json.dumps({"n": n, "k": k, "ms": a.elapsed_time(b), "host_issue_ms": (t1 - t0) * 1e3}))
