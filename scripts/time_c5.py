"""C5 render-only timing for n poses (debug aid): python scripts/time_c5.py [n] [first]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pose_refine_b200 import api, workloads as wl, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W, H = 640, 480
proj = api.compute_proj(wl.LINEMOD_K, W, H)
tris5 = wl.uv_sphere()
verts5, faces5 = api.mesh_index(tris5)
faces5, off5, cv5 = api.mesh_cluster(verts5, faces5)
p5 = torch.as_tensor(wl.shoemake_poses(8192, seed=99)[first:first + n].reshape(-1, 16)).cuda()
v5, f5 = torch.as_tensor(verts5).cuda(), torch.as_tensor(faces5).cuda()
cl5 = (torch.as_tensor(off5).cuda(), torch.as_tensor(cv5).cuda())
depth5 = torch.empty((n, H, W), dtype=torch.int32, device="cuda")
L = _lib.lib()
ws5 = torch.empty(L.pr_render_cloud_workspace_bytes(n, verts5.shape[0], faces5.shape[0], W, H), dtype=torch.uint8, device="cuda")
for rep in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); api.render_clustered_keep_in_gpu(v5, f5, p5, W, H, proj, cl5, out=depth5, ws=ws5); b.record(); torch.cuda.synchronize()
    print(json.dumps({"n": n, "first": first, "ms": a.elapsed_time(b), "valid_px_per_pose": float((depth5 > 0).sum()) / n}))
