"""Time the reference's OWN CUDA build (oracle/_ref/libpose_refine_refcuda.so, built by `make -C oracle refcuda`)
on the C2 workload: render_cuda_keep_in_gpu -> P x depth2cloud_cuda -> P x ICP_Point2Plane_cuda, criteria (0,0,30),
serially and from several host threads (the reference README's recipe), and check its poses against ours.
    python scripts/time_ref_cuda.py [hyp] [threads,threads,...]       -> one JSON line
Measurement infrastructure only: nothing here is on the product path."""
import os, sys, json, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    thread_list = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 8, 16]
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libpose_refine_refcuda.so")
    if not os.path.exists(lib_path):
        print(json.dumps({"ref_cuda": None, "why": "oracle/_ref/libpose_refine_refcuda.so not built"})); return
    import torch
    from pose_refine_b200 import api, workloads as wl
    L = C.CDLL(lib_path)
    L.refcuda_pipeline.restype = C.c_int
    mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
    tris = np.ascontiguousarray(mesh, dtype=np.float32).reshape(-1, 9)
    K = np.ascontiguousarray(wl.LINEMOD_K, dtype=np.float32)
    proj = np.ascontiguousarray(api.compute_proj(K, 640, 480), dtype=np.float32)
    _, scene_pose = wl.fixture_poses()
    scene_depth = np.ascontiguousarray(api.render_cuda(mesh, scene_pose[None], 640, 480, proj)[0], dtype=np.int32)
    poses = np.ascontiguousarray(wl.hypotheses(P, seed=1234), dtype=np.float32).reshape(P, 16)

    # ours, same inputs (for the pose cross-check and the ratio)
    ref = api.PoseRefiner(mesh, 640, 480, K, max_hyp=P)
    ref.set_scene_projective(scene_depth)
    crit = api.ICPConvergenceCriteria(0.0, 0.0, 30)
    ours = ref.run(poses.reshape(P, 4, 4), crit)
    torch.cuda.synchronize()
    t = []
    for _ in range(5):
        a = time.perf_counter(); ours = ref.run(poses.reshape(P, 4, 4), crit); torch.cuda.synchronize(); t.append(time.perf_counter() - a)
    ours_s = float(np.median(t))
    ours = np.asarray(ours, dtype=np.float32).reshape(P, 18).copy()
    ref.close()

    out = {"workload": f"C2: {P} hypotheses, obj_06, 640x480, projective, criteria (0,0,30)", "ours_wall_ms": round(ours_s * 1e3, 3),
           "ours_hyp_per_s": round(P / ours_s, 1), "runs": []}
    p_f = lambda a: a.ctypes.data_as(C.c_void_p)
    for th in thread_list:
        res = np.zeros((P, 18), dtype=np.float32)
        secs = (C.c_double * 3)()
        npts = C.c_long(0)
        best, rc = None, 0
        for rep in range(2):        # first repetition warms up thrust / the context
            a = time.perf_counter()
            rc = L.refcuda_pipeline(p_f(tris), C.c_size_t(tris.shape[0]), p_f(poses), C.c_size_t(P), 640, 480, p_f(K), p_f(proj),
                                    p_f(scene_depth), C.c_float(0.0), C.c_float(0.0), 30, th, p_f(res), secs, C.byref(npts))
            wall = time.perf_counter() - a
            if rc != 0: break
            tot = secs[0] + secs[1] + secs[2]
            best = {"threads": th, "render_ms": round(secs[0] * 1e3, 3), "depth2cloud_ms": round(secs[1] * 1e3, 3),
                    "icp_ms": round(secs[2] * 1e3, 3), "pipeline_ms": round(tot * 1e3, 3), "hyp_per_s": round(P / tot, 1),
                    "wall_ms_incl_scene_init": round(wall * 1e3, 1), "model_points": int(npts.value)}
        if best is None:
            out["runs"].append({"threads": th, "error": int(rc)}); continue
        good = (res[:, 17] > 0.9) & (ours[:, 17] > 0.9)
        best["pose_max_abs_diff_vs_ours_converged"] = float(np.abs(res[good, :16] - ours[good, :16]).max()) if good.any() else None
        best["n_converged_both"] = int(good.sum())
        out["runs"].append(best)
    ok = [r for r in out["runs"] if "hyp_per_s" in r]
    if ok:
        top = max(ok, key=lambda r: r["hyp_per_s"])
        out["ref_cuda_best_hyp_per_s"] = top["hyp_per_s"]; out["ref_cuda_best_threads"] = top["threads"]
        out["speedup_ours_over_ref_cuda"] = round(out["ours_hyp_per_s"] / top["hyp_per_s"], 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
