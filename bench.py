#!/usr/bin/env python
"""bench.py -- pose-hypotheses/sec for (render + 30-iteration projective ICP), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hyp P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path over one batch of P hypotheses (P per GPU: weak
scaling): render P poses of obj_06 (31,468 triangles) at 640x480 -> depth2cloud -> 30-iteration
point-to-plane ICP (criteria (0,0,30): 31 reduction passes) against a projective scene.  Workload
= BASELINE.json configs[1] ("C2"), synthetic inputs of SURVEY.md section 8(d).

One JSON line on stdout (rank 0):
  value   whole-job hypotheses/s with poses resident in HBM (pr_refiner_run_device), CUDA events
  e2e     the same through the host-buffer entry point pr_refiner_run: pinned-host poses H2D and
          results D2H inside the timed region
  roofline    the ICP pass kernel: algorithmic bytes per launch / mean launch time vs measured HBM peak
  cpu_baseline  the CPU oracle (reference build when present) on a bounded sample, same run
  ref_cuda_build  the reference's own CUDA path (its .cu files compiled unmodified for sm_100, oracle/_ref) on a bounded
          sample of the same workload on this GPU, in a child process -- the "reference CUDA build" of north_star's 10x target

`--impl reference` times the reference's own CPU path on the host cores instead (oracle/_ref when
it was built, else the oracle port) -- with cpu_baseline and ref_cuda_build the only places bench.py executes oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pose_refine_b200 import workloads as wl  # noqa: E402

W, H = 640, 480
ITERS = 30
METRIC = "pose-hypotheses/sec (render+30-iter ICP)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def load_inputs(n_hyp, seed):
    mesh = wl.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "obj_06_mesh.npz"))
    _, scene_pose = wl.fixture_poses()
    poses = wl.hypotheses(n_hyp, seed=seed, scene_pose=scene_pose)
    return mesh, scene_pose, poses


def cpu_pipeline(kind_pref, mesh, scene_pose, poses, threads=None, target_s=12.0, sample=None):
    """Times the CPU pipeline (render_cpu -> depth2cloud_cpu -> ICP_Point2Plane_cpu) on a bounded sample."""
    from oracle import binding
    kind = "reference" if (kind_pref == "reference" and binding.available("reference")) else "port"
    if not binding.available(kind):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    chk = binding.load(kind)
    cores = threads or os.cpu_count() or 1
    chk.set_threads(cores)
    K = wl.LINEMOD_K
    proj = chk.compute_proj(K, W, H)
    scene_depth = chk.render(mesh, scene_pose[None], W, H, proj)[0]
    scene = chk.scene_projective(scene_depth, K)
    if sample is None:
        n0 = min(len(poses), max(2 * cores, 8))
        t0, _, _ = chk.pipeline(scene, mesh, poses[:n0], W, H, proj, K, 0.0, 0.0, ITERS, schedule=1)
        sample = int(min(len(poses), max(n0, n0 * target_s / max(t0, 1e-3))))
    best = None
    for schedule in (1, 0):   # outer-parallel over hypotheses, then as shipped (inner OpenMP); keep the faster
        n = sample if schedule == 1 else min(sample, 4 * cores)
        t, _, _ = chk.pipeline(scene, mesh, poses[:n], W, H, proj, K, 0.0, 0.0, ITERS, schedule=schedule)
        rate = n / t
        if best is None or rate > best["value"]:
            best = {"value": rate, "unit": "hypotheses/s", "cores": cores, "kind": kind,
                    "sample": f"{n} of {len(poses)} hypotheses, schedule={'omp over hypotheses' if schedule else 'inner omp as shipped'}, {t:.2f} s"}
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh, scene_pose, poses = load_inputs(args.hyp, 1234)
    from oracle import binding
    kind = "reference" if binding.available("reference") else "port"
    chk = binding.load(kind)
    cores = os.cpu_count() or 1
    chk.set_threads(cores)
    K = wl.LINEMOD_K
    proj = chk.compute_proj(K, W, H)
    scene_depth = chk.render(mesh, scene_pose[None], W, H, proj)[0]
    scene = chk.scene_projective(scene_depth, K)
    n = min(len(poses), max(8 * cores, 64))     # bounded sample per step
    times = []
    for i in range(args.warmup + args.steps):
        t, _, _ = chk.pipeline(scene, mesh, poses[:n], W, H, proj, K, 0.0, 0.0, ITERS, schedule=1)
        if i >= args.warmup:
            times.append(t)
    total = sum(times)
    value = n * len(times) / total
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "hypotheses/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2: {args.hyp} pose hypotheses, obj_06 (31,468 tris), 640x480, projective ICP, criteria (0,0,30)",
                   "step": f"bounded sample: {n} hypotheses per step on the CPU"},
        "cpu_baseline": {"value": value, "unit": "hypotheses/s", "cores": cores, "kind": kind,
                         "sample": f"{n} hypotheses per step, omp parallel over hypotheses"},
        "e2e": {"value": value, "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pose_refine_b200 import api, dist as prd, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- pose_refine_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    P = args.hyp
    mesh, scene_pose, poses = load_inputs(P, 1234 + rank)     # weak scaling: P hypotheses per GPU, different per rank
    K = wl.LINEMOD_K
    proj = api.compute_proj(K, W, H)

    # one-time setup (not timed; reported): mesh upload, scene render on rank 0, NCCL broadcast, scene preparation
    t_setup = time.perf_counter()
    scene_depth = api.render_cuda(mesh, scene_pose[None], W, H, proj)[0] if rank == 0 else None
    scene_depth = prd.broadcast_scene(scene_depth, (H, W), np.int32, src=0, device=torch.device("cuda", local_rank) if world > 1 else None)
    ref = api.PoseRefiner(mesh, W, H, K, max_hyp=P)
    ref.set_scene_projective(scene_depth)
    torch.cuda.synchronize()
    setup_ms = 1e3 * (time.perf_counter() - t_setup)

    crit = api.ICPConvergenceCriteria(0.0, 0.0, ITERS)
    poses_dev = torch.as_tensor(poses.reshape(P, 16)).cuda()
    results_dev = torch.empty((P, 18), dtype=torch.float32, device="cuda")
    poses_pin = torch.as_tensor(poses.reshape(P, 16)).pin_memory()
    results_pin = torch.empty((P, 18), dtype=torch.float32).pin_memory()
    plan = prd.shard_plan(P * world, world)
    shard_sizes = [e - b for b, e in plan]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        ref.run_device(poses_dev, crit, results_dev)
        if world > 1:
            prd.gather_results(results_dev, shard_sizes)

    def step_host():
        ref.run(poses_pin, crit, results_pin)
        if world > 1:
            prd.gather_results(torch.as_tensor(results_pin).cuda(), shard_sizes)

    # ---- device-resident throughput ("value") ---------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib().pr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = _lib.lib().pr_launch_count() - launches0
    dev_ms = e0.elapsed_time(e1)

    # ---- end to end through the host-buffer entry point -------------------------------------------
    for _ in range(args.warmup):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    host_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, host_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, host_ms = float(t[0]), float(t[1])

    # ---- per-stage timing + roofline of the ICP kernel (rank 0, same inputs, live CUDA events on the launch stream).
    # Stages are called through the C ABI with buffers allocated once, exactly as pr_refiner does internally.
    roofline = stage = None
    cpu = None
    ref_cuda = None
    if rank == 0:
        import ctypes as C
        L = _lib.lib()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        n_tris = len(mesh)
        tris_dev = torch.as_tensor(mesh).cuda()
        depth = torch.empty((P, H, W), dtype=torch.int32, device="cuda")
        ws_r_bytes = L.pr_render_workspace_bytes(P, n_tris, W, H)
        ws_r = torch.empty(ws_r_bytes, dtype=torch.uint8, device="cuda")
        ws_c_bytes = L.pr_depth2cloud_workspace_bytes(P, W, H)
        ws_c = torch.empty(max(ws_c_bytes, 256), dtype=torch.uint8, device="cuda")
        counts = torch.empty(P, dtype=torch.int32, device="cuda")
        offsets = torch.empty(P + 1, dtype=torch.int32, device="cuda")
        proj_c = np.ascontiguousarray(proj, np.float32).reshape(16)
        K_c = np.ascontiguousarray(K, np.float32).reshape(9)

        def render():
            _lib.check(L.pr_render_batch(tris_dev.data_ptr(), n_tris, poses_dev.data_ptr(), 1, P, W, H, proj_c.ctypes.data,
                                         _lib.Roi(0, 0, 0, 0), depth.data_ptr(), ws_r.data_ptr(), ws_r_bytes, stream), "pr_render_batch")

        def cloud_count():
            _lib.check(L.pr_depth2cloud_count(depth.data_ptr(), 1, P, W, H, 1, 4, 0, counts.data_ptr(), offsets.data_ptr(), None,
                                              ws_c.data_ptr(), ws_c_bytes, stream), "pr_depth2cloud_count")

        render(); cloud_count()
        cap = int(offsets[P].item())
        n_pts = int(counts.sum().item())
        pts = torch.empty((cap + 8, 3), dtype=torch.float32, device="cuda")

        def cloud():
            cloud_count()
            _lib.check(L.pr_depth2cloud_fill(depth.data_ptr(), 1, P, W, H, K_c.ctypes.data, 1, 0, 0, offsets.data_ptr(), pts.data_ptr(),
                                             cap, ws_c.data_ptr(), ws_c_bytes, stream), "pr_depth2cloud_fill")

        scene = api.SceneProjective().init_cuda(scene_depth, K)
        sc = scene.c()
        ws_i_bytes = L.pr_icp_workspace_bytes(P, cap, W * H)
        ws_i = torch.empty(ws_i_bytes, dtype=torch.uint8, device="cuda")
        res_dev = torch.empty((P, 18), dtype=torch.float32, device="cuda")

        def icp():
            _lib.check(L.pr_icp_projective_batch(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, cap, C.byref(sc), crit.c(),
                                                 res_dev.data_ptr(), 0, ws_i.data_ptr(), ws_i_bytes, stream), "pr_icp_projective_batch")

        def timed(fn, reps=10):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.median(ts))

        ms_render, ms_cloud, ms_icp = timed(render), timed(cloud), timed(icp)
        passes = ITERS + 1
        # per pass 12 B per model point + the scene once (W*H*24 B) + 72 B per hypothesis; the persistent
        # kernel runs all 31 passes in ONE launch, so algorithmic bytes per launch = 31 x that.
        bytes_per_pass = 12 * n_pts + W * H * 24 + 72 * P
        bytes_per_launch = passes * bytes_per_pass
        peak, peak_src = measured_peaks()
        achieved = bytes_per_launch / (ms_icp * 1e-3) / 1e9
        traffic = None     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture
        tp = os.path.join(ROOT, "profiles", "r01_icp_traffic.json")
        if os.path.exists(tp) and P == 512:
            with open(tp) as f:
                traffic = json.load(f)["traffic_bytes_per_launch"]
        roofline = {"bound": "hbm", "kernel": "icp_persistent_kernel<PackedScene> (one launch = 31 passes over all hypotheses)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": ms_icp,
                    "note": "launch_ms = CUDA-event time of pr_icp_projective_batch (scene pack 5 us + plan 13 us + the persistent kernel)"}
        stage = {"render_ms": ms_render, "depth2cloud_ms": ms_cloud, "icp_ms": ms_icp, "model_points": n_pts,
                 "setup_ms_one_time": setup_ms}
        del depth, pts, ws_r, ws_i
        if world == 1 and not args.no_cpu:
            cpu = cpu_pipeline("reference", mesh, scene_pose, poses, target_s=args.cpu_seconds)
            ref_cuda = ref_cuda_build(min(P, 256))

    if rank == 0:
        total_hyp = P * world * args.steps
        out = {
            "metric": METRIC, "value": total_hyp / (dev_ms * 1e-3), "unit": "hypotheses/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: {P} pose hypotheses per GPU, obj_06 (31,468 tris), 640x480, projective ICP, criteria (0,0,30) = 31 passes",
                       "l2": "per-step working set (depth batch 629 MB + clouds ~160 MB) is larger than the 126 MB L2; no explicit flush",
                       "parallelism": f"hypothesis shards x{world}, NCCL scene broadcast + result all-gather" if world > 1 else "single GPU"},
            "e2e": {"value": total_hyp / (host_ms * 1e-3), "unit": "hypotheses/s", "ms_per_step": host_ms / args.steps,
                    "h2d_bytes_per_step": P * 64, "d2h_bytes_per_step": P * 72 + 4},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roofline, "stages": stage,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        if ref_cuda is not None:
            out["ref_cuda_build"] = ref_cuda
        print(json.dumps(out), flush=True)
    ref.close()
    if world > 1:
        dist.destroy_process_group()


def ref_cuda_build(n_hyp):
    """The reference's own CUDA path (render_cuda_keep_in_gpu -> depth2cloud_cuda -> ICP_Point2Plane_cuda per hypothesis,
    oracle/_ref/libpose_refine_refcuda.so = its .cu files compiled unmodified for sm_100) timed on this GPU on a bounded
    sample of the same workload, in a child process (scripts/time_ref_cuda.py).  Reported next to cpu_baseline; it is the
    "reference CUDA build" of north_star's 10x target.  None when the library was not built."""
    import subprocess
    script = os.path.join(ROOT, "scripts", "time_ref_cuda.py")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpose_refine_refcuda.so")):
        return None
    try:
        r = subprocess.run([sys.executable, script, str(n_hyp), "1,8"], capture_output=True, text=True, timeout=240)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        runs = [x for x in d.get("runs", []) if "hyp_per_s" in x]
        if not runs:
            return {"unavailable": "reference CUDA build failed to run", "detail": r.stderr[-200:]}
        top = max(runs, key=lambda x: x["hyp_per_s"])
        return {"value": top["hyp_per_s"], "unit": "hypotheses/s", "host_threads": top["threads"],
                "serial_value": next((x["hyp_per_s"] for x in runs if x["threads"] == 1), None),
                "sample": f"{n_hyp} hypotheses of the same workload, best of 1 / 8 host threads (reference README recipe)",
                "pose_max_abs_diff_vs_ours_converged": top.get("pose_max_abs_diff_vs_ours_converged")}
    except Exception as e:      # a measurement aid must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hyp", type=int, default=512, help="hypotheses per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
