#!/usr/bin/env python
"""bench.py -- pose-hypotheses/sec for (render + 30-iteration projective ICP), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hyp P] [--no-cpu] [--no-configs]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path over one batch of P hypotheses (P per GPU: weak scaling): render P poses
of obj_06 (31,468 triangles) at 640x480 -> depth2cloud -> 30-iteration point-to-plane ICP (criteria (0,0,30): 31
reduction passes) against a projective scene.  Workload = BASELINE.json configs[1] ("C2"), synthetic inputs of
SURVEY.md section 8(d).

One JSON line on stdout (rank 0):
  value     whole-job hypotheses/s with poses resident in HBM (pr_refiner_run_device), CUDA events, max over ranks
  e2e       the same through the host-buffer entry point pr_refiner_run: pinned-host poses H2D and results D2H inside
  roofline  the ICP kernel of the step (icp_hyp_kernel): algorithmic bytes per launch / mean launch time (CUDA events
            around pr_icp_projective_batch_packed on the refiner's own clouds) vs the measured HBM peak
  stages    CUDA-event times of the two calls a step consists of (fused render->cloud, ICP)
  configs   the other BASELINE.json configurations, each with its own roofline block:
              C3  512 hypotheses per GPU against a ~100k-point kd-tree scene (weak)
              C4  4096 hypotheses FIXED at 1280x720, sharded over the N ranks (strong)
              C5  8192 poses FIXED of the 49,920-triangle sphere, render only, sharded over the N ranks (strong)
  cpu_baseline    the reference's CPU path (oracle/_ref, else the oracle port) on a bounded sample, same run (N = 1)
  ref_cuda_build  the reference's own CUDA path (its .cu files compiled unmodified for sm_100, oracle/_ref) on a
                  bounded sample of the same workload on this GPU, in a child process (N = 1)

Multi-GPU goes through the C ABI (pr_comm_*, pr_broadcast_scene, pr_gather_results: NCCL inside the library); the
all-gather of a step's results runs on the communicator's own stream and overlaps the next step.

`--impl reference` times the reference's own CPU path on the host cores on the SAME 512-hypothesis step.
"""
import argparse
import ctypes as C
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


def workload_string(P):
    return f"C2: {P} pose hypotheses per GPU, obj_06 (31,468 tris), 640x480, projective ICP, criteria (0,0,30) = 31 passes"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
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
        time.sleep(0.1)
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


def cpu_checker():
    from oracle import binding
    kind = "reference" if binding.available("reference") else "port"
    if not binding.available(kind):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    return binding.load(kind), kind


def cpu_pipeline(mesh, scene_pose, poses, target_s=12.0):
    """Times the CPU pipeline (render_cpu -> depth2cloud_cpu -> ICP_Point2Plane_cpu) on a bounded sample."""
    chk, kind = cpu_checker()
    cores = os.cpu_count() or 1
    chk.set_threads(cores)
    K = wl.LINEMOD_K
    proj = chk.compute_proj(K, W, H)
    scene_depth = chk.render(mesh, scene_pose[None], W, H, proj)[0]
    scene = chk.scene_projective(scene_depth, K)
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
    """The reference's own CPU implementation of the path on the host cores, the same 512-hypothesis step as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mesh, scene_pose, poses = load_inputs(args.hyp, 1234)
    chk, kind = cpu_checker()
    cores = os.cpu_count() or 1
    chk.set_threads(cores)
    K = wl.LINEMOD_K
    proj = chk.compute_proj(K, W, H)
    scene_depth = chk.render(mesh, scene_pose[None], W, H, proj)[0]
    scene = chk.scene_projective(scene_depth, K)
    n = len(poses)                     # the whole step: ~1 s on 16 cores
    times = []
    for i in range(args.warmup + args.steps):
        t, _, _ = chk.pipeline(scene, mesh, poses, W, H, proj, K, 0.0, 0.0, ITERS, schedule=1)
        if i >= args.warmup:
            times.append(t)
    total = sum(times)
    value = n * len(times) / total
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "hypotheses/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.hyp),
                   "l2": "n/a (CPU arm)", "parallelism": f"{cores} host threads, omp parallel over hypotheses"},
        "cpu_baseline": {"value": value, "unit": "hypotheses/s", "cores": cores, "kind": kind,
                         "sample": f"all {n} hypotheses of the step, omp parallel over hypotheses"},
        "e2e": {"value": value, "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------
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
    comm = prd.Comm.create()
    L = _lib.lib()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def timed_loop(step, steps, warmup, after=None):
        """warm-up, barrier + synchronize, K steps between CUDA events, barrier + synchronize; max over ranks (ms)."""
        for _ in range(warmup):
            step()
        if after:
            after()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        if after:
            after()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def median_step(step, steps, warmup, after=None):
        """the extra configurations: one CUDA event pair per step, MEDIAN over the steps (a shared host now and then stalls
        a single launch several-fold), max over ranks (ms per step)"""
        for _ in range(warmup):
            step()
        if after:
            after()
        barrier()
        ts = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            if after:
                after()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        barrier()
        return max_over_ranks(float(np.median(ts)))

    def scene_on_every_rank(mesh, scene_pose, w, h, proj):
        """rank 0 renders the scene; pr_broadcast_scene (NCCL, device to device) hands it to the others"""
        if rank == 0:
            d = api.render_cuda_keep_in_gpu(mesh, scene_pose[None], w, h, proj)[0].contiguous()
        else:
            d = torch.empty((h, w), dtype=torch.int32, device="cuda")
        comm.broadcast_scene(d, 0)
        return d

    # ================================ C2: the headline ==========================================================
    P = args.hyp
    mesh, scene_pose, poses = load_inputs(P, 1234 + rank)     # weak scaling: P hypotheses per GPU, different per rank
    K = wl.LINEMOD_K
    proj = api.compute_proj(K, W, H)
    t_setup = time.perf_counter()
    scene_depth = scene_on_every_rank(mesh, scene_pose, W, H, proj)
    ref = api.PoseRefiner(mesh, W, H, K, max_hyp=P)
    ref.set_scene_projective_device(scene_depth)
    torch.cuda.synchronize()
    setup_ms = 1e3 * (time.perf_counter() - t_setup)

    crit = api.ICPConvergenceCriteria(0.0, 0.0, ITERS)
    poses_dev = torch.as_tensor(poses.reshape(P, 16)).cuda()
    results_dev = [torch.empty((P, 18), dtype=torch.float32, device="cuda") for _ in range(2)]   # double buffered: the gather of step i overlaps step i+1
    all_dev = [torch.empty((P * world, 18), dtype=torch.float32, device="cuda") for _ in range(2)]
    poses_pin = torch.as_tensor(poses.reshape(P, 16)).pin_memory()
    results_pin = torch.empty((P, 18), dtype=torch.float32).pin_memory()
    flip = [0]

    def step_device():
        i = flip[0] = flip[0] ^ 1
        ref.run_device(poses_dev, crit, results_dev[i])
        comm.gather_results(results_dev[i], all_dev[i])

    def step_host():
        # host buffers in and out through pr_refiner_run (H2D + D2H + synchronise inside); the cross-GPU gather reads the
        # refiner's DEVICE results of the same step
        i = flip[0] = flip[0] ^ 1
        ref.run(poses_pin, crit, results_pin)
        if world > 1:
            comm.gather_results(ref.results_device(P), all_dev[i])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.pr_launch_count()
    for _ in range(args.warmup):
        step_device()
    comm.wait()
    ref.stage_ms()                     # forget the warm-up runs
    dev_ms = timed_loop(step_device, args.steps, 0, after=lambda: comm.wait())
    launches = (L.pr_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    live_render_ms, live_icp_ms, live_runs = ref.stage_ms()     # the K timed steps: events the refiner recorded on the launch stream

    # end to end: wall clock around host calls that synchronise themselves
    for _ in range(args.warmup):
        step_host()
    comm.wait(host=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    comm.wait(host=True)
    barrier()
    host_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop() if rank == 0 else None

    # ---- stages + roofline of the ICP kernel, on the refiner's own clouds (rank 0; L2 flushed before every repetition)
    roofline = stage = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    def roofline_block(ms, n_pts, n_hyp, w, h, what, how):
        # per pass 12 B per model point + the scene once (W*H*24 B) + 72 B per hypothesis; one launch = 31 passes
        alg = (ITERS + 1) * (12 * n_pts + w * h * 24 + 72 * n_hyp)
        achieved = alg / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": f"icp_hyp_kernel<PackedScene> ({what}; one launch = 31 passes over all hypotheses)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_dram_bytes(os.path.join(ROOT, "profiles", "r02_ncu_icp_hyp.txt")) if what.startswith("C2") else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this workload, read from the committed "
                                  "`ncu --set full` summary profiles/r02_ncu_icp_hyp.txt (scripts/gpu_profiles.sh); DRAM bytes cannot be measured in-run",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "launch_ms": ms, "note": how}

    def icp_roofline(refiner, n_hyp, w, h, what):
        """CUDA-event time of the ICP call alone on the clouds the refiner's last run left in its buffers."""
        _, pts, offsets, counts = refiner.buffers(n_hyp)
        n_pts = int(counts.sum().item())
        cap = int(offsets[n_hyp].item())
        sp = api.SceneProjective()
        sp.width, sp.height, sp.max_dist_diff, sp.K = w, h, 0.1, np.asarray(refiner.K, np.float32).reshape(3, 3)
        sp.pcd, sp.normal = refiner.scene_buffers()
        sc = sp.c()
        packed = torch.empty(L.pr_scene_projective_packed_bytes(w, h), dtype=torch.uint8, device="cuda")
        _lib.check(L.pr_scene_projective_pack(C.byref(sc), packed.data_ptr(), stream), "pr_scene_projective_pack")
        ws_bytes = L.pr_icp_workspace_bytes(n_hyp, cap, 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        res = torch.empty((n_hyp, 18), dtype=torch.float32, device="cuda")

        def icp():
            _lib.check(L.pr_icp_projective_batch_packed(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), n_hyp, cap, C.byref(sc),
                                                        packed.data_ptr(), crit.c(), res.data_ptr(), 0, ws.data_ptr(), ws_bytes, stream),
                       "pr_icp_projective_batch_packed")

        ms = timed(icp)
        # per pass 12 B per model point + the scene once (W*H*24 B) + 72 B per hypothesis; one launch = 31 passes
        alg = (ITERS + 1) * (12 * n_pts + w * h * 24 + 72 * n_hyp)
        achieved = alg / (ms * 1e-3) / 1e9
        return ms, n_pts, {"bound": "hbm", "kernel": f"icp_hyp_kernel<PackedScene> ({what}; one launch = 31 passes over all hypotheses)",
                           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                           "traffic_source": "no ncu capture of this configuration (the C2 capture is profiles/r02_ncu_icp_hyp.txt)",
                           "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "launch_ms": ms,
                           "note": "mean CUDA-event time of pr_icp_projective_batch_packed on the clouds the fused render->cloud call "
                                   "produced, 256 MB written between repetitions so that the clouds start in HBM"}

    if rank == 0:
        ref.run_device(poses_dev, crit, results_dev[0])
        torch.cuda.synchronize()
        ref.stage_ms()
        ms_icp_alone, n_pts, _ = icp_roofline(ref, P, W, H, "C2")
        # the roofline of the line: the ICP call as it ran INSIDE the K timed steps (mean of the refiner's own events)
        roofline = roofline_block(live_icp_ms, n_pts, P, W, H, "C2",
                                  f"mean over the {live_runs} timed steps of the CUDA events pr_refiner records on the launch stream around its ICP "
                                  "call (scene packed once per scene; the claim-order kernel + the ICP kernel)")
        stage = {"render_cloud_ms": live_render_ms, "icp_ms": live_icp_ms, "runs": live_runs, "icp_ms_alone_l2_flushed": ms_icp_alone,
                 "model_points": n_pts, "setup_ms_one_time": setup_ms,
                 "note": "per-stage means over the timed steps (pr_refiner_stage_ms); icp_ms_alone: the same call timed by itself, "
                         "256 MB written before each repetition so that the clouds start in HBM"}
    ref.close()
    del ref

    # ================================ the other configurations ==================================================
    configs = None
    if not args.no_configs:
        configs = {}
        ksteps = max(3, min(args.steps, 7))

        # ---- C3: kd-tree scene of ~100k points, 512 hypotheses per GPU (weak)
        chk_scene = scene_depth.cpu().numpy()
        c3_depth = torch.as_tensor(wl.plane_scene_depth(chk_scene, target_valid=100000)).cuda()
        r3 = api.PoseRefiner(mesh, W, H, K, max_hyp=P)
        r3.set_scene_nn_device(c3_depth)
        res3 = torch.empty((P, 18), dtype=torch.float32, device="cuda")
        ms3 = median_step(lambda: r3.run_device(poses_dev, crit, res3), ksteps, 1)
        if rank == 0:
            _, pts3, off3, cnt3 = r3.buffers(P)
            n3 = int(cnt3.sum().item())
            sn = api.SceneNN().init_cuda(c3_depth, K)
            stats = torch.zeros(4, dtype=torch.int64, device="cuda")
            nq = int(cnt3[0].item())
            wsq, wsq_bytes = api._icp_workspace(1, nq, sn)
            snc = sn.c()
            _lib.check(L.pr_nn_walk_stats(pts3.data_ptr(), nq, C.byref(snc), stats.data_ptr(), wsq.data_ptr(), wsq_bytes, stream), "pr_nn_walk_stats")
            st = stats.cpu().numpy()
            n_scene, n_nodes = sn.pcd.shape[0], len(sn.nodes_host)
            alg3 = (ITERS + 1) * (12 * n3 + n_scene * 24 + n_nodes * 52 + 72 * P)
            configs["C3"] = {"workload": f"{P} hypotheses per GPU, kd-tree scene of {n_scene} points / {n_nodes} nodes, 640x480, 31 passes",
                             "scaling": "weak", "value": P * world / (ms3 * 1e-3), "unit": "hypotheses/s", "ms_per_step": ms3,
                             "nn_queries_per_s": (ITERS + 1) * n3 * world / (ms3 * 1e-3),
                             "tree_walk_pass0": {"node_fetches_per_query": float(st[0]) / nq, "leaf_points_tested_per_query": float(st[1]) / nq},
                             "hash_grid_pass0": {"queries_answered_frac": float(st[2]) / nq, "points_tested_per_query": float(st[3]) / nq},
                             "roofline": {"bound": "hbm", "kernel": "icp_hyp_kernel<PackedNnScene> + render/cloud (whole step)",
                                          "achieved": alg3 / (ms3 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": alg3 / (ms3 * 1e-3) / 1e9 / peak, "traffic": None,
                                          "note": "the tree walk is bound by L1/L2 latency and divergence, not HBM (SURVEY.md 8d): the HBM "
                                                  "fraction is small by construction; lts__t_bytes of the kernel: profiles/r02_ncu_icp_nn.txt"}}
        r3.close()
        del r3, res3
        torch.cuda.empty_cache()

        # ---- C4: 4096 hypotheses at 1280x720, FIXED, sharded over the ranks (strong)
        W4, H4, P4 = 1280, 720, 4096
        K4 = wl.k_1280x720()
        proj4 = api.compute_proj(K4, W4, H4)
        d4 = scene_on_every_rank(mesh, scene_pose, W4, H4, proj4)
        b4, e4 = comm.shard(P4)
        poses4 = wl.hypotheses(P4, seed=4321, scene_pose=scene_pose)[b4:e4]
        chunk = 512
        r4 = api.PoseRefiner(mesh, W4, H4, K4, max_hyp=chunk, capacity_points=chunk * 140000)
        r4.set_scene_projective_device(d4)
        p4_dev = torch.as_tensor(poses4.reshape(-1, 16)).cuda()
        res4 = torch.empty((len(poses4), 18), dtype=torch.float32, device="cuda")
        pad4 = -(-P4 // world)
        gat4 = torch.zeros((pad4, 18), dtype=torch.float32, device="cuda")
        all4 = torch.empty((pad4 * world, 18), dtype=torch.float32, device="cuda")

        def step4():
            for b in range(0, len(poses4), chunk):
                r4.run_device(p4_dev[b:b + chunk], crit, res4[b:b + chunk])
            gat4[: len(poses4)].copy_(res4)
            comm.gather_results(gat4, all4)

        ms4 = median_step(step4, ksteps, 1, after=lambda: comm.wait())
        if rank == 0:
            nb = min(chunk, len(poses4))
            r4.run_device(p4_dev[:nb], crit, res4[:nb])
            torch.cuda.synchronize()
            ms_icp4, n_pts4, roof4 = icp_roofline(r4, nb, W4, H4, f"C4, a {nb}-hypothesis batch of this rank's shard")
            configs["C4"] = {"workload": f"{P4} hypotheses FIXED, obj_06, 1280x720, projective, 31 passes, sharded over {world} GPU(s) "
                                         f"({len(poses4)} on rank 0, refined in batches of {chunk})",
                             "scaling": "strong", "value": P4 / (ms4 * 1e-3), "unit": "hypotheses/s", "ms_per_step": ms4,
                             "model_points_per_hypothesis": n_pts4 / nb, "icp_ms_per_batch": ms_icp4, "roofline": roof4}
        r4.close()
        del r4, res4, p4_dev, d4
        torch.cuda.empty_cache()

        # ---- C5: render only, 8192 poses FIXED of the 49,920-triangle sphere, sharded over the ranks (strong)
        P5 = 8192
        tris5 = wl.uv_sphere()
        verts5, faces5 = api.mesh_index(tris5)
        faces5, off5, cv5 = api.mesh_cluster(verts5, faces5)
        b5, e5 = comm.shard(P5)
        n5 = e5 - b5
        p5 = torch.as_tensor(wl.shoemake_poses(P5, seed=99)[b5:e5].reshape(-1, 16)).cuda()
        v5, f5 = torch.as_tensor(verts5).cuda(), torch.as_tensor(faces5).cuda()
        cl5 = (torch.as_tensor(off5).cuda(), torch.as_tensor(cv5).cuda())
        depth5 = torch.empty((n5, H, W), dtype=torch.int32, device="cuda")
        ws5 = torch.empty(L.pr_render_cloud_workspace_bytes(n5, verts5.shape[0], faces5.shape[0], W, H), dtype=torch.uint8, device="cuda")
        ms5 = median_step(lambda: api.render_clustered_keep_in_gpu(v5, f5, p5, W, H, proj, cl5, out=depth5, ws=ws5), ksteps, 1)
        if rank == 0:
            alg5 = n5 * W * H * 4
            configs["C5"] = {"workload": f"{P5} poses FIXED, UV sphere of {tris5.shape[0]} triangles, 640x480 int32 depth kept on the device, "
                                         f"render only, sharded over {world} GPU(s)",
                             "scaling": "strong", "value": P5 / (ms5 * 1e-3), "unit": "poses/s", "ms_per_step": ms5,
                             "roofline": {"bound": "hbm", "kernel": "vertex + cluster binning + raster_tile_kernel (whole render call, per GPU)",
                                          "achieved": alg5 / (ms5 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": alg5 / (ms5 * 1e-3) / 1e9 / peak, "traffic": None,
                                          "note": "algorithmic bytes = 4 B per output pixel per pose of this rank's shard, written once; the "
                                                  "kernel is instruction-bound (profiles/r02_ncu_raster_tile.txt)"}}
        del depth5, ws5
        torch.cuda.empty_cache()

    cpu = ref_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_pipeline(mesh, scene_pose, poses, target_s=args.cpu_seconds)
        ref_cuda = ref_cuda_build(min(P, 256))

    if rank == 0:
        total_hyp = P * world * args.steps
        out = {
            "metric": METRIC, "value": total_hyp / (dev_ms * 1e-3), "unit": "hypotheses/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(P),
                       "l2": "per-step working set (depth batch 629 MB + clouds ~135 MB) is larger than the 126 MB L2; no explicit flush",
                       "parallelism": f"hypothesis shards x{world}: pr_broadcast_scene + pr_gather_results (NCCL inside the C ABI, "
                                      "gather on a side stream)" if world > 1 else "single GPU"},
            "e2e": {"value": total_hyp / (host_ms * 1e-3), "unit": "hypotheses/s", "ms_per_step": host_ms / args.steps,
                    "h2d_bytes_per_step": P * 64, "d2h_bytes_per_step": P * 72 + 4},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roofline, "stages": stage,
        }
        if configs:
            out["configs"] = configs
        if cpu is not None:
            out["cpu_baseline"] = cpu
        if ref_cuda is not None:
            out["ref_cuda_build"] = ref_cuda
        print(json.dumps(out), flush=True)
    comm.close()
    if world > 1:
        dist.destroy_process_group()


def ref_cuda_build(n_hyp):
    """The reference's own CUDA path (render_cuda_keep_in_gpu -> depth2cloud_cuda -> ICP_Point2Plane_cuda per hypothesis,
    oracle/_ref/libpose_refine_refcuda.so = its .cu files compiled unmodified for sm_100) timed on this GPU on a bounded
    sample of the same workload, in a child process (scripts/time_ref_cuda.py).  Reported next to cpu_baseline; it is the
    "reference CUDA build" of north_star's 10x target.  None when the library was not built."""
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


def ncu_dram_bytes(path):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel in a profiles/*.txt summary written by
    scripts/collect_profiles_r02.py from an `ncu --set full` capture (one launch); None when the file or the lines are missing."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    try:
        for ln in open(path):
            parts = ln.split()
            if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[-1] in scale:
                total += float(parts[-2].replace(",", "")) * scale[parts[-1]]
                seen += 1
    except OSError:
        return None
    return total if seen == 2 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hyp", type=int, default=512, help="hypotheses per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / ref_cuda_build legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C4 / C5 blocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
