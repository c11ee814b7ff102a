// A C++ host sharding a batch of pose hypotheses over the GPUs of one box through the C ABI only
// (pr_shard_plan, pr_comm_create, pr_broadcast_scene, pr_gather_results): one host thread per GPU, the pattern the
// reference's README suggests for a single GPU (README.md:15).  Rank 0 renders the scene and broadcasts the DEVICE depth
// image; every rank prepares the scene on its own GPU, refines its contiguous shard and all-gathers the results.
// Checked: the gathered results of every rank equal, bit for bit, what one GPU computes for the whole batch.
//   usage: multi_gpu_demo model.ply [nranks]      (nranks defaults to the number of GPUs, at most 8)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "pose_refine_b200.h"
#include "pose_refine/cuda_renderer/renderer.h"

namespace {
const int W = 640, H = 480, N_HYP = 11;       // 11: shards of unequal size
const float K[9] = {572.4114f, 0.f, 325.2611f, 0.f, 573.57043f, 242.04899f, 0.f, 0.f, 1.f};
std::atomic<int> g_fail{0};
std::atomic<int> g_arrived{0};
unsigned char g_uid[128];

#define OK(expr) do { int rc_ = (expr); if (rc_ != PR_OK) { std::fprintf(stderr, "%s -> %s\n", #expr, pr_error_string(rc_)); g_fail++; return; } } while (0)

void pose_of(int i, float* m) {     // small deterministic perturbations of the reference's scene pose
    const float R[9] = {0.34768538f, 0.93761126f, 0.f, 0.70540612f, -0.26157897f, -0.65877056f, -0.61767070f, 0.22904489f, -0.75234390f};
    const float a = 0.02f * (float)(i - 5), c = std::cos(a), s = std::sin(a);
    const float Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) { float v = 0; for (int j = 0; j < 3; j++) v += Rz[3 * r + j] * R[3 * j + k]; m[4 * r + k] = v; }
    m[3] = 2.f * (float)(i % 3); m[7] = -1.5f * (float)(i % 4); m[11] = 300.f + (float)i;
    m[12] = m[13] = m[14] = 0.f; m[15] = 1.f;
}

void refine_shard(const std::vector<cuda_renderer::Model::Triangle>& tris, pr_comm* comm, int nranks, int rank,
                  std::vector<pr_registration_result>& all_out) {
    float proj[16], scene_pose[16];
    pr_compute_proj(K, W, H, 10.f, 10000.f, proj);
    pose_of(5, scene_pose);
    void* scene_depth = nullptr;
    OK(pr_device_malloc(&scene_depth, (size_t)W * H * 4));
    if (rank == 0) {     // the scene exists on rank 0 only
        float* tris_dev = nullptr; void* ws = nullptr;
        OK(pr_device_malloc((void**)&tris_dev, tris.size() * 36));
        OK(pr_memcpy_h2d(tris_dev, tris.data(), tris.size() * 36, nullptr));
        const size_t wsb = pr_render_workspace_bytes(1, tris.size(), W, H);
        OK(pr_device_malloc(&ws, wsb));
        pr_roi none = {0, 0, 0, 0};
        OK(pr_render_batch(tris_dev, tris.size(), scene_pose, 0, 1, W, H, proj, none, (int32_t*)scene_depth, ws, wsb, nullptr));
        OK(pr_stream_synchronize(nullptr));
        pr_device_free(ws); pr_device_free(tris_dev);
    }
    if (comm) OK(pr_broadcast_scene(comm, scene_depth, (size_t)W * H * 4, 0, nullptr));
    pr_refiner* ref = nullptr;
    OK(pr_refiner_create(&ref, reinterpret_cast<const float*>(tris.data()), tris.size(), W, H, K, N_HYP, 0));
    OK(pr_refiner_set_scene_projective_device(ref, scene_depth, 1, 0.1f, nullptr));
    size_t begin = 0, count = N_HYP;
    OK(pr_shard_plan(N_HYP, nranks, rank, &begin, &count));
    const size_t per_rank = (N_HYP + nranks - 1) / nranks;       // shards padded to the largest one
    std::vector<float> poses(16 * per_rank, 0.f);
    for (size_t i = 0; i < count; i++) pose_of((int)(begin + i), &poses[16 * i]);
    float* poses_dev = nullptr; pr_registration_result *local_dev = nullptr, *all_dev = nullptr;
    OK(pr_device_malloc((void**)&poses_dev, poses.size() * 4));
    OK(pr_device_malloc((void**)&local_dev, per_rank * sizeof(pr_registration_result)));
    OK(pr_device_malloc((void**)&all_dev, per_rank * nranks * sizeof(pr_registration_result)));
    OK(pr_memcpy_h2d(poses_dev, poses.data(), poses.size() * 4, nullptr));
    const pr_icp_criteria crit = {0.f, 0.f, 30};
    if (count) OK(pr_refiner_run_device(ref, poses_dev, count, crit, local_dev, nullptr));
    std::vector<pr_registration_result> gathered(per_rank * nranks);
    if (comm) {
        OK(pr_gather_results(comm, local_dev, per_rank, all_dev, nullptr));
        OK(pr_gather_wait(comm, nullptr, 1));
        OK(pr_memcpy_d2h(gathered.data(), all_dev, gathered.size() * sizeof(pr_registration_result), nullptr));
    } else {
        OK(pr_memcpy_d2h(gathered.data(), local_dev, per_rank * sizeof(pr_registration_result), nullptr));
    }
    // drop the padding: rank r's block holds shard r
    all_out.clear();
    for (int r = 0; r < nranks; r++) {
        size_t b, c;
        pr_shard_plan(N_HYP, nranks, r, &b, &c);
        for (size_t i = 0; i < c; i++) all_out.push_back(gathered[r * per_rank + i]);
    }
    pr_refiner_destroy(ref);
    pr_device_free(poses_dev); pr_device_free(local_dev); pr_device_free(all_dev); pr_device_free(scene_depth);
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s model.ply [nranks]\n", argv[0]); return 2; }
    cuda_renderer::Model model(argv[1]);
    int ndev = 0;
    if (pr_device_count(&ndev) != PR_OK || ndev < 1) { std::fprintf(stderr, "no CUDA device\n"); return 2; }
    int nranks = argc > 2 ? std::atoi(argv[2]) : (ndev < 8 ? ndev : 8);
    if (nranks > ndev) nranks = ndev;
    // the whole batch on GPU 0: what the shards must add up to
    std::vector<pr_registration_result> want;
    pr_set_device(0);
    refine_shard(model.tris, nullptr, 1, 0, want);
    if (g_fail) return 1;
    std::vector<std::vector<pr_registration_result>> got(nranks);
    if (nranks > 1 && pr_nccl_unique_id(g_uid) != PR_OK) { std::fprintf(stderr, "NCCL not available\n"); return 3; }
    std::vector<std::thread> threads;
    for (int r = 0; r < nranks; r++)
        threads.emplace_back([&, r] {
            OK(pr_set_device(r));
            pr_comm* comm = nullptr;
            if (nranks > 1) OK(pr_comm_create(&comm, g_uid, nranks, r));
            refine_shard(model.tris, comm, nranks, r, got[r]);
            if (comm) pr_comm_destroy(comm);
        });
    for (auto& t : threads) t.join();
    if (g_fail) return 1;
    int bad = 0;
    for (int r = 0; r < nranks; r++) {
        if (got[r].size() != want.size()) { bad++; continue; }
        bad += std::memcmp(got[r].data(), want.data(), want.size() * sizeof(pr_registration_result)) != 0;
    }
    std::printf("ranks %d hypotheses %d mismatching_ranks %d fitness[0] %.6f\n", nranks, N_HYP, bad, want[0].fitness);
    return bad ? 1 : 0;
}
