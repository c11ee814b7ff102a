// refiner.cu -- the whole hot path behind one call: render -> depth2cloud -> ICP.
//
// pr_refiner is the batch counterpart of what the reference wires by hand in test.cpp:143-172
// (render_cuda_keep_in_gpu -> depth2cloud_cuda -> init_Scene_*_cuda -> ICP_Point2Plane_cuda) and
// of PoseRenderer (pose_renderer.h:9-32), which uploads the mesh once and renders batches of poses.
// It owns every device buffer it needs (allocated once at create / set_scene), so a run is a fixed
// sequence of launches on one stream with no allocation and no host round trip; the host variant
// adds the pose upload (H2D) and the result download (D2H) and synchronises once at the end.
//
// Also here: library-level entry points (version, error strings, device check).
#include "common.cuh"
#include <cstdio>
#include <cstring>
#include <vector>
#include <new>

struct pr_refiner {
    uint32_t W = 0, H = 0;
    float K[9];
    float proj[16];
    size_t n_tris = 0, max_hyp = 0, capacity_points = 0;
    float* d_tris = nullptr;
    float* d_verts = nullptr;     // deduplicated mesh (pr_mesh_index): the refiner renders the indexed form
    int32_t* d_faces = nullptr;
    size_t n_verts = 0;
    int32_t *d_cl_off = nullptr, *d_cl_verts = nullptr;
    pr_mesh_clusters clusters = {0, nullptr, nullptr};
    float* d_poses = nullptr;
    int32_t* d_depth = nullptr;
    float* d_pts = nullptr;
    uint32_t *d_counts = nullptr, *d_offsets = nullptr, *d_overflow = nullptr;
    pr_registration_result* d_results = nullptr;
    void *ws_render = nullptr, *ws_cloud = nullptr, *ws_icp = nullptr;
    size_t ws_render_bytes = 0, ws_cloud_bytes = 0, ws_icp_bytes = 0;
    // scene
    int scene_kind = -1;   // 0 projective, 1 nn
    float* d_scene_pcd = nullptr;
    float* d_scene_nrm = nullptr;
    pr_node_kdtree* d_nodes = nullptr;
    void* d_scene_packed = nullptr;   // 32-byte records of the projective scene, packed once per scene
    void* d_scene_depth = nullptr;    // staging for a host depth image (4 bytes per pixel)
    void* d_scene_ws = nullptr;       // kd-tree build workspace
    size_t scene_ws_bytes = 0;
    pr_scene_projective sp;
    pr_scene_nn sn;
    uint32_t pending_hyp = 0;
    // stage timing: three events per run (start, after render->cloud, after ICP), a ring of the last kTimedRuns runs
    static constexpr int kTimedRuns = 256;
    cudaEvent_t ev[kTimedRuns][3] = {};
    unsigned ev_next = 0, ev_count = 0;
    uint32_t* h_overflow = nullptr;   // pinned: [0] overflow flag, [1] total points of the last batch (padded)
};

namespace {

void free_scene(pr_refiner* r) {
    cudaFree(r->d_scene_pcd); cudaFree(r->d_scene_nrm); cudaFree(r->d_nodes);
    cudaFree(r->d_scene_packed); cudaFree(r->d_scene_depth); cudaFree(r->d_scene_ws);
    r->d_scene_pcd = nullptr; r->d_scene_nrm = nullptr; r->d_nodes = nullptr;
    r->d_scene_packed = nullptr; r->d_scene_depth = nullptr; r->d_scene_ws = nullptr;
    r->scene_kind = -1;
}
// scene buffers are allocated on first use and kept: setting a scene never allocates after that
int ensure(void** p, size_t bytes) {
    if (*p) return PR_OK;
    PR_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 256));
    return PR_OK;
}

int run_device(pr_refiner* r, const float* poses_dev, size_t n_hyp, pr_icp_criteria crit,
               pr_registration_result* results_dev, cudaStream_t stream) {
    pr_stream_t s = reinterpret_cast<pr_stream_t>(stream);
    cudaEvent_t* ev = r->ev[r->ev_next];
    if (!ev[0]) for (int i = 0; i < 3; i++) PR_CUDA_TRY(cudaEventCreate(&ev[i]));
    struct StageEnd {       // the closing event of a run is recorded on every way out
        pr_refiner* r; cudaEvent_t e; cudaStream_t st;
        ~StageEnd() { cudaEventRecord(e, st); r->ev_next = (r->ev_next + 1) % pr_refiner::kTimedRuns; if (r->ev_count < (unsigned)pr_refiner::kTimedRuns) r->ev_count++; }
    } stage_end{r, ev[2], stream};
    PR_CUDA_TRY(cudaEventRecord(ev[0], stream));
    // render + clouds in one pass over the depth batch (tile-ordered clouds; the reduction does not care about order)
    int rc = pr_render_cloud_batch(r->d_verts, r->n_verts, r->d_faces, r->n_tris, poses_dev, 1, n_hyp, r->W, r->H, r->proj, r->K,
                                   r->d_depth, r->d_pts, r->capacity_points, 4, r->d_counts, r->d_offsets, r->d_overflow,
                                   &r->clusters, r->ws_render, r->ws_render_bytes, s);
    if (rc != PR_OK) return rc;
    PR_CUDA_TRY(cudaEventRecord(ev[1], stream));
    if (r->scene_kind == 0) {
        // average cloud size, for the cluster size of the ICP launch: what the previous batch had (its total is copied to
        // pinned memory behind every run, asynchronously: a stale or missing value only costs a less suitable cluster
        // size); the first batch assumes the object covers 1/14 of the image
        volatile uint32_t* hint = r->h_overflow;
        const size_t est = (hint[1] != 0 && r->pending_hyp != 0) ? (size_t)hint[1] / r->pending_hyp : ((size_t)r->W * r->H) / 14;
        rc = prb::icp_projective_packed(r->d_pts, r->d_offsets, r->d_counts, n_hyp, r->capacity_points, &r->sp, r->d_scene_packed,
                                        crit, results_dev, 0, r->ws_icp, r->ws_icp_bytes, s, est ? est : 1);
        if (rc != PR_OK) return rc;
        if (cudaMemcpyAsync(r->h_overflow + 1, r->d_offsets + n_hyp, 4, cudaMemcpyDeviceToHost, stream) == cudaSuccess)
            r->pending_hyp = (uint32_t)n_hyp;      // batch size that total belongs to (a total that has not landed yet pairs an
                                                   // older total with this size for one call: a heuristic, not a contract)
        return PR_OK;
    }
    return pr_icp_nn_batch(r->d_pts, r->d_offsets, r->d_counts, n_hyp, r->capacity_points, &r->sn, crit, results_dev, 0,
                           r->ws_icp, r->ws_icp_bytes, s);
}

}  // namespace

extern "C" {

int pr_version(void) { return 100; }   // 0.1.0

const char* pr_error_string(int status) {
    switch (status) {
    case PR_OK: return "ok";
    case PR_ERR_INVALID_ARGUMENT: return "invalid argument";
    case PR_ERR_UNSUPPORTED: return "unsupported (e.g. depth2cloud stride != 1)";
    case PR_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case PR_ERR_CAPACITY: return "output capacity exceeded";
    case PR_ERR_IO: return "i/o error";
    case PR_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    case PR_ERR_COMM: return "NCCL unavailable or an NCCL call failed";
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown error";
}

int pr_device_check(void) {
    // attribute query (microseconds), not cudaGetDeviceProperties (milliseconds, and worse with several processes on
    // the host): every Python entry point calls this first
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return PR_ERR_NO_DEVICE; }
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return PR_ERR_NO_DEVICE; }
    return (major == 10) ? PR_OK : PR_ERR_NO_DEVICE;
}

int pr_device_malloc(void** ptr, size_t bytes) {
    if (!ptr) return PR_ERR_INVALID_ARGUMENT;
    PR_CUDA_TRY(cudaMalloc(ptr, bytes ? bytes : 1));
    return PR_OK;
}
int pr_device_free(void* ptr) {
    PR_CUDA_TRY(cudaFree(ptr));
    return PR_OK;
}
int pr_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, pr_stream_t stream) {
    if (bytes == 0) return PR_OK;
    PR_CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, prb::as_stream(stream)));
    PR_CUDA_TRY(cudaStreamSynchronize(prb::as_stream(stream)));
    return PR_OK;
}
int pr_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, pr_stream_t stream) {
    if (bytes == 0) return PR_OK;
    PR_CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, prb::as_stream(stream)));
    PR_CUDA_TRY(cudaStreamSynchronize(prb::as_stream(stream)));
    return PR_OK;
}
int pr_stream_synchronize(pr_stream_t stream) {
    PR_CUDA_TRY(cudaStreamSynchronize(prb::as_stream(stream)));
    return PR_OK;
}

int pr_refiner_create(pr_refiner** out, const float* tris_host, size_t n_tris, uint32_t width, uint32_t height,
                      const float K[9], size_t max_hyp, size_t capacity_points) {
    if (!out || !tris_host || !K || n_tris == 0 || width == 0 || height == 0 || max_hyp == 0) return PR_ERR_INVALID_ARGUMENT;
    int rc = pr_device_check();
    if (rc != PR_OK) return rc;
    pr_refiner* r = new (std::nothrow) pr_refiner();
    if (!r) return PR_ERR_INVALID_ARGUMENT;
    r->W = width; r->H = height; r->n_tris = n_tris; r->max_hyp = max_hyp;
    memcpy(r->K, K, 36);
    pr_compute_proj(K, (int)width, (int)height, 10.f, 10000.f, r->proj);
    const size_t n_px = (size_t)width * height;
    // default: room for every hypothesis covering a quarter of the image
    r->capacity_points = capacity_points ? capacity_points : (max_hyp * (n_px / 4 + 4));
    std::vector<float> verts(n_tris * 9);
    std::vector<int32_t> faces(n_tris * 3);
    rc = pr_mesh_index(tris_host, n_tris, verts.data(), faces.data(), &r->n_verts);
    if (rc != PR_OK) { delete r; return rc; }
    // Morton-ordered faces + clusters: the rasteriser bins 64-triangle clusters instead of triangles
    std::vector<int32_t> cl_off(n_tris + 2), cl_verts(3 * n_tris);
    size_t n_clusters = 0;
    rc = pr_mesh_cluster(verts.data(), r->n_verts, faces.data(), n_tris, cl_off.data(), cl_verts.data(), &n_clusters);
    if (rc != PR_OK) { delete r; return rc; }
#ifndef PR_NO_VERTEX_RENUMBER
    {   // Renumber the vertices in the order the clusters first use them: a cluster's vertex list then reads (mostly) consecutive
        // projected vertices -- whole 32-byte sectors instead of one 16-byte vertex per sector in cluster_span_kernel, and the
        // three vertices of a triangle sit near each other for the tile kernel.  Same triangles, same coordinates: same depth.
        std::vector<int32_t> new_id(r->n_verts, -1);
        int32_t next = 0;
        for (int32_t i = 0; i < cl_off[n_clusters]; i++) if (new_id[cl_verts[i]] < 0) new_id[cl_verts[i]] = next++;
        for (size_t v = 0; v < r->n_verts; v++) if (new_id[v] < 0) new_id[v] = next++;       // vertices no face uses
        std::vector<float> moved(r->n_verts * 3);
        for (size_t v = 0; v < r->n_verts; v++) memcpy(&moved[3 * (size_t)new_id[v]], &verts[3 * v], 12);
        memcpy(verts.data(), moved.data(), r->n_verts * 12);
        for (size_t i = 0; i < 3 * n_tris; i++) faces[i] = new_id[faces[i]];
        for (int32_t i = 0; i < cl_off[n_clusters]; i++) cl_verts[i] = new_id[cl_verts[i]];
    }
#endif
    r->ws_render_bytes = pr_render_cloud_workspace_bytes(max_hyp, r->n_verts, n_tris, width, height);
    r->ws_cloud_bytes = pr_depth2cloud_workspace_bytes(max_hyp, width, height);
    r->ws_icp_bytes = pr_icp_workspace_bytes(max_hyp, r->capacity_points, 3 * n_px + 16);   // projective: n_px; kd-tree: <= n_px points + 2 * (2 n_px + 1) nodes... bounded by 3 n_px for leaf >= 2
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 256); };
    alloc((void**)&r->d_tris, n_tris * 36);
    alloc((void**)&r->d_verts, r->n_verts * 12);
    alloc((void**)&r->d_faces, n_tris * 12);
    alloc((void**)&r->d_cl_off, (n_clusters + 1) * 4);
    alloc((void**)&r->d_cl_verts, (size_t)cl_off[n_clusters] * 4);
    alloc((void**)&r->d_poses, max_hyp * 64);
    alloc((void**)&r->d_depth, max_hyp * n_px * 4);
    alloc((void**)&r->d_pts, r->capacity_points * 12 + 64);
    alloc((void**)&r->d_counts, max_hyp * 4);
    alloc((void**)&r->d_offsets, (max_hyp + 1) * 4);
    alloc((void**)&r->d_overflow, 256);
    alloc((void**)&r->d_results, max_hyp * sizeof(pr_registration_result));
    alloc(&r->ws_render, r->ws_render_bytes);
    alloc(&r->ws_cloud, r->ws_cloud_bytes);
    alloc(&r->ws_icp, r->ws_icp_bytes);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&r->h_overflow, 256);
    if (e == cudaSuccess) memset(r->h_overflow, 0, 256);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_tris, tris_host, n_tris * 36, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_verts, verts.data(), r->n_verts * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_faces, faces.data(), n_tris * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_cl_off, cl_off.data(), (n_clusters + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(r->d_cl_verts, cl_verts.data(), (size_t)cl_off[n_clusters] * 4, cudaMemcpyHostToDevice);
    r->clusters.n_clusters = n_clusters; r->clusters.vert_off_dev = r->d_cl_off; r->clusters.verts_dev = r->d_cl_verts;
    if (e == cudaSuccess) e = cudaMemset(r->d_overflow, 0, 256);
    if (e != cudaSuccess) { pr_refiner_destroy(r); return (int)e; }
    *out = r;
    return PR_OK;
}

void pr_refiner_destroy(pr_refiner* r) {
    if (!r) return;
    free_scene(r);
    cudaFree(r->d_tris); cudaFree(r->d_verts); cudaFree(r->d_faces); cudaFree(r->d_cl_off); cudaFree(r->d_cl_verts); cudaFree(r->d_poses); cudaFree(r->d_depth); cudaFree(r->d_pts);
    cudaFree(r->d_counts); cudaFree(r->d_offsets); cudaFree(r->d_overflow); cudaFree(r->d_results);
    cudaFree(r->ws_render); cudaFree(r->ws_cloud); cudaFree(r->ws_icp);
    for (auto& e3 : r->ev) for (auto& e : e3) if (e) cudaEventDestroy(e);
    if (r->h_overflow) cudaFreeHost(r->h_overflow);
    delete r;
}

int pr_refiner_set_scene_projective_device(pr_refiner* r, const void* depth_dev, int depth_is_int32, float max_dist_diff,
                                           pr_stream_t stream) {
    if (!r || !depth_dev) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)r->W * r->H;
    int rc = ensure((void**)&r->d_scene_pcd, n_px * 12 + 256);
    if (rc == PR_OK) rc = ensure((void**)&r->d_scene_nrm, n_px * 12 + 256);
    if (rc == PR_OK) rc = ensure(&r->d_scene_packed, n_px * 32);
    if (rc != PR_OK) return rc;
    r->scene_kind = -1;
    rc = pr_scene_projective_init(depth_dev, depth_is_int32, r->W, r->H, r->K, r->d_scene_pcd, r->d_scene_nrm, stream);
    if (rc != PR_OK) return rc;
    r->sp.width = r->W; r->sp.height = r->H; r->sp.max_dist_diff = max_dist_diff;
    memcpy(r->sp.K, r->K, 36);
    r->sp.pcd_dev = r->d_scene_pcd; r->sp.normal_dev = r->d_scene_nrm;
    rc = pr_scene_projective_pack(&r->sp, r->d_scene_packed, stream);
    if (rc != PR_OK) return rc;
    r->scene_kind = 0;
    return PR_OK;
}

int pr_refiner_set_scene_projective(pr_refiner* r, const void* depth_host, int depth_is_int32, float max_dist_diff) {
    if (!r || !depth_host) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)r->W * r->H;
    int rc = ensure(&r->d_scene_depth, n_px * 4);
    if (rc != PR_OK) return rc;
    PR_CUDA_TRY(cudaMemcpy(r->d_scene_depth, depth_host, n_px * (depth_is_int32 ? 4 : 2), cudaMemcpyHostToDevice));
    rc = pr_refiner_set_scene_projective_device(r, r->d_scene_depth, depth_is_int32, max_dist_diff, nullptr);
    if (rc != PR_OK) return rc;
    PR_CUDA_TRY(cudaStreamSynchronize(nullptr));
    return PR_OK;
}

int pr_refiner_set_scene_nn_device(pr_refiner* r, const void* depth_dev, int depth_is_int32, pr_stream_t stream) {
    if (!r || !depth_dev) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)r->W * r->H;
    // everything is built on the device (pr_scene_nn_build): points, normals, kd-tree
    r->scene_ws_bytes = pr_scene_nn_build_workspace_bytes(r->W, r->H);
    int rc = ensure(&r->d_scene_ws, r->scene_ws_bytes);
    if (rc == PR_OK) rc = ensure((void**)&r->d_scene_pcd, n_px * 12 + 256);
    if (rc == PR_OK) rc = ensure((void**)&r->d_scene_nrm, n_px * 12 + 256);
    if (rc == PR_OK) rc = ensure((void**)&r->d_nodes, (2 * n_px + 1) * sizeof(pr_node_kdtree));
    if (rc != PR_OK) return rc;
    r->scene_kind = -1;
    size_t n_pts = 0, n_nodes = 0;
    rc = pr_scene_nn_build(depth_dev, depth_is_int32, r->W, r->H, r->K, 10, r->d_scene_pcd, r->d_scene_nrm, n_px, r->d_nodes,
                           2 * n_px + 1, &n_pts, &n_nodes, r->d_scene_ws, r->scene_ws_bytes, stream);
    if (rc != PR_OK) return rc;
    {   // room for the nearest-neighbour cache (one int per model point): grow the ICP workspace once
        const size_t need = pr_icp_nn_workspace_bytes(r->max_hyp, r->capacity_points, n_pts, n_nodes);
        if (need > r->ws_icp_bytes) {
            PR_CUDA_TRY(cudaStreamSynchronize(prb::as_stream(stream)));
            cudaFree(r->ws_icp); r->ws_icp = nullptr; r->ws_icp_bytes = 0;
            PR_CUDA_TRY(cudaMalloc(&r->ws_icp, need));
            r->ws_icp_bytes = need;
        }
    }
    r->sn.max_dist_diff = 0.1f;   // Scene_nn has no setter upstream (pcd_scene.h:49)
    r->sn.pcd_dev = r->d_scene_pcd; r->sn.normal_dev = r->d_scene_nrm; r->sn.nodes_dev = r->d_nodes;
    r->sn.n_points = n_pts; r->sn.n_nodes = n_nodes;
    r->scene_kind = 1;
    return PR_OK;
}

int pr_refiner_set_scene_nn(pr_refiner* r, const void* depth_host, int depth_is_int32) {
    if (!r || !depth_host) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)r->W * r->H;
    int rc = ensure(&r->d_scene_depth, n_px * 4);
    if (rc != PR_OK) return rc;
    PR_CUDA_TRY(cudaMemcpy(r->d_scene_depth, depth_host, n_px * (depth_is_int32 ? 4 : 2), cudaMemcpyHostToDevice));
    return pr_refiner_set_scene_nn_device(r, r->d_scene_depth, depth_is_int32, nullptr);
}

int pr_refiner_run_device(pr_refiner* r, const float* poses_dev, size_t n_hyp, pr_icp_criteria criteria,
                          pr_registration_result* results_dev, pr_stream_t stream) {
    if (!r || !poses_dev || !results_dev || n_hyp > r->max_hyp || r->scene_kind < 0) return PR_ERR_INVALID_ARGUMENT;
    if (n_hyp == 0) return PR_OK;
    return run_device(r, poses_dev, n_hyp, criteria, results_dev, prb::as_stream(stream));
}

int pr_refiner_run(pr_refiner* r, const float* poses_host, size_t n_hyp, pr_icp_criteria criteria,
                   pr_registration_result* results_host, pr_stream_t stream_) {
    if (!r || !poses_host || !results_host || n_hyp > r->max_hyp || r->scene_kind < 0) return PR_ERR_INVALID_ARGUMENT;
    if (n_hyp == 0) return PR_OK;
    cudaStream_t stream = prb::as_stream(stream_);
    PR_CUDA_TRY(cudaMemcpyAsync(r->d_poses, poses_host, n_hyp * 64, cudaMemcpyHostToDevice, stream));
    int rc = run_device(r, r->d_poses, n_hyp, criteria, r->d_results, stream);
    if (rc != PR_OK) return rc;
    PR_CUDA_TRY(cudaMemcpyAsync(results_host, r->d_results, n_hyp * sizeof(pr_registration_result), cudaMemcpyDeviceToHost, stream));
    PR_CUDA_TRY(cudaMemcpyAsync(r->h_overflow, r->d_overflow, 4, cudaMemcpyDeviceToHost, stream));
    PR_CUDA_TRY(cudaStreamSynchronize(stream));
    return r->h_overflow[0] ? PR_ERR_CAPACITY : PR_OK;
}

int pr_refiner_buffers(pr_refiner* r, const int32_t** depth_dev, const float** pts_dev,
                       const uint32_t** offsets_dev, const uint32_t** counts_dev) {
    if (!r) return PR_ERR_INVALID_ARGUMENT;
    if (depth_dev) *depth_dev = r->d_depth;
    if (pts_dev) *pts_dev = r->d_pts;
    if (offsets_dev) *offsets_dev = r->d_offsets;
    if (counts_dev) *counts_dev = r->d_counts;
    return PR_OK;
}

int pr_refiner_overflow_flag(pr_refiner* r, const uint32_t** flag_dev) {
    if (!r || !flag_dev) return PR_ERR_INVALID_ARGUMENT;
    *flag_dev = r->d_overflow;
    return PR_OK;
}

int pr_refiner_stage_ms(pr_refiner* r, float* render_cloud_ms, float* icp_ms, uint32_t* n_runs) {
    if (!r) return PR_ERR_INVALID_ARGUMENT;
    double a = 0.0, b = 0.0;
    unsigned n = 0;
    for (unsigned k = 0; k < r->ev_count; k++) {
        cudaEvent_t* ev = r->ev[(r->ev_next + pr_refiner::kTimedRuns - 1 - k) % pr_refiner::kTimedRuns];
        if (!ev[0]) continue;
        PR_CUDA_TRY(cudaEventSynchronize(ev[2]));
        float t0 = 0.f, t1 = 0.f;
        if (cudaEventElapsedTime(&t0, ev[0], ev[1]) != cudaSuccess || cudaEventElapsedTime(&t1, ev[1], ev[2]) != cudaSuccess) { cudaGetLastError(); continue; }
        a += t0; b += t1; n++;
    }
    if (render_cloud_ms) *render_cloud_ms = n ? (float)(a / n) : 0.f;
    if (icp_ms) *icp_ms = n ? (float)(b / n) : 0.f;
    if (n_runs) *n_runs = n;
    r->ev_count = 0;
    return PR_OK;
}

int pr_refiner_scene_buffers(pr_refiner* r, const float** scene_pcd_dev, const float** scene_normal_dev,
                             const pr_registration_result** results_dev) {
    if (!r) return PR_ERR_INVALID_ARGUMENT;
    if (scene_pcd_dev) *scene_pcd_dev = r->d_scene_pcd;
    if (scene_normal_dev) *scene_normal_dev = r->d_scene_nrm;
    if (results_dev) *results_dev = r->d_results;
    return PR_OK;
}

}  // extern "C"
