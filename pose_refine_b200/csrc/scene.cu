// scene.cu -- one-time scene preparation for sm_100a.
//
// Projective scene (Scene_projective::init_Scene_projective_cuda, depth_scene.cu:3-20 ->
// depth_scene.cpp:3-35): organised cloud via dep2pcd (scene/common.h:47-61) and LINEMOD-style
// integer normals (get_normal, scene/common.cpp:17-107), both computed ON THE DEVICE from the
// device depth image (upstream: on the host, then two H2D copies).  All sums in get_normal are
// small integers (|A| <= 200, |b| <= 2000, |det*d| < 2^27), so 32-bit integer arithmetic is exact
// and equals the reference's `long` arithmetic; the three float operations that follow use
// non-contractable IEEE ops.  Results equal the reference's CPU arrays bit for bit.
//
// Nearest-neighbour scene (Scene_nn::init_Scene_nn_cuda, pcd_scene.cu:3-20 -> pcd_scene.cpp:4-184):
// normals and back-projection on the device, compaction + kd-tree build on the host with the
// reference's level-by-level midpoint-split algorithm, so the 52-byte node array is
// interchangeable with KDTree_cpu's.
#include "common.cuh"
#include <float.h>
#include <limits.h>
#include <vector>
#include <numeric>

namespace prb {

struct SceneK { float fx, fy, cx, cy; };

template <class T> __device__ __forceinline__ int depth_as_u16(T d);
template <> __device__ __forceinline__ int depth_as_u16<uint16_t>(uint16_t d) { return d; }
// cv::Mat::convertTo(CV_16U) saturates (common.cpp:22-23)
template <> __device__ __forceinline__ int depth_as_u16<int32_t>(int32_t d) { return d < 0 ? 0 : (d > 65535 ? 65535 : d); }

// one thread per pixel; grid covers the whole image
template <class T>
__global__ void __launch_bounds__(256)
scene_prep_kernel(const T* __restrict__ depth, int W, int H, SceneK K, int saturate_pcd,
                  float* __restrict__ pcd, float* __restrict__ normal) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t idx = (size_t)y * W + x;
    const T raw = depth[idx];

    // ---- dep2pcd (common.h:47-61); depth_scene.cpp:26 reads CV_32S through at<uint32_t>
    if (pcd) {
        float px = 0.f, py = 0.f, pz = 0.f;
        const unsigned dep = saturate_pcd ? (unsigned)depth_as_u16<T>(raw) : (unsigned)raw;
        if (dep != 0) {
            pz = divf(__uint2float_rn(dep), 1000.0f);
            px = mulf(divf(subf((float)x, K.cx), K.fx), pz);
            py = mulf(divf(subf((float)y, K.cy), K.fy), pz);
        }
        pcd[3 * idx + 0] = px; pcd[3 * idx + 1] = py; pcd[3 * idx + 2] = pz;
    }

    // ---- get_normal (common.cpp:17-107)
    if (normal) {
        float nx = 0.f, ny = 0.f, nz = 0.f;
        const int r = 5;
        if (y >= r && y < H - r - 1 && x >= r && x < W - r - 1) {
            const int d = depth_as_u16<T>(raw);
            if (d < 2000) {
                int A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
#pragma unroll
                for (int j = -r; j <= r; j += r) {
#pragma unroll
                    for (int i = -r; i <= r; i += r) {
                        if (i == 0 && j == 0) continue;
                        const int delta = depth_as_u16<T>(depth[idx + (ptrdiff_t)j * W + i]) - d;
                        const int f = (abs(delta) < 50) ? 1 : 0;
                        A0 += f * i * i; A1 += f * i * j; A3 += f * j * j;
                        b0 += f * i * delta; b1 += f * j * delta;
                    }
                }
                const int det = A0 * A3 - A1 * A1;
                const int ddx = A3 * b0 - A1 * b1;
                const int ddy = -A1 * b0 + A0 * b1;
                float lx = mulf(K.fx, (float)ddx);
                float ly = mulf(K.fy, (float)ddy);
                float lz = (float)(-det * d);
                const float len = __fsqrt_rn(addf(addf(mulf(lx, lx), mulf(ly, ly)), mulf(lz, lz)));
                if (len > 0.f) {
                    const float inv = divf(1.0f, len);
                    nx = mulf(lx, inv); ny = mulf(ly, inv); nz = mulf(lz, inv);
                }
            }
        }
        normal[3 * idx + 0] = nx; normal[3 * idx + 1] = ny; normal[3 * idx + 2] = nz;
    }
}

int launch_scene_prep(const void* depth_dev, int is_i32, uint32_t W, uint32_t H, const float K[9], int saturate_pcd,
                      float* pcd_dev, float* normal_dev, cudaStream_t stream) {
    const SceneK k = {K[0], K[4], K[2], K[5]};
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    if (is_i32) scene_prep_kernel<int32_t><<<grid, 256, 0, stream>>>((const int32_t*)depth_dev, (int)W, (int)H, k, saturate_pcd, pcd_dev, normal_dev);
    else scene_prep_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)depth_dev, (int)W, (int)H, k, saturate_pcd, pcd_dev, normal_dev);
    count_launch();
    PR_LAUNCH_CHECK();
    return PR_OK;
}

// ---------------------------------------------------------------------------------------------
// kd-tree build on the host: KDTree_cpu::build_tree (pcd_scene.cpp:45-184).
// Breadth-first by generation; a node with more than max_leaf points gets: bbox of its points,
// widest axis (first wins), provisional split at the bbox midpoint, stable partition with the
// right part written from the back (reversed) and points equal to the split alternating sides,
// final split value = midpoint between the two sides' nearest coordinates.  Afterwards points and
// normals are permuted into leaf order.
// ---------------------------------------------------------------------------------------------
struct P3 { float c[3]; };

static void kdtree_build_host(std::vector<P3>& pts, std::vector<P3>& nrm, std::vector<pr_node_kdtree>& nodes, int max_leaf) {
    const int n = (int)pts.size();
    std::vector<int> order(n), scratch(n);
    std::iota(order.begin(), order.end(), 0);
    auto fresh = []() {
        pr_node_kdtree nd;
        nd.parent = nd.child1 = nd.child2 = -1;
        nd.split_v = 0.f;
        for (float& b : nd.bbox) b = 0.f;
        nd.split_dim = 0; nd.left = 0; nd.right = 0;
        return nd;
    };
    nodes.clear();
    nodes.push_back(fresh());
    nodes[0].right = n;
    size_t level_begin = 0, level_end = 1;
    while (level_begin < level_end) {
        for (size_t ni = level_begin; ni < level_end; ni++) {
            const int lo = nodes[ni].left, hi = nodes[ni].right;
            if (hi - lo <= max_leaf) continue;
            float lo3[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi3[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int k = lo; k < hi; k++)
                for (int a = 0; a < 3; a++) {
                    const float v = pts[order[k]].c[a];
                    if (v > hi3[a]) hi3[a] = v;
                    if (v < lo3[a]) lo3[a] = v;
                }
            int axis = 0; float cut = 0.f, best_span = -FLT_MAX;
            for (int a = 0; a < 3; a++) {
                const float span = hi3[a] - lo3[a];
                if (span > best_span) { best_span = span; axis = a; cut = (lo3[a] + hi3[a]) / 2; }
            }
            int front = lo, back = hi - 1;
            float left_max = -FLT_MAX, right_min = FLT_MAX;
            bool tie_left = true;
            for (int k = lo; k < hi; k++) {
                const float v = pts[order[k]].c[axis];
                if (v == cut) tie_left = !tie_left;
                if (v < cut || (v == cut && tie_left)) { scratch[front++] = order[k]; if (v > left_max) left_max = v; }
                else { scratch[back--] = order[k]; if (v < right_min) right_min = v; }
            }
            for (int k = lo; k < hi; k++) order[k] = scratch[k];
            const int c1 = (int)nodes.size();
            pr_node_kdtree a = fresh(), b = fresh();
            a.left = lo; a.right = front; a.parent = (int)ni;
            b.left = front; b.right = hi; b.parent = (int)ni;
            nodes.push_back(a); nodes.push_back(b);
            pr_node_kdtree& me = nodes[ni];
            me.child1 = c1; me.child2 = c1 + 1;
            me.split_v = (left_max + right_min) / 2;
            me.split_dim = axis;
            for (int ax = 0; ax < 3; ax++) { me.bbox[2 * ax] = lo3[ax]; me.bbox[2 * ax + 1] = hi3[ax]; }
        }
        level_begin = level_end;
        level_end = nodes.size();
    }
    std::vector<P3> tmp(n);
    for (int i = 0; i < n; i++) tmp[i] = pts[order[i]];
    pts.swap(tmp);
    for (int i = 0; i < n; i++) tmp[i] = nrm[order[i]];
    nrm.swap(tmp);
}

}  // namespace prb

using namespace prb;

extern "C" {

int pr_scene_projective_init(const void* depth_dev, int depth_is_int32, uint32_t width, uint32_t height,
                             const float K[9], float* pcd_dev, float* normal_dev, pr_stream_t stream) {
    if (!depth_dev || !K || (!pcd_dev && !normal_dev) || width == 0 || height == 0) return PR_ERR_INVALID_ARGUMENT;
    return launch_scene_prep(depth_dev, depth_is_int32, width, height, K, /*saturate_pcd=*/0, pcd_dev, normal_dev, as_stream(stream));
}

int pr_scene_nn_build_host(const void* depth_host, int depth_is_int32, uint32_t width, uint32_t height,
                           const float K[9], int max_leaf, float* pcd_host, float* normal_host, size_t capacity_points,
                           pr_node_kdtree* nodes_host, size_t capacity_nodes, size_t* n_points, size_t* n_nodes) {
    if (!depth_host || !K || !n_points || !n_nodes || width == 0 || height == 0 || max_leaf < 1) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)width * height;
    const size_t px_bytes = depth_is_int32 ? 4 : 2;
    void* d_depth = nullptr; float *d_pcd = nullptr, *d_nrm = nullptr;
    int rc = PR_OK;
    std::vector<float> pcd(n_px * 3), nrm(n_px * 3);
    cudaError_t e;
    if ((e = cudaMalloc(&d_depth, n_px * px_bytes)) != cudaSuccess) return (int)e;
    if ((e = cudaMalloc((void**)&d_pcd, n_px * 12)) != cudaSuccess) { cudaFree(d_depth); return (int)e; }
    if ((e = cudaMalloc((void**)&d_nrm, n_px * 12)) != cudaSuccess) { cudaFree(d_depth); cudaFree(d_pcd); return (int)e; }
    do {
        if ((e = cudaMemcpy(d_depth, depth_host, n_px * px_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) { rc = (int)e; break; }
        // pcd_scene.cpp:9-14: the CV_32S image is saturate-converted to CV_16U before anything else
        rc = launch_scene_prep(d_depth, depth_is_int32, width, height, K, /*saturate_pcd=*/1, d_pcd, d_nrm, 0);
        if (rc != PR_OK) break;
        if ((e = cudaMemcpy(pcd.data(), d_pcd, n_px * 12, cudaMemcpyDeviceToHost)) != cudaSuccess) { rc = (int)e; break; }
        if ((e = cudaMemcpy(nrm.data(), d_nrm, n_px * 12, cudaMemcpyDeviceToHost)) != cudaSuccess) { rc = (int)e; break; }
    } while (0);
    cudaFree(d_depth); cudaFree(d_pcd); cudaFree(d_nrm);
    if (rc != PR_OK) return rc;

    // pcd_scene.cpp:21-30: keep pixels with depth > 0, row-major
    std::vector<P3> pts, nr;
    pts.reserve(n_px); nr.reserve(n_px);
    for (size_t i = 0; i < n_px; i++) {
        int d;
        if (depth_is_int32) { const int32_t v = ((const int32_t*)depth_host)[i]; d = v < 0 ? 0 : (v > 65535 ? 65535 : v); }
        else d = ((const uint16_t*)depth_host)[i];
        if (d > 0) {
            pts.push_back({{pcd[3 * i], pcd[3 * i + 1], pcd[3 * i + 2]}});
            nr.push_back({{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]}});
        }
    }
    std::vector<pr_node_kdtree> nodes;
    if (!pts.empty()) kdtree_build_host(pts, nr, nodes, max_leaf);
    *n_points = pts.size();
    *n_nodes = nodes.size();
    if (!pcd_host && !normal_host && !nodes_host) return PR_OK;   // size query
    if (pts.size() > capacity_points || nodes.size() > capacity_nodes) return PR_ERR_CAPACITY;
    if (pcd_host) memcpy(pcd_host, pts.data(), pts.size() * 12);
    if (normal_host) memcpy(normal_host, nr.data(), nr.size() * 12);
    if (nodes_host) memcpy(nodes_host, nodes.data(), nodes.size() * sizeof(pr_node_kdtree));
    return PR_OK;
}

}  // extern "C"
