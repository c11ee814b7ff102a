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
// normals and back-projection on the device; compaction + kd-tree build either on the device
// (pr_scene_nn_build: kd_* kernels below, the same tree node for node) or on the host with the
// reference's level-by-level algorithm (pr_scene_nn_build_host, upstream's way).
#include "common.cuh"
#include <float.h>
#include <limits.h>
#include <vector>
#include <numeric>
#include <algorithm>

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

// ---------------------------------------------------------------------------------------------
// kd-tree build on the DEVICE: the same tree, node for node and point for point, as KDTree_cpu::build_tree
// (pcd_scene.cpp:45-184) / kdtree_build_host above.  The reference builds breadth-first by generation and
// the children of a generation are numbered in node order, so a generation is: (plan) prefix sum over its
// nodes of "is split" -> child ids; (split) one CTA per split node: bounding box by reduction, widest axis,
// provisional cut, then the stable two-sided partition expressed with prefix sums -- element i goes left iff
// v < cut or (v == cut and it is an even-numbered tie: lr_switch toggles BEFORE it is used), its slot is
// left + #left before it, or right-1-#right before it (the right part is written from the back, i.e. reversed).
// min / max are order independent, so bbox and the final cut equal the sequential ones bit for bit (a tie
// between -0 and +0 as an extreme could differ in sign; cannot occur for back-projected pixels, x = (u-cx)/fx*z).
// ---------------------------------------------------------------------------------------------
constexpr int kKdThreads = 1024;

__device__ __forceinline__ unsigned kd_block_excl_scan(unsigned v, unsigned* s_warp, unsigned* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned wprefix = 0, all = 0;
    for (int w = 0; w < kKdThreads / 32; w++) { const unsigned x = s_warp[w]; if (w < warp) wprefix += x; all += x; }
    *total = all;
    return wprefix + incl - v;
}
__device__ __forceinline__ float kd_block_min(float v, float* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float r = s_red[0];
    for (int w = 1; w < kKdThreads / 32; w++) r = fminf(r, s_red[w]);
    return r;
}

// valid pixels (saturated depth > 0) in row-major order -> compact points / normals, identity order; one CTA
template <class T>
__global__ void __launch_bounds__(kKdThreads)
kd_compact_kernel(const T* __restrict__ depth, unsigned n_px, const float* __restrict__ pcd, const float* __restrict__ nrm,
                  float* __restrict__ pts_c, float* __restrict__ nrm_c, int* __restrict__ order, unsigned capacity,
                  unsigned* __restrict__ n_points) {
    __shared__ unsigned s_warp[kKdThreads / 32];
    unsigned carry = 0;
    for (unsigned b = 0; b < n_px; b += kKdThreads) {
        const unsigned i = b + threadIdx.x;
        const bool valid = (i < n_px) && (depth_as_u16<T>(depth[i]) > 0);      // pcd_scene.cpp:9-30
        unsigned total;
        const unsigned slot = carry + kd_block_excl_scan(valid ? 1u : 0u, s_warp, &total);
        if (valid && slot < capacity) {
#pragma unroll
            for (int a = 0; a < 3; a++) { pts_c[3 * (size_t)slot + a] = pcd[3 * (size_t)i + a]; nrm_c[3 * (size_t)slot + a] = nrm[3 * (size_t)i + a]; }
            order[slot] = (int)slot;
        }
        carry += total;
    }
    if (threadIdx.x == 0) *n_points = carry;
}

struct KdCtl { unsigned n_points, n_nodes, new_nodes, pad; };

__global__ void kd_root_kernel(pr_node_kdtree* nodes, KdCtl* ctl) {
    pr_node_kdtree nd;
    nd.parent = nd.child1 = nd.child2 = -1; nd.split_v = 0.f;
    for (int i = 0; i < 6; i++) nd.bbox[i] = 0.f;
    nd.split_dim = 0; nd.left = 0; nd.right = (int)ctl->n_points;
    nodes[0] = nd;
    ctl->n_nodes = 1;
}

// one CTA: nodes [begin, end) of a generation -> child ids (children numbered in node order, pcd_scene.cpp:153-166)
__global__ void __launch_bounds__(kKdThreads)
kd_plan_kernel(pr_node_kdtree* __restrict__ nodes, unsigned begin, unsigned end, int max_leaf, unsigned capacity_nodes, KdCtl* ctl) {
    __shared__ unsigned s_warp[kKdThreads / 32];
    const unsigned base = ctl->n_nodes;
    unsigned carry = 0;
    for (unsigned b = begin; b < end; b += kKdThreads) {
        const unsigned ni = b + threadIdx.x;
        const bool split = (ni < end) && (nodes[ni].right - nodes[ni].left > max_leaf);
        unsigned total;
        const unsigned rank = carry + kd_block_excl_scan(split ? 1u : 0u, s_warp, &total);
        if (split) {
            const unsigned c1 = base + 2 * rank;
            if (c1 + 1 < capacity_nodes) { nodes[ni].child1 = (int)c1; nodes[ni].child2 = (int)c1 + 1; }
        }
        carry += total;
    }
    if (threadIdx.x == 0) { ctl->new_nodes = 2 * carry; ctl->n_nodes = base + 2 * carry; }
}

// one CTA per node of the generation; leaves return at once
__global__ void __launch_bounds__(kKdThreads)
kd_split_kernel(pr_node_kdtree* __restrict__ nodes, unsigned begin, int max_leaf, unsigned capacity_nodes,
                const float* __restrict__ pts, int* __restrict__ order, int* __restrict__ scratch) {
    __shared__ unsigned s_warp[kKdThreads / 32];
    __shared__ float s_red[kKdThreads / 32];
    __shared__ float s_box[6];
    __shared__ int s_axis;
    __shared__ float s_cut;
    const unsigned ni = begin + blockIdx.x;
    const int lo = nodes[ni].left, hi = nodes[ni].right;
    if (hi - lo <= max_leaf) return;
    const int c1 = nodes[ni].child1;
    if (c1 < 0) return;                                   // node capacity exhausted (reported by the host code)
    // ---- bounding box (pcd_scene.cpp:82-97)
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = lo + (int)threadIdx.x; k < hi; k += kKdThreads) {
        const float* p = pts + 3 * (size_t)order[k];
#pragma unroll
        for (int a = 0; a < 3; a++) { const float v = p[a]; if (v > mx[a]) mx[a] = v; if (v < mn[a]) mn[a] = v; }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float l = kd_block_min(mn[a], s_red);
        const float h = -kd_block_min(-mx[a], s_red);
        if (threadIdx.x == 0) { s_box[2 * a] = l; s_box[2 * a + 1] = h; }
    }
    if (threadIdx.x == 0) {
        // widest axis, first wins; provisional cut at the middle of the box (pcd_scene.cpp:99-113)
        int axis = 0; float cut = 0.f, best = -FLT_MAX;
        for (int a = 0; a < 3; a++) {
            const float span = subf(s_box[2 * a + 1], s_box[2 * a]);
            if (span > best) { best = span; axis = a; cut = divf(addf(s_box[2 * a], s_box[2 * a + 1]), 2.0f); }
        }
        s_axis = axis; s_cut = cut;
    }
    __syncthreads();
    const int axis = s_axis;
    const float cut = s_cut;
    // ---- stable two-sided partition (pcd_scene.cpp:115-137)
    unsigned carry_left = 0, carry_ties = 0;
    float left_max = -FLT_MAX, right_min = FLT_MAX;
    for (int b = lo; b < hi; b += kKdThreads) {
        const int k = b + (int)threadIdx.x;
        const bool in = k < hi;
        int id = 0; float v = 0.f;
        if (in) { id = order[k]; v = pts[3 * (size_t)id + axis]; }
        const bool tie = in && (v == cut);
        unsigned t_total, l_total;
        const unsigned tie_rank = carry_ties + kd_block_excl_scan(tie ? 1u : 0u, s_warp, &t_total) + 1;   // 1-based
        // lr_switch starts true and toggles before use: the 1st tie goes right, the 2nd left, ...
        const bool left = in && (v < cut || (tie && (tie_rank % 2 == 0)));
        const unsigned l_before = carry_left + kd_block_excl_scan(left ? 1u : 0u, s_warp, &l_total);
        if (in) {
            if (left) { scratch[lo + (int)l_before] = id; if (v > left_max) left_max = v; }
            else { scratch[hi - 1 - ((k - lo) - (int)l_before)] = id; if (v < right_min) right_min = v; }
        }
        carry_ties += t_total; carry_left += l_total;
    }
    const float lmax = -kd_block_min(-left_max, s_red);
    const float rmin = kd_block_min(right_min, s_red);
    __syncthreads();
    for (int k = lo + (int)threadIdx.x; k < hi; k += kKdThreads) order[k] = scratch[k];
    if (threadIdx.x == 0) {
        const int front = lo + (int)carry_left;
        pr_node_kdtree me = nodes[ni];
        me.split_v = divf(addf(lmax, rmin), 2.0f);
        me.split_dim = axis;
        for (int i = 0; i < 6; i++) me.bbox[i] = s_box[i];
        nodes[ni] = me;
        pr_node_kdtree a;
        a.parent = (int)ni; a.child1 = a.child2 = -1; a.split_v = 0.f;
        for (int i = 0; i < 6; i++) a.bbox[i] = 0.f;
        a.split_dim = 0;
        pr_node_kdtree bnode = a;
        a.left = lo; a.right = front;
        bnode.left = front; bnode.right = hi;
        nodes[c1] = a; nodes[c1 + 1] = bnode;
    }
}

__global__ void __launch_bounds__(256)
kd_gather_kernel(const float* __restrict__ pts_c, const float* __restrict__ nrm_c, const int* __restrict__ order, unsigned n,
                 float* __restrict__ pcd_out, float* __restrict__ nrm_out) {
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const size_t s = 3 * (size_t)order[i];
#pragma unroll
    for (int a = 0; a < 3; a++) { pcd_out[3 * (size_t)i + a] = pts_c[s + a]; nrm_out[3 * (size_t)i + a] = nrm_c[s + a]; }
}

inline size_t scene_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

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

size_t pr_scene_nn_build_workspace_bytes(uint32_t width, uint32_t height) {
    const size_t n_px = (size_t)width * height;
    // organised cloud + normals, compacted cloud + normals, order + scratch, control block
    return 4 * scene_align_up(n_px * 12, 256) + 2 * scene_align_up(n_px * 4, 256) + 256;
}

int pr_scene_nn_build(const void* depth_dev, int depth_is_int32, uint32_t width, uint32_t height, const float K[9], int max_leaf,
                      float* pcd_dev, float* normal_dev, size_t capacity_points, pr_node_kdtree* nodes_dev, size_t capacity_nodes,
                      size_t* n_points, size_t* n_nodes, void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    if (!depth_dev || !K || !pcd_dev || !normal_dev || !nodes_dev || !n_points || !n_nodes || !workspace_dev) return PR_ERR_INVALID_ARGUMENT;
    if (width == 0 || height == 0 || max_leaf < 1 || capacity_nodes < 1) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)width * height;
    if (n_px > 0x7FFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < pr_scene_nn_build_workspace_bytes(width, height)) return PR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = as_stream(stream_);
    char* w = (char*)workspace_dev;
    const size_t a3 = scene_align_up(n_px * 12, 256), a1 = scene_align_up(n_px * 4, 256);
    float* pcd_full = (float*)w; float* nrm_full = (float*)(w + a3);
    float* pts_c = (float*)(w + 2 * a3); float* nrm_c = (float*)(w + 3 * a3);
    int* order = (int*)(w + 4 * a3); int* scratch = (int*)(w + 4 * a3 + a1);
    KdCtl* ctl = (KdCtl*)(w + 4 * a3 + 2 * a1);
    // pcd_scene.cpp:9-14: the CV_32S image is saturate-converted to CV_16U before anything else
    int rc = launch_scene_prep(depth_dev, depth_is_int32, width, height, K, /*saturate_pcd=*/1, pcd_full, nrm_full, stream);
    if (rc != PR_OK) return rc;
    if (depth_is_int32) kd_compact_kernel<int32_t><<<1, kKdThreads, 0, stream>>>((const int32_t*)depth_dev, (unsigned)n_px, pcd_full, nrm_full, pts_c, nrm_c, order, (unsigned)n_px, &ctl->n_points);
    else kd_compact_kernel<uint16_t><<<1, kKdThreads, 0, stream>>>((const uint16_t*)depth_dev, (unsigned)n_px, pcd_full, nrm_full, pts_c, nrm_c, order, (unsigned)n_px, &ctl->n_points);
    count_launch();
    KdCtl h;
    PR_CUDA_TRY(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, stream));
    PR_CUDA_TRY(cudaStreamSynchronize(stream));
    const size_t n = h.n_points;
    *n_points = n; *n_nodes = 0;
    if (n == 0) return PR_OK;
    if (n > capacity_points) return PR_ERR_CAPACITY;
    kd_root_kernel<<<1, 1, 0, stream>>>(nodes_dev, ctl);
    count_launch();
    unsigned begin = 0, end = 1;
    while (begin < end) {                                   // one generation per round (pcd_scene.cpp:68-175)
        kd_plan_kernel<<<1, kKdThreads, 0, stream>>>(nodes_dev, begin, end, max_leaf, (unsigned)std::min<size_t>(capacity_nodes, 0xFFFFFFFFu), ctl);
        PR_CUDA_TRY(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, stream));
        PR_CUDA_TRY(cudaStreamSynchronize(stream));
        if (h.n_nodes > capacity_nodes) { *n_nodes = h.n_nodes; return PR_ERR_CAPACITY; }
        if (h.new_nodes) kd_split_kernel<<<end - begin, kKdThreads, 0, stream>>>(nodes_dev, begin, max_leaf, (unsigned)capacity_nodes, pts_c, order, scratch);
        count_launch(2);
        begin = end; end = h.n_nodes;
    }
    kd_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pts_c, nrm_c, order, (unsigned)n, pcd_dev, normal_dev);
    count_launch();
    PR_LAUNCH_CHECK();
    PR_CUDA_TRY(cudaStreamSynchronize(stream));
    *n_nodes = h.n_nodes;
    return PR_OK;
}

}  // extern "C"
