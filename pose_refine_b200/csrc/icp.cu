// icp.cu -- batched point-to-plane ICP for sm_100a.
//
// Replaces, for a ragged batch of pose hypotheses and entirely on the device:
//   ICP_Point2Plane_cuda<Scene>     cuda_icp/icp.cu:156-223   (iteration driver)
//   thrust__pcd2Ab<Scene>           cuda_icp/icp.h:128-209    (per-point 29-float term)
//   Scene_projective::query         scene/depth_scene/depth_scene.h:30-48 + common.h:63-73
//   Scene_nn::query                 scene/pcd_scene/pcd_scene.h:61-136
//   transform_pcd_cuda              cuda_icp/icp.cu:142-153   (fused away: see below)
//   eigen_slover_666                cuda_icp/icp.cpp:29-45    (solver.cuh)
//
// Upstream runs, per hypothesis and per iteration: a thrust::transform_reduce (two cub kernels +
// a cudaMalloc/cudaFree), a stream sync, a 116-byte D2H copy, a host Eigen solve, a transform
// kernel that rewrites every point, another sync.  Here one launch per pass covers ALL hypotheses:
// every CTA owns a chunk of one hypothesis' points, applies that hypothesis' ACCUMULATED 4x4 to
// the ORIGINAL points on the fly (points are only read: 12 B/point/pass), looks up the
// correspondence, accumulates the 29 sums in registers, reduces them with a register-transposing
// warp butterfly + shared memory, and deposits one partial per chunk.  The last CTA of a
// hypothesis to deposit (ticket counter) adds the partials in chunk order, evaluates
// fitness / rmse / the stop tests exactly as icp.cu:181-194 and solves the 6x6 system on the
// spot -- no host round trip anywhere in the loop, deterministic summation order.
//
// Two drivers share those device functions:
//   * icp_persistent_kernel (default): ONE launch for all passes of all hypotheses.  Work items
//     (pass, chunk) are claimed in order from a global counter; an item of pass p waits (acquire
//     spin by one thread) until its hypothesis has finished pass p-1 -- a per-hypothesis flag
//     replaces the per-pass kernel boundary, so there is no per-pass tail and the solve of one
//     hypothesis overlaps the point work of the others.  Point tiles are staged global -> shared
//     with TMA bulk copies (cp.async.bulk + mbarrier), double buffered and prefetched across items
//     (the points never change, only the 4x4 does); each thread pulls four points with three
//     128-bit shared loads and keeps four scene gathers in flight.  The projective scene is
//     repacked once per call into two 16-byte-aligned float4 per pixel so a correspondence is two
//     128-bit loads instead of six scalar ones.
//   * icp_pass_kernel: one launch per pass (first generation); kept for pr_pcd2ab_* and as the
//     cross-check (PR_ICP_IMPL=pass).
#include "common.cuh"
#include "solver.cuh"
#include <float.h>
#include <limits.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace prb {

constexpr int kIcpThreads = 256;
constexpr int kIcpWarps = kIcpThreads / 32;
constexpr int kPartialStride = 32;   // floats per chunk partial (29 used)

struct alignas(128) HypState {
    float T[12];            // accumulated transform, rows 0..2 (row 3 = 0 0 0 1)
    float fitness, rmse;    // values of the previous pass ("backup", icp.cu:179)
    int done;               // hypothesis has returned
    int pass;               // passes evaluated so far (= upstream's `iter`)
    unsigned arrived;       // chunk CTAs that deposited in the current pass
    unsigned n_chunks;
    unsigned chunk_begin;   // first chunk id of this hypothesis
    unsigned pad[9];
};
static_assert(sizeof(HypState) == 128, "HypState");

struct ProjScene {
    int W, H;
    float fW, fH;
    float max_dist;
    float fx, fy, cx, cy;
    const float* pcd;
    const float* nrm;
};
struct NnScene {
    float max_dist_sq;
    const float* pcd;
    const float* nrm;
    const pr_node_kdtree* nodes;
    int n_nodes;
};

struct Corr { float qx, qy, qz, nx, ny, nz; };

// Scene_projective::query (depth_scene.h:30-48).  The pixel selection uses non-contractable ops so
// that, for equal p, it picks the same pixel as the CPU build.  int(v) of pcd2dep (common.h:63-73)
// is truncation; "0 <= int(v) < W" is tested in the float domain as -1 < v < W, which is the same
// set for finite v and also rejects NaN / out-of-int-range values (x86 gives INT_MIN there).
__device__ __forceinline__ bool query(const ProjScene& s, float px, float py, float pz, Corr& c) {
    const float uf = addf(addf(mulf(divf(px, pz), s.fx), s.cx), 0.5f);
    const float vf = addf(addf(mulf(divf(py, pz), s.fy), s.cy), 0.5f);
    if (!(uf > -1.0f && uf < s.fW && vf > -1.0f && vf < s.fH)) return false;
    const size_t idx = (size_t)(int)uf + (size_t)(int)vf * (size_t)s.W;
    const float* q = s.pcd + 3 * idx;
    c.qx = __ldg(q); c.qy = __ldg(q + 1); c.qz = __ldg(q + 2);
    const float dz = pz - c.qz;
    const float adz = (dz > 0.f) ? dz : -dz;
    if (c.qz <= 0.f || adz > s.max_dist) return false;
    const float* n = s.nrm + 3 * idx;
    c.nx = __ldg(n); c.ny = __ldg(n + 1); c.nz = __ldg(n + 2);
    return true;
}

// Scene_nn::query (pcd_scene.h:61-136): the reference's stackless descend / backtrack walk over the
// 52-byte nodes, including its pruning rule (distance to the RE-VISITED node's box) and its
// strict-< tie rule (first visited wins).
__device__ __forceinline__ bool query(const NnScene& s, float px, float py, float pz, Corr& c) {
    if (s.n_nodes <= 0) return false;
    bool backtrack = false;
    int last = -1, cur = 0, best = 0;
    float best_d2 = FLT_MAX;
    while (cur >= 0) {
        const pr_node_kdtree* nd = s.nodes + cur;
        const int child1 = __ldg(&nd->child1), child2 = __ldg(&nd->child2);
        if (!backtrack) {
            if (child1 < 0 || child2 < 0) {
                const int lo = __ldg(&nd->left), hi = __ldg(&nd->right);
                for (int i = lo; i < hi; i++) {
                    const float dx = px - __ldg(s.pcd + 3 * i), dy = py - __ldg(s.pcd + 3 * i + 1), dz = pz - __ldg(s.pcd + 3 * i + 2);
                    const float d2 = addf(addf(mulf(dx, dx), mulf(dy, dy)), mulf(dz, dz));
                    if (d2 < best_d2) { best_d2 = d2; best = i; }
                }
                backtrack = true; last = cur; cur = __ldg(&nd->parent);
            } else {
                const int dim = __ldg(&nd->split_dim);
                const float sv = __ldg(&nd->split_v);
                const float diff = (dim == 0 ? px : (dim == 1 ? py : pz)) - sv;
                last = cur; cur = (diff < 0.f) ? child1 : child2;
            }
        } else {
            const int dim = __ldg(&nd->split_dim);
            const float sv = __ldg(&nd->split_v);
            const float diff = (dim == 0 ? px : (dim == 1 ? py : pz)) - sv;
            const int near_child = (diff < 0.f) ? child1 : child2;
            const int far_child = (diff < 0.f) ? child2 : child1;
            float lb = 0.f;
            const float b0 = __ldg(&nd->bbox[0]), b1 = __ldg(&nd->bbox[1]), b2 = __ldg(&nd->bbox[2]);
            const float b3 = __ldg(&nd->bbox[3]), b4 = __ldg(&nd->bbox[4]), b5 = __ldg(&nd->bbox[5]);
            if (px < b0) lb = addf(lb, mulf(b0 - px, b0 - px)); else if (px > b1) lb = addf(lb, mulf(b1 - px, b1 - px));
            if (py < b2) lb = addf(lb, mulf(b2 - py, b2 - py)); else if (py > b3) lb = addf(lb, mulf(b3 - py, b3 - py));
            if (pz < b4) lb = addf(lb, mulf(b4 - pz, b4 - pz)); else if (pz > b5) lb = addf(lb, mulf(b5 - pz, b5 - pz));
            if (last == near_child && lb <= best_d2) { last = cur; cur = far_child; backtrack = false; }
            else { last = cur; cur = __ldg(&nd->parent); }
        }
    }
    if (!(best_d2 < s.max_dist_sq)) return false;
    c.qx = __ldg(s.pcd + 3 * best); c.qy = __ldg(s.pcd + 3 * best + 1); c.qz = __ldg(s.pcd + 3 * best + 2);
    c.nx = __ldg(s.nrm + 3 * best); c.ny = __ldg(s.nrm + 3 * best + 1); c.nz = __ldg(s.nrm + 3 * best + 2);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Packed kd-tree for the persistent driver.  Same tree (same nodes, same leaf ranges, same points) as
// the reference's Node_kdtree array, re-laid out per ICP call so that a node is two aligned float4:
//     {lo.x, lo.y, lo.z, a}   {hi.x, hi.y, hi.z, unused}
// a >= 0: internal node, children a and a+1 (build_tree appends them together, pcd_scene.cpp:160-170);
// a <  0: leaf, a = 0x80000000 | count << 24 | left.  [lo,hi] is the box of the node's OWN points --
// computed here for leaves too (the reference stores none for leaves, pcd_scene.h:14-19).
// The query is an exact nearest-neighbour search like Scene_nn::query, but it prunes with the box of
// the CHILD it is about to enter (the reference prunes with the box of the node it re-visits, which is
// much weaker: 385 node visits per query on the fixture vs ~20-40 here), starts
// from best = max_dist^2 (anything farther is invalid anyway, pcd_scene.h:127) and keeps the far
// children on a small explicit stack.  Distances use the reference's operation order, so the winner is
// the same point except for exact distance ties between points of different leaves.
struct PackedNnScene {
    float max_dist_sq;
    const float4* nodes;      // 2 per node
    const float4* pts4;       // {x, y, z, 0}
    const float* nrm;         // original Vec3f normals
    int n_nodes;
    NnScene ref;              // the reference layout (fallback walk when the stack would overflow)
};

__global__ void __launch_bounds__(256)
nn_pack_points_kernel(const float* __restrict__ pcd, size_t n, float4* __restrict__ pts4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) pts4[i] = make_float4(pcd[3 * i], pcd[3 * i + 1], pcd[3 * i + 2], 0.f);
}
// sets *unsupported when a leaf does not fit the packed encoding (more than 127 points or left >= 2^24)
__global__ void __launch_bounds__(256)
nn_pack_nodes_kernel(const pr_node_kdtree* __restrict__ nodes, int n_nodes, const float* __restrict__ pcd,
                     float4* __restrict__ out, unsigned* __restrict__ unsupported) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_nodes) return;
    const pr_node_kdtree nd = nodes[i];
    float lo[3], hi[3];
    int a;
    if (nd.child1 < 0 || nd.child2 < 0) {
        const int cnt = nd.right - nd.left;
        if (cnt < 0 || cnt > 127 || nd.left < 0 || nd.left >= (1 << 24)) { *unsupported = 1; return; }
        for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
        for (int j = nd.left; j < nd.right; j++)
            for (int k = 0; k < 3; k++) { const float v = pcd[3 * j + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
        a = (int)(0x80000000u | ((unsigned)cnt << 24) | (unsigned)nd.left);
    } else {
        if (nd.child2 != nd.child1 + 1) { *unsupported = 1; return; }
        for (int k = 0; k < 3; k++) { lo[k] = nd.bbox[2 * k]; hi[k] = nd.bbox[2 * k + 1]; }
        a = nd.child1;
    }
    out[2 * i] = make_float4(lo[0], lo[1], lo[2], __int_as_float(a));
    out[2 * i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

__device__ __forceinline__ float box_dist_sq(const float4& lo, const float4& hi, float px, float py, float pz) {
    const float dx = fmaxf(fmaxf(lo.x - px, px - hi.x), 0.f);
    const float dy = fmaxf(fmaxf(lo.y - py, py - hi.y), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - pz, pz - hi.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ bool query(const PackedNnScene& s, float px, float py, float pz, Corr& c) {
    if (s.n_nodes <= 0) return false;
    constexpr int kStack = 40;
    int stack_n[kStack];
    float stack_lb[kStack];
    int sp = 0;
    float best = s.max_dist_sq;
    int best_i = -1;
    bool overflow = false;
    float4 lo = __ldg(s.nodes), hi = __ldg(s.nodes + 1);
    // a box lower bound is rounded, so it is trusted only with a 1e-5 margin
    bool go = box_dist_sq(lo, hi, px, py, pz) * 0.99999f < best;
    while (go) {
        const int a = __float_as_int(lo.w);
        if (a < 0) {
            const int left = a & 0xFFFFFF, cnt = (a >> 24) & 127;
            for (int i = left; i < left + cnt; i++) {
                const float4 q = __ldg(s.pts4 + i);
                const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
                const float d2 = addf(addf(mulf(dx, dx), mulf(dy, dy)), mulf(dz, dz));    // pcd_scene.h:86-89
                if (d2 < best) { best = d2; best_i = i; }
            }
            go = false;
        } else {
            const float4 lo1 = __ldg(s.nodes + 2 * a), hi1 = __ldg(s.nodes + 2 * a + 1);
            const float4 lo2 = __ldg(s.nodes + 2 * a + 2), hi2 = __ldg(s.nodes + 2 * a + 3);
            const float lb1 = box_dist_sq(lo1, hi1, px, py, pz) * 0.99999f, lb2 = box_dist_sq(lo2, hi2, px, py, pz) * 0.99999f;
            const bool first1 = lb1 <= lb2;
            const float lb_near = first1 ? lb1 : lb2, lb_far = first1 ? lb2 : lb1;
            if (lb_far < best) {
                if (sp < kStack) { stack_n[sp] = first1 ? a + 1 : a; stack_lb[sp] = lb_far; sp++; }
                else overflow = true;
            }
            if (lb_near < best) { lo = first1 ? lo1 : lo2; hi = first1 ? hi1 : hi2; continue; }
            go = false;
        }
        while (sp > 0) {
            --sp;
            if (stack_lb[sp] < best) {
                const int n = stack_n[sp];
                lo = __ldg(s.nodes + 2 * n); hi = __ldg(s.nodes + 2 * n + 1);
                go = true;
                break;
            }
        }
    }
    if (overflow) return query(s.ref, px, py, pz, c);     // deeper than the stack: the reference walk
    if (best_i < 0) return false;
    const float4 q = __ldg(s.pts4 + best_i);
    c.qx = q.x; c.qy = q.y; c.qz = q.z;
    c.nx = __ldg(s.nrm + 3 * best_i); c.ny = __ldg(s.nrm + 3 * best_i + 1); c.nz = __ldg(s.nrm + 3 * best_i + 2);
    return true;
}

// thrust__pcd2Ab::operator() (icp.h:138-208): adds one correspondence into the 29 running sums.
__device__ __forceinline__ void accumulate(float* acc, float px, float py, float pz, const Corr& c, float w = 1.0f) {
    const float dx = c.qx - px, dy = c.qy - py, dz = c.qz - pz;
    const float r = dx * c.nx + dy * c.ny + dz * c.nz;
    float J[6];
    J[0] = c.nz * py - c.ny * pz;
    J[1] = c.nx * pz - c.nz * px;
    J[2] = c.ny * px - c.nx * py;
    J[3] = c.nx; J[4] = c.ny; J[5] = c.nz;
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) { acc[k] = fmaf(J[i], J[j], acc[k]); k++; }
#pragma unroll
    for (int i = 0; i < 6; i++) acc[21 + i] = fmaf(J[i], r, acc[21 + i]);
    acc[27] += dx * dx + dy * dy + dz * dz;
    acc[28] += w;
}

// ---- packed accumulation (sm_100 FFMA2) ------------------------------------------------------------
// Blackwell has a two-wide FP32 FMA (PTX fma.rn.f32x2, SASS FFMA2) whose first multiplicand may be a
// scalar broadcast.  The 21 + 6 products J_i*J_j, J_i*r are rows "J_i x (J_i..J_5, r)", so with J and r
// parked in the pairs E0=(J0,J1) E1=(J2,J3) E2=(J4,J5) E3=(r,0) they take 18 FFMA2 instead of 27 FFMA
// (row 1, 3, 5 start on an odd element: that lane recomputes the symmetric product and is ignored).
struct Acc2 {
    float2 p[18];      // see unpack_acc2 for the slot -> Vec29f index map
    float2 dd;         // sum dx^2, sum dy^2
    float dz2, cnt;
};
__device__ __forceinline__ float2 ffma2(float a, float2 b, float2 c) {     // a * b + c, a broadcast
    float2 d;
    asm("{\n"
        ".reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %2};\n"
        "mov.b64 rb, {%3, %4};\n"
        "mov.b64 rc, {%5, %6};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;\n"
        "}\n" : "=f"(d.x), "=f"(d.y) : "f"(a), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 ffma2v(float2 a, float2 b, float2 c) {   // element-wise a * b + c
    float2 d;
    asm("{\n"
        ".reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %5};\n"
        "mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;\n"
        "}\n" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ void zero_acc2(Acc2& a) {
#pragma unroll
    for (int i = 0; i < 18; i++) a.p[i] = make_float2(0.f, 0.f);
    a.dd = make_float2(0.f, 0.f); a.dz2 = 0.f; a.cnt = 0.f;
}
__device__ __forceinline__ void accumulate2(Acc2& a, float px, float py, float pz, const Corr& c, float w = 1.0f) {
    const float dx = c.qx - px, dy = c.qy - py, dz = c.qz - pz;
    const float r = dx * c.nx + dy * c.ny + dz * c.nz;
    const float2 E0 = make_float2(c.nz * py - c.ny * pz, c.nx * pz - c.nz * px);
    const float2 E1 = make_float2(c.ny * px - c.nx * py, c.nx);
    const float2 E2 = make_float2(c.ny, c.nz);
    const float2 E3 = make_float2(r, 0.f);
    a.p[0] = ffma2(E0.x, E0, a.p[0]); a.p[1] = ffma2(E0.x, E1, a.p[1]); a.p[2] = ffma2(E0.x, E2, a.p[2]); a.p[3] = ffma2(E0.x, E3, a.p[3]);
    a.p[4] = ffma2(E0.y, E0, a.p[4]); a.p[5] = ffma2(E0.y, E1, a.p[5]); a.p[6] = ffma2(E0.y, E2, a.p[6]); a.p[7] = ffma2(E0.y, E3, a.p[7]);
    a.p[8] = ffma2(E1.x, E1, a.p[8]); a.p[9] = ffma2(E1.x, E2, a.p[9]); a.p[10] = ffma2(E1.x, E3, a.p[10]);
    a.p[11] = ffma2(E1.y, E1, a.p[11]); a.p[12] = ffma2(E1.y, E2, a.p[12]); a.p[13] = ffma2(E1.y, E3, a.p[13]);
    a.p[14] = ffma2(E2.x, E2, a.p[14]); a.p[15] = ffma2(E2.x, E3, a.p[15]);
    a.p[16] = ffma2(E2.y, E2, a.p[16]); a.p[17] = ffma2(E2.y, E3, a.p[17]);
    a.dd = ffma2v(make_float2(dx, dy), make_float2(dx, dy), a.dd);
    a.dz2 = fmaf(dz, dz, a.dz2);
    a.cnt += w;
}
// packed slots -> the 29 sums in thrust__pcd2Ab's order (icp.h:165-206), padded to 32
__device__ __forceinline__ void unpack_acc2(const Acc2& a, float (&v)[32]) {
    v[0] = a.p[0].x;  v[1] = a.p[0].y;  v[2] = a.p[1].x;  v[3] = a.p[1].y;  v[4] = a.p[2].x;  v[5] = a.p[2].y;   // J0 * J0..J5
    v[6] = a.p[4].y;  v[7] = a.p[5].x;  v[8] = a.p[5].y;  v[9] = a.p[6].x;  v[10] = a.p[6].y;                    // J1 * J1..J5
    v[11] = a.p[8].x; v[12] = a.p[8].y; v[13] = a.p[9].x; v[14] = a.p[9].y;                                      // J2 * J2..J5
    v[15] = a.p[11].y; v[16] = a.p[12].x; v[17] = a.p[12].y;                                                     // J3 * J3..J5
    v[18] = a.p[14].x; v[19] = a.p[14].y;                                                                        // J4 * J4..J5
    v[20] = a.p[16].y;                                                                                           // J5 * J5
    v[21] = a.p[3].x; v[22] = a.p[7].x; v[23] = a.p[10].x; v[24] = a.p[13].x; v[25] = a.p[15].x; v[26] = a.p[17].x;   // J * r
    v[27] = a.dd.x + a.dd.y + a.dz2;
    v[28] = a.cnt;
    v[29] = 0.f; v[30] = 0.f; v[31] = 0.f;
}


// ---- two points per instruction ---------------------------------------------------------------------
// The second generation of the packed path: instead of packing two SUMS of one point into an FFMA2
// (which needs register moves to form the operand pairs), every quantity of the point pipeline is a
// pair (value for point A, value for point B) of the two points a lane processes together --
// transform, projection, residual, Jacobian and all 29 sums run as FFMA2 / FMUL2 / FADD2 with no
// packing moves: the per-point selects that reject a correspondence write straight into the halves of
// the pair registers.  Sum i is kept as (sum over "A" points, sum over "B" points) and folded at the end
// of the item.  Pairs are carried as 64-bit values so that ptxas allocates them as aligned register
// pairs once; it folds negation and scalar broadcast into the FFMA2 operands.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2_t bc2(float a) { return pk2(a, a); }
__device__ __forceinline__ void unpk2(f2_t a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// acc += a * b with the accumulator as a read-write operand: input and output are the same register pair by
// construction, so ptxas has no loop-carried copies to insert at the back edge of the group loop
__device__ __forceinline__ void fma2_acc(f2_t& acc, f2_t a, f2_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t sub2(f2_t a, f2_t b) { f2_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t neg2(f2_t a) {
    f2_t d;
    asm("{\n.reg .f32 l, h;\nmov.b64 {l, h}, %1;\nneg.f32 l, l;\nneg.f32 h, h;\nmov.b64 %0, {l, h};\n}" : "=l"(d) : "l"(a));
    return d;
}
struct AccP { f2_t s[28]; float cnt_a, cnt_b; };   // s[i] = Vec29f entry i (icp.h:165-206) as (sum over A points, sum over B points)
__device__ __forceinline__ void acc_zero(AccP& a) {
#pragma unroll
    for (int i = 0; i < 28; i++) a.s[i] = pk2(0.f, 0.f);
    a.cnt_a = 0.f; a.cnt_b = 0.f;
}
__device__ __forceinline__ void acc_unpack(const AccP& a, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 28; i++) { float lo, hi; unpk2(a.s[i], lo, hi); v[i] = lo + hi; }
    v[28] = a.cnt_a + a.cnt_b;
    v[29] = 0.f; v[30] = 0.f; v[31] = 0.f;
}
// thrust__pcd2Ab::operator() (icp.h:138-208) for two points at once; a rejected point arrives as q = p, n = 0
// (so d = 0, r = 0, J = 0: all 28 float sums get +0); the caller counts the accepted points.
__device__ __forceinline__ void accumulate_pair(AccP& a, f2_t px, f2_t py, f2_t pz, f2_t qx, f2_t qy, f2_t qz,
                                                f2_t nx, f2_t ny, f2_t nz) {
    const f2_t dx = sub2(qx, px), dy = sub2(qy, py), dz = sub2(qz, pz);
    const f2_t r = fma2(dz, nz, fma2(dy, ny, mul2(dx, nx)));
    f2_t J[6];
    J[0] = fma2(nz, py, neg2(mul2(ny, pz)));
    J[1] = fma2(nx, pz, neg2(mul2(nz, px)));
    J[2] = fma2(ny, px, neg2(mul2(nx, py)));
    J[3] = nx; J[4] = ny; J[5] = nz;
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) { fma2_acc(a.s[k], J[i], J[j]); k++; }
#pragma unroll
    for (int i = 0; i < 6; i++) fma2_acc(a.s[21 + i], J[i], r);
    fma2_acc(a.s[27], dx, dx); fma2_acc(a.s[27], dy, dy); fma2_acc(a.s[27], dz, dz);
}

__device__ __forceinline__ void acc_zero(Acc2& a) { zero_acc2(a); }
__device__ __forceinline__ void acc_add(Acc2& a, float px, float py, float pz, const Corr& c) { accumulate2(a, px, py, pz, c); }
__device__ __forceinline__ void acc_unpack(const Acc2& a, float (&v)[32]) { unpack_acc2(a, v); }
typedef Acc2 AccT;     // per-point accumulation of the nearest-neighbour scenes; the projective driver uses AccP

// Warp reduction of 32 values per lane that leaves, in lane L, the warp-wide sum of value L:
// at each butterfly step a lane keeps one half of its values and ships the other half, so the
// whole thing costs 16+8+4+2+1 = 31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32]) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; i++) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// Cross-CTA signalling inside the persistent kernel.  Everything a waiter reads after the signal
// (HypState::T, the chunk partials) is read with ld.global.cg, i.e. from L2, and every producer
// publishes with a gpu-scope RELEASE (its earlier stores are in L2 before the flag / ticket is
// visible).  The waiter therefore polls with a RELAXED load: an acquire load would make ptxas add
// CCTL.IVALL -- an invalidate of the SM's whole L1 -- to every poll, which evicts the scene lines
// the other warps of the SM are gathering from (measured: L1 hit rate 10%, 17% of all stall samples
// on CCTL.IVALL).  The dependent loads are issued only after the poll's branch resolves and bypass L1.
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_add_release(unsigned* p, unsigned v) {
    unsigned old;
#ifdef PR_DBG_NOFENCE       // what-if: no release fence in front of the ticket (incorrect ordering, timing only)
    asm volatile("atom.add.relaxed.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
#else
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
#endif
    return old;
}

// The stop logic of one hypothesis after its sums S[29] are known (icp.cu:179-212), run by one
// thread.  Updates the state and, when the hypothesis returns, its result.  RELEASE = true is the
// persistent driver's flavour: state is read past L1 and `pass` / `done` are published with release
// stores after everything else, because other CTAs of the SAME launch are waiting on them.
template <bool RELEASE>
__device__ __noinline__ void finish_pass_impl(HypState* st, const float* S, unsigned n_points, pr_icp_criteria crit,
                                 pr_registration_result* res) {
    const float count = S[28], total = S[27];
    const int iter = RELEASE ? __ldcg(&st->pass) : st->pass;
    bool ret = false;
    float fitness = RELEASE ? __ldcg(&st->fitness) : st->fitness, rmse = RELEASE ? __ldcg(&st->rmse) : st->rmse;
    if (count == 0.f) {
        ret = true;                                            // icp.cu:183 (result keeps the previous values)
    } else {
        const float prev_fit = fitness, prev_rmse = rmse;
        fitness = divf(count, (float)n_points);                // icp.cu:185
        rmse = __fsqrt_rn(divf(total, count));                 // icp.cu:186
        if (iter == crit.max_iteration) ret = true;            // icp.cu:189
        else if (fabsf(fitness - prev_fit) < crit.relative_fitness && fabsf(rmse - prev_rmse) < crit.relative_rmse)
            ret = true;                                        // icp.cu:191-194
    }
    if (RELEASE) { __stcg(&st->fitness, fitness); __stcg(&st->rmse, rmse); }
    else { st->fitness = fitness; st->rmse = rmse; }
    float T[16];
#pragma unroll
    for (int i = 0; i < 12; i++) T[i] = RELEASE ? __ldcg(&st->T[i]) : st->T[i];
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
    if (!ret) {
        float Sr[29], E[16];
#pragma unroll
        for (int i = 0; i < 29; i++) Sr[i] = S[i];
        solve_666_unrolled(Sr, E);                             // unpack icp.cu:198-205 + solve icp.cu:207
        // result.transformation_ = extrinsic * result.transformation_ (icp.cu:212); geometry.h:107-111
        // sums each dot product from index 3 down to 0.
        float Tn[12];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float acc = 0.f;
#pragma unroll
                for (int k = 3; k >= 0; k--) acc = addf(acc, mulf(E[4 * i + k], T[4 * k + j]));
                Tn[4 * i + j] = acc;
            }
#pragma unroll
        for (int i = 0; i < 12; i++) { if (RELEASE) __stcg(&st->T[i], Tn[i]); else st->T[i] = Tn[i]; }
        if (RELEASE) st_release(reinterpret_cast<unsigned*>(&st->pass), (unsigned)(iter + 1));
        else st->pass = iter + 1;
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) res->transformation[i] = T[i];
        res->inlier_rmse = rmse; res->fitness = fitness;
        if (RELEASE) st_release(reinterpret_cast<unsigned*>(&st->done), 1u);
        else st->done = 1;
    }
}
__device__ __forceinline__ void finish_pass(HypState* st, const float* S, unsigned n_points, pr_icp_criteria crit,
                                            pr_registration_result* res) {
    finish_pass_impl<false>(st, S, n_points, crit, res);
}
__device__ __forceinline__ void finish_pass_release(HypState* st, const float* S, unsigned n_points, pr_icp_criteria crit,
                                                    pr_registration_result* res) {
    finish_pass_impl<true>(st, S, n_points, crit, res);
}

// plan: chunk table + state initialisation.  One CTA; n_hyp is at most a few thousand.
__global__ void __launch_bounds__(kIcpThreads)
icp_plan_kernel(const uint32_t* __restrict__ counts, uint32_t n_hyp, uint32_t chunk_points, HypState* __restrict__ state,
                uint32_t* __restrict__ chunk_hyp, uint32_t max_chunks, uint32_t* __restrict__ total_chunks,
                pr_registration_result* __restrict__ results, unsigned* __restrict__ next_item,
                const uint32_t* __restrict__ offsets = nullptr, uint4* __restrict__ chunk_info = nullptr) {
    __shared__ unsigned s_warp[kIcpWarps];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t b = 0; b < n_hyp; b += kIcpThreads) {
        const uint32_t h = b + threadIdx.x;
        const unsigned cnt = (h < n_hyp) ? counts[h] : 0u;
        const unsigned v = (cnt + chunk_points - 1) / chunk_points;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned wprefix = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kIcpWarps; w++) { const unsigned s = s_warp[w]; if (w < warp) wprefix += s; all += s; }
        const unsigned begin = s_carry + wprefix + incl - v;
        if (h < n_hyp) {
            HypState st;
#pragma unroll
            for (int i = 0; i < 12; i++) st.T[i] = (i % 5 == 0) ? 1.f : 0.f;
            st.fitness = 0.f; st.rmse = 0.f; st.pass = 0; st.arrived = 0;
            st.done = (cnt == 0) ? 1 : 0;   // empty cloud: count == 0 on the first pass (icp.cu:183)
            st.n_chunks = v; st.chunk_begin = begin;
#pragma unroll
            for (int i = 0; i < 9; i++) st.pad[i] = 0;
            state[h] = st;
            pr_registration_result r;
#pragma unroll
            for (int i = 0; i < 16; i++) r.transformation[i] = (i % 5 == 0) ? 1.f : 0.f;
            r.inlier_rmse = 0.f; r.fitness = 0.f;
            results[h] = r;
            const unsigned off = chunk_info ? offsets[h] : 0u;
            for (unsigned j = 0; j < v; j++) {
                if (begin + j >= max_chunks) break;
                chunk_hyp[begin + j] = h;
                // everything a worker needs to know about a chunk in one 16-byte record:
                // hypothesis, first point (absolute), number of points, points of the hypothesis
                if (chunk_info) chunk_info[begin + j] = make_uint4(h, off + j * chunk_points, min(chunk_points, cnt - j * chunk_points), cnt);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total_chunks = min(s_carry, max_chunks);
        if (next_item) *next_item = 0;
    }
}

// One pass over all hypotheses.  Persistent grid: every CTA walks the chunk table with a stride of
// gridDim.x; one chunk = up to chunk_points points of one hypothesis.
// out29 != nullptr: "reduce only" mode used by pr_pcd2ab_* (single hypothesis, identity transform).
template <class SceneT>
__global__ void __launch_bounds__(kIcpThreads, 3)
icp_pass_kernel(const float* __restrict__ pts, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts,
                const uint32_t* __restrict__ chunk_hyp, const uint32_t* __restrict__ total_chunks, uint32_t chunk_points,
                HypState* __restrict__ state, float* __restrict__ partials, SceneT scene, pr_icp_criteria crit,
                pr_registration_result* __restrict__ results, float* __restrict__ out29) {
    __shared__ float s_part[kIcpWarps][32];
    __shared__ float s_sum[32];
    __shared__ int s_last;
    const unsigned total = *total_chunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned c = blockIdx.x; c < total; c += gridDim.x) {
        const unsigned h = chunk_hyp[c];
        HypState* st = state + h;
        if (st->done) continue;     // uniform over the CTA
        float T[12];
#pragma unroll
        for (int i = 0; i < 12; i++) T[i] = st->T[i];
        const unsigned lc = c - st->chunk_begin;
        const unsigned n_h = counts[h];
        const unsigned first = lc * chunk_points;
        const unsigned n = min(chunk_points, n_h - first);
        const float* p0 = pts + 3 * ((size_t)offsets[h] + first);

        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = 0.f;
        for (unsigned i = threadIdx.x; i < n; i += kIcpThreads) {
            const float x = p0[3 * i], y = p0[3 * i + 1], z = p0[3 * i + 2];
            // transform_pcd_cuda (icp.cu:147-149) with the accumulated transform
            const float px = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
            const float py = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
            const float pz = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
            Corr cr;
            if (query(scene, px, py, pz, cr)) accumulate(acc, px, py, pz, cr);
        }
        const float mine = warp_transpose_reduce(acc);
        s_part[warp][lane] = mine;
        __syncthreads();
        if (warp == 0) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kIcpWarps; w++) s += s_part[w][lane];
            partials[(size_t)c * kPartialStride + lane] = s;
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(&st->arrived, 1u);
            s_last = (ticket == st->n_chunks - 1) ? 1 : 0;
        }
        __syncthreads();
        if (s_last && warp == 0) {
            // last CTA of this hypothesis: add the chunk partials in chunk order and finish the pass
            __threadfence();
            float s = 0.f;
            const unsigned cb = st->chunk_begin, nc = st->n_chunks;
            for (unsigned j = 0; j < nc; j++) s += __ldcg(partials + (size_t)(cb + j) * kPartialStride + lane);
            s_sum[lane] = s;
            __syncwarp();
            if (lane == 0) {
                st->arrived = 0;
                if (out29) {
                    for (int i = 0; i < 29; i++) out29[i] = s_sum[i];
                } else {
                    finish_pass(st, s_sum, n_h, crit, results + h);
                }
                __threadfence();
            }
        }
        __syncthreads();   // s_part / s_last are reused by the next chunk
    }
}

// ---------------------------------------------------------------------------------------------
// persistent driver: warps are independent workers
// ---------------------------------------------------------------------------------------------
// Every warp runs its own loop: claim a work item (pass, chunk of kPersistChunk points of one
// hypothesis) from the global counter, stream the chunk through its own double-buffered
// shared-memory tiles (TMA bulk copies issued by lane 0, completion on the warp's own mbarriers),
// reduce its 29 sums with the transposing butterfly and deposit one partial per item.  There is no
// CTA-wide barrier anywhere; the only cross-warp synchronisation is the per-hypothesis
// acquire/release on HypState::pass / done and the ticket counter.
#ifndef PR_WTILE
#define PR_WTILE 1024
#endif
#ifndef PR_CHUNK
#define PR_CHUNK 4096
#endif
// Tuning (measured on B200, 512 hypotheses x 31 passes, ICP only): ILP 4 / 2 CTAs per SM 2.83 ms;
// ILP 8 / 1 CTA of 256 threads 2.29 ms; ILP 8 / 384 threads 2.68 ms; tile 256 2.62 ms.  Eight gathers in
// flight per lane with few, register-rich warps beats more warps with fewer gathers each.
#ifndef PR_ILP
#define PR_ILP 8
#endif
#ifndef PR_MINB
#define PR_MINB 1
#endif
#ifndef PR_PTHREADS
#define PR_PTHREADS 256
#endif
constexpr int kPThreads = PR_PTHREADS;         // threads per CTA of the persistent kernel
constexpr int kPWarps = kPThreads / 32;
constexpr int kWTile = PR_WTILE;               // points per warp tile of the projective driver (12 KB at 1024)
// nearest-neighbour driver (measured, 512 hypotheses against a 99k-point tree): tile 512 / 2 CTAs per SM 181 ms,
// tile 256 / 2 CTAs 125 ms, tile 128 / 4 CTAs 102 ms -- the tree walk lives in L1, so every KB of shared memory
// given back to L1 and every extra resident warp counts, even at 64 registers per thread
#ifndef PR_NN_TILE
#define PR_NN_TILE 128
#endif
#ifndef PR_NN_MINB
#define PR_NN_MINB 4
#endif
constexpr int kWTileNn = PR_NN_TILE;           // nearest-neighbour scenes: small tiles, so that several CTAs fit an SM
constexpr int kWStages = 2;
constexpr uint32_t kPersistChunk = PR_CHUNK;   // points per work item of a large batch (see persist_chunk_points)
constexpr int kIlp = PR_ILP;                   // points per lane per group (gathers in flight per lane)
struct PackedScene;
template <class SceneT> struct TileOf { static constexpr int kPoints = std::is_same<SceneT, PackedScene>::value ? kWTile : kWTileNn; };
template <class SceneT> constexpr int persist_smem() { return kPWarps * kWStages * (TileOf<SceneT>::kPoints * 12) + kPWarps * kWStages * 8; }

struct IcpCtl {            // device-side control block
    unsigned next_item;    // work-item claim counter
    unsigned total_chunks;
    unsigned pad[30];
};

// packed projective scene: one 32-byte record per pixel (= one L2 sector per correspondence).  Two separate
// arrays ({qx,qy,qz,nx} 16 B + {ny,nz} 8 B) cost 3x the time: measured 2.08 ms vs 0.71 ms with the second gather removed.
struct PackedScene {
    int W, H;
    float fW, fH;
    float max_dist;
    float fx, fy, cx05, cy05;     // cx + 0.5, cy + 0.5
    const float4* rec;            // pixel i: rec[2i] = {qx,qy,qz,nx}, rec[2i+1] = {ny,nz,0,0}
};

__global__ void __launch_bounds__(256)
scene_pack_kernel(const float* __restrict__ pcd, const float* __restrict__ nrm, size_t n_px, float4* __restrict__ rec) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_px) return;
    rec[2 * i] = make_float4(pcd[3 * i], pcd[3 * i + 1], pcd[3 * i + 2], nrm[3 * i]);
    rec[2 * i + 1] = make_float4(nrm[3 * i + 1], nrm[3 * i + 2], 0.f, 0.f);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
// TMA bulk copy global -> shared (1-D), completion signalled on an mbarrier
__device__ __forceinline__ void tma_load_1d(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// one scene record with a single 256-bit load (LDG.E.256, sm_100); two 128-bit loads of the same sector: +5 %
__device__ __forceinline__ void load_rec(const PackedScene& s, int idx, float4& A, float2& B) {
    const float4* r = s.rec + 2 * (size_t)(unsigned)idx;
    float u0 = 0.f, u1 = 0.f;       // padding words of the record
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(A.x), "=f"(A.y), "=f"(A.z), "=f"(A.w), "=f"(B.x), "=f"(B.y), "=f"(u0), "=f"(u1) : "l"(r));
    (void)u0; (void)u1;
}
__device__ __forceinline__ void transform(const float* T, float x, float y, float z, float& px, float& py, float& pz) {
    // transform_pcd_cuda (icp.cu:147-149) with the accumulated transform
    px = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
    py = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
    pz = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
}

// 1/z for the projection: hardware reciprocal + one Newton step (faithful to within 1 ulp for the
// normal-range depths a point cloud holds); any other input still yields a value the bounds test
// below classifies safely.
__device__ __forceinline__ float fast_rcp(float z) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
    const float e = fmaf(-z, r, 1.0f);
    return fmaf(r, e, r);
}

// 32*kIlp consecutive points of a tile against the packed projective scene: lane l owns points
// l, 32+l, 64+l, ..., so each gather instruction covers 32 CONSECUTIVE model points -- neighbouring scene
// pixels.  The kIlp gathers of a lane are issued before any of them is consumed.
// Pixel selection: u = int(px/pz*fx + cx + 0.5) (common.h:63-73) evaluated as
// fma(px*(1/pz), fx, cx+0.5), within 2 ulp of the reference's operation order.  "0 <= int(v) < W" is
// tested on the truncated integers as unsigned compares; fmaxf(v, -2) first turns NaN into a
// rejected value (a plain float->int conversion would turn NaN into pixel 0).

// 1/z with the sign folded in: returns -(1/z) refined by one Newton step.  Bit for bit the negation of
// fast_rcp(z) (rcp.approx is odd, (-z)*r == z*(-r), and round-to-nearest is symmetric).
__device__ __forceinline__ f2_t fast_nrcp2(f2_t z) {
    float z0, z1, r0, r1;
    unpk2(z, z0, z1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(-z0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(-z1));
    const f2_t nr = pk2(r0, r1);
    const f2_t e = fma2(z, nr, bc2(1.0f));
    return fma2(nr, e, nr);
}
// group_projective with two points per instruction (see AccP).  Lane l owns points l + 32k of the group;
// points 2j and 2j+1 form pair j.  No branches: a rejected point is turned into q = p, n = 0 by the
// selects that also move the gathered values into the pair registers.  A non-finite transformed point
// would turn into NaN sums here (0 * inf); the item loop detects that and redoes the item with the
// per-point path (slow_item), so such clouds stay correct and everything else pays nothing for them.
template <bool TAIL>
__device__ __forceinline__ void group_projective(const PackedScene& s, unsigned addr, unsigned first, unsigned n,
                                                 const float* T, AccP& acc) {
    constexpr int NP = kIlp / 2;
    static_assert(kIlp % 2 == 0, "pairs");
    f2_t px[NP], py[NP], pz[NP];
    int idx[kIlp];
    bool ok[kIlp];
    const float nfx = -s.fx, nfy = -s.fy;
#pragma unroll
    for (int j = 0; j < NP; j++) {
        float x0 = lds32(addr + 384 * (2 * j)), y0 = lds32(addr + 384 * (2 * j) + 4), z0 = lds32(addr + 384 * (2 * j) + 8);
        float x1 = lds32(addr + 384 * (2 * j + 1)), y1 = lds32(addr + 384 * (2 * j + 1) + 4), z1 = lds32(addr + 384 * (2 * j + 1) + 8);
        bool in0 = true, in1 = true;
        if (TAIL) {     // the tile holds stale data past the end of the item
            in0 = first + 32 * (2 * j) < n; in1 = first + 32 * (2 * j + 1) < n;
            x0 = in0 ? x0 : 0.f; y0 = in0 ? y0 : 0.f; z0 = in0 ? z0 : 0.f;
            x1 = in1 ? x1 : 0.f; y1 = in1 ? y1 : 0.f; z1 = in1 ? z1 : 0.f;
        }
        const f2_t x = pk2(x0, x1), y = pk2(y0, y1), z = pk2(z0, z1);
        // transform_pcd_cuda (icp.cu:147-149) with the accumulated transform, same FMA chain as transform()
        px[j] = fma2(bc2(T[2]), z, fma2(bc2(T[1]), y, fma2(bc2(T[0]), x, bc2(T[3]))));
        py[j] = fma2(bc2(T[6]), z, fma2(bc2(T[5]), y, fma2(bc2(T[4]), x, bc2(T[7]))));
        pz[j] = fma2(bc2(T[10]), z, fma2(bc2(T[9]), y, fma2(bc2(T[8]), x, bc2(T[11]))));
        const f2_t nrz = fast_nrcp2(pz[j]);
        // fma(px*rz, fx, cx+0.5) == fma(px*(-rz), -fx, cx+0.5)
        float uf0, uf1, vf0, vf1;
        unpk2(fma2(bc2(nfx), mul2(px[j], nrz), bc2(s.cx05)), uf0, uf1);
        unpk2(fma2(bc2(nfy), mul2(py[j], nrz), bc2(s.cy05)), vf0, vf1);
        const int u0 = __float2int_rz(fmaxf(uf0, -2.0f)), v0 = __float2int_rz(fmaxf(vf0, -2.0f));
        const int u1 = __float2int_rz(fmaxf(uf1, -2.0f)), v1 = __float2int_rz(fmaxf(vf1, -2.0f));
        ok[2 * j] = ((unsigned)u0 < (unsigned)s.W) & ((unsigned)v0 < (unsigned)s.H) & in0;
        ok[2 * j + 1] = ((unsigned)u1 < (unsigned)s.W) & ((unsigned)v1 < (unsigned)s.H) & in1;
        idx[2 * j] = v0 * s.W + u0;
        idx[2 * j + 1] = v1 * s.W + u1;
#ifdef PR_DBG_NOGATHER      // what-if: every gather hits the same few L1 lines (wrong results, timing only)
        idx[2 * j] = (threadIdx.x & 31) + 32 * (2 * j); idx[2 * j + 1] = (threadIdx.x & 31) + 32 * (2 * j + 1);
#endif
#ifdef PR_DBG_WRAP          // what-if (use with PR_DBG_NOACC so that every pass runs): same access pattern folded
                            // into a PR_DBG_WRAP-pixel window in the middle of the object
        idx[2 * j] = 280 * 640 + 300 + (idx[2 * j] & (PR_DBG_WRAP - 1)); idx[2 * j + 1] = 280 * 640 + 300 + (idx[2 * j + 1] & (PR_DBG_WRAP - 1));
#endif
    }
    float4 A[kIlp];
    float2 B[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; k++) {
        if (ok[k]) {
#ifdef PR_DBG_NOLOAD        // what-if (with PR_DBG_NOACC): no scene access at all, every in-image point "valid"
            A[k] = make_float4(1.f, 2.f, 0.3f, 0.5f); B[k] = make_float2(0.5f, 0.7f);
#else
            load_rec(s, idx[k], A[k], B[k]);
#endif
        }
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
        const float4 A0 = A[2 * j], A1 = A[2 * j + 1];
        const float2 B0 = B[2 * j], B1 = B[2 * j + 1];
        float px0, px1, py0, py1, pz0, pz1;
        unpk2(px[j], px0, px1); unpk2(py[j], py0, py1); unpk2(pz[j], pz0, pz1);
        const bool v0 = ok[2 * j] && A0.z > 0.f && fabsf(pz0 - A0.z) <= s.max_dist;          // depth_scene.h:42
        const bool v1 = ok[2 * j + 1] && A1.z > 0.f && fabsf(pz1 - A1.z) <= s.max_dist;
        const f2_t qx = pk2(v0 ? A0.x : px0, v1 ? A1.x : px1);
        const f2_t qy = pk2(v0 ? A0.y : py0, v1 ? A1.y : py1);
        const f2_t qz = pk2(v0 ? A0.z : pz0, v1 ? A1.z : pz1);
        const f2_t nx = pk2(v0 ? A0.w : 0.f, v1 ? A1.w : 0.f);
        const f2_t ny = pk2(v0 ? B0.x : 0.f, v1 ? B1.x : 0.f);
        const f2_t nz = pk2(v0 ? B0.y : 0.f, v1 ? B1.y : 0.f);
#ifdef PR_DBG_NOACC         // what-if: no Jacobian / sums (wrong results, timing only)
        acc.s[0] = fma2(qx, nx, acc.s[0]); acc.s[1] = fma2(qy, ny, acc.s[1]); acc.s[2] = fma2(qz, nz, acc.s[2]);
#else
        accumulate_pair(acc, px[j], py[j], pz[j], qx, qy, qz, nx, ny, nz);
#endif
        if (v0) acc.cnt_a += 1.0f;
        if (v1) acc.cnt_b += 1.0f;
    }
}

__device__ __forceinline__ void compute_tile(const PackedScene& s, unsigned tile, unsigned n, const float* T, AccP& acc) {
    const unsigned lane = threadIdx.x & 31;
    unsigned addr = tile + 12 * lane;
    unsigned first = lane;
    constexpr unsigned kGroup = 32 * kIlp;
    static_assert(kWTile % kGroup == 0, "a tail group must not read past the tile");
    const unsigned n_full = n - n % kGroup;
#pragma unroll 1
    for (; first < n_full; first += kGroup, addr += 12 * kGroup) group_projective<false>(s, addr, first, n, T, acc);
    if (n_full < n) group_projective<true>(s, addr, first, n, T, acc);
}

// per-point query against the packed scene with the pixel selection of group_projective (robust path)
__device__ __forceinline__ bool query(const PackedScene& s, float px, float py, float pz, Corr& c) {
    const float rz = fast_rcp(pz);
    const int ui = __float2int_rz(fmaxf(fmaf(px * rz, s.fx, s.cx05), -2.0f));
    const int vi = __float2int_rz(fmaxf(fmaf(py * rz, s.fy, s.cy05), -2.0f));
    if (!(((unsigned)ui < (unsigned)s.W) & ((unsigned)vi < (unsigned)s.H))) return false;
    float4 A; float2 B;
    load_rec(s, vi * s.W + ui, A, B);
    if (!(A.z > 0.f && fabsf(pz - A.z) <= s.max_dist)) return false;
    c.qx = A.x; c.qy = A.y; c.qz = A.z; c.nx = A.w; c.ny = B.x; c.nz = B.y;
    return true;
}
// the whole item again, point by point from global memory (taken only when the packed path produced NaN sums)
__device__ __noinline__ float slow_item(const PackedScene& s, const float* __restrict__ g, unsigned n, const float* T) {
    const unsigned lane = threadIdx.x & 31;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = 0.f;
    for (unsigned i = lane; i < n; i += 32) {
        float px, py, pz;
        transform(T, g[3 * i], g[3 * i + 1], g[3 * i + 2], px, py, pz);
        Corr c;
        if (query(s, px, py, pz, c)) accumulate(v, px, py, pz, c);
    }
    return warp_transpose_reduce(v);
}

// one warp tile, any scene with a per-point query() (nearest neighbour)
template <class SceneT>
__device__ __forceinline__ void compute_tile(const SceneT& s, unsigned tile, unsigned n, const float* T, AccT& acc) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll 1
    for (unsigned i = lane; i < n; i += 32) {
        float px, py, pz;
        transform(T, lds32(tile + 12 * i), lds32(tile + 12 * i + 4), lds32(tile + 12 * i + 8), px, py, pz);
        Corr c;
        if (query(s, px, py, pz, c)) acc_add(acc, px, py, pz, c);
    }
}

// the warp that deposited the last partial of a hypothesis: add the partials in chunk order (lane l
// owns sum l), hand the 29 sums to lane 0 and run the stop logic / solve there.
__device__ __noinline__ void warp_finish_hypothesis(HypState* st, const float* partials, unsigned n_h, pr_icp_criteria crit,
                                                    pr_registration_result* res) {
    const unsigned lane = threadIdx.x & 31;
    float sum = 0.f;
    const unsigned cb = __ldcg(&st->chunk_begin), nc = __ldcg(&st->n_chunks);
    // chunk order (deterministic); four loads in flight per step
    unsigned j = 0;
    const float* pp = partials + (size_t)cb * kPartialStride + lane;
    for (; j + 4 <= nc; j += 4) {
        const float a0 = __ldcg(pp + (size_t)j * kPartialStride), a1 = __ldcg(pp + (size_t)(j + 1) * kPartialStride);
        const float a2 = __ldcg(pp + (size_t)(j + 2) * kPartialStride), a3 = __ldcg(pp + (size_t)(j + 3) * kPartialStride);
        sum = (((sum + a0) + a1) + a2) + a3;
    }
    for (; j < nc; j++) sum += __ldcg(pp + (size_t)j * kPartialStride);
    float S[29];
#pragma unroll
    for (int i = 0; i < 29; i++) S[i] = __shfl_sync(0xffffffffu, sum, i);
    if (lane == 0) {
        __stcg(&st->arrived, 0u);
        finish_pass_release(st, S, n_h, crit, res);
    }
    __syncwarp();
}

// global -> shared copy of one tile: TMA when the source is 16-byte aligned and the copy, rounded up
// to 16 bytes, stays inside the point buffer; otherwise the warp copies it with plain loads.
// Returns true when the TMA path was taken (the caller then waits on the stage's mbarrier).
__device__ __forceinline__ bool stage_tile(const float* src, unsigned n, uintptr_t pts_end, unsigned tile, unsigned bar) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned bytes = (n * 12 + 15) & ~15u;
    const bool tma = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (reinterpret_cast<uintptr_t>(src) + bytes <= pts_end);
    if (tma) {
        if (lane == 0) {
            mbar_expect_tx(bar, bytes);
            tma_load_1d(tile, src, bytes, bar);
        }
    } else {
        for (unsigned i = lane; i < n * 3; i += 32) sts32(tile + 4 * i, src[i]);
        __syncwarp();
    }
    return tma;
}

// projective: one CTA of register-rich warps per SM; nearest neighbour: four CTAs (the tree walk hides latency with warps)
template <class SceneT> struct MinBlocksOf { static constexpr int kValue = std::is_same<SceneT, PackedScene>::value ? PR_MINB : PR_NN_MINB; };
template <class SceneT>
__global__ void __launch_bounds__(kPThreads, MinBlocksOf<SceneT>::kValue)
icp_persistent_kernel(const float* __restrict__ pts, size_t capacity_points, const uint4* __restrict__ chunk_info, IcpCtl* ctl,
                      HypState* state, float* partials, SceneT scene, pr_icp_criteria crit,
                      pr_registration_result* results) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using AccK = typename std::conditional<std::is_same<SceneT, PackedScene>::value, AccP, AccT>::type;
    constexpr int kWTile = TileOf<SceneT>::kPoints;     // shadows the projective constant on purpose
    constexpr int kWTileFloats = kWTile * 3;
    constexpr int kWTileBytes = kWTileFloats * 4;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned tile0 = smem_u32(smem_raw) + warp * (kWStages * kWTileBytes);
    const unsigned bar0 = smem_u32(smem_raw) + kPWarps * kWStages * kWTileBytes + warp * (kWStages * 8);
    const uintptr_t pts_end = reinterpret_cast<uintptr_t>(pts) + capacity_points * 12;

    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const unsigned total = __ldcg(&ctl->total_chunks);
    const unsigned n_items = total * (unsigned)(crit.max_iteration + 1);
    auto claim = [&]() {
        unsigned v = 0;
        if (lane == 0) v = atomicAdd(&ctl->next_item, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    // item -> its chunk record {hypothesis, first point, points in chunk, points of the hypothesis}
    auto locate = [&](unsigned item) {
        const unsigned c = item % total;
        return (item < n_items) ? __ldg(chunk_info + c) : make_uint4(0u, 0u, 0u, 0u);
    };

    unsigned stage = 0;                 // stage that holds (or will hold) the next tile to consume
    unsigned parity = 0;                // bit s = parity to wait for on stage s
    unsigned item = claim();
    uint4 info = locate(item);
    bool tma_cur = false;               // was the next tile to consume fetched by TMA?
    if (item < n_items)
        tma_cur = stage_tile(pts + 3 * (size_t)info.y, min((unsigned)kWTile, info.z), pts_end, tile0 + stage * kWTileBytes, bar0 + 8 * stage);
    // arrival ticket of a deposited partial; the warp that takes the last ticket of a hypothesis finishes its pass
    auto take_ticket = [&](HypState* tst, unsigned n_h, unsigned th) {
        __syncwarp();            // all 32 partial stores precede lane 0's release below
        int is_last = 0;
        if (lane == 0) is_last = (atom_add_release(&tst->arrived, 1u) == __ldcg(&tst->n_chunks) - 1) ? 1 : 0;
        is_last = __shfl_sync(0xffffffffu, is_last, 0);
        if (is_last) warp_finish_hypothesis(tst, partials, n_h, crit, results + th);
    };
    while (item < n_items) {
        const unsigned pass = item / total;
        const unsigned c = item - pass * total;
        const unsigned h = info.x, n_pts = info.z;
        const float* g = pts + 3 * (size_t)info.y;
        HypState* st = state + h;
        const unsigned n_tiles = (n_pts + kWTile - 1) / kWTile;
        unsigned next_item = 0xFFFFFFFFu;
        uint4 next_info = make_uint4(0u, 0u, 0u, 0u);

        // ---- wait until the hypothesis has finished the previous pass (or has returned)
        int flag = 0;
        if (lane == 0) {
            unsigned backoff = 32;
            for (;;) {
                if (ld_relaxed(reinterpret_cast<const unsigned*>(&st->pass)) >= pass) { flag = 0; break; }
                if (ld_relaxed(reinterpret_cast<const unsigned*>(&st->done))) { flag = 1; break; }
                __nanosleep(backoff);
                if (backoff < 1024) backoff <<= 1;
            }
        }
        const bool skip = __shfl_sync(0xffffffffu, flag, 0) != 0;
        float T[12];
        AccK acc;
#pragma unroll
        for (int i = 0; i < 12; i++) T[i] = skip ? 0.f : __ldcg(&st->T[i]);
        acc_zero(acc);

        // tile bookkeeping shared by both loop shapes: on entering tile t, prefetch its successor (next
        // tile of this item, or the first tile of the next claimed item) and wait for tile t itself
        bool tma_next = false;
        auto enter_tile = [&](unsigned t) {
            const unsigned other = stage ^ 1;
            tma_next = false;
            if (t + 1 < n_tiles) {
                tma_next = stage_tile(g + (size_t)(t + 1) * kWTileFloats, min((unsigned)kWTile, n_pts - (t + 1) * kWTile), pts_end,
                                      tile0 + other * kWTileBytes, bar0 + 8 * other);
            } else {
                // claim the next item only now: claiming a whole item earlier parks ~1 item per warp in front of the
                // workers, which pushes them a pass ahead of their dependencies (measured: 2.21 -> 2.33 ms)
                next_item = claim();
                next_info = locate(next_item);
                if (next_item < n_items)
                    tma_next = stage_tile(pts + 3 * (size_t)next_info.y, min((unsigned)kWTile, next_info.z), pts_end,
                                          tile0 + other * kWTileBytes, bar0 + 8 * other);
            }
            if (tma_cur) {
                mbar_wait(bar0 + 8 * stage, (parity >> stage) & 1u);
                parity ^= (1u << stage);
            }
        };
        auto leave_tile = [&]() {
            __syncwarp();            // every lane is done reading this stage before it is refilled
            stage ^= 1;
            tma_cur = tma_next;
        };
        // (tried, no gain: taking the arrival ticket one tile into the next item so that the release fence finds the
        // partial stores already performed -- 1 % at chunk 3072 and a circular wait at chunk 4096; splitting the claim
        // over the last two tiles to hide the atomic and the chunk-record load -- 1.50 -> 1.52 ms.)
        // (tried: issuing the gathers of group g+1 before consuming group g from a second register set.  It
        // does not overlap anything: ptxas tracks both sets on the same scoreboard, so the first consume
        // waits for the newest gathers as well -- measured 2.19 -> 2.39..2.54 ms.)
        for (unsigned t = 0; t < n_tiles; t++) {
            enter_tile(t);
            if (!skip) compute_tile(scene, tile0 + stage * kWTileBytes, min((unsigned)kWTile, n_pts - t * kWTile), T, acc);
            leave_tile();
        }

        // ---- item complete: reduce over the warp, deposit, maybe finish the pass of this hypothesis
        if (!skip) {
            float v[32];
            acc_unpack(acc, v);
            float mine = warp_transpose_reduce(v);             // lane l = sum of value l
            if constexpr (std::is_same<SceneT, PackedScene>::value) {
                if (__any_sync(0xffffffffu, mine != mine)) mine = slow_item(scene, g, n_pts, T);
            }
            __stcg(partials + (size_t)c * kPartialStride + lane, mine);
            take_ticket(st, info.w, h);
        }
        item = next_item;
        info = next_info;
    }
}

// p <- T_h * p for every point of every hypothesis (PR_ICP_UPDATE_POINTS): what the reference's
// in-place transform_pcd_cuda calls add up to.
__global__ void __launch_bounds__(256)
icp_apply_kernel(float* __restrict__ pts, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts,
                 const uint32_t* __restrict__ chunk_hyp, const uint32_t* __restrict__ total_chunks, uint32_t chunk_points,
                 const HypState* __restrict__ state) {
    const unsigned total = *total_chunks;
    for (unsigned c = blockIdx.x; c < total; c += gridDim.x) {
        const unsigned h = chunk_hyp[c];
        const HypState* st = state + h;
        float T[12];
#pragma unroll
        for (int i = 0; i < 12; i++) T[i] = st->T[i];
        const unsigned first = (c - st->chunk_begin) * chunk_points;
        const unsigned n = min(chunk_points, counts[h] - first);
        float* p0 = pts + 3 * ((size_t)offsets[h] + first);
        for (unsigned i = threadIdx.x; i < n; i += 256) {
            const float x = p0[3 * i], y = p0[3 * i + 1], z = p0[3 * i + 2];
            p0[3 * i + 0] = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
            p0[3 * i + 1] = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
            p0[3 * i + 2] = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
        }
    }
}

inline size_t icp_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct IcpWs {
    HypState* state; uint32_t* chunk_hyp; uint4* chunk_info; IcpCtl* ctl; float* partials; float4* packed;
    size_t max_chunks, bytes;
};

// chunk size of the per-pass driver: big chunks amortise the per-CTA reduction; small batches need
// more CTAs than SMs
inline uint32_t pick_chunk_points(size_t n_hyp, size_t capacity_points) {
    uint32_t chunk = 4096;
    while (chunk > 512 && capacity_points / chunk + n_hyp < (size_t)kNumSMs * 4) chunk >>= 1;
    return chunk;
}

inline IcpWs carve_icp_ws(void* base, size_t n_hyp, size_t capacity_points, size_t scene_pixels) {
    IcpWs ws;
    // sized for the smallest chunk either driver uses
    ws.max_chunks = capacity_points / 512 + n_hyp + 1;
    char* w = (char*)base;
    size_t used = 0;
    auto take = [&](size_t bytes) { char* p = w + used; used += icp_align_up(bytes, 256); return p; };
    ws.state = (HypState*)take(n_hyp * sizeof(HypState));
    ws.chunk_hyp = (uint32_t*)take(ws.max_chunks * 4);
    ws.chunk_info = (uint4*)take(ws.max_chunks * 16);
    ws.ctl = (IcpCtl*)take(sizeof(IcpCtl));
    ws.partials = (float*)take(ws.max_chunks * kPartialStride * 4);
    ws.packed = (float4*)take(scene_pixels * 32);
    ws.bytes = used;
    return ws;
}

// PR_ICP_IMPL=pass selects the first-generation driver (one launch per pass) for cross-checks
inline bool use_pass_driver() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PR_ICP_IMPL"); v = (e && strcmp(e, "pass") == 0) ? 1 : 0; }
    return v == 1;
}

template <class SceneT>
int persistent_grid(int* grid_out) {
    static int cached = 0;
    if (!cached) {
        const int smem = persist_smem<SceneT>();
        PR_CUDA_TRY(cudaFuncSetAttribute(icp_persistent_kernel<SceneT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 0, dev = 0, sms = 0;
        PR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_persistent_kernel<SceneT>, kPThreads, smem));
        PR_CUDA_TRY(cudaGetDevice(&dev));
        PR_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cached = std::max(1, occ) * std::max(1, sms);
    }
    *grid_out = cached;
    return PR_OK;
}

// Points per work item of the persistent driver.  Large items amortise the per-item cost (claim, reduction,
// release fence, ticket: 4096 beats 2048 by 10 % on 512 hypotheses), but a pass must still consist of a
// few items per resident warp or the warps run into the pass-to-pass dependency of their hypotheses
// (8192: +30 %), and a small batch needs enough items to occupy the machine at all.
#ifndef PR_NN_CHUNK
#define PR_NN_CHUNK 1024
#endif
// nearest-neighbour scenes: a point costs ~50x more (tree walk), so the per-item cost is irrelevant and small items
// balance better
template <class PScene>
inline uint32_t persist_chunk_points(size_t n_hyp, size_t capacity_points) {
    uint32_t chunk = std::is_same<PScene, PackedScene>::value ? kPersistChunk : (uint32_t)PR_NN_CHUNK;
    while (chunk > 512 && capacity_points / chunk + n_hyp < (size_t)kNumSMs * kPWarps * 2) chunk >>= 1;
    return chunk;
}

// SceneT: the scene as the per-pass driver consumes it; PScene: as the persistent driver consumes it
template <class SceneT, class PScene>
int run_icp(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp, size_t capacity_points,
            const SceneT& scene, const PScene& pscene, pr_icp_criteria crit, pr_registration_result* results_dev, int flags,
            const IcpWs& ws, cudaStream_t stream) {
    if (use_pass_driver()) {
        const uint32_t chunk = pick_chunk_points(n_hyp, capacity_points);
        // persistent grid: 3 CTAs per SM (register-limited), never more CTAs than chunks can exist
        const unsigned grid = (unsigned)std::min<size_t>(capacity_points / chunk + n_hyp + 1, (size_t)kNumSMs * 3);
        icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts_dev, (uint32_t)n_hyp, chunk, ws.state, ws.chunk_hyp,
                                                       (uint32_t)ws.max_chunks, &ws.ctl->total_chunks, results_dev, &ws.ctl->next_item);
        for (int it = 0; it <= crit.max_iteration; it++)
            icp_pass_kernel<SceneT><<<grid, kIcpThreads, 0, stream>>>(pts_dev, offsets_dev, counts_dev, ws.chunk_hyp, &ws.ctl->total_chunks,
                                                                      chunk, ws.state, ws.partials, scene, crit, results_dev, nullptr);
        count_launch(2 + (uint64_t)crit.max_iteration);
        if (flags & PR_ICP_UPDATE_POINTS) {
            icp_apply_kernel<<<grid, 256, 0, stream>>>(pts_dev, offsets_dev, counts_dev, ws.chunk_hyp, &ws.ctl->total_chunks, chunk, ws.state);
            count_launch();
        }
        PR_LAUNCH_CHECK();
        return PR_OK;
    }
    const uint32_t chunk_pts = persist_chunk_points<PScene>(n_hyp, capacity_points);
    const size_t max_items = (capacity_points / chunk_pts + n_hyp + 1) * (size_t)(crit.max_iteration + 1);
    if (max_items > 0x7FFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts_dev, (uint32_t)n_hyp, chunk_pts, ws.state, ws.chunk_hyp,
                                                   (uint32_t)ws.max_chunks, &ws.ctl->total_chunks, results_dev, &ws.ctl->next_item,
                                                   offsets_dev, ws.chunk_info);
    int grid = 0;
    int rc = persistent_grid<PScene>(&grid);
    if (rc != PR_OK) return rc;
    grid = (int)std::min<size_t>((size_t)grid, (max_items + kPWarps - 1) / kPWarps);
    int threads = kPThreads;
    if (const char* e = getenv("PR_ICP_WARPS")) threads = std::max(1, std::min(kPWarps, atoi(e))) * 32;   // experiments
    icp_persistent_kernel<PScene><<<grid, threads, persist_smem<PScene>(), stream>>>(
        pts_dev, capacity_points, ws.chunk_info, ws.ctl, ws.state, ws.partials, pscene, crit, results_dev);
    count_launch(2);
    if (flags & PR_ICP_UPDATE_POINTS) {
        icp_apply_kernel<<<kNumSMs * 4, 256, 0, stream>>>(pts_dev, offsets_dev, counts_dev, ws.chunk_hyp, &ws.ctl->total_chunks,
                                                          chunk_pts, ws.state);
        count_launch();
    }
    PR_LAUNCH_CHECK();
    return PR_OK;
}

inline int check_icp_args(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                          size_t capacity_points, pr_icp_criteria crit, pr_registration_result* results_dev, void* workspace_dev) {
    if (!pts_dev || !offsets_dev || !counts_dev || !results_dev || !workspace_dev) return PR_ERR_INVALID_ARGUMENT;
    if (crit.max_iteration < 0 || n_hyp > 0x7FFFFFFFull / 64 || capacity_points > 0xFFFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    return PR_OK;
}

inline int make_proj_scene(const pr_scene_projective* s, ProjScene& o) {
    if (!s || !s->pcd_dev || !s->normal_dev || s->width == 0 || s->height == 0 || s->width > 32768 || s->height > 32768)
        return PR_ERR_INVALID_ARGUMENT;
    o.W = (int)s->width; o.H = (int)s->height; o.fW = (float)s->width; o.fH = (float)s->height;
    o.max_dist = s->max_dist_diff;
    o.fx = s->K[0]; o.fy = s->K[4]; o.cx = s->K[2]; o.cy = s->K[5];
    o.pcd = s->pcd_dev; o.nrm = s->normal_dev;
    return PR_OK;
}
inline int make_nn_scene(const pr_scene_nn* s, NnScene& o) {
    if (!s || (s->n_nodes && (!s->pcd_dev || !s->normal_dev || !s->nodes_dev)) || s->n_nodes > 0x7FFFFFFFull)
        return PR_ERR_INVALID_ARGUMENT;
    o.max_dist_sq = s->max_dist_diff * s->max_dist_diff;   // pow2(max_dist_diff), pcd_scene.h:127
    o.pcd = s->pcd_dev; o.nrm = s->normal_dev; o.nodes = s->nodes_dev; o.n_nodes = (int)s->n_nodes;
    return PR_OK;
}

// single cloud, identity transform, one reduction pass -> out29 (parity / debug entry point)
template <class SceneT>
int run_pcd2ab(const float* pts_dev, size_t n, const SceneT& scene, float* out29_dev, cudaStream_t stream) {
    if (!pts_dev || !out29_dev || n > 0xFFFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    // scratch: counts/offsets (2 words) + workspace, allocated here because this is a debug call
    const size_t ws_bytes = carve_icp_ws(nullptr, 1, n, 0).bytes;
    char* scratch = nullptr;
    PR_CUDA_TRY(cudaMalloc((void**)&scratch, 256 + 256 + ws_bytes));
    uint32_t* counts = (uint32_t*)scratch;
    uint32_t* offsets = (uint32_t*)(scratch + 128);
    pr_registration_result* res = (pr_registration_result*)(scratch + 256);
    const uint32_t h_counts = (uint32_t)n, h_off = 0;
    cudaMemcpyAsync(counts, &h_counts, 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(offsets, &h_off, 4, cudaMemcpyHostToDevice, stream);
    IcpWs ws = carve_icp_ws(scratch + 512, 1, n, 0);
    const uint32_t chunk = pick_chunk_points(1, n);
    const unsigned grid = (unsigned)std::min<size_t>(n / chunk + 2, (size_t)kNumSMs * 3);
    pr_icp_criteria crit = {0.f, 0.f, 0};
    cudaMemsetAsync(out29_dev, 0, 29 * 4, stream);
    icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts, 1, chunk, ws.state, ws.chunk_hyp, (uint32_t)ws.max_chunks,
                                                   &ws.ctl->total_chunks, res, &ws.ctl->next_item);
    icp_pass_kernel<SceneT><<<grid, kIcpThreads, 0, stream>>>(pts_dev, offsets, counts, ws.chunk_hyp, &ws.ctl->total_chunks, chunk, ws.state,
                                                              ws.partials, scene, crit, res, out29_dev);
    count_launch(2);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(scratch);
    return e == cudaSuccess ? PR_OK : (int)e;
}

}  // namespace prb

using namespace prb;

extern "C" {

size_t pr_icp_workspace_bytes(size_t n_hyp, size_t capacity_points, size_t scene_pixels) {
    return carve_icp_ws(nullptr, n_hyp, capacity_points, scene_pixels).bytes;
}

int pr_icp_projective_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                            size_t capacity_points, const pr_scene_projective* scene, pr_icp_criteria criteria,
                            pr_registration_result* results_dev, int flags,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    ProjScene s;
    int rc = make_proj_scene(scene, s);
    if (rc != PR_OK) return rc;
    rc = check_icp_args(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, criteria, results_dev, workspace_dev);
    if (rc != PR_OK) return rc;
    if (n_hyp == 0) return PR_OK;
    const size_t n_px = (size_t)s.W * s.H;
    IcpWs ws = carve_icp_ws(workspace_dev, n_hyp, capacity_points, n_px);
    if (workspace_bytes < ws.bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = as_stream(stream_);
    PackedScene ps;
    ps.W = s.W; ps.H = s.H; ps.fW = s.fW; ps.fH = s.fH; ps.max_dist = s.max_dist;
    ps.fx = s.fx; ps.fy = s.fy; ps.cx05 = s.cx + 0.5f; ps.cy05 = s.cy + 0.5f; ps.rec = ws.packed;
    if (!use_pass_driver()) {
        scene_pack_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, stream>>>(s.pcd, s.nrm, n_px, ws.packed);
        count_launch();
    }
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, ps, criteria, results_dev, flags, ws, stream);
}

int pr_icp_nn_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                    size_t capacity_points, const pr_scene_nn* scene, pr_icp_criteria criteria,
                    pr_registration_result* results_dev, int flags,
                    void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    NnScene s;
    int rc = make_nn_scene(scene, s);
    if (rc != PR_OK) return rc;
    rc = check_icp_args(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, criteria, results_dev, workspace_dev);
    if (rc != PR_OK) return rc;
    if (n_hyp == 0) return PR_OK;
    cudaStream_t stream = as_stream(stream_);
    // the packed tree needs 16 B per scene point + 32 B per node (+ a flag word)
    const size_t units = scene->n_points + 2 * scene->n_nodes + 16;
    const IcpWs ws_packed = carve_icp_ws(workspace_dev, n_hyp, capacity_points, units);
    if (!use_pass_driver() && s.n_nodes > 0 && workspace_bytes >= ws_packed.bytes) {
        float4* pts4 = ws_packed.packed;
        float4* pnodes = pts4 + scene->n_points;
        unsigned* flag = reinterpret_cast<unsigned*>(pnodes + 2 * scene->n_nodes);
        PR_CUDA_TRY(cudaMemsetAsync(flag, 0, 4, stream));
        nn_pack_points_kernel<<<(unsigned)((scene->n_points + 255) / 256), 256, 0, stream>>>(s.pcd, scene->n_points, pts4);
        nn_pack_nodes_kernel<<<(unsigned)((scene->n_nodes + 255) / 256), 256, 0, stream>>>(s.nodes, s.n_nodes, s.pcd, pnodes, flag);
        count_launch(2);
        // the encoding limits (leaf <= 127 points, < 2^24 points, sibling children) hold for every tree
        // KDTree_cpu::build_tree / pr_scene_nn_build_host produce; a foreign tree that breaks them is detected
        // on the device and reported after a one-word read-back.
        unsigned h_flag = 0;
        PR_CUDA_TRY(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, stream));
        PR_CUDA_TRY(cudaStreamSynchronize(stream));
        if (!h_flag) {
            PackedNnScene ps;
            ps.max_dist_sq = s.max_dist_sq; ps.nodes = pnodes; ps.pts4 = pts4; ps.nrm = s.nrm; ps.n_nodes = s.n_nodes; ps.ref = s;
            return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, ps, criteria, results_dev, flags, ws_packed, stream);
        }
    }
    IcpWs ws = carve_icp_ws(workspace_dev, n_hyp, capacity_points, 0);
    if (workspace_bytes < ws.bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, s, criteria, results_dev, flags, ws, stream);
}

int pr_solve_666(const float A[36], const float b[6], float T[16]) {
    if (!A || !b || !T) return PR_ERR_INVALID_ARGUMENT;
    // the register-resident form the device runs (same arithmetic as solve_666): pack the lower triangle
    // the way thrust__pcd2Ab orders it (icp.h:165-197)
    float S[29], E[16];
    int shift = 0;
    for (int y = 0; y < 6; y++)
        for (int x = y; x < 6; x++) S[shift++] = A[x + 6 * y];
    for (int i = 0; i < 6; i++) S[21 + i] = b[i];
    S[27] = 0.f; S[28] = 0.f;
    solve_666_unrolled(S, E);
    for (int i = 0; i < 16; i++) T[i] = E[i];
    return PR_OK;
}

int pr_pcd2ab_projective(const float* pts_dev, size_t n, const pr_scene_projective* scene, float* out29_dev, pr_stream_t stream) {
    ProjScene s;
    int rc = make_proj_scene(scene, s);
    if (rc != PR_OK) return rc;
    return run_pcd2ab(pts_dev, n, s, out29_dev, as_stream(stream));
}

int pr_pcd2ab_nn(const float* pts_dev, size_t n, const pr_scene_nn* scene, float* out29_dev, pr_stream_t stream) {
    NnScene s;
    int rc = make_nn_scene(scene, s);
    if (rc != PR_OK) return rc;
    return run_pcd2ab(pts_dev, n, s, out29_dev, as_stream(stream));
}

}  // extern "C"
