// oracle/oracle.cpp -- TEST INFRASTRUCTURE ONLY. CPU restatement ("port") of the reference's
// hot path: batched depth rasteriser -> depth2cloud -> scene preparation -> point-to-plane ICP.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker.  The product (pose_refine_b200/, include/) never
// links, loads or falls back to it.
//
// Parity pin: every entry point here is checked bit-for-bit (integer outputs) or value-for-value
// at 1 thread (float outputs) against the reference's own sources compiled verbatim into
// oracle/_ref/libpose_refine_ref.so (oracle/Makefile, oracle/ref_glue.cpp) by
// tests/test_oracle_vs_ref.py, and against the committed fixtures in tests/golden/ (generated
// from that same build by scripts/make_golden.py).  The one third-party piece the reference
// relies on that is absent from /root/reference is Eigen (unpinned version,
// cuda_icp/CMakeLists.txt:18): its LDLT and AngleAxis/Quaternion arithmetic is restated from
// Eigen's published algorithm (see solve_666 below); that boundary is pinned only end-to-end.
//
// All citations are relative to /root/reference.  Plain float arithmetic, evaluated in the
// reference's order; built with -ffp-contract=off so no FMA contraction can change rounding.
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cfloat>
#include <climits>
#include <cmath>
#include <vector>
#include <numeric>
#include <algorithm>
#include <chrono>
#include <omp.h>

namespace {

// ---- small helpers ---------------------------------------------------------------------------

// C++ float->int32 conversion of a value outside int32 (or NaN) is undefined; the reference's
// x86-64 build gets cvttss2si's "integer indefinite" (INT_MIN).  Stated explicitly here so the
// CUDA path has a defined target for the same corner (last paragraph).
inline int32_t f2i_x86(float v) {
    if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT_MIN;
    return (int32_t)v;
}
inline float sel_max(float a, float b) { return (a > b) ? a : b; }  // renderer.h:335-336
inline float sel_min(float a, float b) { return (a < b) ? a : b; }  // renderer.h:337-338

// renderer.h:296-303 mat_mul_v: rows a,b,c of a row-major 4x4 applied to (x,y,z,1)
inline void xform3(const float* m, const float* v, float* o) {
    o[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
    o[1] = m[4] * v[0] + m[5] * v[1] + m[6] * v[2] + m[7];
    o[2] = m[8] * v[0] + m[9] * v[1] + m[10] * v[2] + m[11];
}
// renderer.h:314-317 calculateSignedArea
inline float signed_area(const float* A, const float* B, const float* C) {
    return 0.5f * ((C[0] - A[0]) * (B[1] - A[1]) - (B[0] - A[0]) * (C[1] - A[1]));
}

// ---- rasteriser: renderer.cpp:190-257 (rasterization) ---------------------------------------
// tri_cam: camera-space triangle (after the pose), proj: 4x4, zbuf: W' x H' INT_MAX-initialised.
void raster_one(const float* tri_model, const float* pose, const float* proj, int32_t* zbuf,
                size_t width, size_t height, const int* roi) {
    float cam[9], clip[9];
    for (int v = 0; v < 3; v++) xform3(pose, tri_model + 3 * v, cam + 3 * v);   // renderer.cpp:278
    const float z[3] = {cam[2], cam[5], cam[8]};                                // renderer.cpp:282-286
    for (int v = 0; v < 3; v++) xform3(proj, cam + 3 * v, clip + 3 * v);        // renderer.cpp:288

    float pts[3][2];
    for (int v = 0; v < 3; v++) {                                               // renderer.cpp:197-204
        pts[v][0] = clip[3 * v + 0] / z[v] * width / 2.0f + width / 2.0f;
        pts[v][1] = clip[3 * v + 1] / z[v] * height / 2.0f + height / 2.0f;
    }
    // Defined behaviour for what is UB upstream: triangles whose screen coordinates are not
    // finite, or whose screen area has no finite reciprocal, are skipped.
    for (int v = 0; v < 3; v++)
        if (!std::isfinite(pts[v][0]) || !std::isfinite(pts[v][1])) return;

    float bbmin[2] = {FLT_MAX, FLT_MAX}, bbmax[2] = {-FLT_MAX, -FLT_MAX};
    float cmax[2] = {float(width - 1), float(height - 1)}, cmin[2] = {0, 0};
    size_t real_width = width;
    if (roi[2] > 0 && roi[3] > 0) {                                             // renderer.cpp:213-219
        cmin[0] = roi[0];
        cmin[1] = height - 1 - (roi[1] + roi[3] - 1);
        cmax[0] = (roi[0] + roi[2]) - 1;
        cmax[1] = height - 1 - roi[1];
        real_width = roi[2];
    }
    for (int i = 0; i < 3; i++)                                                 // renderer.cpp:222-227
        for (int j = 0; j < 2; j++) {
            bbmin[j] = sel_max(cmin[j], sel_min(bbmin[j], pts[i][j]));
            bbmax[j] = sel_min(cmax[j], sel_max(bbmax[j], pts[i][j]));
        }

    const float base_inv = 1 / signed_area(pts[0], pts[1], pts[2]);             // renderer.h:324
    if (!std::isfinite(base_inv)) return;

    for (size_t py = size_t(bbmin[1] + 0.5f); py <= bbmax[1]; py++) {           // renderer.cpp:230-231
        for (size_t px = size_t(bbmin[0] + 0.5f); px <= bbmax[0]; px++) {
            const float P[2] = {float(px), float(py)};
            const float beta = signed_area(pts[0], P, pts[2]) * base_inv;       // renderer.h:325-326
            const float gamma = signed_area(pts[0], pts[1], P) * base_inv;
            const float bc[3] = {1.0f - beta - gamma, beta, gamma};
            if (bc[0] < -0.0f || bc[1] < -0.0f || bc[2] < -0.0f ||             // renderer.cpp:234-235
                bc[0] > 1.0f || bc[1] > 1.0f || bc[2] > 1.0f) continue;
            const float oz[3] = {bc[0] / z[0], bc[1] / z[1], bc[2] / z[2]};     // renderer.cpp:237
            const float frag = (bc[0] + bc[1] + bc[2]) / (oz[0] + oz[1] + oz[2]);  // :244-245
            const size_t xw = px - roi[0];                                       // renderer.cpp:247-248
            const size_t yw = height - 1 - py - roi[1];
            const int32_t d = f2i_x86(frag + 0.5f);                              // renderer.cpp:250
            int32_t& slot = zbuf[xw + yw * real_width];
            if (d < slot) slot = d;                                              // renderer.cpp:253-254
        }
    }
}

// ---- geometry used by the ICP side ------------------------------------------------------------
struct V3 { float x, y, z; };

// scene/common.h:47-61 dep2pcd (dep in mm, integer type T; size_t pixel coordinates)
template <class T> inline V3 dep2pcd(size_t x, size_t y, T dep, const float* K) {
    if (dep == 0) return {0, 0, 0};
    const float z = dep / 1000.0f;
    const float xp = (x + size_t(0) - K[2]) / K[0] * z;
    const float yp = (y + size_t(0) - K[5]) / K[4] * z;
    return {xp, yp, z};
}

// scene/common.cpp:3-15 accumBilateral
inline void accum_bilateral(long delta, long i, long j, long* A, long* b, int threshold) {
    const long f = std::labs(delta) < threshold ? 1 : 0;
    const long fi = f * i, fj = f * j;
    A[0] += fi * i; A[1] += fi * j; A[3] += fj * j;
    b[0] += fi * delta; b[1] += fj * delta;
}

// scene/common.cpp:17-107 get_normal on a CV_16U image (CV_32S is saturate-converted first)
void normals_u16(const uint16_t* depth, int W, int H, const float* K, V3* normals) {
    for (size_t i = 0; i < size_t(W) * H; i++) normals[i] = {0, 0, 0};
    const int r = 5, dist_thr = 2000, diff_thr = 50;
    const int off[8][2] = {{-r, -r}, {0, -r}, {+r, -r}, {-r, 0}, {+r, 0}, {-r, +r}, {0, +r}, {+r, +r}};
    for (int y = r; y < H - r - 1; ++y) {
        for (int x = r; x < W - r - 1; ++x) {
            const uint16_t* c = depth + (size_t(y) * W + x);
            const long d = c[0];
            if (d >= dist_thr) continue;
            long A[4] = {0, 0, 0, 0}, b[2] = {0, 0};
            for (int k = 0; k < 8; k++)
                accum_bilateral(long(c[off[k][0] + off[k][1] * W]) - d, off[k][0], off[k][1], A, b, diff_thr);
            const long det = A[0] * A[3] - A[1] * A[1];
            const long ddx = A[3] * b[0] - A[1] * b[1];
            const long ddy = -A[1] * b[0] + A[0] * b[1];
            float nx = static_cast<float>(K[0] * ddx);
            float ny = static_cast<float>(K[4] * ddy);
            float nz = static_cast<float>(-det * d);
            const float len = sqrtf(nx * nx + ny * ny + nz * nz);
            if (len > 0) {
                const float inv = 1.0f / len;
                nx *= inv; ny *= inv; nz *= inv;
                normals[size_t(y) * W + x] = {nx, ny, nz};
            }
        }
    }
}

inline uint16_t sat_u16(int32_t v) { return (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v)); }

// scene/pcd_scene/pcd_scene.h:5-25 -- 52-byte node, identical field order
struct Node {
    int parent = -1, child1 = -1, child2 = -1;
    float split_v = 0;
    float bbox[6] = {0, 0, 0, 0, 0, 0};
    int split_dim = 0;
    int left = 0, right = 0;
};
static_assert(sizeof(Node) == 52, "Node_kdtree layout");

// scene/pcd_scene/pcd_scene.cpp:45-184 KDTree_cpu::build_tree
void build_tree(std::vector<V3>& pcd, std::vector<V3>& nrm, std::vector<Node>& nodes, int max_leaf) {
    const size_t n = pcd.size();
    std::vector<int> index(n), scratch(n);
    std::iota(index.begin(), index.end(), 0);
    nodes.assign(1, Node());
    nodes[0].left = 0; nodes[0].right = (int)n;

    size_t count = 1, gen_begin = 0, gen_end = 0;
    for (;;) {
        nodes.resize(count * 2 + 1);
        bool grew = false;
        gen_begin = gen_end; gen_end = count;
        for (size_t ni = gen_begin; ni < gen_end; ni++) {
            const int lo = nodes[ni].left, hi = nodes[ni].right;
            if (hi - lo <= max_leaf) continue;
            float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int k = lo; k < hi; k++) {
                const V3& p = pcd[index[k]];
                const float c[3] = {p.x, p.y, p.z};
                for (int a = 0; a < 3; a++) { if (c[a] > mx[a]) mx[a] = c[a]; if (c[a] < mn[a]) mn[a] = c[a]; }
            }
            int dim = 0; float split = 0, widest = -FLT_MAX;
            for (int a = 0; a < 3; a++) {
                const float span = mx[a] - mn[a], mid = (mn[a] + mx[a]) / 2;
                if (span > widest) { widest = span; dim = a; split = mid; }
            }
            int li = lo, ri = hi - 1;
            float low = -FLT_MAX, high = FLT_MAX;
            bool flip = true;
            for (int k = lo; k < hi; k++) {
                const V3& q = pcd[index[k]];
                const float p = dim == 0 ? q.x : (dim == 1 ? q.y : q.z);
                if (p == split) flip = !flip;
                if (p < split || (p == split && flip)) { scratch[li++] = index[k]; if (p > low) low = p; }
                else { scratch[ri--] = index[k]; if (p < high) high = p; }
            }
            split = (low + high) / 2;
            for (int k = lo; k < hi; k++) index[k] = scratch[k];

            Node& me = nodes[ni];
            me.child1 = (int)count; me.child2 = (int)count + 1;
            me.split_v = split; me.split_dim = dim;
            for (int a = 0; a < 3; a++) { me.bbox[2 * a] = mn[a]; me.bbox[2 * a + 1] = mx[a]; }
            nodes[count].left = lo; nodes[count].right = li; nodes[count].parent = (int)ni;
            nodes[count + 1].left = li; nodes[count + 1].right = hi; nodes[count + 1].parent = (int)ni;
            count += 2;
            grew = true;
        }
        if (!grew) break;
    }
    nodes.resize(count);
    std::vector<V3> tmp(n);
    for (size_t i = 0; i < n; i++) tmp[i] = pcd[index[i]];
    pcd = tmp;
    for (size_t i = 0; i < n; i++) tmp[i] = nrm[index[i]];
    nrm = tmp;
}

struct Scene {
    int kind = 0;  // 0 projective (depth_scene.h), 1 nearest-neighbour (pcd_scene.h)
    size_t W = 640, H = 480;
    float max_dist = 0.1f;
    float K[9];
    std::vector<V3> pcd, nrm;
    std::vector<Node> nodes;
};

inline float sq(float v) { return v * v; }                        // common.h:79-81 pow2
inline float absf(float v) { return (v > 0) ? v : (-v); }         // common.h:75-77 std__abs

// depth_scene.h:30-48 + common.h:63-73
inline bool query_projective(const Scene& s, const V3& p, V3& q, V3& n) {
    const int u = f2i_x86(p.x / p.z * s.K[0] + s.K[2] - size_t(0) + 0.5f);
    const int v = f2i_x86(p.y / p.z * s.K[4] + s.K[5] - size_t(0) + 0.5f);
    if (size_t(u) >= s.W || size_t(v) >= s.H || u < 0 || v < 0) return false;
    const size_t idx = size_t(u) + size_t(v) * s.W;
    q = s.pcd[idx];
    if (q.z <= 0 || absf(p.z - q.z) > s.max_dist) return false;
    n = s.nrm[idx];
    return true;
}

// pcd_scene.h:61-136: stackless descend / backtrack walk.  visits/tests are optional counters.
inline bool query_nn(const Scene& s, const V3& p, V3& q, V3& n, int* best_idx, long* visits, long* tests) {
    bool backtrack = false;
    int last = -1, cur = 0, best = 0;
    float best_d2 = FLT_MAX;
    while (cur >= 0) {
        const Node& nd = s.nodes[cur];
        if (visits) ++*visits;
        float diff = 0;
        if (nd.split_dim == 0) diff = p.x - nd.split_v;
        if (nd.split_dim == 1) diff = p.y - nd.split_v;
        if (nd.split_dim == 2) diff = p.z - nd.split_v;
        int near_child = nd.child1, far_child = nd.child1;
        if (diff < 0) far_child = nd.child2; else near_child = nd.child2;
        if (!backtrack) {
            if (nd.child1 < 0 || nd.child2 < 0) {
                for (int i = nd.left; i < nd.right; i++) {
                    if (tests) ++*tests;
                    const float d2 = sq(p.x - s.pcd[i].x) + sq(p.y - s.pcd[i].y) + sq(p.z - s.pcd[i].z);
                    if (d2 < best_d2) { best_d2 = d2; best = i; }
                }
                backtrack = true; last = cur; cur = nd.parent;
            } else { last = cur; cur = near_child; }
        } else {
            float lb = 0;  // distance to THIS node's box (pcd_scene.h:104-115)
            if (p.x < nd.bbox[0]) lb += sq(nd.bbox[0] - p.x); else if (p.x > nd.bbox[1]) lb += sq(nd.bbox[1] - p.x);
            if (p.y < nd.bbox[2]) lb += sq(nd.bbox[2] - p.y); else if (p.y > nd.bbox[3]) lb += sq(nd.bbox[3] - p.y);
            if (p.z < nd.bbox[4]) lb += sq(nd.bbox[4] - p.z); else if (p.z > nd.bbox[5]) lb += sq(nd.bbox[5] - p.z);
            if (last == near_child && lb <= best_d2) { last = cur; cur = far_child; backtrack = false; }
            else { last = cur; cur = nd.parent; }
        }
    }
    if (best_idx) *best_idx = best;
    if (best_d2 < sq(s.max_dist)) { q = s.pcd[best]; n = s.nrm[best]; return true; }
    return false;
}

inline bool query(const Scene& s, const V3& p, V3& q, V3& n) {
    return s.kind == 0 ? query_projective(s, p, q, n) : query_nn(s, p, q, n, nullptr, nullptr, nullptr);
}

// icp.h:138-208 thrust__pcd2Ab: adds this point's 29 terms into acc
inline void pcd2ab_add(const Scene& s, const V3& p, float* acc) {
    V3 q, n;
    if (!query(s, p, q, n)) return;
    const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
    const float r = dx * n.x + dy * n.y + dz * n.z;
    const float J[6] = {n.z * p.y - n.y * p.z, n.x * p.z - n.z * p.x, n.y * p.x - n.x * p.y, n.x, n.y, n.z};
    int k = 0;
    for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++) acc[k++] += J[i] * J[j];
    for (int i = 0; i < 6; i++) acc[21 + i] += J[i] * r;
    acc[27] += sq(dx) + sq(dy) + sq(dz);
    acc[28] += 1;
}

// icp.cpp:29-45 eigen_slover_666 + :7-17 TransformVector6dToMatrix4d, with Eigen's arithmetic
// restated: LDLT = Eigen/src/Cholesky/LDLT.h (unblocked, lower, diagonal pivoting) and
// AngleAxis -> Quaternion product -> toRotationMatrix (Eigen/src/Geometry).  A is 6x6 symmetric.
void solve_666(const float* Af, const float* bf, float* T16) {
    const int n = 6;
    double a[6][6], x[6];
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i][j] = (double)Af[i + 6 * j] + (i == j ? 0.01 * 1.0 : 0.01 * 0.0);
    for (int i = 0; i < n; i++) x[i] = (double)bf[i];
    int tr[6];
    double tmp[6];
    for (int k = 0; k < n; k++) {
        int piv = k; double big = std::fabs(a[k][k]);
        for (int i = k + 1; i < n; i++) if (std::fabs(a[i][i]) > big) { big = std::fabs(a[i][i]); piv = i; }
        tr[k] = piv;
        if (piv != k) {
            for (int j = 0; j < k; j++) std::swap(a[k][j], a[piv][j]);
            for (int i = piv + 1; i < n; i++) std::swap(a[i][k], a[i][piv]);
            std::swap(a[k][k], a[piv][piv]);
            for (int i = k + 1; i < piv; i++) std::swap(a[i][k], a[piv][i]);
        }
        if (k > 0) {
            for (int j = 0; j < k; j++) tmp[j] = a[j][j] * a[k][j];
            double acc = 0;
            for (int j = 0; j < k; j++) acc += a[k][j] * tmp[j];
            a[k][k] -= acc;
            for (int i = k + 1; i < n; i++) {
                double acc2 = 0;
                for (int j = 0; j < k; j++) acc2 += a[i][j] * tmp[j];
                a[i][k] -= acc2;
            }
        }
        if (k + 1 < n && std::fabs(a[k][k]) > 0.0)
            for (int i = k + 1; i < n; i++) a[i][k] /= a[k][k];
    }
    for (int k = 0; k < n; k++) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) x[i] -= a[i][j] * x[j];
    for (int i = 0; i < n; i++) { if (std::fabs(a[i][i]) > DBL_MIN) x[i] /= a[i][i]; else x[i] = 0; }
    for (int i = n - 1; i >= 0; i--) for (int j = i + 1; j < n; j++) x[i] -= a[j][i] * x[j];
    for (int k = n - 1; k >= 0; k--) if (tr[k] != k) std::swap(x[k], x[tr[k]]);

    // q = qz(x2) * qy(x1) * qx(x0)
    struct Q { double w, x, y, z; };
    auto mul = [](const Q& p, const Q& q) {
        return Q{p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
                 p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z, p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x};
    };
    const Q qz{std::cos(0.5 * x[2]), 0, 0, std::sin(0.5 * x[2])};
    const Q qy{std::cos(0.5 * x[1]), 0, std::sin(0.5 * x[1]), 0};
    const Q qx{std::cos(0.5 * x[0]), std::sin(0.5 * x[0]), 0, 0};
    const Q q = mul(mul(qz, qy), qx);
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    const double R[9] = {1.0 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.0 - (txx + tzz), tyz - twx,
                         txz - twy, tyz + twx, 1.0 - (txx + tyy)};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) T16[4 * i + j] = (float)R[3 * i + j];
        T16[4 * i + 3] = (float)x[3 + i];
    }
    T16[12] = 0; T16[13] = 0; T16[14] = 0; T16[15] = 1;
}

// icp.cpp:125-188 ICP_Point2Plane_cpu.  Returns the index of the pass that returned (0-based).
int icp_run(const Scene& s, V3* pts, size_t n, float rel_fit, float rel_rmse, int max_iter, float* out18) {
    float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float fitness = 0, rmse = 0;
    auto emit = [&]() { std::memcpy(out18, T, 64); out18[16] = rmse; out18[17] = fitness; };
    for (uint32_t iter = 0; iter <= (uint32_t)max_iter; iter++) {
        float S[29];
        for (int k = 0; k < 29; k++) S[k] = 0;
#pragma omp parallel
        {
            float loc[29];
            for (int k = 0; k < 29; k++) loc[k] = 0;
#pragma omp for schedule(static) nowait
            for (long i = 0; i < (long)n; i++) pcd2ab_add(s, pts[i], loc);
#pragma omp critical
            for (int k = 0; k < 29; k++) S[k] += loc[k];
        }
        const float prev_fit = fitness, prev_rmse = rmse;
        const float count = S[28], total = S[27];
        if (count == 0) { emit(); return (int)iter; }
        fitness = float(count) / n;
        rmse = std::sqrt(total / count);
        if (iter == (uint32_t)max_iter) { emit(); return (int)iter; }
        if (std::abs(fitness - prev_fit) < rel_fit && std::abs(rmse - prev_rmse) < rel_rmse) { emit(); return (int)iter; }

        float A[36], b[6];
        for (int i = 0; i < 6; i++) b[i] = S[21 + i];
        int shift = 0;
        for (int y = 0; y < 6; y++) for (int x = y; x < 6; x++) { A[x + y * 6] = S[shift]; A[y + x * 6] = S[shift]; shift++; }
        float E[16];
        solve_666(A, b, E);
        // icp.cpp:47-59 transform_pcd
#pragma omp parallel for
        for (long i = 0; i < (long)n; i++) {
            const V3 p = pts[i];
            pts[i].x = E[0] * p.x + E[1] * p.y + E[2] * p.z + E[3];
            pts[i].y = E[4] * p.x + E[5] * p.y + E[6] * p.z + E[7];
            pts[i].z = E[8] * p.x + E[9] * p.y + E[10] * p.z + E[11];
        }
        // geometry.h:292-298 + :107-111: T <- E * T, dot products summed from index 3 down to 0
        float Tn[16];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
            float acc = 0;
            for (int k = 3; k >= 0; k--) acc += E[4 * i + k] * T[4 * k + j];
            Tn[4 * i + j] = acc;
        }
        std::memcpy(T, Tn, 64);
    }
    emit();
    return max_iter;
}

// icp.cpp:73-117 depth2cloud_cpu, stride 1 (stride > 1 indexes out of bounds upstream, App. B-3)
template <class T>
long depth2cloud(const T* depth, uint32_t W, uint32_t H, const float* K, uint32_t tl_x, uint32_t tl_y, float* out, long cap) {
    long n = 0;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            const T d = depth[x + size_t(y) * W];
            if (d <= 0) continue;
            if (n < cap) {
                const float z = d / 1000.0f;
                out[3 * n + 0] = (x + tl_x - K[2]) / K[0] * z;
                out[3 * n + 1] = (y + tl_y - K[5]) / K[4] * z;
                out[3 * n + 2] = z;
            }
            n++;
        }
    return n;
}

void render_batch(const float* tris, size_t T, const float* poses, size_t P, size_t W, size_t H,
                  const float* proj, const int* roi, int32_t* out) {
    size_t rw = W, rh = H;
    if (roi[2] > 0 && roi[3] > 0) { rw = roi[2]; rh = roi[3]; }
    const size_t per = rw * rh;
    for (size_t i = 0; i < P * per; i++) out[i] = INT_MAX;                      // renderer.cpp:270
#pragma omp parallel for
    for (long p = 0; p < (long)P; p++)                                           // renderer.cpp:272-291
        for (size_t t = 0; t < T; t++) raster_one(tris + 9 * t, poses + 16 * p, proj, out + p * per, W, H, roi);
    for (size_t i = 0; i < P * per; i++) if (out[i] == INT_MAX) out[i] = 0;     // renderer.cpp:293-295
}

}  // namespace

extern "C" {

const char* orc_kind() { return "port"; }
void orc_set_threads(int n) { omp_set_num_threads(n); }
int orc_max_threads() { return omp_get_max_threads(); }

// renderer.cpp:161-185 compute_proj
void orc_compute_proj(const float* K, int W, int H, float near_, float far_, float* p) {
    for (int i = 0; i < 16; i++) p[i] = 0;
    p[0] = 2 * K[0] / W;
    p[1] = -(-2 * K[1] / W);
    p[2] = -(-2 * K[2] / W + 1);
    p[5] = -(2 * K[4] / H);
    p[6] = -(2 * K[5] / H - 1);
    p[10] = -(-(far_ + near_) / (far_ - near_));
    p[11] = -2 * far_ * near_ / (far_ - near_);
    p[14] = 1;
}

// renderer.cpp:259-298 render_cpu
void orc_render(const float* tris, size_t T, const float* poses, size_t P, size_t W, size_t H,
                const float* proj, const int* roi, int32_t* out) {
    render_batch(tris, T, poses, P, W, H, proj, roi, out);
}

// renderer.cpp:300-366 raw2depth_uint16_cpu / raw2mask_uint8_cpu / raw2depth_mask_cpu
void orc_raw2depth_mask(const int32_t* raw, size_t n, uint16_t* depth, uint8_t* mask) {
    for (size_t i = 0; i < n; i++) {
        if (depth) depth[i] = uint16_t(raw[i]);
        if (mask) mask[i] = (raw[i] > 0) ? 255 : 0;
    }
}

long orc_depth2cloud(const void* depth, int is_i32, uint32_t W, uint32_t H, const float* K,
                     uint32_t stride, uint32_t tl_x, uint32_t tl_y, float* out_pts, long cap) {
    if (stride != 1) return -1;
    return is_i32 ? depth2cloud((const int32_t*)depth, W, H, K, tl_x, tl_y, out_pts, cap)
                  : depth2cloud((const uint16_t*)depth, W, H, K, tl_x, tl_y, out_pts, cap);
}

static std::vector<uint16_t> to_u16(const void* depth, int is_i32, size_t n) {
    std::vector<uint16_t> d(n);
    if (is_i32) for (size_t i = 0; i < n; i++) d[i] = sat_u16(((const int32_t*)depth)[i]);   // common.cpp:22-23
    else std::memcpy(d.data(), depth, n * 2);
    return d;
}

void orc_get_normal(const void* depth, int is_i32, int W, int H, const float* K, float* normals) {
    std::vector<uint16_t> d = to_u16(depth, is_i32, size_t(W) * H);
    normals_u16(d.data(), W, H, K, (V3*)normals);
}

// depth_scene.cpp:3-35
void* orc_scene_projective_create(const void* depth, int is_i32, const float* K, size_t W, size_t H, float max_dist) {
    Scene* s = new Scene();
    s->kind = 0; s->W = W; s->H = H; s->max_dist = max_dist;
    std::memcpy(s->K, K, 36);
    s->pcd.resize(W * H); s->nrm.resize(W * H);
    for (size_t r = 0; r < H; r++)
        for (size_t c = 0; c < W; c++)
            s->pcd[c + r * W] = is_i32 ? dep2pcd(c, r, ((const uint32_t*)depth)[c + r * W], K)   // depth_scene.cpp:26
                                       : dep2pcd(c, r, ((const uint16_t*)depth)[c + r * W], K);
    std::vector<uint16_t> d = to_u16(depth, is_i32, W * H);
    normals_u16(d.data(), (int)W, (int)H, K, s->nrm.data());
    return s;
}
// pcd_scene.cpp:4-37 (+ build_tree)
void* orc_scene_nn_create(const void* depth, int is_i32, const float* K, size_t W, size_t H) {
    Scene* s = new Scene();
    s->kind = 1; s->W = W; s->H = H; s->max_dist = 0.1f;
    std::memcpy(s->K, K, 36);
    std::vector<uint16_t> d = to_u16(depth, is_i32, W * H);
    std::vector<V3> normal(W * H);
    normals_u16(d.data(), (int)W, (int)H, K, normal.data());
    for (size_t r = 0; r < H; r++)
        for (size_t c = 0; c < W; c++) {
            const uint16_t v = d[c + r * W];
            if (v > 0) { s->pcd.push_back(dep2pcd(c, r, v, K)); s->nrm.push_back(normal[c + r * W]); }
        }
    build_tree(s->pcd, s->nrm, s->nodes, 10);
    return s;
}
// a Scene_nn over caller-provided arrays (already leaf-ordered points + 52-byte nodes)
void* orc_scene_nn_from_arrays(const float* pcd, const float* nrm, long n_pts, const void* nodes, long n_nodes, float max_dist) {
    Scene* s = new Scene();
    s->kind = 1; s->max_dist = max_dist;
    s->pcd.resize(n_pts); s->nrm.resize(n_pts); s->nodes.resize(n_nodes);
    std::memcpy(s->pcd.data(), pcd, n_pts * 12); std::memcpy(s->nrm.data(), nrm, n_pts * 12);
    std::memcpy(s->nodes.data(), nodes, n_nodes * sizeof(Node));
    return s;
}
void orc_scene_destroy(void* h) { delete (Scene*)h; }
void orc_scene_sizes(void* h, long* n_pts, long* n_nodes) {
    Scene* s = (Scene*)h;
    *n_pts = (long)s->pcd.size(); *n_nodes = (long)s->nodes.size();
}
void orc_scene_get(void* h, float* pcd, float* normal, void* nodes) {
    Scene* s = (Scene*)h;
    std::memcpy(pcd, s->pcd.data(), s->pcd.size() * 12);
    std::memcpy(normal, s->nrm.data(), s->nrm.size() * 12);
    if (s->kind == 1 && nodes) std::memcpy(nodes, s->nodes.data(), s->nodes.size() * sizeof(Node));
}

void orc_query(void* h, const float* pts, size_t n, float* dst, float* nrm, uint8_t* valid) {
    const Scene& s = *(Scene*)h;
    for (size_t i = 0; i < n; i++) {
        V3 q, nn;
        const bool v = query(s, ((const V3*)pts)[i], q, nn);
        valid[i] = v;
        if (v) { ((V3*)dst)[i] = q; ((V3*)nrm)[i] = nn; }
    }
}
// nearest-neighbour index + traversal statistics of the reference walk (Scene_nn only)
void orc_query_nn_stats(void* h, const float* pts, size_t n, int* idx, long* visits, long* tests) {
    const Scene& s = *(Scene*)h;
    *visits = 0; *tests = 0;
    for (size_t i = 0; i < n; i++) { V3 q, nn; query_nn(s, ((const V3*)pts)[i], q, nn, idx ? idx + i : nullptr, visits, tests); }
}
void orc_pcd2ab(void* h, const float* pts, size_t n, float* out29) {
    const Scene& s = *(Scene*)h;
    for (int k = 0; k < 29; k++) out29[k] = 0;
    for (size_t i = 0; i < n; i++) pcd2ab_add(s, ((const V3*)pts)[i], out29);
}
void orc_solve_666(const float* A, const float* b, float* T16) { solve_666(A, b, T16); }

int orc_icp(void* h, float* pts, size_t n, float rel_fit, float rel_rmse, int max_iter, float* out18) {
    return icp_run(*(Scene*)h, (V3*)pts, n, rel_fit, rel_rmse, max_iter, out18);
}

// Same pipeline and schedules as ref_pipeline (oracle/ref_glue.cpp); see BASELINE.md section 3.
double orc_pipeline(void* hv, const float* tris, size_t T, const float* poses, size_t P, size_t W, size_t H,
                    const float* proj, const float* K, float rel_fit, float rel_rmse, int max_iter,
                    int schedule, float* results, long* n_pts) {
    const Scene& s = *(Scene*)hv;
    const int roi[4] = {0, 0, 0, 0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<int32_t> depth(P * W * H);
    render_batch(tris, T, poses, P, W, H, proj, roi, depth.data());
    auto one = [&](size_t i) {
        std::vector<float> cloud(W * H * 3);
        long n = depth2cloud(depth.data() + i * W * H, (uint32_t)W, (uint32_t)H, K, 0, 0, cloud.data(), (long)(W * H));
        if (n_pts) n_pts[i] = n;
        icp_run(s, (V3*)cloud.data(), (size_t)n, rel_fit, rel_rmse, max_iter, results + 18 * i);
    };
    if (schedule == 0) {
        for (size_t i = 0; i < P; i++) one(i);
    } else {
        omp_set_max_active_levels(1);
#pragma omp parallel for schedule(dynamic, 1)
        for (long i = 0; i < (long)P; i++) one((size_t)i);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
