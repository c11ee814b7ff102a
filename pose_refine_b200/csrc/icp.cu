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

// thrust__pcd2Ab::operator() (icp.h:138-208): adds one correspondence into the 29 running sums.
__device__ __forceinline__ void accumulate(float* acc, float px, float py, float pz, const Corr& c) {
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
    acc[28] += 1.0f;
}

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

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The stop logic of one hypothesis after its sums S[29] are known (icp.cu:179-212), run by one
// thread.  Updates the state and, when the hypothesis returns, its result.  RELEASE = true is the
// persistent driver's flavour: state is read past L1 and `pass` / `done` are published with release
// stores after everything else, because other CTAs of the SAME launch are waiting on them.
template <bool RELEASE>
__device__ void finish_pass_impl(HypState* st, const float* S, unsigned n_points, pr_icp_criteria crit,
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
    st->fitness = fitness; st->rmse = rmse;
    float T[16];
#pragma unroll
    for (int i = 0; i < 12; i++) T[i] = RELEASE ? __ldcg(&st->T[i]) : st->T[i];
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
    if (!ret) {
        float A[36], b[6], E[16];
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = S[21 + i];
        int shift = 0;
        for (int y = 0; y < 6; y++)
            for (int x = y; x < 6; x++) { A[x + y * 6] = S[shift]; A[y + x * 6] = S[shift]; shift++; }   // icp.cu:198-205
        solve_666(A, b, E);                                    // icp.cu:207
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
        for (int i = 0; i < 12; i++) st->T[i] = Tn[i];
        if (RELEASE) { __threadfence(); st_release(reinterpret_cast<unsigned*>(&st->pass), (unsigned)(iter + 1)); }
        else st->pass = iter + 1;
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) res->transformation[i] = T[i];
        res->inlier_rmse = rmse; res->fitness = fitness;
        if (RELEASE) { __threadfence(); st_release(reinterpret_cast<unsigned*>(&st->done), 1u); }
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
                pr_registration_result* __restrict__ results, unsigned* __restrict__ next_item) {
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
            for (unsigned j = 0; j < v; j++) if (begin + j < max_chunks) chunk_hyp[begin + j] = h;
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
// persistent driver
// ---------------------------------------------------------------------------------------------
constexpr int kTilePts = 2048;                 // points per shared-memory tile (24 KB)
constexpr int kTileFloats = kTilePts * 3;
constexpr uint32_t kPersistChunk = 8192;       // points per work item (4 tiles)

struct IcpCtl {            // device-side control block
    unsigned next_item;    // work-item claim counter
    unsigned total_chunks;
    unsigned pad[30];
};

// packed projective scene: two aligned arrays, per pixel {qx,qy,qz,nx} (16 B) and {ny,nz} (8 B)
struct PackedScene {
    int W, H;
    float fW, fH;
    float max_dist;
    float fx, fy, cx, cy;
    const float4* qn;
    const float2* n2;
};

__global__ void __launch_bounds__(256)
scene_pack_kernel(const float* __restrict__ pcd, const float* __restrict__ nrm, size_t n_px, float4* __restrict__ qn,
                  float2* __restrict__ n2) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_px) return;
    qn[i] = make_float4(pcd[3 * i], pcd[3 * i + 1], pcd[3 * i + 2], nrm[3 * i]);
    n2[i] = make_float2(nrm[3 * i + 1], nrm[3 * i + 2]);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
// TMA bulk copy global -> shared (1-D), completion signalled on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
struct ItemDesc {
    unsigned item, h, pass, n_h, n_pts, n_tiles;
    const float* g;        // first point of the item
    bool valid;
};

__device__ __forceinline__ ItemDesc describe_item(unsigned item, unsigned n_items, unsigned total, const float* pts,
                                                  const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts,
                                                  const uint32_t* __restrict__ chunk_hyp, const HypState* state) {
    ItemDesc d;
    d.item = item;
    d.valid = item < n_items;
    d.h = d.pass = d.n_h = d.n_pts = d.n_tiles = 0;
    d.g = pts;
    if (d.valid) {
        d.pass = item / total;
        const unsigned c = item - d.pass * total;
        d.h = __ldg(chunk_hyp + c);
        d.n_h = __ldg(counts + d.h);
        const unsigned first = (c - __ldcg(&state[d.h].chunk_begin)) * kPersistChunk;
        d.n_pts = min(kPersistChunk, d.n_h - first);
        d.n_tiles = (d.n_pts + kTilePts - 1) / kTilePts;
        d.g = pts + 3 * ((size_t)__ldg(offsets + d.h) + first);
    }
    return d;
}

// the four points a thread owns in a tile: floats [12g, 12g+12) -> three conflict-free LDS.128
struct Quad { float x[4], y[4], z[4]; };
__device__ __forceinline__ Quad load_quad(const float* tile, unsigned g) {
    const float4 a = *reinterpret_cast<const float4*>(tile + 12 * g);
    const float4 b = *reinterpret_cast<const float4*>(tile + 12 * g + 4);
    const float4 c = *reinterpret_cast<const float4*>(tile + 12 * g + 8);
    Quad q;
    q.x[0] = a.x; q.y[0] = a.y; q.z[0] = a.z;
    q.x[1] = a.w; q.y[1] = b.x; q.z[1] = b.y;
    q.x[2] = b.z; q.y[2] = b.w; q.z[2] = c.x;
    q.x[3] = c.y; q.y[3] = c.z; q.z[3] = c.w;
    return q;
}

__device__ __forceinline__ void transform(const float* T, float x, float y, float z, float& px, float& py, float& pz) {
    // transform_pcd_cuda (icp.cu:147-149) with the accumulated transform
    px = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
    py = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
    pz = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
}

// one tile, packed projective scene: per thread two quads; the four gathers of a quad are issued
// before any of them is consumed.
__device__ __forceinline__ void compute_tile(const PackedScene& s, const float* tile, unsigned n, const float* T, float* acc) {
#pragma unroll 1
    for (unsigned g = threadIdx.x; 4 * g < n; g += kIcpThreads) {
        const Quad q = load_quad(tile, g);
        float px[4], py[4], pz[4];
        int idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            transform(T, q.x[k], q.y[k], q.z[k], px[k], py[k], pz[k]);
            // Scene_projective::query pixel selection (depth_scene.h:30-36, common.h:63-73), exact ops
            const float uf = addf(addf(mulf(divf(px[k], pz[k]), s.fx), s.cx), 0.5f);
            const float vf = addf(addf(mulf(divf(py[k], pz[k]), s.fy), s.cy), 0.5f);
            const bool in = (uf > -1.0f && uf < s.fW && vf > -1.0f && vf < s.fH) && (4 * g + k < n);
            idx[k] = in ? ((int)uf + (int)vf * s.W) : -1;
        }
        float4 A[4];
        float2 B[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (idx[k] >= 0) {
                A[k] = __ldg(s.qn + idx[k]);
                B[k] = __ldg(s.n2 + idx[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (idx[k] >= 0) {
                const float dz = pz[k] - A[k].z;
                const float adz = (dz > 0.f) ? dz : -dz;
                if (!(A[k].z <= 0.f || adz > s.max_dist)) {       // depth_scene.h:42
                    Corr c;
                    c.qx = A[k].x; c.qy = A[k].y; c.qz = A[k].z; c.nx = A[k].w; c.ny = B[k].x; c.nz = B[k].y;
                    accumulate(acc, px[k], py[k], pz[k], c);
                }
            }
        }
    }
}

// one tile, any scene with a per-point query() (nearest neighbour)
template <class SceneT>
__device__ __forceinline__ void compute_tile(const SceneT& s, const float* tile, unsigned n, const float* T, float* acc) {
#pragma unroll 1
    for (unsigned i = threadIdx.x; i < n; i += kIcpThreads) {
        float px, py, pz;
        transform(T, tile[3 * i], tile[3 * i + 1], tile[3 * i + 2], px, py, pz);
        Corr c;
        if (query(s, px, py, pz, c)) accumulate(acc, px, py, pz, c);
    }
}

template <class SceneT>
__global__ void __launch_bounds__(kIcpThreads, 2)
icp_persistent_kernel(const float* __restrict__ pts, size_t capacity_points, const uint32_t* __restrict__ offsets,
                      const uint32_t* __restrict__ counts, const uint32_t* __restrict__ chunk_hyp, IcpCtl* ctl,
                      HypState* state, float* partials, SceneT scene, pr_icp_criteria crit,
                      pr_registration_result* results) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile_buf[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw) + kTileFloats};
    __shared__ __align__(8) uint64_t s_full[2];
    __shared__ float s_part[kIcpWarps][32];
    __shared__ float s_sum[32];
    __shared__ unsigned s_claim;
    __shared__ int s_flag;

    const unsigned total = __ldcg(&ctl->total_chunks);
    const unsigned n_items = total * (unsigned)(crit.max_iteration + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_claim = atomicAdd(&ctl->next_item, 1u);
    }
    __syncthreads();
    ItemDesc cur = describe_item(s_claim, n_items, total, pts, offsets, counts, chunk_hyp, state);
    __syncthreads();
    if (threadIdx.x == 0) s_claim = atomicAdd(&ctl->next_item, 1u);
    __syncthreads();
    ItemDesc nxt = describe_item(s_claim, n_items, total, pts, offsets, counts, chunk_hyp, state);

    // a tile may be fetched by TMA when its global address is 16-byte aligned and the copy, rounded up
    // to 16 bytes, stays inside the point buffer; otherwise all threads copy it.
    auto tile_bytes = [&](const ItemDesc& d, unsigned t, const float*& src, unsigned& n) -> unsigned {
        src = d.g + (size_t)t * kTileFloats;
        n = min((unsigned)kTilePts, d.n_pts - t * kTilePts);
        const unsigned bytes = (n * 12 + 15) & ~15u;
        const bool ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                        (reinterpret_cast<uintptr_t>(src) + bytes <= reinterpret_cast<uintptr_t>(pts) + capacity_points * 12);
        return ok ? bytes : 0u;
    };
    unsigned parity[2] = {0, 0};
    bool by_tma[2] = {false, false};
    auto issue = [&](const ItemDesc& d, unsigned t, int b) {
        const float* src; unsigned n;
        const unsigned bytes = tile_bytes(d, t, src, n);
        by_tma[b] = bytes != 0;
        if (by_tma[b] && threadIdx.x == 0) {
            mbar_expect_tx(&s_full[b], bytes);
            tma_load_1d(tile_buf[b], src, bytes, &s_full[b]);
        }
    };

    int buf = 0;
    unsigned t_cur = 0;
    if (cur.valid) issue(cur, 0, 0);
    float T[12];
    float acc[32];
    bool skip = false;
    while (cur.valid) {
        // ---- prefetch the tile after this one (same item, or first tile of the next claimed item)
        const bool last_tile = (t_cur + 1 == cur.n_tiles);
        if (!last_tile) issue(cur, t_cur + 1, buf ^ 1);
        else if (nxt.valid) issue(nxt, 0, buf ^ 1);

        // ---- first tile of an item: wait until the hypothesis has finished the previous pass
        if (t_cur == 0) {
            if (threadIdx.x == 0) {
                const HypState* st = state + cur.h;
                int flag;
                for (;;) {
                    if (ld_acquire(reinterpret_cast<const unsigned*>(&st->done))) { flag = 1; break; }
                    if (ld_acquire(reinterpret_cast<const unsigned*>(&st->pass)) >= cur.pass) { flag = 0; break; }
                    __nanosleep(100);
                }
                s_flag = flag;
            }
            __syncthreads();
            skip = s_flag != 0;
            if (!skip) {
#pragma unroll
                for (int i = 0; i < 12; i++) T[i] = __ldcg(&state[cur.h].T[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = 0.f;
        }

        // ---- consume the tile
        const float* src; unsigned n;
        tile_bytes(cur, t_cur, src, n);
        if (by_tma[buf]) {
            mbar_wait(&s_full[buf], parity[buf]);
            parity[buf] ^= 1;
        } else if (!skip) {
            for (unsigned i = threadIdx.x; i < n * 3; i += kIcpThreads) tile_buf[buf][i] = src[i];
            __syncthreads();
        }
        if (!skip) compute_tile(scene, tile_buf[buf], n, T, acc);

        // ---- last tile of the item: reduce, deposit, maybe finish the pass of this hypothesis
        if (last_tile) {
            if (!skip) {
                HypState* st = state + cur.h;
                const float mine = warp_transpose_reduce(acc);
                s_part[warp][lane] = mine;
                __syncthreads();
                const unsigned c = cur.item - cur.pass * total;
                if (warp == 0) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < kIcpWarps; w++) s += s_part[w][lane];
                    __stcg(partials + (size_t)c * kPartialStride + lane, s);
                    __threadfence();
                    __syncwarp();
                    int is_last = 0;
                    if (lane == 0) is_last = (atomicAdd(&st->arrived, 1u) == __ldcg(&st->n_chunks) - 1) ? 1 : 0;
                    is_last = __shfl_sync(0xffffffffu, is_last, 0);
                    if (is_last) {
                        __threadfence();
                        float sum = 0.f;
                        const unsigned cb = __ldcg(&st->chunk_begin), nc = __ldcg(&st->n_chunks);
                        for (unsigned j = 0; j < nc; j++) sum += __ldcg(partials + (size_t)(cb + j) * kPartialStride + lane);
                        s_sum[lane] = sum;
                        __syncwarp();
                        if (lane == 0) {
                            st->arrived = 0;
                            finish_pass_release(st, s_sum, cur.n_h, crit, results + cur.h);
                        }
                    }
                }
            }
            cur = nxt;
            t_cur = 0;
            __syncthreads();
            if (threadIdx.x == 0) s_claim = atomicAdd(&ctl->next_item, 1u);
            __syncthreads();
            nxt = describe_item(s_claim, n_items, total, pts, offsets, counts, chunk_hyp, state);
        } else {
            t_cur++;
            __syncthreads();   // everybody is done with tile_buf[buf] before it is refilled
        }
        buf ^= 1;
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
    HypState* state; uint32_t* chunk_hyp; IcpCtl* ctl; float* partials; float4* packed; float2* packed2;
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
    ws.ctl = (IcpCtl*)take(sizeof(IcpCtl));
    ws.partials = (float*)take(ws.max_chunks * kPartialStride * 4);
    ws.packed = (float4*)take(scene_pixels * 16);
    ws.packed2 = (float2*)take(scene_pixels * 8);
    ws.bytes = used;
    return ws;
}

inline bool use_pass_driver() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PR_ICP_IMPL"); v = (e && strcmp(e, "pass") == 0) ? 1 : 0; }
    return v == 1;
}

template <class SceneT>
int persistent_grid(int* grid_out) {
    static int cached = 0;
    if (!cached) {
        const int smem = 2 * kTileFloats * 4;
        PR_CUDA_TRY(cudaFuncSetAttribute(icp_persistent_kernel<SceneT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 0, dev = 0, sms = 0;
        PR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_persistent_kernel<SceneT>, kIcpThreads, smem));
        PR_CUDA_TRY(cudaGetDevice(&dev));
        PR_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cached = std::max(1, occ) * std::max(1, sms);
    }
    *grid_out = cached;
    return PR_OK;
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
    int grid = 0;
    int rc = persistent_grid<PScene>(&grid);
    if (rc != PR_OK) return rc;
    const size_t max_items = (capacity_points / kPersistChunk + n_hyp + 1) * (size_t)(crit.max_iteration + 1);
    if (max_items > 0x7FFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    grid = (int)std::min<size_t>((size_t)grid, max_items);
    icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts_dev, (uint32_t)n_hyp, kPersistChunk, ws.state, ws.chunk_hyp,
                                                   (uint32_t)ws.max_chunks, &ws.ctl->total_chunks, results_dev, &ws.ctl->next_item);
    icp_persistent_kernel<PScene><<<grid, kIcpThreads, 2 * kTileFloats * 4, stream>>>(
        pts_dev, capacity_points, offsets_dev, counts_dev, ws.chunk_hyp, ws.ctl, ws.state, ws.partials, pscene, crit, results_dev);
    count_launch(2);
    if (flags & PR_ICP_UPDATE_POINTS) {
        icp_apply_kernel<<<kNumSMs * 4, 256, 0, stream>>>(pts_dev, offsets_dev, counts_dev, ws.chunk_hyp, &ws.ctl->total_chunks,
                                                          kPersistChunk, ws.state);
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
    ps.fx = s.fx; ps.fy = s.fy; ps.cx = s.cx; ps.cy = s.cy; ps.qn = ws.packed; ps.n2 = ws.packed2;
    if (!use_pass_driver()) {
        scene_pack_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, stream>>>(s.pcd, s.nrm, n_px, ws.packed, ws.packed2);
        count_launch();
    }
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, ps, criteria, results_dev, flags, ws, stream);
}

int pr_icp_nn_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                    size_t capacity_points, const pr_scene_nn* scene, pr_icp_criteria criteria,
                    pr_registration_result* results_dev, int flags,
                    void* workspace_dev, size_t workspace_bytes, pr_stream_t stream) {
    NnScene s;
    int rc = make_nn_scene(scene, s);
    if (rc != PR_OK) return rc;
    rc = check_icp_args(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, criteria, results_dev, workspace_dev);
    if (rc != PR_OK) return rc;
    if (n_hyp == 0) return PR_OK;
    IcpWs ws = carve_icp_ws(workspace_dev, n_hyp, capacity_points, 0);
    if (workspace_bytes < ws.bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, s, criteria, results_dev, flags, ws, as_stream(stream));
}

int pr_solve_666(const float A[36], const float b[6], float T[16]) {
    if (!A || !b || !T) return PR_ERR_INVALID_ARGUMENT;
    solve_666(A, b, T);
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
