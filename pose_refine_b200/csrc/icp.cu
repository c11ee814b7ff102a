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
#include "common.cuh"
#include "solver.cuh"
#include <float.h>
#include <limits.h>
#include <algorithm>

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

// The stop logic of one hypothesis after its sums S[29] are known (icp.cu:179-212), run by one
// thread.  Returns nothing; updates the state and, when the hypothesis returns, its result.
__device__ void finish_pass(HypState* st, const float* S, unsigned n_points, pr_icp_criteria crit,
                            pr_registration_result* res) {
    const float count = S[28], total = S[27];
    const int iter = st->pass;
    bool ret = false;
    float fitness = st->fitness, rmse = st->rmse;
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
        float T[16], Tn[12];
#pragma unroll
        for (int i = 0; i < 12; i++) T[i] = st->T[i];
        T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
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
        st->pass = iter + 1;
    } else {
#pragma unroll
        for (int i = 0; i < 12; i++) res->transformation[i] = st->T[i];
        res->transformation[12] = 0.f; res->transformation[13] = 0.f; res->transformation[14] = 0.f; res->transformation[15] = 1.f;
        res->inlier_rmse = rmse; res->fitness = fitness;
        st->done = 1;
    }
}

// plan: chunk table + state initialisation.  One CTA; n_hyp is at most a few thousand.
__global__ void __launch_bounds__(kIcpThreads)
icp_plan_kernel(const uint32_t* __restrict__ counts, uint32_t n_hyp, uint32_t chunk_points, HypState* __restrict__ state,
                uint32_t* __restrict__ chunk_hyp, uint32_t max_chunks, uint32_t* __restrict__ total_chunks,
                pr_registration_result* __restrict__ results) {
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
    if (threadIdx.x == 0) *total_chunks = min(s_carry, max_chunks);
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
    HypState* state; uint32_t* chunk_hyp; uint32_t* total_chunks; float* partials;
    size_t max_chunks, bytes;
};

// chunk size: big chunks amortise the per-CTA reduction; small batches need more CTAs than SMs
inline uint32_t pick_chunk_points(size_t n_hyp, size_t capacity_points) {
    uint32_t chunk = 4096;
    while (chunk > 512 && capacity_points / chunk + n_hyp < (size_t)kNumSMs * 4) chunk >>= 1;
    return chunk;
}

inline IcpWs carve_icp_ws(void* base, size_t n_hyp, size_t capacity_points) {
    IcpWs ws;
    // sized for the smallest chunk pick_chunk_points can return
    ws.max_chunks = capacity_points / 512 + n_hyp + 1;
    char* w = (char*)base;
    size_t used = 0;
    auto take = [&](size_t bytes) { char* p = w + used; used += icp_align_up(bytes, 256); return p; };
    ws.state = (HypState*)take(n_hyp * sizeof(HypState));
    ws.chunk_hyp = (uint32_t*)take(ws.max_chunks * 4);
    ws.total_chunks = (uint32_t*)take(4);
    ws.partials = (float*)take(ws.max_chunks * kPartialStride * 4);
    ws.bytes = used;
    return ws;
}

template <class SceneT>
int run_icp(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp, size_t capacity_points,
            const SceneT& scene, pr_icp_criteria crit, pr_registration_result* results_dev, int flags,
            void* workspace_dev, size_t workspace_bytes, cudaStream_t stream) {
    if (!pts_dev || !offsets_dev || !counts_dev || !results_dev || !workspace_dev) return PR_ERR_INVALID_ARGUMENT;
    if (crit.max_iteration < 0 || n_hyp > 0x7FFFFFFFull / 64 || capacity_points > 0xFFFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    if (n_hyp == 0) return PR_OK;
    IcpWs ws = carve_icp_ws(workspace_dev, n_hyp, capacity_points);
    if (workspace_bytes < ws.bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    const uint32_t chunk = pick_chunk_points(n_hyp, capacity_points);
    // persistent grid: 3 CTAs per SM (register-limited), never more CTAs than chunks can exist
    const unsigned grid = (unsigned)std::min<size_t>(capacity_points / chunk + n_hyp + 1, (size_t)kNumSMs * 3);
    icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts_dev, (uint32_t)n_hyp, chunk, ws.state, ws.chunk_hyp,
                                                   (uint32_t)ws.max_chunks, ws.total_chunks, results_dev);
    for (int it = 0; it <= crit.max_iteration; it++)
        icp_pass_kernel<SceneT><<<grid, kIcpThreads, 0, stream>>>(pts_dev, offsets_dev, counts_dev, ws.chunk_hyp, ws.total_chunks, chunk,
                                                                  ws.state, ws.partials, scene, crit, results_dev, nullptr);
    count_launch(2 + (uint64_t)crit.max_iteration);
    if (flags & PR_ICP_UPDATE_POINTS) {
        icp_apply_kernel<<<grid, 256, 0, stream>>>(pts_dev, offsets_dev, counts_dev, ws.chunk_hyp, ws.total_chunks, chunk, ws.state);
        count_launch();
    }
    PR_LAUNCH_CHECK();
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
    const size_t ws_bytes = carve_icp_ws(nullptr, 1, n).bytes;
    char* scratch = nullptr;
    PR_CUDA_TRY(cudaMalloc((void**)&scratch, 256 + 256 + ws_bytes));
    uint32_t* counts = (uint32_t*)scratch;
    uint32_t* offsets = (uint32_t*)(scratch + 128);
    pr_registration_result* res = (pr_registration_result*)(scratch + 256);
    const uint32_t h_counts = (uint32_t)n, h_off = 0;
    cudaMemcpyAsync(counts, &h_counts, 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(offsets, &h_off, 4, cudaMemcpyHostToDevice, stream);
    IcpWs ws = carve_icp_ws(scratch + 512, 1, n);
    const uint32_t chunk = pick_chunk_points(1, n);
    const unsigned grid = (unsigned)std::min<size_t>(n / chunk + 2, (size_t)kNumSMs * 3);
    pr_icp_criteria crit = {0.f, 0.f, 0};
    cudaMemsetAsync(out29_dev, 0, 29 * 4, stream);
    icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts, 1, chunk, ws.state, ws.chunk_hyp, (uint32_t)ws.max_chunks, ws.total_chunks, res);
    icp_pass_kernel<SceneT><<<grid, kIcpThreads, 0, stream>>>(pts_dev, offsets, counts, ws.chunk_hyp, ws.total_chunks, chunk, ws.state,
                                                              ws.partials, scene, crit, res, out29_dev);
    count_launch(2);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(scratch);
    return e == cudaSuccess ? PR_OK : (int)e;
}

}  // namespace prb

using namespace prb;

extern "C" {

size_t pr_icp_workspace_bytes(size_t n_hyp, size_t capacity_points) {
    return carve_icp_ws(nullptr, n_hyp, capacity_points).bytes;
}

int pr_icp_projective_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                            size_t capacity_points, const pr_scene_projective* scene, pr_icp_criteria criteria,
                            pr_registration_result* results_dev, int flags,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream) {
    ProjScene s;
    int rc = make_proj_scene(scene, s);
    if (rc != PR_OK) return rc;
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, criteria, results_dev, flags,
                   workspace_dev, workspace_bytes, as_stream(stream));
}

int pr_icp_nn_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                    size_t capacity_points, const pr_scene_nn* scene, pr_icp_criteria criteria,
                    pr_registration_result* results_dev, int flags,
                    void* workspace_dev, size_t workspace_bytes, pr_stream_t stream) {
    NnScene s;
    int rc = make_nn_scene(scene, s);
    if (rc != PR_OK) return rc;
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, criteria, results_dev, flags,
                   workspace_dev, workspace_bytes, as_stream(stream));
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
