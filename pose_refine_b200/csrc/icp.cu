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
// kernel that rewrites every point, another sync.  Here ONE launch runs all passes of all hypotheses.
//
// icp_hyp_kernel (the driver that ships): a HYPOTHESIS is owned, for all of its passes, by one thread-block
// CLUSTER of C CTAs (C = 2, 4 or 8, chosen per launch).  Every warp of the cluster owns a fixed slice of the
// hypothesis' points and re-reads it every pass with plain coalesced loads, one group ahead -- from L2, not HBM:
// nothing else touches those lines between the passes of a hypothesis, and the cluster size is chosen so that the
// clouds of all hypotheses in flight fit the L2 (pick_cluster).  A pass applies the hypothesis' ACCUMULATED 4x4 to the
// ORIGINAL points on the fly (no write-back, the fusion the reference's notes.md:3 asks for), looks the correspondences
// up, accumulates the 29 sums in registers -- two points per FFMA2 --, reduces them with a register-transposing warp
// butterfly, adds the warps of a CTA through shared memory and the CTAs of the cluster through DISTRIBUTED shared
// memory (st.async onto the peers' mbarriers, one 4-byte store per lane and peer), and every CTA then evaluates
// fitness / rmse / the stop tests exactly as icp.cu:181-194 and solves the 6x6 system redundantly (identical inputs,
// identical bits) -- no broadcast of the new 4x4, no global-memory flags, no polling, no tickets.  Clusters claim
// hypotheses from one global counter, longest first (icp_order_kernel); several CTAs of different clusters share an
// SM, so the serial reduce -> solve section of one hypothesis overlaps the point pass of another.
// (Tried and measured, git history "icp_pipe_kernel": two hypotheses in flight per cluster, every warp alternating between
// them, the serial section on one warp behind mbarriers only -- the barrier stalls (14 % of warp time) vanish, but either all
// 512 clouds are in flight at once (135 MB > L2: hit rate 97 % -> 62 %, 1.66 ms) or the clusters must be twice as wide, which
// halves the slices and doubles the per-pass overhead per point (1.58 ms).  One hypothesis per cluster of two CTAs: 1.42 ms.)
//
// icp_pass_kernel (first generation): one launch per pass with the reference's exact arithmetic; kept as the
// in-library cross-check (flags & PR_ICP_REFERENCE_ARITHMETIC) and for pr_pcd2ab_*.
#include "icp_device.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <cstddef>

namespace prb {

// ---------------------------------------------------------------------------------------------
// stop logic shared by both drivers
// ---------------------------------------------------------------------------------------------
// The stop logic of one hypothesis after its sums S[29] are known (icp.cu:179-212), run by one thread.
// T (12 floats, rows 0..2), fit_rmse (previous pass' fitness, rmse: the "backup" of icp.cu:179) are updated in
// place; returns true when the hypothesis returns (res is then written when non-null).
template <bool FAST_SOLVER>
__device__ __forceinline__ bool finish_pass_core(float* T12, float* fit_rmse, const float* S, unsigned n_points, int iter,
                                                 pr_icp_criteria crit, pr_registration_result* res) {
    const float count = S[28], total = S[27];
    bool ret = false;
    float fitness = fit_rmse[0], rmse = fit_rmse[1];
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
    fit_rmse[0] = fitness; fit_rmse[1] = rmse;
    float T[16];
#pragma unroll
    for (int i = 0; i < 12; i++) T[i] = T12[i];
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
    if (!ret) {
        float Sr[29], E[16];
#pragma unroll
        for (int i = 0; i < 29; i++) Sr[i] = S[i];
#ifdef PR_DBG_NOSOLVE      // what-if build (timing only, wrong results): the cost of the solve in the pass-to-pass chain
        for (int i = 0; i < 16; i++) E[i] = (i % 5 == 0) ? 1.f : 0.f;
        E[3] = Sr[21] * 1e-9f;
#else
        if (FAST_SOLVER) solve_666_fast(Sr, E);                // unpack icp.cu:198-205 + solve icp.cu:207
        else solve_666_unrolled(Sr, E);
#endif
        // result.transformation_ = extrinsic * result.transformation_ (icp.cu:212); geometry.h:107-111
        // sums each dot product from index 3 down to 0.
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float acc = 0.f;
#pragma unroll
                for (int k = 3; k >= 0; k--) acc = addf(acc, mulf(E[4 * i + k], T[4 * k + j]));
                T12[4 * i + j] = acc;
            }
    } else if (res) {
#pragma unroll
        for (int i = 0; i < 16; i++) res->transformation[i] = T[i];
        res->inlier_rmse = rmse; res->fitness = fitness;
    }
    return ret;
}

__device__ __noinline__ void finish_pass(HypState* st, const float* S, unsigned n_points, pr_icp_criteria crit,
                                         pr_registration_result* res) {
    float fr[2] = {st->fitness, st->rmse};
    const int iter = st->pass;
    const bool ret = finish_pass_core<false>(st->T, fr, S, n_points, iter, crit, res);
    st->fitness = fr[0]; st->rmse = fr[1];
    if (ret) st->done = 1; else st->pass = iter + 1;
}

// plan of the per-pass driver: chunk table + state initialisation.  One CTA; n_hyp is at most a few thousand.
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
            for (unsigned j = 0; j < v; j++) {
                if (begin + j >= max_chunks) break;
                chunk_hyp[begin + j] = h;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_chunks = min(s_carry, max_chunks);
}

// One pass over all hypotheses with the reference's arithmetic.  Persistent grid: every CTA walks the chunk table
// with a stride of gridDim.x; one chunk = up to chunk_points points of one hypothesis.
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
// hypothesis-resident driver
// ---------------------------------------------------------------------------------------------
#ifndef PR_HYP_WARPS
#define PR_HYP_WARPS 4
#endif
#ifndef PR_HYP_MINB
#define PR_HYP_MINB 4
#endif
#ifndef PR_HYP_ILP
#define PR_HYP_ILP 4
#endif
#ifndef PR_HYP_CLUSTER          // 0: chosen per launch (pick_cluster)
#define PR_HYP_CLUSTER 0
#endif
#ifndef PR_NN_WARPS
#define PR_NN_WARPS 8
#endif
#ifndef PR_NN_MINB
#define PR_NN_MINB 4          // the tree walk hides latency with warps: 32 per SM at 64 registers (measured: 2 CTAs 141 ms, 3 124 ms, 4 112 ms on C3)
#endif
constexpr int kMaxCluster = 8;                 // portable cluster limit
constexpr int kIlp = PR_HYP_ILP;               // points per lane per group of the projective loop (gathers in flight per lane)

struct HypCtl {            // device-side control block
    unsigned next_hyp;     // claim counter
    unsigned pad[31];
};

// Claim order of a launch: hypotheses by descending point count (ties by index), and the control block reset.  The clusters
// take hypotheses from one counter; with more hypotheses than clusters the kernel ends when the last cluster does, and
// longest-first keeps that end short (512 clouds of 18-26 k points on 296 clusters: the two hypotheses of the unluckiest
// cluster add up to 43 k points instead of 48.5 k).  Rank by counting: n_hyp^2 compares, a few microseconds.
__global__ void __launch_bounds__(256)
icp_order_kernel(const uint32_t* __restrict__ counts, unsigned n_hyp, uint32_t* __restrict__ order, HypCtl* __restrict__ ctl) {
    __shared__ uint32_t s_c[256];
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) ctl->next_hyp = 0u;
    const uint32_t mine = i < n_hyp ? counts[i] : 0u;
    unsigned rank = 0;
    for (unsigned base = 0; base < n_hyp; base += 256) {
        __syncthreads();
        s_c[threadIdx.x] = base + threadIdx.x < n_hyp ? counts[base + threadIdx.x] : 0u;
        __syncthreads();
        const unsigned m = min(256u, n_hyp - base);
        for (unsigned j = 0; j < m; j++) {
            const uint32_t c = s_c[j];
            rank += (c > mine || (c == mine && base + j < i)) ? 1u : 0u;
        }
    }
    if (i < n_hyp) order[rank] = i;
}

// packed projective scene: one 32-byte record per pixel (= one L2 sector per correspondence).  Two separate
// arrays ({qx,qy,qz,nx} 16 B + {ny,nz} 8 B) cost 3x the time: measured 2.08 ms vs 0.71 ms with the second gather removed.
struct PackedScene {
    int W, H;
    float max_dist;
    float fx, fy, cx, cy;
    float one;                    // 1.0f the compiler cannot see (project_pair)
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
__device__ __forceinline__ float lds32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// thread-block cluster plumbing (raw PTX: the kernel is launched with a runtime cluster size)
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned mapa_u32(unsigned smem_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(unsigned addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// remote store that signals the destination CTA's mbarrier (complete_tx of 4 bytes): data and notification in one
// asynchronous operation -- no cluster-wide barrier, no gpu-scope fence (barrier.cluster.arrive.release costs a MEMBAR.GPU)
__device__ __forceinline__ void st_async_f32(unsigned remote_addr, float v, unsigned remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
                 ::"r"(remote_addr), "f"(v), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(unsigned addr, unsigned v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// one scene record with a single 256-bit load (LDG.E.256, sm_100); two 128-bit loads of the same sector: +5 %.
// The byte address is formed by one mad.wide.u32 (index * 32 + base): left to the compiler, the "rejected -> record 0"
// select is applied to both halves of the 64-bit offset and the base is added by an IADD3 / IADD3.X pair.
__device__ __forceinline__ void load_rec(const PackedScene& s, int idx, float4& A, float2& B) {
    unsigned long long addr;
    asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(addr) : "r"(idx), "l"(s.rec));
    float u0, u1;       // padding words of the record
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(A.x), "=f"(A.y), "=f"(A.z), "=f"(A.w), "=f"(B.x), "=f"(B.y), "=f"(u0), "=f"(u1) : "l"(addr));
    (void)u0; (void)u1;
}
__device__ __forceinline__ void transform(const float* T, float x, float y, float z, float& px, float& py, float& pz) {
    // transform_pcd_cuda (icp.cu:147-149) with the accumulated transform
    px = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
    py = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
    pz = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
}

// ---- pixel selection of the hot loop: pcd2dep (common.h:63-73), EXACT ---------------------------------------
// u = int(px / pz * fx + cx + 0.5f): an IEEE division, a product and two sums, none of them fused (the CPU build
// has no FMA).  The division is the textbook sequence nvcc itself emits for div.rn.f32 on its fast path --
// r = rcp.approx(b), one Newton step, q0 = a*r, q = fma(fma(-b, q0, a), r, q0) -- which returns the correctly
// rounded quotient whenever no intermediate leaves the normal range; two points per instruction (f32x2) and the
// reciprocal shared by the x and the y quotient.  |pz| outside [2^-60, 2^60] (a range in which a = px, py of any
// magnitude that can still land inside the image keeps every intermediate normal) raises `odd`, and the warp then
// redoes its slice with the per-point path below, which uses div.rn.f32 itself.  So for every point the pixel is
// the one the CPU oracle picks, bit for bit (tests/test_gpu_parity.py::test_correspondences_*).
__device__ __forceinline__ f2_t rcp2_refined(f2_t z) {
    float z0, z1, r0, r1;
    unpk2(z, z0, z1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(z0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(z1));
    const f2_t r = pk2(r0, r1);
    const f2_t e = fma2(neg2(z), r, bc2(1.0f));
    return fma2(r, e, r);
}
__device__ __forceinline__ f2_t div2_rn(f2_t a, f2_t b, f2_t r) {       // r = rcp2_refined(b)
    const f2_t q0 = mul2(a, r);
    const f2_t e = fma2(neg2(b), q0, a);
    return fma2(e, r, q0);
}
__device__ __forceinline__ bool pz_in_exact_range(float pz) {           // false for NaN, too
    // Only the lower side needs a test: a point with |pz| > 2^60 can never pass the depth gate (|pz - qz| <= max_dist,
    // qz a depth in metres), whatever pixel it is given, and rcp.approx flushes to zero beyond 2^126, which lands on
    // pixel (cx, cy) -- inside the image, gate still false.
    return fabsf(pz) >= 8.673617e-19f;                                  // 2^-60
}
// pixel coordinates (as floats, before truncation) of a pair of points.
// ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with -fmad=false (it honours .rn only for the
// scalar forms), which would round q*fx + cx once instead of twice.  The sum is therefore written as fma(t, one, cx) with
// `one` = 1.0f read from the kernel parameters: the same value as add.rn(t, cx), but with the product as a MULTIPLICAND,
// where nothing can be contracted into it.  (SASS: FMUL2, FFMA2, FADD2 per coordinate -- checked by
// tests/test_cabi_symbols.py::test_pixel_selection_is_not_contracted.)
__device__ __forceinline__ void project_pair(const PackedScene& s, f2_t px, f2_t py, f2_t pz, f2_t& uf, f2_t& vf) {
    const f2_t r = rcp2_refined(pz);
    const f2_t one = bc2(s.one);
    uf = add2(fma2(mul2(div2_rn(px, pz, r), bc2(s.fx)), one, bc2(s.cx)), bc2(0.5f));
    vf = add2(fma2(mul2(div2_rn(py, pz, r), bc2(s.fy)), one, bc2(s.cy)), bc2(0.5f));
}
// int(v) of pcd2dep is truncation toward zero; "0 <= int(v) < W" (depth_scene.h:33-37) as one unsigned compare.
// cvt.rzi saturates out-of-range values (rejected by the compare) and maps NaN to 0 -- a NaN point reaches
// pixel (0,0) and is then rejected by the depth gate (|NaN - qz| <= max_dist is false), as upstream rejects it
// (x86 turns NaN into INT_MIN).
__device__ __forceinline__ bool pixel_of(const PackedScene& s, float uf, float vf, int& idx) {
    const int u = __float2int_rz(uf), v = __float2int_rz(vf);
    idx = v * s.W + u;
    return ((unsigned)u < (unsigned)s.W) & ((unsigned)v < (unsigned)s.H);
}

// 32*NP*2 consecutive points of a tile against the packed projective scene: lane l owns points l, 32+l, 64+l, ...
// so each gather instruction covers 32 CONSECUTIVE model points -- neighbouring scene pixels; points 2j and 2j+1
// of a lane form pair j.  The gathers of a lane are issued before any of them is consumed.  No branches: a
// rejected point is turned into q = p, n = 0 by the selects that also move the gathered values into the pair
// registers (selects run on the ALU pipe; the FP32 pipe is the busy one).  A non-finite transformed point would
// turn into NaN sums here (0 * inf); the caller detects that and redoes the slice with the per-point path.
// The 4x4 is re-read from shared memory by every group (three broadcast LDS.128): its 12 registers are then live only
// while the group's points are transformed, not across the gathers and the 28 x 2 accumulators (which is what decides
// whether 16 warps fit an SM without spilling).
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// The points of one group as a lane holds them: point k of the lane is point (group start + 32 k + lane) of the slice.
template <int NP> struct GroupPts { float x[2 * NP], y[2 * NP], z[2 * NP]; };
// Plain coalesced loads, straight from global memory (L2-resident: the cluster re-reads its hypothesis every pass and
// nothing else touches those lines in between): the three 4-byte loads of a point instruction cover 384 contiguous
// bytes per warp, and x / y / z of the same lines hit in L1.  The next group is requested as soon as the current one has
// been transformed, into the same registers, so its latency hides behind projection, gathers and accumulation; no
// shared-memory tiles, no mbarriers, no per-tile bookkeeping in the loop (the TMA ring this replaces spent 10 % of the
// kernel's instructions and a quarter of its shared memory there; measured 1.62 -> see DESIGN.md 6.1).
template <int NP, bool CLAMP>
__device__ __forceinline__ void load_group(GroupPts<NP>& P, const float* __restrict__ g, unsigned first, unsigned n) {
#pragma unroll
    for (int k = 0; k < 2 * NP; k++) {
        unsigned i = first + 32 * k;
        if (CLAMP) i = (i < n) ? i : 0u;
        const float* p = g + 3 * (size_t)i;
        P.x[k] = __ldg(p); P.y[k] = __ldg(p + 1); P.z[k] = __ldg(p + 2);
    }
}
// One group: transform, project, gather, accumulate.  TAIL: lanes whose points lie past the end of the slice (n) are
// masked.  PREFETCH: 0 none, 1 the next group unclamped, 2 clamped to the slice.
template <int NP, bool TAIL, int PREFETCH>
__device__ __forceinline__ void group_projective(const PackedScene& s, GroupPts<NP>& P, const float* __restrict__ g, unsigned first,
                                                 unsigned n, unsigned t_addr, AccP& acc, bool& odd) {
    f2_t px[NP], py[NP], pz[NP];
    int idx[2 * NP];
    bool ok[2 * NP];
    {
        float T[12];
        const float4 r0 = lds128(t_addr), r1 = lds128(t_addr + 16), r2 = lds128(t_addr + 32);
        T[0] = r0.x; T[1] = r0.y; T[2] = r0.z; T[3] = r0.w; T[4] = r1.x; T[5] = r1.y; T[6] = r1.z; T[7] = r1.w;
        T[8] = r2.x; T[9] = r2.y; T[10] = r2.z; T[11] = r2.w;
#pragma unroll
        for (int j = 0; j < NP; j++) {
            float x0 = P.x[2 * j], y0 = P.y[2 * j], z0 = P.z[2 * j], x1 = P.x[2 * j + 1], y1 = P.y[2 * j + 1], z1 = P.z[2 * j + 1];
            if (TAIL) {     // a masked lane holds whatever lies behind the slice: make it the finite point (0, 0, 1)
                const bool in0 = first + 32 * (2 * j) < n, in1 = first + 32 * (2 * j + 1) < n;
                x0 = in0 ? x0 : 0.f; y0 = in0 ? y0 : 0.f; z0 = in0 ? z0 : 1.f;
                x1 = in1 ? x1 : 0.f; y1 = in1 ? y1 : 0.f; z1 = in1 ? z1 : 1.f;
            }
            const f2_t x = pk2(x0, x1), y = pk2(y0, y1), z = pk2(z0, z1);
            // transform_pcd_cuda (icp.cu:147-149) with the accumulated transform, same FMA chain as transform()
            px[j] = fma2(bc2(T[2]), z, fma2(bc2(T[1]), y, fma2(bc2(T[0]), x, bc2(T[3]))));
            py[j] = fma2(bc2(T[6]), z, fma2(bc2(T[5]), y, fma2(bc2(T[4]), x, bc2(T[7]))));
            pz[j] = fma2(bc2(T[10]), z, fma2(bc2(T[9]), y, fma2(bc2(T[8]), x, bc2(T[11]))));
        }
    }
    if (PREFETCH == 1) load_group<NP, false>(P, g, first + 64 * NP, n);
    if (PREFETCH == 2) load_group<NP, true>(P, g, first + 64 * NP, n);
#pragma unroll
    for (int j = 0; j < NP; j++) {
        bool in0 = true, in1 = true;
        if (TAIL) { in0 = first + 32 * (2 * j) < n; in1 = first + 32 * (2 * j + 1) < n; }
        f2_t uf, vf;
        project_pair(s, px[j], py[j], pz[j], uf, vf);
        float u0, u1, v0, v1, pz0, pz1;
        unpk2(uf, u0, u1); unpk2(vf, v0, v1); unpk2(pz[j], pz0, pz1);
        ok[2 * j] = pixel_of(s, u0, v0, idx[2 * j]) & in0;
        ok[2 * j + 1] = pixel_of(s, u1, v1, idx[2 * j + 1]) & in1;
        odd |= (in0 && !pz_in_exact_range(pz0)) | (in1 && !pz_in_exact_range(pz1));
    }
    // every lane gathers: a rejected point reads record 0 (a cache hit) instead of predicating the load, which keeps
    // the 8 destination registers of each LDG.256 out of the compiler's "may hold an older value" bookkeeping
    float4 A[2 * NP];
    float2 B[2 * NP];
#pragma unroll
    for (int k = 0; k < 2 * NP; k++) load_rec(s, ok[k] ? idx[k] : 0, A[k], B[k]);
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
        accumulate_pair(acc, px[j], py[j], pz[j], qx, qy, qz, nx, ny, nz);
        acc.cnt += (v0 ? 1 : 0) + (v1 ? 1 : 0);
    }
}

// a warp's slice of one hypothesis, one pass: n points at g (global memory)
#ifndef PR_HYP_UNROLL
#define PR_HYP_UNROLL 1
#endif
constexpr int kHypUnroll = PR_HYP_UNROLL;
__device__ __forceinline__ void compute_slice(const PackedScene& s, const float* __restrict__ g, unsigned n, bool may_overread,
                                              unsigned t_addr, AccP& acc, bool& odd) {
    const unsigned lane = threadIdx.x & 31;
    constexpr int NP = kIlp / 2;
    constexpr unsigned kGroup = 32 * kIlp;
    static_assert(kIlp % 2 == 0, "groups of pairs");
    if (n == 0) return;
    GroupPts<NP> P;
    unsigned first = lane;
    if (may_overread) {
        // the prefetch of the group behind a full group may run up to one group past the slice: other clouds' points or
        // the buffer's padding, never consumed unmasked
        const unsigned n_full = n - n % kGroup;
        load_group<NP, false>(P, g, first, n);
#pragma unroll kHypUnroll
        for (; first < n_full; first += kGroup) group_projective<NP, false, 1>(s, P, g, first, n, t_addr, acc, odd);
        if (n_full < n) group_projective<NP, true, 0>(s, P, g, first, n, t_addr, acc, odd);
    } else {
        load_group<NP, true>(P, g, first, n);
#pragma unroll 1
        for (; first - lane < n; first += kGroup) group_projective<NP, true, 2>(s, P, g, first, n, t_addr, acc, odd);
    }
}

// per-point query against the packed scene with the reference's own operations (div.rn.f32): the robust path
__device__ __forceinline__ bool query(const PackedScene& s, float px, float py, float pz, Corr& c) {
    const float uf = addf(addf(mulf(divf(px, pz), s.fx), s.cx), 0.5f);
    const float vf = addf(addf(mulf(divf(py, pz), s.fy), s.cy), 0.5f);
    if (!(uf > -1.0f && uf < (float)s.W && vf > -1.0f && vf < (float)s.H)) return false;
    float4 A; float2 B;
    load_rec(s, (int)vf * s.W + (int)uf, A, B);
    if (!(A.z > 0.f && fabsf(pz - A.z) <= s.max_dist)) return false;
    c.qx = A.x; c.qy = A.y; c.qz = A.z; c.nx = A.w; c.ny = B.x; c.nz = B.y;
    return true;
}
// a warp's whole slice again, point by point from global memory (taken only for non-finite / out-of-range points)
__device__ __noinline__ float slow_slice(const PackedScene& s, const float* __restrict__ g, unsigned n, const float* T) {
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

// A warp's share of one hypothesis against the kd-tree: 32-point rows dealt round-robin to the G warps of the cluster (row
// k goes to warp k mod G), not one contiguous slice per warp -- the cost of a walk varies over the object, neighbouring points
// cost about the same, and with contiguous slices the warps of a pass finished far apart (a quarter of all stall samples at
// the pass-end barrier; 107 -> 63 ms per C3 step together with the two-phase walk of nn_search_packed_t).  `cache`
// (nullable): one int per point of the hypothesis, the winner of the previous pass, which seeds this pass' search; pass 0
// only writes it.
// (Tried: lanes as persistent workers -- a lane whose walk is over takes the warp's next point at once instead of waiting
// for the slowest walk of its row.  Lane occupancy rises, but the 32 lanes then work on points of different rows, their node
// and leaf fetches stop sharing cache lines, and the kernel is bound by exactly those fetches: 63 -> 70 ms.)
#ifndef PR_NN_TABLE
#define PR_NN_TABLE 2
#endif
constexpr int kNnQueue = 64;      // per warp: point indices waiting for the tree walk (<= 31 left over + 32 new)
__device__ __forceinline__ void nn_corr_of(const PackedNnScene& s, int best_i, Corr& c) {
    const float4 q = __ldg(s.pts4 + best_i);
    c.qx = q.x; c.qy = q.y; c.qz = q.z;
    c.nx = __ldg(s.nrm + 3 * best_i); c.ny = __ldg(s.nrm + 3 * best_i + 1); c.nz = __ldg(s.nrm + 3 * best_i + 2);
}
// (Tried: the 29 running sums in shared memory instead of registers -- the walk then compiles without its 316 bytes of spills
// and 5 or 6 CTAs fit an SM, but 30 KB of shared memory per CTA come out of the L1 the walk lives in: 56 -> 61 ms at 4 CTAs
// per SM, 63 ms at 5, 72 ms at 6.  L1 capacity, not occupancy, is what this kernel is short of.)
// `queue` (shared memory, kNnQueue ints per warp): the grid (nn_grid_query) answers most points at once; the others are
// compacted into the queue and walked through the tree 32 at a time, so the long walks run on full warps instead of
// holding up the 30 lanes of their row that were done after a dozen distance tests.
__device__ __forceinline__ void compute_share_nn(const PackedNnScene& s, const float* __restrict__ g, unsigned n, unsigned me, unsigned G,
                                                 unsigned t_addr, AccT& acc, int* __restrict__ cache, bool use_cache, int* __restrict__ queue) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    // the accumulated transform stays in shared memory (three LDS.128 per point): the walk needs the registers
    auto xform = [&](unsigned i, float& px, float& py, float& pz) {
        const float4 r0 = lds128(t_addr), r1 = lds128(t_addr + 16), r2 = lds128(t_addr + 32);
        const float T[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
        transform(T, __ldg(g + 3 * i), __ldg(g + 3 * i + 1), __ldg(g + 3 * i + 2), px, py, pz);
    };
    NnGridParams gp;
    gp.enabled = 0u;
    if (s.grid.params && s.nodes) gp = *s.grid.params;
    unsigned qn = 0;
    // the tree walk of the first `count` (<= 32) queued points
    auto drain = [&](unsigned count) {
        if (lane < count) {
            const unsigned i = (unsigned)queue[lane];
            float px, py, pz;
            xform(i, px, py, pz);
            Corr c;
            int found;
            const int hint = (cache && use_cache) ? cache[i] : -1;
            if (query(s, px, py, pz, c, hint, found)) acc_add(acc, px, py, pz, c);
            if (cache) cache[i] = found;
        }
        __syncwarp();
        if (qn > 32) {
            int t = 0;
            if (lane + 32 < qn) t = queue[lane + 32];
            __syncwarp();
            if (lane + 32 < qn) queue[lane] = t;
            __syncwarp();
        }
        qn -= count;
    };
#pragma unroll 1
    for (unsigned base = me * 32; base < n; base += G * 32) {
        const unsigned i = base + lane;
        bool walk = false;
        if (i < n) {
            float px, py, pz;
            xform(i, px, py, pz);
            int best_i;
            unsigned tests = 0;
            if (gp.enabled && nn_grid_query<false>(s, gp, px, py, pz, best_i, tests)) {
                Corr c;
                nn_corr_of(s, best_i, c);
                acc_add(acc, px, py, pz, c);
                if (cache) cache[i] = best_i;
            } else {
                walk = true;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, walk);
        if (walk) queue[qn + __popc(m & lt_mask)] = (int)i;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) drain(32);
    }
    if (qn) drain(qn);
}

// a foreign kd-tree the packed encoding cannot hold (flag raised by nn_pack_nodes_kernel): walk the reference layout
__device__ __forceinline__ void resolve_scene(PackedScene&) {}
__device__ __forceinline__ void resolve_scene(PackedNnScene& s) {
    if (s.unsupported && *s.unsupported) s.nodes = nullptr;
}

template <class SceneT> struct HypTraits {
    static constexpr bool kProjective = std::is_same<SceneT, PackedScene>::value;
    static constexpr int kWarps = kProjective ? PR_HYP_WARPS : PR_NN_WARPS;
    static constexpr int kMinBlocks = kProjective ? PR_HYP_MINB : PR_NN_MINB;
    using Acc = typename std::conditional<kProjective, AccP, AccT>::type;
    static constexpr size_t kExtraSmem = kProjective ? 0 : (size_t)kTopNodes * 32 + (size_t)kWarps * kNnQueue * 4;     // top levels of the kd-tree + the warps' walk queues
};
// dynamic shared memory of a CTA: state | sums | spare | mbarriers | per-warp partials | cluster slots (2 parities) | tree top
template <int kWarps> __host__ __device__ constexpr size_t hyp_fixed_smem() { return 64 + 128 + 64 + 64 + (size_t)kWarps * 128 + 2 * kMaxCluster * 128; }
template <int kWarps> constexpr size_t hyp_smem(size_t extra) { return hyp_fixed_smem<kWarps>() + extra; }

// state words behind the 12 floats of T
enum { kStFit = 12, kStRmse = 13, kStDone = 14, kStHyp = 15 };

// one thread per CTA: stop tests + solve on the CTA's copy of the hypothesis state (st: T[12], fitness, rmse, done).
// Every CTA of a cluster runs it on identical inputs; only the leader (res != nullptr) writes the result.
__device__ __noinline__ void finish_hyp(float* st, const float* S_in, unsigned n_points, int pass, pr_icp_criteria crit,
                                        pr_registration_result* res, float* final_T12) {
    float T12[12], fr[2] = {st[kStFit], st[kStRmse]}, S[29];
#pragma unroll
    for (int i = 0; i < 12; i++) T12[i] = st[i];
#pragma unroll
    for (int i = 0; i < 29; i++) S[i] = S_in[i];
    const bool ret = finish_pass_core<true>(T12, fr, S, n_points, pass, crit, res);
    st[kStFit] = fr[0]; st[kStRmse] = fr[1];
    if (ret) {
        reinterpret_cast<unsigned*>(st)[kStDone] = 1u;
        if (final_T12) {        // what the result holds: the transform before this pass' (never applied) update
#pragma unroll
            for (int i = 0; i < 12; i++) final_T12[i] = T12[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = T12[i];
    }
}

template <class SceneT>
__global__ void __launch_bounds__(HypTraits<SceneT>::kWarps * 32, HypTraits<SceneT>::kMinBlocks)
icp_hyp_kernel(const float* __restrict__ pts, size_t capacity_points, const uint32_t* __restrict__ offsets,
               const uint32_t* __restrict__ counts, unsigned n_hyp, HypCtl* ctl, const uint32_t* __restrict__ order,
               float* __restrict__ final_T, SceneT scene, pr_icp_criteria crit, pr_registration_result* __restrict__ results,
               float* __restrict__ out32) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Tr = HypTraits<SceneT>;
    constexpr int kWarps = Tr::kWarps;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned C = cluster_nctarank(), rank = cluster_ctarank();
    // dynamic shared memory: state | sums | spare point | mbarriers | per-warp partials | cluster slots | tree top
    float* s_T = reinterpret_cast<float*>(smem_raw);     // [12] T, then fitness, rmse, done, hypothesis id
    float* s_S = s_T + 16;                               // [32]: the hypothesis' sums of this pass
    float* s_safe = s_S + 32;                            // [3] the spare point (0, 0, 1) masked lanes read (+ padding to 64 bytes)
    float* s_xbar = s_safe + 16;                         // three mbarriers: the peers' partials have arrived (one per pass parity), tree top copied
    float* s_part = s_xbar + 16;                         // [kWarps][32]
    float* s_cl = s_part + kWarps * 32;                  // [2][kMaxCluster][32]: CTA partials of the cluster, by pass parity
    volatile unsigned* s_w = reinterpret_cast<volatile unsigned*>(s_T);
    const unsigned smem0 = smem_u32(smem_raw);
    const unsigned extra0 = smem0 + (unsigned)hyp_fixed_smem<kWarps>();
    const uintptr_t pts_end = reinterpret_cast<uintptr_t>(pts) + capacity_points * 12;

    SceneT sc = scene;
    resolve_scene(sc);
    if (threadIdx.x < 3) s_safe[threadIdx.x] = (threadIdx.x == 2) ? 1.f : 0.f;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(s_xbar), 1); mbar_init(smem_u32(s_xbar) + 8, 1); mbar_init(smem_u32(s_xbar) + 16, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned xph = 0;                   // warp 0: bit p = parity to wait for on the exchange barrier of pass parity p
    if constexpr (!Tr::kProjective) {
        // kd-tree: nodes [0, n_top) = the top levels (breadth-first numbering) -> shared memory, one TMA bulk copy
        const int n_top = sc.nodes ? min(sc.n_nodes, kTopNodes) : 0;
        if (n_top > 0) {
            if (warp == 0) {
                const unsigned tb = smem_u32(s_xbar) + 16;
                if (lane == 0) {
                    mbar_expect_tx(tb, (unsigned)n_top * 32);
                    tma_load_1d(extra0, sc.nodes, (unsigned)n_top * 32, tb);
                }
                mbar_wait(tb, 0);
            }
            sc.top = reinterpret_cast<const float4*>(smem_raw + (extra0 - smem0));
            sc.n_top = n_top;
        }
        __syncthreads();
    }
    if (C > 1) cluster_sync_all();      // every CTA of the cluster is running before its shared memory is written remotely

    for (;;) {
        // ---- claim a hypothesis for the cluster
        if (C > 1) {
            if (rank == 0 && threadIdx.x == 0) {
                const unsigned k = atomicAdd(&ctl->next_hyp, 1u);
                const unsigned hh = k < n_hyp ? order[k] : n_hyp;
                for (unsigned r = 0; r < C; r++) st_cluster_u32(mapa_u32(smem_u32(s_T + kStHyp), r), hh);
            }
            cluster_sync_all();
        } else {
            if (threadIdx.x == 0) {
                const unsigned k = atomicAdd(&ctl->next_hyp, 1u);
                s_w[kStHyp] = k < n_hyp ? order[k] : n_hyp;
            }
            __syncthreads();
        }
        const unsigned h = s_w[kStHyp];
        if (h >= n_hyp) break;
        if (threadIdx.x < 12) s_T[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.f : 0.f;       // identity (icp.h:29-31)
        if (threadIdx.x == 12) { s_T[kStFit] = 0.f; s_T[kStRmse] = 0.f; s_w[kStDone] = 0u; }
        __syncthreads();

        // ---- this warp's slice: whole 64-point units, spread evenly over the C * kWarps warps of the cluster
        const unsigned n = counts[h];
        const unsigned units = (n + 63) >> 6;
        const unsigned G = C * kWarps, g = rank * kWarps + warp;
        const unsigned ubase = units / G, urem = units - ubase * G;
        const unsigned my_units = ubase + (g < urem ? 1u : 0u);
        const unsigned first_pt = (g * ubase + min(g, urem)) << 6;
        const unsigned n_mine = my_units ? min(my_units << 6, n - first_pt) : 0u;
        const float* gsl = pts + 3 * ((size_t)offsets[h] + first_pt);
        // the slice is read straight from global memory every pass (L2-resident: nothing else touches those lines between
        // the passes of a hypothesis): one group ahead in registers for the projective loop (compute_slice), point by
        // point for the tree walk, whose cost per point is two orders of magnitude above the 12 bytes
        [[maybe_unused]] bool may_overread = false;
        if constexpr (Tr::kProjective) {
            constexpr unsigned kGroup = 32 * kIlp;
            may_overread = reinterpret_cast<uintptr_t>(gsl) + 12ull * (n_mine - n_mine % kGroup + 2 * kGroup) <= pts_end;
        }

        for (int pass = 0;; pass++) {
            const unsigned t_addr = smem_u32(s_T);
            typename Tr::Acc acc;
            acc_zero(acc);
            bool odd = false;
            if constexpr (Tr::kProjective) compute_slice(sc, gsl, n_mine, may_overread, t_addr, acc, odd);
            else compute_share_nn(sc, pts + 3 * (size_t)offsets[h], n, g, G, t_addr, acc, sc.cache ? sc.cache + (size_t)offsets[h] : nullptr, pass > 0,
                                  reinterpret_cast<int*>(smem_raw + (extra0 - smem0) + kTopNodes * 32) + warp * kNnQueue);
            // ---- the warp's 29 sums -> lane l holds sum l
            float v[32];
            acc_unpack(acc, v);
            float mine = warp_transpose_reduce(v);
            if constexpr (Tr::kProjective) {
                if (__any_sync(0xffffffffu, odd || mine != mine)) mine = slow_slice(sc, gsl, n_mine, s_T);
            }
            s_part[warp * 32 + lane] = mine;
            __syncthreads();
            const unsigned pp = (unsigned)(pass & 1);
            const unsigned slot = pp * (kMaxCluster * 32);
            if (warp == 0) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kWarps; w++) s += s_part[w * 32 + lane];         // fixed order
                s_cl[slot + rank * 32 + lane] = s;
                if (C > 1) {
                    // every CTA sends its partial to every peer (st.async: the store completes 4 bytes on the peer's
                    // mbarrier of this pass parity) and waits until its own barrier has seen the peers' 128 bytes each.
                    // A peer can be at most one pass ahead (it needs this CTA's partial to finish its pass), so two
                    // slots / two barriers by pass parity never collide.
                    const unsigned xb = smem_u32(s_xbar) + 8 * pp;
                    if (lane == 0) mbar_expect_tx(xb, (C - 1) * 128);
                    const unsigned mine_addr = smem_u32(s_cl + slot + rank * 32 + lane);
                    for (unsigned r = 0; r < C; r++)
                        if (r != rank) st_async_f32(mapa_u32(mine_addr, r), s, mapa_u32(xb, r));
                    mbar_wait(xb, (xph >> pp) & 1u);
                    xph ^= (1u << pp);
                }
                __syncwarp();
                s = 0.f;
                for (unsigned r = 0; r < C; r++) s += s_cl[slot + r * 32 + lane];      // rank order: same bits in every CTA
                s_S[lane] = s;
                if (out32 && pass == 0 && rank == 0) out32[(size_t)h * 32 + lane] = s;
                __syncwarp();
                if (lane == 0)
                    finish_hyp(s_T, s_S, n, pass, crit, rank == 0 ? results + h : nullptr,
                               (rank == 0 && final_T) ? final_T + (size_t)h * 12 : nullptr);
            }
            __syncthreads();
            if (s_w[kStDone]) break;
        }
    }
}

// p <- T_h * p for every point of every hypothesis (PR_ICP_UPDATE_POINTS): what the reference's
// in-place transform_pcd_cuda calls add up to.  kApplySlices CTAs per hypothesis.
constexpr int kApplySlices = 8;
__global__ void __launch_bounds__(256)
icp_apply_kernel(float* __restrict__ pts, const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts,
                 unsigned n_hyp, const float* __restrict__ final_T) {
    const unsigned h = blockIdx.x / kApplySlices, sl = blockIdx.x % kApplySlices;
    if (h >= n_hyp) return;
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; i++) T[i] = final_T[(size_t)h * 12 + i];
    const unsigned n = counts[h];
    float* p0 = pts + 3 * (size_t)offsets[h];
    for (unsigned i = sl * 256 + threadIdx.x; i < n; i += kApplySlices * 256) {
        const float x = p0[3 * i], y = p0[3 * i + 1], z = p0[3 * i + 2];
        p0[3 * i + 0] = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
        p0[3 * i + 1] = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
        p0[3 * i + 2] = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
    }
}
// per-pass driver: the accumulated transforms live in HypState
__global__ void __launch_bounds__(256)
icp_state_to_final_kernel(const HypState* __restrict__ state, unsigned n_hyp, float* __restrict__ final_T) {
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_hyp * 12) final_T[i] = state[i / 12].T[i % 12];
}

// ---- parity / debug kernels: the hot loop's correspondence search, point by point ---------------------------
// out_idx[i] = scene index (pixel, or leaf-ordered point) the SHIPPED selection code picks for point i, -1 = none.
// The projective variant runs project_pair / pixel_of -- the functions group_projective runs -- on pairs (i, i+1).
__global__ void __launch_bounds__(256)
corr_projective_kernel(const float* __restrict__ pts, unsigned n, PackedScene s, int* __restrict__ out_idx) {
    const unsigned pair = blockIdx.x * 256 + threadIdx.x;
    const unsigned i0 = 2 * pair, i1 = 2 * pair + 1;
    if (i0 >= n) return;
    const bool has1 = i1 < n;
    const float x0 = pts[3 * i0], y0 = pts[3 * i0 + 1], z0 = pts[3 * i0 + 2];
    const float x1 = has1 ? pts[3 * i1] : 0.f, y1 = has1 ? pts[3 * i1 + 1] : 0.f, z1 = has1 ? pts[3 * i1 + 2] : 1.f;
    f2_t uf, vf;
    project_pair(s, pk2(x0, x1), pk2(y0, y1), pk2(z0, z1), uf, vf);
    float u0, u1, v0, v1;
    unpk2(uf, u0, u1); unpk2(vf, v0, v1);
    const float zz[2] = {z0, z1};
    const float uu[2] = {u0, u1}, vv[2] = {v0, v1};
    for (int k = 0; k < (has1 ? 2 : 1); k++) {
        int idx;
        bool ok = pixel_of(s, uu[k], vv[k], idx);
        if (!pz_in_exact_range(zz[k])) {        // what the kernel does for such a point: the per-point path
            const float px = pts[3 * (i0 + k)], py = pts[3 * (i0 + k) + 1];
            const float ufs = addf(addf(mulf(divf(px, zz[k]), s.fx), s.cx), 0.5f), vfs = addf(addf(mulf(divf(py, zz[k]), s.fy), s.cy), 0.5f);
            ok = (ufs > -1.0f && ufs < (float)s.W && vfs > -1.0f && vfs < (float)s.H);
            idx = ok ? (int)vfs * s.W + (int)ufs : 0;
        }
        if (ok) {
            float4 A; float2 B;
            load_rec(s, idx, A, B);
            ok = A.z > 0.f && fabsf(zz[k] - A.z) <= s.max_dist;
        }
        out_idx[i0 + k] = ok ? idx : -1;
    }
}
__global__ void __launch_bounds__(256)
corr_nn_kernel(const float* __restrict__ pts, unsigned n, PackedNnScene s, int* __restrict__ out_idx) {
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    const bool packed_ok = !(s.unsupported && *s.unsupported);
    int b = -2;
    unsigned tests = 0;
    bool done = false;
    if (packed_ok && s.grid.params) {            // the grid first, exactly as compute_share_nn decides
        const NnGridParams gp = *s.grid.params;
        done = gp.enabled && nn_grid_query<false>(s, gp, px, py, pz, b, tests);
    }
    if (!done) {
        b = packed_ok ? nn_search_packed(s, px, py, pz) : -2;
        if (b == -2) b = nn_search_reference(s.ref, px, py, pz);
    }
    out_idx[i] = b < 0 ? -1 : b;
}

// node fetches (box tests) and leaf-point distance tests of the packed walk, summed over the queries
__global__ void __launch_bounds__(256)
walk_stats_kernel(const float* __restrict__ pts, unsigned n, PackedNnScene s, unsigned long long* __restrict__ stats) {
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    unsigned v = 0, t = 0, r = 0, gt = 0;
    if (i < n) {
        nn_search_packed_t<true>(s, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], v, t);
        if (s.grid.params) {
            const NnGridParams gp = *s.grid.params;
            int b;
            r = (gp.enabled && nn_grid_query<true>(s, gp, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], b, gt)) ? 1u : 0u;
        }
    }
    v = __reduce_add_sync(0xffffffffu, v); t = __reduce_add_sync(0xffffffffu, t);
    r = __reduce_add_sync(0xffffffffu, r); gt = __reduce_add_sync(0xffffffffu, gt);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats, (unsigned long long)v); atomicAdd(stats + 1, (unsigned long long)t);
        atomicAdd(stats + 2, (unsigned long long)r); atomicAdd(stats + 3, (unsigned long long)gt);
    }
}

__global__ void __launch_bounds__(64)
solve_kernel(const float* __restrict__ S29, unsigned n, int fast, float* __restrict__ E16) {
    const unsigned i = blockIdx.x * 64 + threadIdx.x;
    if (i >= n) return;
    float S[29], E[16];
#pragma unroll
    for (int k = 0; k < 29; k++) S[k] = S29[(size_t)i * 29 + k];
    if (fast) solve_666_fast(S, E); else solve_666_unrolled(S, E);
#pragma unroll
    for (int k = 0; k < 16; k++) E16[(size_t)i * 16 + k] = E[k];
}

inline size_t icp_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct IcpWs {
    HypState* state; uint32_t* chunk_hyp; HypCtl* ctl; uint32_t* total_chunks; float* partials; float* final_T; float4* packed;
    int* nn_cache;       // capacity_points ints (kd-tree scenes, optional)
    size_t max_chunks, bytes;
};

// chunk size of the per-pass driver: big chunks amortise the per-CTA reduction; small batches need
// more CTAs than SMs
inline uint32_t pick_chunk_points(size_t n_hyp, size_t capacity_points, int sms) {
    uint32_t chunk = 4096;
    while (chunk > 512 && capacity_points / chunk + n_hyp < (size_t)sms * 4) chunk >>= 1;
    return chunk;
}

inline IcpWs carve_icp_ws(void* base, size_t n_hyp, size_t capacity_points, size_t scene_pixels, bool with_nn_cache = false) {
    IcpWs ws;
    ws.max_chunks = capacity_points / 512 + n_hyp + 1;      // sized for the smallest chunk the per-pass driver uses
    char* w = (char*)base;
    size_t used = 0;
    auto take = [&](size_t bytes) { char* p = w + used; used += icp_align_up(bytes, 256); return p; };
    ws.state = (HypState*)take(n_hyp * sizeof(HypState));
    ws.chunk_hyp = (uint32_t*)take(ws.max_chunks * 4);
    ws.ctl = (HypCtl*)take(sizeof(HypCtl));
    ws.total_chunks = (uint32_t*)take(256);
    ws.partials = (float*)take(ws.max_chunks * kPartialStride * 4);
    ws.final_T = (float*)take(n_hyp * 12 * 4);
    ws.packed = (float4*)take(scene_pixels * 32);
    ws.nn_cache = with_nn_cache ? (int*)take(capacity_points * 4 + 256) : nullptr;
    ws.bytes = used;
    return ws;
}

// ---- launch configuration of icp_hyp_kernel, per device ----------------------------------------------------
struct DeviceInfo { int sms = 0; size_t smem_optin = 0; };
inline int device_info(DeviceInfo& out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    int dev = 0;
    PR_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64) return PR_ERR_INVALID_ARGUMENT;
    if (cache[dev].sms == 0) {
        int sms = 0, optin = 0;
        PR_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PR_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cache[dev].sms = sms; cache[dev].smem_optin = (size_t)optin;
    }
    out = cache[dev];
    return PR_OK;
}

// Cluster size of a launch.  A hypothesis is split over C * kWarps warps; the per-pass cost that does not shrink with
// C is the reduce -> cluster barrier -> solve chain (~1.5 us), so C is kept as small as the batch allows: enough
// hypotheses per cluster slot that the claim counter balances the SMs, larger clusters only for small batches,
// where otherwise most SMs would have nothing to do.
#ifndef PR_HYP_CLUSTER_MIN
#define PR_HYP_CLUSTER_MIN 2
#endif
constexpr size_t kL2BudgetBytes = 80u << 20;    // what the clouds of the hypotheses in flight may occupy of the 126 MB L2
inline int pick_cluster(size_t n_hyp, int sms, int min_blocks, size_t pts_per_hyp) {
#ifdef PR_DEBUG       // experiment builds only (scripts/build_variants.py)
    if (const char* e = getenv("PR_HYP_CLUSTER")) { const int c = atoi(e); if (c == 1 || c == 2 || c == 4 || c == 8) return c; }
#endif
    if (PR_HYP_CLUSTER > 0) return PR_HYP_CLUSTER;
    // (1) the smallest cluster that still fills the machine: n_hyp * C >= resident CTAs (a small batch spreads each
    //     hypothesis over more SMs), never below 2 -- measured on 512 hypotheses x 22k points (592 resident CTAs):
    //     C = 1 1.65 ms, C = 2 1.42 ms, C = 4 1.54 ms, C = 8 1.93 ms; 64 hypotheses: C = 2 0.50 ms, 4 0.36 ms, 8 0.29 ms.
    // (2) the clouds of the hypotheses in flight (resident CTAs / C of them) must stay in L2, or every pass streams
    //     from HBM again: 512 hypotheses x 88k points (1280x720): C = 2 5.86 ms, C = 4 5.77 ms, C = 8 5.30 ms.
    const size_t slots = (size_t)sms * (size_t)min_blocks;
    int c = PR_HYP_CLUSTER_MIN;
    while (c < kMaxCluster && n_hyp * (size_t)c < slots) c <<= 1;
    while (c < kMaxCluster && (slots / (size_t)c) * pts_per_hyp * 12 > kL2BudgetBytes) c <<= 1;
    return c;
}

// kd-tree scenes: the same rule (measured on C3: C = 2 112 ms, 4 113 ms, 8 116 ms)
inline int pick_cluster_nn(size_t n_hyp, int sms, int min_blocks) {
#ifdef PR_DEBUG
    if (const char* e = getenv("PR_NN_CLUSTER")) { const int c = atoi(e); if (c == 1 || c == 2 || c == 4 || c == 8) return c; }
#endif
    return pick_cluster(n_hyp, sms, min_blocks, 0);
}

template <class SceneT>
int launch_hyp(const float* pts_dev, size_t capacity_points, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
               const IcpWs& ws, const SceneT& scene, pr_icp_criteria crit, pr_registration_result* results_dev, float* out32,
               size_t pts_per_hyp, cudaStream_t stream) {
    using Tr = HypTraits<SceneT>;
    DeviceInfo di;
    int rc = device_info(di);
    if (rc != PR_OK) return rc;
    const size_t smem = hyp_smem<Tr::kWarps>(Tr::kExtraSmem);
    auto kernel = icp_hyp_kernel<SceneT>;
    PR_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int C = Tr::kProjective ? pick_cluster(n_hyp, di.sms, Tr::kMinBlocks, pts_per_hyp) : pick_cluster_nn(n_hyp, di.sms, Tr::kMinBlocks);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(Tr::kWarps * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(di.sms * Tr::kMinBlocks / C * C), 1, 1);
    int max_clusters = 0;
    {   // resident clusters of this instantiation, per (device, cluster size); the query costs tens of microseconds
        static std::mutex mu;
        static int cache[64][kMaxCluster + 1];
        int dev = 0;
        PR_CUDA_TRY(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 0 || dev >= 64) return PR_ERR_INVALID_ARGUMENT;
        if (cache[dev][C] == 0) {
            int q = 0;
            PR_CUDA_TRY(cudaOccupancyMaxActiveClusters(&q, kernel, &cfg));
            cache[dev][C] = q > 0 ? q : -1;
        }
        max_clusters = cache[dev][C];
    }
    if (max_clusters < 1) return PR_ERR_UNSUPPORTED;
    const size_t clusters = std::min<size_t>((size_t)max_clusters, n_hyp);
    cfg.gridDim = dim3((unsigned)(clusters * C), 1, 1);
    const unsigned n_hyp_u = (unsigned)n_hyp;
    HypCtl* ctl = ws.ctl;
    const uint32_t* order = ws.chunk_hyp;        // n_hyp words (the table of the per-pass driver, unused by this one)
    float* final_T = ws.final_T;
    icp_order_kernel<<<(n_hyp_u + 255) / 256, 256, 0, stream>>>(counts_dev, n_hyp_u, ws.chunk_hyp, ctl);
    PR_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, pts_dev, capacity_points, offsets_dev, counts_dev, n_hyp_u, ctl, order, final_T, scene, crit,
                                   results_dev, out32));
    count_launch(2);
    return PR_OK;
}

// SceneT: the scene as the per-pass (reference arithmetic) driver consumes it; PScene: as the shipped driver consumes it
template <class SceneT, class PScene>
int run_icp(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp, size_t capacity_points,
            const SceneT& scene, const PScene& pscene, pr_icp_criteria crit, pr_registration_result* results_dev, int flags,
            const IcpWs& ws, cudaStream_t stream, size_t pts_per_hyp = 0) {
    if (pts_per_hyp == 0) pts_per_hyp = n_hyp ? capacity_points / n_hyp : 0;      // the caller's buffer bounds the clouds
    if (flags & PR_ICP_REFERENCE_ARITHMETIC) {
        DeviceInfo di;
        int rc = device_info(di);
        if (rc != PR_OK) return rc;
        const uint32_t chunk = pick_chunk_points(n_hyp, capacity_points, di.sms);
        // persistent grid: 3 CTAs per SM (register-limited), never more CTAs than chunks can exist
        const unsigned grid = (unsigned)std::min<size_t>(capacity_points / chunk + n_hyp + 1, (size_t)di.sms * 3);
        icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts_dev, (uint32_t)n_hyp, chunk, ws.state, ws.chunk_hyp,
                                                       (uint32_t)ws.max_chunks, ws.total_chunks, results_dev);
        for (int it = 0; it <= crit.max_iteration; it++)
            icp_pass_kernel<SceneT><<<grid, kIcpThreads, 0, stream>>>(pts_dev, offsets_dev, counts_dev, ws.chunk_hyp, ws.total_chunks,
                                                                      chunk, ws.state, ws.partials, scene, crit, results_dev, nullptr);
        count_launch(2 + (uint64_t)crit.max_iteration);
        if (flags & PR_ICP_UPDATE_POINTS) {
            icp_state_to_final_kernel<<<(unsigned)((n_hyp * 12 + 255) / 256), 256, 0, stream>>>(ws.state, (unsigned)n_hyp, ws.final_T);
            icp_apply_kernel<<<(unsigned)(n_hyp * kApplySlices), 256, 0, stream>>>(pts_dev, offsets_dev, counts_dev, (unsigned)n_hyp, ws.final_T);
            count_launch(2);
        }
        PR_LAUNCH_CHECK();
        return PR_OK;
    }
    int rc = launch_hyp(pts_dev, capacity_points, offsets_dev, counts_dev, n_hyp, ws, pscene, crit, results_dev, nullptr, pts_per_hyp, stream);
    if (rc != PR_OK) return rc;
    if (flags & PR_ICP_UPDATE_POINTS) {
        icp_apply_kernel<<<(unsigned)(n_hyp * kApplySlices), 256, 0, stream>>>(pts_dev, offsets_dev, counts_dev, (unsigned)n_hyp, ws.final_T);
        count_launch();
    }
    PR_LAUNCH_CHECK();
    return PR_OK;
}

inline int check_icp_args(const float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                          size_t capacity_points, pr_icp_criteria crit, pr_registration_result* results_dev, void* workspace_dev) {
    if (!pts_dev || !offsets_dev || !counts_dev || !results_dev || !workspace_dev) return PR_ERR_INVALID_ARGUMENT;
    if (crit.max_iteration < 0 || crit.max_iteration > (1 << 20) || n_hyp > 0x7FFFFFFFull / 64 || capacity_points > 0xFFFFFFFFull)
        return PR_ERR_INVALID_ARGUMENT;
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
inline PackedScene make_packed_scene(const ProjScene& s, const float4* rec) {
    PackedScene ps;
    ps.W = s.W; ps.H = s.H; ps.max_dist = s.max_dist;
    ps.fx = s.fx; ps.fy = s.fy; ps.cx = s.cx; ps.cy = s.cy; ps.one = 1.0f; ps.rec = rec;
    return ps;
}
inline int make_nn_scene(const pr_scene_nn* s, NnScene& o) {
    if (!s || (s->n_nodes && (!s->pcd_dev || !s->normal_dev || !s->nodes_dev)) || s->n_nodes > 0x7FFFFFFFull)
        return PR_ERR_INVALID_ARGUMENT;
    o.max_dist_sq = s->max_dist_diff * s->max_dist_diff;   // pow2(max_dist_diff), pcd_scene.h:127
    o.pcd = s->pcd_dev; o.nrm = s->normal_dev; o.nodes = s->nodes_dev; o.n_nodes = (int)s->n_nodes;
    return PR_OK;
}

// the packed kd-tree lives behind the generic workspace: 16 B per scene point + 32 B per node (+ a flag word)
struct PackedTree { float4* pts4; float4* nodes; unsigned* flag; };
inline PackedTree carve_packed_tree(float4* base, size_t n_points, size_t n_nodes) {
    PackedTree t;
    t.pts4 = base; t.nodes = base + n_points; t.flag = reinterpret_cast<unsigned*>(t.nodes + 2 * n_nodes);
    return t;
}
inline int pack_tree(const NnScene& s, size_t n_points, const PackedTree& t, cudaStream_t stream) {
    PR_CUDA_TRY(cudaMemsetAsync(t.flag, 0, 4, stream));
    if (n_points) nn_pack_points_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, stream>>>(s.pcd, n_points, t.pts4);
    nn_pack_nodes_kernel<<<(unsigned)((s.n_nodes + 255) / 256), 256, 0, stream>>>(s.nodes, s.n_nodes, s.pcd, t.nodes, t.flag);
    count_launch(2);
    return PR_OK;
}

// the hash grid (icp_device.cuh) lives behind the packed tree, when the workspace has room for it
struct GridWs {
    NnGridParams* params; uint2* table; unsigned* keys; unsigned* counts; unsigned* cursor; unsigned* pt_slot; float4* gpts;
    unsigned log2_slots;
    size_t bytes;
};
inline GridWs carve_grid(void* base, size_t n_points) {
    GridWs g;
    // 8 n block memberships in about n / 3 distinct blocks for a sampled surface (2 n slots: a sixth full); a scene of
    // isolated points (up to 8 n blocks) overfills the table, which the count kernel notices and answers by switching the
    // grid off (the tree then answers everything)
    unsigned lg = 10;
    while (lg < 28 && ((size_t)1 << lg) < PR_NN_TABLE * n_points) lg++;
    g.log2_slots = lg;
    const size_t slots = (size_t)1 << lg;
    char* w = (char*)base;
    size_t used = 0;
    auto take = [&](size_t bytes) { char* p = w + used; used += icp_align_up(bytes, 256); return p; };
    g.params = (NnGridParams*)take(sizeof(NnGridParams));
    g.table = (uint2*)take(slots * 8);
    g.keys = (unsigned*)take(slots * 4);
    g.counts = (unsigned*)take(slots * 4);
    g.cursor = (unsigned*)take(slots * 4);
    g.pt_slot = (unsigned*)take(8 * n_points * 4);
    g.gpts = (float4*)take((8 * n_points + 8) * 16);      // + the elements the last list's vector loads may touch
    g.bytes = used;
    return g;
}
inline int build_grid(const PackedTree& t, size_t n_points, int n_nodes, const GridWs& g, cudaStream_t stream) {
    const size_t slots = (size_t)1 << g.log2_slots;
    PR_CUDA_TRY(cudaMemsetAsync(g.keys, 0xFF, slots * 4, stream));
    PR_CUDA_TRY(cudaMemsetAsync(g.counts, 0, slots * 4, stream));
    nn_grid_params_kernel<<<1, 256, 0, stream>>>(t.nodes, n_nodes, (unsigned)n_points, g.log2_slots, t.flag, g.params);
    const unsigned blocks = (unsigned)((8 * n_points + 255) / 256);
    nn_grid_count_kernel<<<blocks, 256, 0, stream>>>(t.pts4, (unsigned)n_points, g.params, g.keys, g.counts, g.pt_slot);
    nn_grid_scan_kernel<<<1, 1024, 0, stream>>>(g.keys, g.counts, (unsigned)slots, g.cursor, g.table);
    nn_grid_fill_kernel<<<blocks, 256, 0, stream>>>(t.pts4, (unsigned)n_points, g.params, g.pt_slot, g.cursor, g.gpts);
    count_launch(4);
    return PR_OK;
}
// packed scene of a kd-tree scene inside `region` (region_bytes): packed tree always, the grid when it fits
inline void init_packed_nn(const NnScene& s, PackedNnScene& ps) {        // nothing packed: the kernel walks the reference layout
    ps.max_dist_sq = s.max_dist_sq; ps.nrm = s.nrm; ps.n_nodes = s.n_nodes; ps.ref = s;
    ps.nodes = nullptr; ps.pts4 = nullptr; ps.unsupported = nullptr; ps.top = nullptr; ps.n_top = 0; ps.cache = nullptr;
    ps.grid.params = nullptr; ps.grid.table = nullptr; ps.grid.gpts = nullptr;
}
inline int make_packed_nn(const NnScene& s, const pr_scene_nn* scene, float4* region, size_t region_bytes, PackedNnScene& ps, cudaStream_t stream) {
    init_packed_nn(s, ps);
    if (s.n_nodes <= 0) return PR_OK;
    const size_t tree_bytes = icp_align_up((scene->n_points + 2 * scene->n_nodes) * 16 + 4, 256);
    if (region_bytes < tree_bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    const PackedTree t = carve_packed_tree(region, scene->n_points, scene->n_nodes);
    int rc = pack_tree(s, scene->n_points, t, stream);
    if (rc != PR_OK) return rc;
    ps.nodes = t.nodes; ps.pts4 = t.pts4; ps.unsupported = t.flag;
    if (scene->n_points > 0 && scene->n_points < (1u << 21)) {      // 8 n list entries behind a 24-bit start
        const GridWs g = carve_grid((char*)region + tree_bytes, scene->n_points);
        if (tree_bytes + g.bytes <= region_bytes) {
            rc = build_grid(t, scene->n_points, s.n_nodes, g, stream);
            if (rc != PR_OK) return rc;
            ps.grid.params = g.params; ps.grid.table = g.table; ps.grid.gpts = g.gpts;
        }
    }
    return PR_OK;
}
// float4-pairs ("scene pixels" of carve_icp_ws, 32 bytes each) a kd-tree scene needs: packed tree + grid
inline size_t nn_scene_units(size_t n_points, size_t n_nodes) {
    const size_t tree_bytes = icp_align_up((n_points + 2 * n_nodes) * 16 + 4, 256);
    return (tree_bytes + carve_grid(nullptr, n_points).bytes + 31) / 32 + 16;
}
// workspace layout for a kd-tree scene, the richest that fits: grid + search cache, grid, cache, neither
// (pr_icp_nn_workspace_bytes has room for everything; a workspace sized by pr_icp_workspace_bytes still works)
inline bool carve_nn_ws(void* base, size_t bytes, size_t n_hyp, size_t capacity_points, const pr_scene_nn* scene, bool want_cache,
                        IcpWs& ws, size_t& region_bytes) {
    const size_t full = nn_scene_units(scene->n_points, scene->n_nodes), small = scene->n_points + 2 * scene->n_nodes + 16;
    const size_t units[4] = {full, full, small, small};
    const bool cache[4] = {true, false, true, false};
    for (int k = 0; k < 4; k++) {
        if (cache[k] && !want_cache) continue;
        ws = carve_icp_ws(base, n_hyp, capacity_points, units[k], cache[k]);
        if (ws.bytes <= bytes) { region_bytes = units[k] * 32; return true; }
    }
    return false;
}

// single cloud, identity transform, one reduction pass -> out29 (parity / debug entry point).  FAST: through the
// shipped driver (criteria (0,0,0): one evaluation pass; the sums of pass 0 are copied out), else the reference-arithmetic kernel.
template <class SceneT>
int run_pcd2ab(const float* pts_dev, size_t n, const SceneT& scene, float* out29_dev, cudaStream_t stream) {
    if (!pts_dev || !out29_dev || n > 0xFFFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    DeviceInfo di;
    int rc = device_info(di);
    if (rc != PR_OK) return rc;
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
    const uint32_t chunk = pick_chunk_points(1, n, di.sms);
    const unsigned grid = (unsigned)std::min<size_t>(n / chunk + 2, (size_t)di.sms * 3);
    pr_icp_criteria crit = {0.f, 0.f, 0};
    cudaMemsetAsync(out29_dev, 0, 29 * 4, stream);
    icp_plan_kernel<<<1, kIcpThreads, 0, stream>>>(counts, 1, chunk, ws.state, ws.chunk_hyp, (uint32_t)ws.max_chunks,
                                                   ws.total_chunks, res);
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

size_t pr_icp_workspace_bytes(size_t n_hyp, size_t capacity_points, size_t scene_pixels) {
    return carve_icp_ws(nullptr, n_hyp, capacity_points, scene_pixels).bytes;
}

size_t pr_icp_nn_workspace_bytes(size_t n_hyp, size_t capacity_points, size_t n_scene_points, size_t n_nodes) {
    return carve_icp_ws(nullptr, n_hyp, capacity_points, nn_scene_units(n_scene_points, n_nodes), true).bytes;
}

size_t pr_scene_projective_packed_bytes(uint32_t width, uint32_t height) { return (size_t)width * height * 32; }

int pr_scene_projective_pack(const pr_scene_projective* scene, void* packed_dev, pr_stream_t stream) {
    ProjScene s;
    int rc = make_proj_scene(scene, s);
    if (rc != PR_OK) return rc;
    if (!packed_dev || ((uintptr_t)packed_dev & 31)) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)s.W * s.H;
    scene_pack_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, as_stream(stream)>>>(s.pcd, s.nrm, n_px, (float4*)packed_dev);
    count_launch();
    PR_LAUNCH_CHECK();
    return PR_OK;
}

}  // extern "C"

// pts_per_hyp: the caller's estimate of the average cloud size (0: capacity_points / n_hyp); only steers the cluster size
int prb::icp_projective_packed(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                               size_t capacity_points, const pr_scene_projective* scene, const void* packed_dev,
                               pr_icp_criteria criteria, pr_registration_result* results_dev, int flags,
                               void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_, size_t pts_per_hyp) {
    ProjScene s;
    int rc = make_proj_scene(scene, s);
    if (rc != PR_OK) return rc;
    rc = check_icp_args(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, criteria, results_dev, workspace_dev);
    if (rc != PR_OK) return rc;
    if (n_hyp == 0) return PR_OK;
    const size_t n_px = (size_t)s.W * s.H;
    const bool own_pack = packed_dev == nullptr;
    IcpWs ws = carve_icp_ws(workspace_dev, n_hyp, capacity_points, own_pack ? n_px : 0);
    if (workspace_bytes < ws.bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = as_stream(stream_);
    const bool hot = !(flags & PR_ICP_REFERENCE_ARITHMETIC);
    if (hot && own_pack) {
        scene_pack_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, stream>>>(s.pcd, s.nrm, n_px, ws.packed);
        count_launch();
    }
    const PackedScene ps = make_packed_scene(s, own_pack ? ws.packed : (const float4*)packed_dev);
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, ps, criteria, results_dev, flags, ws, stream, pts_per_hyp);
}

extern "C" {

int pr_icp_projective_batch_packed(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                                   size_t capacity_points, const pr_scene_projective* scene, const void* packed_dev,
                                   pr_icp_criteria criteria, pr_registration_result* results_dev, int flags,
                                   void* workspace_dev, size_t workspace_bytes, pr_stream_t stream) {
    return prb::icp_projective_packed(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, scene, packed_dev, criteria, results_dev,
                                      flags, workspace_dev, workspace_bytes, stream, 0);
}

int pr_icp_projective_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                            size_t capacity_points, const pr_scene_projective* scene, pr_icp_criteria criteria,
                            pr_registration_result* results_dev, int flags,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream) {
    return pr_icp_projective_batch_packed(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, scene, nullptr, criteria,
                                          results_dev, flags, workspace_dev, workspace_bytes, stream);
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
    // Leaves of more than 127 points, more than 2^24 points or children that are not siblings do not fit the packed
    // encoding.  No tree KDTree_cpu::build_tree / pr_scene_nn_build* produce has them (leaf <= max_leaf, children appended
    // together); a foreign tree that does is detected by the packing kernel, which then marks the packed root as an empty
    // leaf-less tree and the kernel walks the reference layout instead -- decided on the device, no host round trip.
    // with room for one int per model point (pr_icp_nn_workspace_bytes) every pass starts its tree walks from the previous
    // pass' winners, and with room for the hash grid most points never walk the tree at all; a workspace sized by
    // pr_icp_workspace_bytes still works, without either
    IcpWs ws;
    size_t region_bytes = 0;
    if (!carve_nn_ws(workspace_dev, workspace_bytes, n_hyp, capacity_points, scene, true, ws, region_bytes)) return PR_ERR_WORKSPACE_TOO_SMALL;
    PackedNnScene ps;
    if (!(flags & PR_ICP_REFERENCE_ARITHMETIC)) {
        rc = make_packed_nn(s, scene, ws.packed, region_bytes, ps, stream);
        if (rc != PR_OK) return rc;
        if (ps.nodes) ps.cache = ws.nn_cache;
    } else {
        init_packed_nn(s, ps);
    }
    return run_icp(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, s, ps, criteria, results_dev, flags, ws, stream);
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

// ---- parity entry points onto the shipped kernel -------------------------------------------------------------
int pr_pass_sums_projective(const float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                            size_t capacity_points, const pr_scene_projective* scene, float* out32_dev,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    ProjScene s;
    int rc = make_proj_scene(scene, s);
    if (rc != PR_OK) return rc;
    pr_icp_criteria crit = {0.f, 0.f, 0};
    if (!out32_dev) return PR_ERR_INVALID_ARGUMENT;
    if (n_hyp == 0) return PR_OK;
    const size_t n_px = (size_t)s.W * s.H;
    IcpWs ws = carve_icp_ws(workspace_dev, n_hyp, capacity_points, n_px);
    // the results of the evaluation pass land in the (otherwise unused) per-pass state area
    rc = check_icp_args(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, crit, (pr_registration_result*)ws.state, workspace_dev);
    if (rc != PR_OK) return rc;
    if (workspace_bytes < ws.bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = as_stream(stream_);
    scene_pack_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, stream>>>(s.pcd, s.nrm, n_px, ws.packed);
    count_launch();
    static_assert(sizeof(HypState) >= sizeof(pr_registration_result), "results fit the state area");
    rc = launch_hyp(pts_dev, capacity_points, offsets_dev, counts_dev, n_hyp, ws, make_packed_scene(s, ws.packed), crit,
                    (pr_registration_result*)ws.state, out32_dev, 0, stream);
    if (rc != PR_OK) return rc;
    PR_LAUNCH_CHECK();
    return PR_OK;
}

int pr_pass_sums_nn(const float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                    size_t capacity_points, const pr_scene_nn* scene, float* out32_dev,
                    void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    NnScene s;
    int rc = make_nn_scene(scene, s);
    if (rc != PR_OK) return rc;
    pr_icp_criteria crit = {0.f, 0.f, 0};
    if (!out32_dev) return PR_ERR_INVALID_ARGUMENT;
    if (n_hyp == 0) return PR_OK;
    if (!workspace_dev) return PR_ERR_INVALID_ARGUMENT;
    IcpWs ws;
    size_t region_bytes = 0;
    if (!carve_nn_ws(workspace_dev, workspace_bytes, n_hyp, capacity_points, scene, false, ws, region_bytes)) return PR_ERR_WORKSPACE_TOO_SMALL;
    rc = check_icp_args(pts_dev, offsets_dev, counts_dev, n_hyp, capacity_points, crit, (pr_registration_result*)ws.state, workspace_dev);
    if (rc != PR_OK) return rc;
    cudaStream_t stream = as_stream(stream_);
    PackedNnScene ps;
    rc = make_packed_nn(s, scene, ws.packed, region_bytes, ps, stream);
    if (rc != PR_OK) return rc;
    rc = launch_hyp(pts_dev, capacity_points, offsets_dev, counts_dev, n_hyp, ws, ps, crit, (pr_registration_result*)ws.state,
                    out32_dev, 0, stream);
    if (rc != PR_OK) return rc;
    PR_LAUNCH_CHECK();
    return PR_OK;
}

int pr_correspondences_projective(const float* pts_dev, size_t n, const pr_scene_projective* scene, int32_t* idx_dev,
                                  void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    ProjScene s;
    int rc = make_proj_scene(scene, s);
    if (rc != PR_OK) return rc;
    if (!pts_dev || !idx_dev || !workspace_dev || n > 0x7FFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    if (n == 0) return PR_OK;
    const size_t n_px = (size_t)s.W * s.H;
    IcpWs ws = carve_icp_ws(workspace_dev, 1, n, n_px);
    if (workspace_bytes < ws.bytes) return PR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = as_stream(stream_);
    scene_pack_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, stream>>>(s.pcd, s.nrm, n_px, ws.packed);
    corr_projective_kernel<<<(unsigned)(((n + 1) / 2 + 255) / 256), 256, 0, stream>>>(pts_dev, (unsigned)n, make_packed_scene(s, ws.packed), idx_dev);
    count_launch(2);
    PR_LAUNCH_CHECK();
    return PR_OK;
}

int pr_correspondences_nn(const float* pts_dev, size_t n, const pr_scene_nn* scene, int32_t* idx_dev,
                          void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    NnScene s;
    int rc = make_nn_scene(scene, s);
    if (rc != PR_OK) return rc;
    if (!pts_dev || !idx_dev || !workspace_dev || n > 0x7FFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    if (n == 0) return PR_OK;
    cudaStream_t stream = as_stream(stream_);
    if (s.n_nodes == 0) { PR_CUDA_TRY(cudaMemsetAsync(idx_dev, 0xFF, n * 4, stream)); return PR_OK; }
    IcpWs ws;
    size_t region_bytes = 0;
    if (!carve_nn_ws(workspace_dev, workspace_bytes, 1, n, scene, false, ws, region_bytes)) return PR_ERR_WORKSPACE_TOO_SMALL;
    PackedNnScene ps;
    rc = make_packed_nn(s, scene, ws.packed, region_bytes, ps, stream);
    if (rc != PR_OK) return rc;
    corr_nn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pts_dev, (unsigned)n, ps, idx_dev);
    count_launch();
    PR_LAUNCH_CHECK();
    return PR_OK;
}

int pr_nn_walk_stats(const float* pts_dev, size_t n, const pr_scene_nn* scene, uint64_t* stats2_dev,
                     void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    NnScene s;
    int rc = make_nn_scene(scene, s);
    if (rc != PR_OK) return rc;
    if (!pts_dev || !stats2_dev || !workspace_dev || n > 0x7FFFFFFFull || s.n_nodes == 0) return PR_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = as_stream(stream_);
    PR_CUDA_TRY(cudaMemsetAsync(stats2_dev, 0, 32, stream));
    if (n == 0) return PR_OK;
    IcpWs ws;
    size_t region_bytes = 0;
    if (!carve_nn_ws(workspace_dev, workspace_bytes, 1, n, scene, false, ws, region_bytes)) return PR_ERR_WORKSPACE_TOO_SMALL;
    PackedNnScene ps;
    rc = make_packed_nn(s, scene, ws.packed, region_bytes, ps, stream);
    if (rc != PR_OK) return rc;
    walk_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pts_dev, (unsigned)n, ps, reinterpret_cast<unsigned long long*>(stats2_dev));
    count_launch();
    PR_LAUNCH_CHECK();
    return PR_OK;
}

int pr_solve_666_device(const float* S29_dev, size_t n, int fast, float* E16_dev, pr_stream_t stream_) {
    if (!S29_dev || !E16_dev || n > 0x7FFFFFFFull) return PR_ERR_INVALID_ARGUMENT;
    if (n == 0) return PR_OK;
    solve_kernel<<<(unsigned)((n + 63) / 64), 64, 0, as_stream(stream_)>>>(S29_dev, (unsigned)n, fast, E16_dev);
    count_launch();
    PR_LAUNCH_CHECK();
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
