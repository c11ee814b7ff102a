// common.cuh -- shared helpers for the sm_100a kernels of libpose_refine_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>

#include "../../include/pose_refine_b200.h"

namespace prb {

// every kernel launch of the library goes through this counter (pr_launch_count)
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define PR_CUDA_TRY(expr)                                  \
    do {                                                   \
        cudaError_t pr_e_ = (expr);                        \
        if (pr_e_ != cudaSuccess) return (int)pr_e_;       \
    } while (0)

#define PR_LAUNCH_CHECK()                                  \
    do {                                                   \
        cudaError_t pr_e_ = cudaPeekAtLastError();         \
        if (pr_e_ != cudaSuccess) return (int)pr_e_;       \
    } while (0)

inline cudaStream_t as_stream(pr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// SMs of the current device (grid sizing of the grid-stride kernels); 148 on a B200
inline int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

// IEEE single ops that the compiler may never contract into FMA: the rasteriser and the cloud /
// scene-preparation kernels use them so their results equal the reference's x86 CPU build
// (no FMA there: -O3 without -march) bit for bit.
__device__ __forceinline__ float mulf(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addf(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subf(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float divf(float a, float b) { return __fdiv_rn(a, b); }

// x86 cvttss2si semantics for float -> int32 (what the reference's CPU build gets where C++
// leaves the conversion undefined): NaN / out of range -> INT_MIN.
__device__ __forceinline__ int f2i_x86(float v) {
    return (v > -2147483904.0f && v < 2147483648.0f) ? __float2int_rz(v) : INT_MIN;
}

struct alignas(16) Mat34 {  // rows 0..2 of a row-major 4x4 rigid transform
    float m[12];
};

// cloud.cu: compacted clouds of a depth batch from per-tile valid-pixel counts (the rasteriser's tiles), tile-major
// point order; tile_valid: n_images * tiles_x * tiles_y counts (in); tile_off: scratch of
// cloud_tiles_scratch_words(n_images, tiles) words (offsets, non-empty tile lists, list lengths).
inline size_t cloud_tiles_scratch_words(size_t n_images, size_t n_tiles) { return 2 * n_images * n_tiles + n_images; }
int cloud_from_tiles(const int32_t* depth_dev, size_t n_images, uint32_t width, uint32_t height, const float K[9],
                     int tile_w, int tile_h, int tiles_x, int tiles_y, const unsigned* tile_valid, unsigned* tile_off,
                     uint32_t* counts_dev, uint32_t* offsets_dev, uint32_t* overflow_dev, size_t capacity_points,
                     uint32_t align_points, float* out_pts_dev, cudaStream_t stream);

// icp.cu: pr_icp_projective_batch_packed with the caller's estimate of the average cloud size (steers the cluster size)
int icp_projective_packed(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                          size_t capacity_points, const pr_scene_projective* scene, const void* packed_dev,
                          pr_icp_criteria criteria, pr_registration_result* results_dev, int flags,
                          void* workspace_dev, size_t workspace_bytes, pr_stream_t stream, size_t pts_per_hyp);

}  // namespace prb
