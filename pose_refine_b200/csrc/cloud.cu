// cloud.cu -- batched depth image -> compacted point cloud for sm_100a.
//
// Replaces depth2cloud_cuda<T> with its depth2mask / exclusive_scan / depth2cloud kernels
// (cuda_icp/icp.cu:228-291; CPU twin icp.cpp:73-117) for a whole batch of images in four
// launches and no host round trip (upstream: 3 kernels + a thrust scan + 2 D2H reads + 2
// cudaMallocs per image).  Output order = ascending pixel index of the valid (depth > 0)
// pixels; arithmetic as SURVEY.md App. A-2 with non-contractable IEEE ops, so the points equal
// depth2cloud_cpu bit for bit.
//
// Layout: an image is cut into segments of kSegPx pixels, one CTA each.
//   count : valid pixels per segment            -> seg[n_images][n_seg]
//   scan  : per image, exclusive scan over its segments (in place) + image total -> counts
//   offs  : exclusive scan over images of the padded totals -> offsets[n_images+1]
//   fill  : per segment, intra-CTA scan + scatter of back-projected points
// Fused with the rasteriser (pr_render_cloud_batch -> cloud_from_tiles): the tile write-out has already counted the
// valid pixels, so tile_scan_kernel + cloud_fill_tiles_kernel only visit the non-empty screen tiles (tile-major order).
#include "common.cuh"
#include <limits.h>

namespace prb {

constexpr int kSegThreads = 256;
constexpr int kPxPerThread = 8;
constexpr int kSegPx = kSegThreads * kPxPerThread;

template <class T> struct Px8 { T v[kPxPerThread]; };

// loads 8 consecutive pixels starting at idx (idx % 8 == 0); pixels past n_px read as 0
template <class T>
__device__ __forceinline__ Px8<T> load8(const T* __restrict__ img, uint32_t idx, uint32_t n_px, bool vec) {
    Px8<T> r;
    if (vec && idx + kPxPerThread <= n_px) {
        if (sizeof(T) == 4) {
            const int4 a = __ldg(reinterpret_cast<const int4*>(img + idx));
            const int4 b = __ldg(reinterpret_cast<const int4*>(img + idx) + 1);
            r.v[0] = (T)a.x; r.v[1] = (T)a.y; r.v[2] = (T)a.z; r.v[3] = (T)a.w;
            r.v[4] = (T)b.x; r.v[5] = (T)b.y; r.v[6] = (T)b.z; r.v[7] = (T)b.w;
        } else {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(img + idx));
            r.v[0] = (T)(a.x & 0xFFFF); r.v[1] = (T)(a.x >> 16); r.v[2] = (T)(a.y & 0xFFFF); r.v[3] = (T)(a.y >> 16);
            r.v[4] = (T)(a.z & 0xFFFF); r.v[5] = (T)(a.z >> 16); r.v[6] = (T)(a.w & 0xFFFF); r.v[7] = (T)(a.w >> 16);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kPxPerThread; k++) r.v[k] = (idx + k < n_px) ? img[idx + k] : (T)0;
    }
    return r;
}

template <class T>
__global__ void __launch_bounds__(kSegThreads)
cloud_count_kernel(const T* __restrict__ depth, uint32_t n_px, uint32_t n_seg, uint32_t* __restrict__ seg, int vec) {
    __shared__ unsigned s_warp[kSegThreads / 32];
    const uint32_t image = blockIdx.y, sid = blockIdx.x;
    const T* img = depth + (size_t)image * n_px;
    const uint32_t idx = sid * kSegPx + threadIdx.x * kPxPerThread;
    unsigned c = 0;
    if (idx < n_px) {
        const Px8<T> p = load8(img, idx, n_px, vec != 0);
#pragma unroll
        for (int k = 0; k < kPxPerThread; k++) c += (p.v[k] > 0) ? 1u : 0u;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < kSegThreads / 32; w++) t += s_warp[w];
        seg[(size_t)image * n_seg + sid] = t;
    }
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix,
// *total receives the block sum.  s_warp: 8 words of shared memory.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* s_warp, unsigned* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();   // protect s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned wprefix = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kSegThreads / 32; w++) { const unsigned s = s_warp[w]; if (w < warp) wprefix += s; all += s; }
    *total = all;
    return wprefix + incl - v;
}

// one CTA per image: seg counts -> exclusive offsets (in place), counts[image] = total
__global__ void __launch_bounds__(kSegThreads)
cloud_scan_kernel(const uint32_t* seg_in, uint32_t* seg_out, uint32_t n_seg, uint32_t* __restrict__ counts) {
    __shared__ unsigned s_warp[kSegThreads / 32];
    const uint32_t* s = seg_in + (size_t)blockIdx.x * n_seg;
    uint32_t* o = seg_out + (size_t)blockIdx.x * n_seg;      // may alias seg_in
    unsigned carry = 0;
    for (uint32_t b = 0; b < n_seg; b += kSegThreads) {
        const uint32_t i = b + threadIdx.x;
        const unsigned v = (i < n_seg) ? s[i] : 0u;
        unsigned total;
        const unsigned excl = block_excl_scan(v, s_warp, &total);
        if (i < n_seg) o[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) counts[blockIdx.x] = carry;
}

// single CTA: offsets[i] = sum_{j<i} align_up(counts[j], align); offsets[n] = padded total.
// A cloud that would end beyond `capacity` points is emptied (counts[i] = 0) and *overflow is set.
__global__ void __launch_bounds__(kSegThreads)
cloud_offsets_kernel(uint32_t* __restrict__ counts, uint32_t n, uint32_t align, unsigned long long capacity,
                     uint32_t* __restrict__ offsets, uint32_t* __restrict__ overflow) {
    __shared__ unsigned s_warp[kSegThreads / 32];
    // offsets are 32-bit: a batch whose padded clouds add up to more than 2^32 - 1 points cannot be addressed, whatever the
    // capacity says ("unlimited" included) -- the clouds past that point are dropped and flagged like any other overflow
    if (capacity > 0xFFFFFFFFull) capacity = 0xFFFFFFFFull;
    unsigned long long carry = 0;
    bool spilled = false;
    for (uint32_t b = 0; b < n; b += kSegThreads) {
        const uint32_t i = b + threadIdx.x;
        unsigned v = (i < n) ? counts[i] : 0u;
        v = (v + align - 1) / align * align;           // the entry points reject batches whose 256-image block totals could wrap
        unsigned total;
        const unsigned excl = block_excl_scan(v, s_warp, &total);
        if (i < n) {
            const unsigned long long at = carry + excl;
            offsets[i] = (uint32_t)(at > 0xFFFFFFFFull ? 0xFFFFFFFFull : at);
            if (at + v > capacity) { counts[i] = 0; spilled = true; }
        }
        carry += total;
    }
    if (threadIdx.x == 0) offsets[n] = (uint32_t)(carry > 0xFFFFFFFFull ? 0xFFFFFFFFull : carry);
    if (overflow) {
        if (threadIdx.x == 0) *overflow = 0;
        __syncthreads();
        if (spilled) *overflow = 1;
    }
}

struct Intrinsics { float fx, fy, cx, cy; };

template <class T>
__global__ void __launch_bounds__(kSegThreads)
cloud_fill_kernel(const T* __restrict__ depth, uint32_t width, uint32_t n_px, uint32_t n_seg,
                  const uint32_t* __restrict__ seg, const uint32_t* __restrict__ offsets, Intrinsics K,
                  uint32_t tl_x, uint32_t tl_y, float* __restrict__ out, size_t capacity, int vec) {
    __shared__ unsigned s_warp[kSegThreads / 32];
    const uint32_t image = blockIdx.y, sid = blockIdx.x;
    const T* img = depth + (size_t)image * n_px;
    const uint32_t idx = sid * kSegPx + threadIdx.x * kPxPerThread;
    Px8<T> p;
    unsigned c = 0;
    if (idx < n_px) {
        p = load8(img, idx, n_px, vec != 0);
#pragma unroll
        for (int k = 0; k < kPxPerThread; k++) c += (p.v[k] > 0) ? 1u : 0u;
    }
    unsigned total;
    const unsigned excl = block_excl_scan(c, s_warp, &total);
    if (c == 0) return;
    size_t dst = (size_t)offsets[image] + seg[(size_t)image * n_seg + sid] + excl;
    uint32_t v = idx / width, u = idx - v * width;
#pragma unroll
    for (int k = 0; k < kPxPerThread; k++) {
        if (p.v[k] > 0) {
            if (dst < capacity) {
                // icp.cu:249-251
                const float z = divf((float)p.v[k], 1000.0f);
                const float x = mulf(divf(subf((float)(u + tl_x), K.cx), K.fx), z);
                const float y = mulf(divf(subf((float)(v + tl_y), K.cy), K.fy), z);
                out[3 * dst + 0] = x; out[3 * dst + 1] = y; out[3 * dst + 2] = z;
            }
            dst++;
        }
        if (++u == width) { u = 0; v++; }
    }
}

// one CTA per image: tile counts -> exclusive offsets, image total, and the compact list of its non-empty tiles
__global__ void __launch_bounds__(kSegThreads)
tile_scan_kernel(const uint32_t* __restrict__ tile_valid, uint32_t* __restrict__ tile_off, uint32_t n_tiles,
                 uint32_t* __restrict__ counts, uint32_t* __restrict__ tile_list, uint32_t* __restrict__ n_list) {
    __shared__ unsigned s_warp[kSegThreads / 32];
    const uint32_t* s = tile_valid + (size_t)blockIdx.x * n_tiles;
    uint32_t* o = tile_off + (size_t)blockIdx.x * n_tiles;
    uint32_t* l = tile_list + (size_t)blockIdx.x * n_tiles;
    unsigned carry = 0, lcarry = 0;
    for (uint32_t b = 0; b < n_tiles; b += kSegThreads) {
        const uint32_t i = b + threadIdx.x;
        const unsigned v = (i < n_tiles) ? s[i] : 0u;
        unsigned total, ltotal;
        const unsigned excl = block_excl_scan(v, s_warp, &total);
        const unsigned lexcl = block_excl_scan(v ? 1u : 0u, s_warp, &ltotal);
        if (i < n_tiles) o[i] = carry + excl;
        if (v) l[lcarry + lexcl] = i;
        carry += total; lcarry += ltotal;
    }
    if (threadIdx.x == 0) { counts[blockIdx.x] = carry; n_list[blockIdx.x] = lcarry; }
}

// Fill from screen tiles (the refiner's fused render -> cloud path): kFillCtas CTAs per image walk the image's list of
// non-empty tiles, so only the ~10 % of tiles the object covers are read again (the segment-based pass above reads
// every pixel of every image twice, and a grid with one CTA per tile spends its time launching 70k empty CTAs).  Points of an image are ordered tile by
// tile (row-major tiles, row-major pixels inside a tile) -- deterministic, but NOT depth2cloud_cuda's row-major
// order; the coordinates themselves are the same values.  A tile is 64 pixels wide and a multiple of 32 rows high; a
// band of 32 rows = kSegPx pixels is handled per step, 8 pixels per thread.
constexpr int kFillCtas = 16;
__global__ void __launch_bounds__(kSegThreads)
cloud_fill_tiles_kernel(const int32_t* __restrict__ depth, uint32_t width, uint32_t height, int tiles_x, uint32_t n_tiles,
                        const unsigned* __restrict__ tile_list, const unsigned* __restrict__ n_list,
                        const unsigned* __restrict__ tile_off,
                        const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts, Intrinsics K,
                        float* __restrict__ out, size_t capacity, int tile_h) {
    __shared__ unsigned s_warp[kSegThreads / 32];
    __shared__ float s_fx[64], s_fy[128];                         // (u - cx) / fx per tile column, (v - cy) / fy per tile row
    __shared__ float s_pts[kSegThreads / 32][32 * kPxPerThread * 3];   // per warp: its compacted points of one band
    const uint32_t image = blockIdx.y;
    if (counts[image] == 0) return;                               // empty, or the cloud did not fit (overflow)
    const uint32_t n_here = n_list[image];
  for (uint32_t li = blockIdx.x; li < n_here; li += gridDim.x) {
    const uint32_t tile = tile_list[(size_t)image * n_tiles + li];
    const size_t t_idx = (size_t)image * n_tiles + tile;
    const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
    // the two pixel-only factors of icp.cu:250-251, one IEEE division per thread instead of two per point
    __syncthreads();
    if (threadIdx.x < 64) s_fx[threadIdx.x] = divf(subf((float)(tx * 64 + threadIdx.x), K.cx), K.fx);
    else if ((int)threadIdx.x < 64 + tile_h) s_fy[threadIdx.x - 64] = divf(subf((float)(ty * tile_h + threadIdx.x - 64), K.cy), K.fy);
    __syncthreads();
    unsigned done = 0;                                            // points of this tile written by earlier 32-row bands
   for (int band = 0; band < tile_h; band += 32) {
    // A warp covers 4 tile rows x 64 columns; lane l takes the 8 pixels of row (l & 3), column chunk (l >> 2), so that 32
    // CONSECUTIVE POINTS of the cloud are an 8 x 4 pixel block, not a 32 x 1 strip: the 32 lanes of an ICP warp then look up
    // neighbours in two dimensions -- the same few tree nodes / grid blocks / scene sectors (PR_CLOUD_STRIPS: the old order).
#ifdef PR_CLOUD_STRIPS
    const uint32_t row_in_band = threadIdx.x >> 3, chunk = threadIdx.x & 7;
#else
    const uint32_t row_in_band = (threadIdx.x >> 5) * 4 + (threadIdx.x & 3), chunk = (threadIdx.x & 31) >> 2;
#endif
    const uint32_t v = ty * tile_h + band + row_in_band;          // 32 rows per band
    const uint32_t u0 = tx * 64 + chunk * 8;
    const int32_t* img = depth + (size_t)image * width * height;
    int d[kPxPerThread];
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < kPxPerThread; k++) d[k] = 0;
    if (v < height) {
        const int32_t* row = img + (size_t)v * width;
        if (u0 + kPxPerThread <= width && ((reinterpret_cast<uintptr_t>(row + u0) & 15) == 0)) {
            const int4 a = __ldg(reinterpret_cast<const int4*>(row + u0));
            const int4 b = __ldg(reinterpret_cast<const int4*>(row + u0) + 1);
            d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < kPxPerThread; k++) d[k] = (u0 + k < width) ? row[u0 + k] : 0;
        }
#pragma unroll
        for (int k = 0; k < kPxPerThread; k++) c += (d[k] > 0) ? 1u : 0u;
    }
    unsigned total;
    const unsigned excl = block_excl_scan(c, s_warp, &total);
    // The points of a warp (4 tile rows) are consecutive in the output.  Each lane parks its points in the warp's
    // shared-memory block at its rank, then the warp copies the block with lane-consecutive stores (a lane writing its
    // own points directly scatters 4-byte stores over 3 KB per instruction).
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned wincl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, wincl, o); if (lane >= (unsigned)o) wincl += t; }
    const unsigned wtotal = __shfl_sync(0xffffffffu, wincl, 31);
    unsigned slot = wincl - c;                                    // rank inside the warp
    const unsigned wfirst = __shfl_sync(0xffffffffu, excl, 0);    // rank of the warp's first point inside the band
    float* sp = s_pts[warp];
#pragma unroll
    for (int k = 0; k < kPxPerThread; k++) {
        if (d[k] > 0) {
            // icp.cu:249-251
            const float z = divf((float)d[k], 1000.0f);
            sp[3 * slot + 0] = mulf(s_fx[chunk * 8 + k], z);
            sp[3 * slot + 1] = mulf(s_fy[band + row_in_band], z);
            sp[3 * slot + 2] = z;
            slot++;
        }
    }
    __syncwarp();
    const size_t wdst = (size_t)offsets[image] + tile_off[t_idx] + done + wfirst;      // first output point of this warp
    const size_t room = (wdst < capacity) ? capacity - wdst : 0;
    const unsigned n_out = (unsigned)min((size_t)wtotal, room) * 3;
    for (unsigned i = lane; i < n_out; i += 32) out[3 * wdst + i] = sp[i];
    __syncwarp();
    done += total;
   }
  }
}

inline size_t cloud_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int cloud_from_tiles(const int32_t* depth_dev, size_t n_images, uint32_t width, uint32_t height, const float K[9],
                     int tile_w, int tile_h, int tiles_x, int tiles_y, const unsigned* tile_valid, unsigned* tile_off,
                     uint32_t* counts_dev, uint32_t* offsets_dev, uint32_t* overflow_dev, size_t capacity_points,
                     uint32_t align_points, float* out_pts_dev, cudaStream_t stream) {
    if (tile_w != 64 || tile_h % 32 != 0 || tile_h > 128) return PR_ERR_UNSUPPORTED;       // the fill kernel's thread -> pixel map
    if (((size_t)width * height + align_points) * (n_images < 256 ? n_images : 256) > 0xFFFFFFFFull) return PR_ERR_INVALID_ARGUMENT;   // 32-bit block totals of the offset scan
    const uint32_t n_tiles = (uint32_t)(tiles_x * tiles_y);
    // scratch behind the offsets: list of non-empty tiles per image, then the list lengths (see cloud_tiles_scratch_words)
    unsigned* tile_list = tile_off + n_images * (size_t)n_tiles;
    unsigned* n_list = tile_list + n_images * (size_t)n_tiles;
    tile_scan_kernel<<<(unsigned)n_images, kSegThreads, 0, stream>>>(tile_valid, tile_off, n_tiles, counts_dev, tile_list, n_list);
    cloud_offsets_kernel<<<1, kSegThreads, 0, stream>>>(counts_dev, (uint32_t)n_images, align_points,
                                                        capacity_points ? (unsigned long long)capacity_points : ~0ull,
                                                        offsets_dev, overflow_dev);
    const Intrinsics Ki = {K[0], K[4], K[2], K[5]};
    cloud_fill_tiles_kernel<<<dim3(kFillCtas, (unsigned)n_images), kSegThreads, 0, stream>>>(
        depth_dev, width, height, tiles_x, n_tiles, tile_list, n_list, tile_off, offsets_dev, counts_dev, Ki, out_pts_dev,
        capacity_points ? capacity_points : ~(size_t)0, tile_h);
    count_launch(3);
    PR_LAUNCH_CHECK();
    return PR_OK;
}

}  // namespace prb

using namespace prb;

extern "C" {

size_t pr_depth2cloud_workspace_bytes(size_t n_images, uint32_t width, uint32_t height) {
    const size_t n_px = (size_t)width * height;
    const size_t n_seg = (n_px + kSegPx - 1) / kSegPx;
    return cloud_align_up(n_images * n_seg * 4, 256);
}

int pr_depth2cloud_count(const void* depth_dev, int depth_is_int32, size_t n_images, uint32_t width, uint32_t height,
                         uint32_t stride, uint32_t align_points, size_t capacity_points, uint32_t* counts_dev,
                         uint32_t* offsets_dev, uint32_t* overflow_dev,
                         void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    if (!depth_dev || !counts_dev || !offsets_dev || !workspace_dev) return PR_ERR_INVALID_ARGUMENT;
    if (stride != 1) return PR_ERR_UNSUPPORTED;   // out of bounds upstream (icp.cpp:77-82, icp.cu:235-245)
    if (align_points == 0 || width == 0 || height == 0) return PR_ERR_INVALID_ARGUMENT;
    const size_t n_px = (size_t)width * height;
    if (n_px > 0x7FFFFFFFull || n_images > 65535) return PR_ERR_INVALID_ARGUMENT;
    if ((n_px + align_points) * (n_images < 256 ? n_images : 256) > 0xFFFFFFFFull) return PR_ERR_INVALID_ARGUMENT;   // 32-bit block totals of the offset scan
    if (workspace_bytes < pr_depth2cloud_workspace_bytes(n_images, width, height)) return PR_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = as_stream(stream_);
    const uint32_t n_seg = (uint32_t)((n_px + kSegPx - 1) / kSegPx);
    uint32_t* seg = (uint32_t*)workspace_dev;
    if (n_images == 0) { PR_CUDA_TRY(cudaMemsetAsync(offsets_dev, 0, 4, stream)); return PR_OK; }
    const int vec = ((n_px % kPxPerThread) == 0) && (((uintptr_t)depth_dev & 15) == 0);
    const dim3 grid(n_seg, (unsigned)n_images);
    if (depth_is_int32) cloud_count_kernel<int32_t><<<grid, kSegThreads, 0, stream>>>((const int32_t*)depth_dev, (uint32_t)n_px, n_seg, seg, vec);
    else cloud_count_kernel<uint16_t><<<grid, kSegThreads, 0, stream>>>((const uint16_t*)depth_dev, (uint32_t)n_px, n_seg, seg, vec);
    cloud_scan_kernel<<<(unsigned)n_images, kSegThreads, 0, stream>>>(seg, seg, n_seg, counts_dev);
    cloud_offsets_kernel<<<1, kSegThreads, 0, stream>>>(counts_dev, (uint32_t)n_images, align_points,
                                                        capacity_points ? (unsigned long long)capacity_points : ~0ull,
                                                        offsets_dev, overflow_dev);
    count_launch(3);
    PR_LAUNCH_CHECK();
    return PR_OK;
}

int pr_depth2cloud_fill(const void* depth_dev, int depth_is_int32, size_t n_images, uint32_t width, uint32_t height,
                        const float K[9], uint32_t stride, uint32_t tl_x, uint32_t tl_y,
                        const uint32_t* offsets_dev, float* out_pts_dev, size_t capacity_points,
                        const void* workspace_dev, size_t workspace_bytes, pr_stream_t stream_) {
    if (!depth_dev || !K || !offsets_dev || !out_pts_dev || !workspace_dev) return PR_ERR_INVALID_ARGUMENT;
    if (stride != 1) return PR_ERR_UNSUPPORTED;
    const size_t n_px = (size_t)width * height;
    if (n_px == 0 || n_px > 0x7FFFFFFFull || n_images > 65535) return PR_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < pr_depth2cloud_workspace_bytes(n_images, width, height)) return PR_ERR_WORKSPACE_TOO_SMALL;
    if (n_images == 0) return PR_OK;
    cudaStream_t stream = as_stream(stream_);
    const uint32_t n_seg = (uint32_t)((n_px + kSegPx - 1) / kSegPx);
    const uint32_t* seg = (const uint32_t*)workspace_dev;
    const int vec = ((n_px % kPxPerThread) == 0) && (((uintptr_t)depth_dev & 15) == 0);
    const Intrinsics Ki = {K[0], K[4], K[2], K[5]};
    const dim3 grid(n_seg, (unsigned)n_images);
    if (depth_is_int32)
        cloud_fill_kernel<int32_t><<<grid, kSegThreads, 0, stream>>>((const int32_t*)depth_dev, width, (uint32_t)n_px, n_seg, seg,
                                                                    offsets_dev, Ki, tl_x, tl_y, out_pts_dev, capacity_points ? capacity_points : ~(size_t)0, vec);
    else
        cloud_fill_kernel<uint16_t><<<grid, kSegThreads, 0, stream>>>((const uint16_t*)depth_dev, width, (uint32_t)n_px, n_seg, seg,
                                                                     offsets_dev, Ki, tl_x, tl_y, out_pts_dev, capacity_points ? capacity_points : ~(size_t)0, vec);
    count_launch();
    PR_LAUNCH_CHECK();
    return PR_OK;
}

}  // extern "C"
