// icp_device.cuh -- device-side building blocks of the ICP kernels (icp.cu): scene views, the two
// Scene_*::query restatements, the per-point 29-sum term in its three forms, the transposing warp reduction.
// Split out of icp.cu so that the drivers and the parity/debug kernels share ONE copy of every function.
#pragma once
#include "common.cuh"
#include "solver.cuh"
#include <float.h>
#include <limits.h>

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
__device__ __forceinline__ int nn_search_reference(const NnScene& s, float px, float py, float pz) {
    if (s.n_nodes <= 0) return -1;
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
    if (!(best_d2 < s.max_dist_sq)) return -1;
    return best;
}
__device__ __forceinline__ bool query(const NnScene& s, float px, float py, float pz, Corr& c) {
    const int best = nn_search_reference(s, px, py, pz);
    if (best < 0) return false;
    c.qx = __ldg(s.pcd + 3 * best); c.qy = __ldg(s.pcd + 3 * best + 1); c.qz = __ldg(s.pcd + 3 * best + 2);
    c.nx = __ldg(s.nrm + 3 * best); c.ny = __ldg(s.nrm + 3 * best + 1); c.nz = __ldg(s.nrm + 3 * best + 2);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Packed kd-tree for the hypothesis-resident driver.  Same tree (same nodes, same leaf ranges, same points) as
// the reference's Node_kdtree array, re-laid out per ICP call so that a node is two aligned float4:
//     {lo.x, lo.y, lo.z, a}   {hi.x, hi.y, hi.z, split_v}
// a >= 0: internal node, a = split_dim << 28 | child1, children child1 and child1 + 1 (build_tree appends them together,
// pcd_scene.cpp:160-170); a < 0: leaf, a = 0x80000000 | count << 24 | left.  [lo,hi] is the box of the node's OWN
// points -- computed here for leaves too (the reference stores none for leaves, pcd_scene.h:14-19).
// build_tree numbers the nodes generation by generation (breadth first), so nodes [0, kTopNodes) ARE the top levels
// of the tree: the kernel keeps them in shared memory (one TMA bulk copy per CTA), deeper nodes come through L1/L2.
//
// The query is an exact nearest-neighbour search that returns the SAME point as Scene_nn::query, ties included.
// It descends into the child whose box is nearer first (that finds a good candidate early, which is what makes the
// pruning bite: the reference's own order -- the query's side of the split plane first -- needs 2-3x the node visits
// on thin-shell scenes when the query is still a few centimetres off the surface), prunes a subtree when the distance
// to the box of its OWN points, scaled by 0.99999 (the bound is a rounded float), is >= the best distance so far, and
// keeps the candidate on strict <.  None of that can change WHICH point wins unless two points are at exactly the same
// float distance; the reference then keeps the one its own walk visits first (strict <, pcd_scene.h:88-90).  So on
// d2 == best the two candidates are put in the reference's visiting order explicitly (nn_visited_first: walk down
// from the root to the node where their index ranges part; the child on the query's side of the split plane is
// visited first, pcd_scene.h:92-100; inside a leaf the lower index).  A subtree that holds a point at exactly the best
// distance is never pruned (its bound is <= that distance, strictly below it after the margin), so every tie is seen.
// The search starts from best = max_dist^2 (anything farther is invalid anyway, pcd_scene.h:127) and parks the far
// children whose box is still within reach (id + bound) on a small explicit stack.
// Measured on C3 (512 hypotheses, 99k-point tree, 4 CTAs per SM): 512 nodes staged 155 ms, 128 nodes 145 ms, 16 nodes 107 ms
// per step, none 136 ms (before the search cache) -- the walk lives in L1, every CTA of an SM would hold its own copy of the
// same top levels, and L1 already keeps ONE copy of them hot; beyond the first levels staging only takes L1 away.
#ifndef PR_NN_TOP
#define PR_NN_TOP 16
#endif
constexpr int kTopNodes = PR_NN_TOP;      // the top 4 levels (15 nodes), 512 bytes of shared memory per CTA

// ---- hash grid over the scene points: the fast path of the exact nearest-neighbour query ----------------------------
// Built per ICP call next to the packed tree (a few small kernels): cubic cells of side c (a small multiple of the scene's
// point spacing, estimated from the tree's leaves).  The unit of the table is a BLOCK of 2 x 2 x 2 cells, one per anchor
// cell that has a point in any of its eight cells: an open-addressing hash table anchor -> (start, count) and, per block,
// a contiguous copy of its points (every point sits in the eight blocks that contain its cell).  A query reads the block
// whose centre is nearest to it: every scene point within 0.5 c of the query lies in that block, so when the best candidate
// of the block is nearer than 0.498 c (the margin covers the float rounding of the cell coordinates) it IS the nearest
// neighbour -- ties included, they are at the same distance and therefore in the block as well, and nn_visited_first puts
// them in the reference's order.  One table probe and one contiguous list of a few dozen points replace a 15-level
// dependent walk.  Anything else (nothing that near, a query outside the grid, an overfull block) is answered by the tree.
// (First version: one list per CELL and eight probes + eight short loops per query -- as many instructions as the walk it
// replaced, because every one of the eight loops ran for the fullest cell among the warp's lanes: 79 ms per C3 step vs 63.)
struct NnGridParams {         // device memory, written by nn_grid_params_kernel
    float ox, oy, oz, inv_c;
    float r_ok_sq;            // (0.498 c)^2
    unsigned mask, shift;     // table size - 1, 32 - log2(table size)
    unsigned enabled;
};
struct NnGrid {
    const NnGridParams* params;   // nullptr: no grid
    const uint2* table;           // {anchor key, start << 8 | count}; key 0xFFFFFFFF = empty; count 255 = overfull block
    const float4* gpts;           // per block, its scene points: {x, y, z, leaf-order index as int bits}
};
constexpr unsigned kGridEmpty = 0xFFFFFFFFu;
constexpr float kGridMaxCells = 1023.0f;          // per axis (10 bits of the key)

struct PackedNnScene {
    float max_dist_sq;
    const float4* nodes;      // 2 per node
    const float4* pts4;       // {x, y, z, 0}
    const float* nrm;         // original Vec3f normals
    int n_nodes;
    const unsigned* unsupported;   // device flag raised by nn_pack_nodes_kernel: this tree does not fit the encoding
    const float4* top;        // shared-memory copy of nodes [0, n_top) (set by the kernel; nullptr: none)
    int n_top;
    int* cache;               // one int per model point of the batch (same indexing as the points): last pass' winner; nullable
    NnScene ref;              // the reference layout (fallback walk when the stack would overflow)
    NnGrid grid;
};

__global__ void __launch_bounds__(256)
nn_pack_points_kernel(const float* __restrict__ pcd, size_t n, float4* __restrict__ pts4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) pts4[i] = make_float4(pcd[3 * i], pcd[3 * i + 1], pcd[3 * i + 2], 0.f);
}
// sets *unsupported when the tree does not fit the packed encoding (a leaf of more than 127 points, left >= 2^24,
// children that are not siblings, a child index >= 2^28)
__global__ void __launch_bounds__(256)
nn_pack_nodes_kernel(const pr_node_kdtree* __restrict__ nodes, int n_nodes, const float* __restrict__ pcd,
                     float4* __restrict__ out, unsigned* __restrict__ unsupported) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_nodes) return;
    const pr_node_kdtree nd = nodes[i];
    float lo[3], hi[3];
    int a;
    float w = 0.f;
    if (nd.child1 < 0 || nd.child2 < 0) {
        const int cnt = nd.right - nd.left;
        if (cnt < 0 || cnt > 127 || nd.left < 0 || nd.left >= (1 << 24)) { *unsupported = 1; return; }
        for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
        for (int j = nd.left; j < nd.right; j++)
            for (int k = 0; k < 3; k++) { const float v = pcd[3 * j + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
        a = (int)(0x80000000u | ((unsigned)cnt << 24) | (unsigned)nd.left);
    } else {
        if (nd.child2 != nd.child1 + 1 || nd.child1 >= (1 << 28) || nd.split_dim < 0 || nd.split_dim > 2) { *unsupported = 1; return; }
        for (int k = 0; k < 3; k++) { lo[k] = nd.bbox[2 * k]; hi[k] = nd.bbox[2 * k + 1]; }
        a = nd.child1 | (nd.split_dim << 28);
        w = nd.split_v;
    }
    out[2 * i] = make_float4(lo[0], lo[1], lo[2], __int_as_float(a));
    out[2 * i + 1] = make_float4(hi[0], hi[1], hi[2], w);
}

__device__ __forceinline__ float box_dist_sq(const float4& lo, const float4& hi, float px, float py, float pz) {
    const float dx = fmaxf(fmaxf(lo.x - px, px - hi.x), 0.f);
    const float dy = fmaxf(fmaxf(lo.y - py, py - hi.y), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - pz, pz - hi.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
}
// children (c, c + 1) of an internal node: four consecutive float4 = one 64-byte fetch.  The top levels come from the
// CTA's shared-memory copy, deeper nodes through L1/L2; the pointer is SELECTED (not branched on), and the loads are
// generic, so lanes of a warp that are at different depths do not diverge here.
__device__ __forceinline__ const float4* node_ptr(const PackedNnScene& s, int i) {
    return (i + 1 < s.n_top) ? s.top + 2 * i : s.nodes + 2 * i;
}

// true when Scene_nn::query's walk reaches point a before point b (a != b) for this query
__device__ __noinline__ bool nn_visited_first(const NnScene& s, float px, float py, float pz, int a, int b) {
    int cur = 0;
    for (;;) {
        const pr_node_kdtree* nd = s.nodes + cur;
        const int c1 = nd->child1, c2 = nd->child2;
        if (c1 < 0 || c2 < 0) return a < b;                                   // same leaf: scanned in index order
        const int mid = s.nodes[c2].left;                                     // child1 owns [left, mid), child2 [mid, right)
        const bool a1 = a < mid, b1 = b < mid;
        if (a1 == b1) { cur = a1 ? c1 : c2; continue; }
        const int dim = nd->split_dim;
        const float diff = (dim == 0 ? px : (dim == 1 ? py : pz)) - nd->split_v;
        return (diff < 0.f) ? a1 : !a1;                                       // the query's side first (pcd_scene.h:92-100)
    }
}

// cell coordinates (in cell units, origin at the grid origin) -- the same expression for scene points and queries
__device__ __forceinline__ void grid_coords(const NnGridParams& g, float x, float y, float z, float& fx, float& fy, float& fz) {
    fx = (x - g.ox) * g.inv_c; fy = (y - g.oy) * g.inv_c; fz = (z - g.oz) * g.inv_c;
}
__device__ __forceinline__ unsigned grid_key(int ix, int iy, int iz) { return (unsigned)ix | ((unsigned)iy << 10) | ((unsigned)iz << 20); }
// Slot of a key.  (Tried: a locality-preserving slot -- the blocks of a 4 x 4 x 4 neighbourhood in 64 consecutive slots, so that
// the probes of a warp share cache lines: slower, its clustered keys probe longer (66 ms vs 58 ms at n slots).  What matters is
// the table's SIZE: 16 n slots 62 ms, 4 n 59 ms, 2 n 57 ms per C3 step.)
__device__ __forceinline__ unsigned grid_hash(unsigned key, unsigned shift) { return (key * 2654435761u) >> shift; }

#ifndef PR_NN_CELL
#define PR_NN_CELL 3.0f       // cell side in units of the leaf-estimated point spacing (C3: 2.0 69 ms, 2.5 57-63, 3.0 57-61, 3.5 59, 4.0 72)
#endif
// one CTA: point spacing from the leaves of the packed tree -> cell size, origin, table geometry
__global__ void __launch_bounds__(256)
nn_grid_params_kernel(const float4* __restrict__ nodes, int n_nodes, unsigned n_points, unsigned log2_slots,
                      const unsigned* __restrict__ unsupported, NnGridParams* __restrict__ out) {
    __shared__ float s_sum[256];
    __shared__ unsigned s_cnt[256];
    float sum = 0.f;
    unsigned cnt = 0;
    for (int i = threadIdx.x; i < n_nodes; i += 256) {
        const float4 lo = nodes[2 * i], hi = nodes[2 * i + 1];
        const int a = __float_as_int(lo.w);
        if (a >= 0) continue;
        const int c = (a >> 24) & 127;
        if (c < 3) continue;
        float e0 = hi.x - lo.x, e1 = hi.y - lo.y, e2 = hi.z - lo.z;      // the two largest extents span the surface patch
        const float mn = fminf(e0, fminf(e1, e2));
        const float prod = (mn == e0) ? e1 * e2 : ((mn == e1) ? e0 * e2 : e0 * e1);
        if (prod > 0.f) { sum += sqrtf(prod / (float)c); cnt++; }
    }
    s_sum[threadIdx.x] = sum; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float4 lo = nodes[0], hi = nodes[1];
        const float ext = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
        float c = s_cnt[0] ? PR_NN_CELL * s_sum[0] / (float)s_cnt[0] : 0.f;
        c = fmaxf(c, ext / (kGridMaxCells - 4.0f));                    // the key holds 10 bits per axis
        NnGridParams g;
        g.ox = lo.x - 1.5f * c; g.oy = lo.y - 1.5f * c; g.oz = lo.z - 1.5f * c;
        g.inv_c = 1.0f / c;
        g.r_ok_sq = (0.498f * c) * (0.498f * c);
        g.mask = (1u << log2_slots) - 1u; g.shift = 32u - log2_slots;
        g.enabled = (c > 0.f && c < 3.0e38f && g.inv_c > 0.f && g.inv_c < 3.0e38f && n_points < (1u << 21) && !(unsupported && *unsupported)) ? 1u : 0u;
        *out = g;
    }
}
// per (scene point, one of the eight blocks that contain its cell): claim / find the block's slot, count the point
__global__ void __launch_bounds__(256)
nn_grid_count_kernel(const float4* __restrict__ pts4, unsigned n, NnGridParams* gp, unsigned* __restrict__ keys,
                     unsigned* __restrict__ counts, unsigned* __restrict__ pt_slot) {
    const unsigned t = blockIdx.x * 256 + threadIdx.x;
    const unsigned i = t >> 3, k = t & 7u;
    if (i >= n) return;
    const NnGridParams g = *gp;
    if (!g.enabled) return;
    const float4 p = pts4[i];
    float fx, fy, fz;
    grid_coords(g, p.x, p.y, p.z, fx, fy, fz);
    // the point's cell is >= 1.5 cells inside the grid by construction; anchors of its blocks: cell - {0,1} per axis
    const unsigned key = grid_key((int)fx - (int)(k & 1u), (int)fy - (int)((k >> 1) & 1u), (int)fz - (int)(k >> 2));
    unsigned h = grid_hash(key, g.shift);
    for (int probes = 0;; probes++) {
        const unsigned old = atomicCAS(keys + h, kGridEmpty, key);
        if (old == kGridEmpty || old == key) break;
        if (probes == 256) {            // the table (sized for ~one block per point) is too full for this scene: no grid, the tree answers
            gp->enabled = 0u;
            pt_slot[t] = kGridEmpty;
            return;
        }
        h = (h + 1) & g.mask;
    }
    atomicAdd(counts + h, 1u);
    pt_slot[t] = h;
}
// one CTA of 1024 threads: exclusive scan of the slot counts -> table entries {key, start << 8 | count}; cursor = start
__global__ void __launch_bounds__(1024)
nn_grid_scan_kernel(const unsigned* __restrict__ keys, const unsigned* __restrict__ counts, unsigned n_slots, unsigned* __restrict__ cursor,
                    uint2* __restrict__ table) {
    __shared__ unsigned s_warp[32];
    const unsigned per = (n_slots + 1023) / 1024;
    const unsigned b = threadIdx.x * per, e = min(b + per, n_slots);
    unsigned sum = 0;
    for (unsigned i = b; i < e; i++) sum += counts[i];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (unsigned)o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (unsigned)o) wi += t; }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    unsigned run = s_warp[warp] + incl - sum;
    for (unsigned i = b; i < e; i++) {
        const unsigned c = counts[i];
        cursor[i] = run;
        table[i] = make_uint2(keys[i], (run << 8) | min(c, 255u));
        run += c;
    }
}
__global__ void __launch_bounds__(256)
nn_grid_fill_kernel(const float4* __restrict__ pts4, unsigned n, const NnGridParams* __restrict__ gp, const unsigned* __restrict__ pt_slot,
                    unsigned* __restrict__ cursor, float4* __restrict__ gpts) {
    const unsigned t = blockIdx.x * 256 + threadIdx.x;
    const unsigned i = t >> 3;
    if (i >= n || !gp->enabled || pt_slot[t] == kGridEmpty) return;
    const float4 p = pts4[i];
    const unsigned at = atomicAdd(cursor + pt_slot[t], 1u);            // order inside a block is arbitrary: the query's result is not
    gpts[at] = make_float4(p.x, p.y, p.z, __int_as_float((int)i));
}

// The grid's answer: true when best_i (>= 0) is certainly the point Scene_nn::query returns; false = ask the tree.
template <bool COUNT>
__device__ __forceinline__ bool nn_grid_query(const PackedNnScene& s, const NnGridParams& g, float px, float py, float pz, int& best_i,
                                              unsigned& tests) {
    best_i = -1;
    float fx, fy, fz;
    grid_coords(g, px, py, pz, fx, fy, fz);
    if (!(fx >= 0.5f && fx < kGridMaxCells - 0.5f && fy >= 0.5f && fy < kGridMaxCells - 0.5f && fz >= 0.5f && fz < kGridMaxCells - 0.5f)) return false;
    const unsigned key = grid_key((int)(fx - 0.5f), (int)(fy - 0.5f), (int)(fz - 0.5f));
    unsigned h = grid_hash(key, g.shift);
    uint2 v = __ldg(s.grid.table + h);
    while (v.x != key && v.x != kGridEmpty) { h = (h + 1) & g.mask; v = __ldg(s.grid.table + h); }
    if (v.x != key) return false;
    const unsigned cnt = v.y & 255u;
    const float4* __restrict__ list = s.grid.gpts + (v.y >> 8);
    if (COUNT) tests += cnt;
    float best = fminf(g.r_ok_sq, s.max_dist_sq);
    // four points per step: their loads are in flight together (the elements behind a list are the next list's points or the
    // pad behind the last list, so the loads need no bound; the tests do)
#ifndef PR_NN_LIST_ILP
#define PR_NN_LIST_ILP 4
#endif
    for (unsigned j = 0; j < cnt; j += PR_NN_LIST_ILP) {
        float4 q[PR_NN_LIST_ILP];
#pragma unroll
        for (int k = 0; k < PR_NN_LIST_ILP; k++) q[k] = __ldg(list + j + k);
#pragma unroll
        for (int k = 0; k < PR_NN_LIST_ILP; k++) {
            const float dx = px - q[k].x, dy = py - q[k].y, dz = pz - q[k].z;
            const float d2 = addf(addf(mulf(dx, dx), mulf(dy, dy)), mulf(dz, dz));    // pcd_scene.h:86-89
            if (d2 <= best && j + k < cnt) {
                const int i = __float_as_int(q[k].w);
                if (d2 < best) { best = d2; best_i = i; }
                else if (best_i >= 0 && i != best_i && nn_visited_first(s.ref, px, py, pz, i, best_i)) best_i = i;
            }
        }
    }
    return cnt != 255u && best_i >= 0;
}

// exact nearest neighbour over the packed tree: index of the winner (leaf order), or -1 when nothing is nearer
// than max_dist.  -2: the explicit stack overflowed (the caller falls back to the reference walk).
// hint: a scene point to start from (>= 0), typically the winner of the previous ICP pass of the same model point -- the
// pose moves little between passes, so its distance is already (nearly) the minimum and the walk only has to prove it:
// every subtree farther than that is pruned at once.  The result does not depend on the hint (it is just a candidate
// "visited" early; ties are ordered by nn_visited_first, not by visiting order).
template <bool COUNT>
__device__ __forceinline__ int nn_search_packed_t(const PackedNnScene& s, float px, float py, float pz, unsigned& visits, unsigned& tests,
                                                  int hint = -1) {
    if (s.n_nodes <= 0) return -1;
    constexpr int kStack = 24;       // far children parked: one per level at most (trees here are <= 20 deep)
    int stack_n[kStack];
    float stack_lb[kStack];
    int sp = 0;
    float best = s.max_dist_sq;
    int best_i = -1;
    if (hint >= 0) {
        const float4 q = __ldg(s.pts4 + hint);
        const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
        const float d2 = addf(addf(mulf(dx, dx), mulf(dy, dy)), mulf(dz, dz));
        if (d2 < best) { best = d2; best_i = hint; }
    }
    bool overflow = false;
    float4 lo, hi;
    { const float4* p = node_ptr(s, 0); lo = p[0]; hi = p[1]; }
    if (COUNT) visits++;
    // a box lower bound is a rounded float, so it is trusted only with a 1e-5 margin
    bool go = box_dist_sq(lo, hi, px, py, pz) * 0.99999f < best;
    // next parked subtree that is still within reach -> (lo, hi); false when the stack is exhausted
    auto pop = [&]() {
        while (sp > 0) {
            --sp;
            if (stack_lb[sp] < best) {
                const float4* p = (stack_n[sp] < s.n_top) ? s.top + 2 * stack_n[sp] : s.nodes + 2 * stack_n[sp];
                lo = p[0]; hi = p[1];
                return true;
            }
        }
        return false;
    };
    // Two phases per round, so that the lanes of a warp scan their leaves TOGETHER: (A) walk internal nodes until the
    // current node is a leaf, (B) scan it.  With the scan inside the walk loop a lane scanned its leaf while the others were
    // still descending: the scan -- half of the kernel's instructions -- ran on 4.8 of 32 lanes (ncu, profiles/r02_ncu_icp_nn.txt).
    while (go) {
        while (go && __float_as_int(lo.w) >= 0) {
            const int c1 = __float_as_int(lo.w) & 0xFFFFFFF;
            const float4* p = node_ptr(s, c1);
            const float4 lo1 = p[0], hi1 = p[1], lo2 = p[2], hi2 = p[3];
            if (COUNT) visits += 2;
            const float lb1 = box_dist_sq(lo1, hi1, px, py, pz) * 0.99999f, lb2 = box_dist_sq(lo2, hi2, px, py, pz) * 0.99999f;
            const bool first1 = lb1 <= lb2;          // the nearer box first
            const float lb_near = first1 ? lb1 : lb2, lb_far = first1 ? lb2 : lb1;
            if (lb_far < best) {
                if (sp < kStack) { stack_n[sp] = first1 ? c1 + 1 : c1; stack_lb[sp] = lb_far; sp++; }
                else overflow = true;
            }
            if (lb_near < best) { lo = first1 ? lo1 : lo2; hi = first1 ? hi1 : hi2; }
            else go = pop();
        }
        if (!go) break;
        {
            const int a = __float_as_int(lo.w);
            const int left = a & 0xFFFFFF, cnt = (a >> 24) & 127;
            if (COUNT) tests += (unsigned)cnt;
#ifndef PR_NN_LEAF_ILP
#define PR_NN_LEAF_ILP 2
#endif
            // PR_NN_LEAF_ILP points per step, their loads in flight together (behind the points of the last leaf lie the
            // packed nodes: the loads need no bound, the tests do)
            for (int i0 = left; i0 < left + cnt; i0 += PR_NN_LEAF_ILP) {
                float4 q[PR_NN_LEAF_ILP];
#pragma unroll
                for (int k = 0; k < PR_NN_LEAF_ILP; k++) q[k] = __ldg(s.pts4 + i0 + k);
#pragma unroll
                for (int k = 0; k < PR_NN_LEAF_ILP; k++) {
                    const int i = i0 + k;
                    const float dx = px - q[k].x, dy = py - q[k].y, dz = pz - q[k].z;
                    const float d2 = addf(addf(mulf(dx, dx), mulf(dy, dy)), mulf(dz, dz));    // pcd_scene.h:86-89
                    if (d2 <= best && i < left + cnt) {
                        if (d2 < best) { best = d2; best_i = i; }
                        else if (best_i >= 0 && i != best_i && nn_visited_first(s.ref, px, py, pz, i, best_i)) best_i = i;
                    }
                }
            }
        }
        go = pop();
    }
    return overflow ? -2 : best_i;
}
__device__ __forceinline__ int nn_search_packed(const PackedNnScene& s, float px, float py, float pz) {
    unsigned v = 0, t = 0;
    return nn_search_packed_t<false>(s, px, py, pz, v, t);
}

// hint / found: see nn_search_packed_t; found = -1 when there is no valid correspondence
__device__ __forceinline__ bool query(const PackedNnScene& s, float px, float py, float pz, Corr& c, int hint, int& found) {
    found = -1;
    if (s.nodes == nullptr) return query(s.ref, px, py, pz, c);
    unsigned v = 0, t = 0;
    const int best_i = nn_search_packed_t<false>(s, px, py, pz, v, t, hint);
    if (best_i == -2) return query(s.ref, px, py, pz, c);     // deeper than the stack: the reference walk
    if (best_i < 0) return false;
    found = best_i;
    const float4 q = __ldg(s.pts4 + best_i);
    c.qx = q.x; c.qy = q.y; c.qz = q.z;
    c.nx = __ldg(s.nrm + 3 * best_i); c.ny = __ldg(s.nrm + 3 * best_i + 1); c.nz = __ldg(s.nrm + 3 * best_i + 2);
    return true;
}
__device__ __forceinline__ bool query(const PackedNnScene& s, float px, float py, float pz, Corr& c) {
    if (s.nodes == nullptr) return query(s.ref, px, py, pz, c);
    const int best_i = nn_search_packed(s, px, py, pz);
    if (best_i == -2) return query(s.ref, px, py, pz, c);     // deeper than the stack: the reference walk
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
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t sub2(f2_t a, f2_t b) { f2_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t neg2(f2_t a) {
    f2_t d;
    asm("{\n.reg .f32 l, h;\nmov.b64 {l, h}, %1;\nneg.f32 l, l;\nneg.f32 h, h;\nmov.b64 %0, {l, h};\n}" : "=l"(d) : "l"(a));
    return d;
}
// s[i] = Vec29f entry i (icp.h:165-206) as (sum over A points, sum over B points); the inlier count is an integer
// (a predicated IADD runs on the ALU pipe, the FP32 pipe is the one this kernel saturates)
struct AccP { f2_t s[28]; int cnt; };
__device__ __forceinline__ void acc_zero(AccP& a) {
#pragma unroll
    for (int i = 0; i < 28; i++) a.s[i] = pk2(0.f, 0.f);
    a.cnt = 0;
}
__device__ __forceinline__ void acc_unpack(const AccP& a, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 28; i++) { float lo, hi; unpk2(a.s[i], lo, hi); v[i] = lo + hi; }
    v[28] = (float)a.cnt;
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

}  // namespace prb
