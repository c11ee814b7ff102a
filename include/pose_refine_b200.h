/* pose_refine_b200.h -- C ABI of libpose_refine_b200.so (sm_100a).
 *
 * This is the drop-in boundary for pose_refine's data-parallel hot path: the batched depth
 * rasteriser (cuda_renderer) and the point-to-plane ICP inner loop (cuda_icp).  The reference
 * has no FFI layer -- its boundary is the C++ header API of two static libraries
 * (cuda_renderer/renderer.h, cuda_icp/icp.h, cuda_icp/scene/...).  Each entry point below names
 * the reference interface (file:line under the reference tree) it replaces; the C++ headers in
 * include/pose_refine/ re-create the reference's own names on top of these calls.
 *
 * Conventions
 *   - plain pointers and sizes only; `*_dev` pointers are device memory on the current device,
 *     `*_host` pointers are host memory; pr_stream_t is a cudaStream_t (0 = default stream).
 *   - every call returns 0 (PR_OK), a negative PR_ERR_* code, or a positive cudaError_t value.
 *     Nothing exits the process (the reference's gpuErrchk does, renderer.cu:4-12).
 *   - calls are asynchronous on `stream` unless their comment says they synchronise.
 *   - no hidden allocation: scratch memory is caller-provided (`*_workspace_bytes` tells how much),
 *     except for the pr_refiner_* convenience object, which owns its buffers.
 *   - units follow the reference: meshes and rendered depth in millimetres, clouds in metres.
 */
#ifndef POSE_REFINE_B200_H
#define POSE_REFINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PR_OK 0
#define PR_ERR_INVALID_ARGUMENT (-1)
#define PR_ERR_UNSUPPORTED (-2)
#define PR_ERR_WORKSPACE_TOO_SMALL (-3)
#define PR_ERR_CAPACITY (-4)
#define PR_ERR_IO (-5)
#define PR_ERR_NO_DEVICE (-6)
#define PR_ERR_COMM (-7)          /* NCCL missing (libnccl.so.2 could not be loaded) or an NCCL call failed */

typedef struct CUstream_st* pr_stream_t;

/* Model::ROI, cuda_renderer/renderer.h:43-48.  width<=0 or height<=0 means "no ROI". */
typedef struct pr_roi { int x, y, width, height; } pr_roi;

/* ICPConvergenceCriteria, cuda_icp/icp.h:38-50 (defaults 1e-5, 1e-5, 30). */
typedef struct pr_icp_criteria {
    float relative_fitness;
    float relative_rmse;
    int max_iteration;
} pr_icp_criteria;

/* RegistrationResult, cuda_icp/icp.h:26-36: row-major 4x4, inlier_rmse_, fitness_ (72 bytes). */
typedef struct pr_registration_result {
    float transformation[16];
    float inlier_rmse;
    float fitness;
} pr_registration_result;

/* Scene_projective, cuda_icp/scene/depth_scene/depth_scene.h:7-15.  pcd/normal: width*height
 * Vec3f (3 packed floats) each, organised row-major, in DEVICE memory owned by the caller. */
typedef struct pr_scene_projective {
    uint64_t width, height;
    float max_dist_diff;
    float K[9];
    const float* pcd_dev;
    const float* normal_dev;
} pr_scene_projective;

/* Node_kdtree, cuda_icp/scene/pcd_scene/pcd_scene.h:5-25 (52 bytes, same field order). */
typedef struct pr_node_kdtree {
    int parent, child1, child2;
    float split_v;
    float bbox[6];
    int split_dim;
    int left, right;
} pr_node_kdtree;

/* Scene_nn, cuda_icp/scene/pcd_scene/pcd_scene.h:47-58: leaf-ordered points + normals + nodes in
 * DEVICE memory owned by the caller (KDTree_cuda, pcd_scene.h:37-43). */
typedef struct pr_scene_nn {
    float max_dist_diff;
    const float* pcd_dev;
    const float* normal_dev;
    const pr_node_kdtree* nodes_dev;
    uint64_t n_points, n_nodes;
} pr_scene_nn;

/* ---------------------------------------------------------------------------------------- */
/* library                                                                                    */
int pr_version(void);
const char* pr_error_string(int status);
/* PR_OK iff the current CUDA device can run this library (compute capability 10.x). */
int pr_device_check(void);

/* Device memory helpers so that C / C++ hosts above this ABI need no CUDA headers: they replace  */
/* the cudaMalloc / cudaFree / thrust::copy calls inside device_vector_holder<T>                   */
/* (cuda_icp/scene/common.cu:3-40, cuda_renderer/renderer.cu:15-50).                              */
int pr_device_malloc(void** ptr, size_t bytes);
int pr_device_free(void* ptr);
int pr_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, pr_stream_t stream);   /* synchronises */
int pr_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, pr_stream_t stream);   /* synchronises */
int pr_stream_synchronize(pr_stream_t stream);

/* ---------------------------------------------------------------------------------------- */
/* mesh ingestion: replaces cuda_renderer::Model::LoadModel (renderer.cpp:16-58, assimp).       */
/* Reads an ASCII or binary_little_endian PLY; writes triangles in face order as 9 floats each  */
/* (Model::Triangle, renderer.h:60-70).  Call with tris_host == NULL to get the count.          */
int pr_load_ply(const char* path, float* tris_host, size_t capacity_tris, size_t* n_tris);

/* compute_proj, cuda_renderer/renderer.cpp:161-185 (host, pure arithmetic). */
int pr_compute_proj(const float K[9], int width, int height, float near_plane, float far_plane, float proj[16]);

/* ---------------------------------------------------------------------------------------- */
/* rasteriser: replaces render_cuda_keep_in_gpu / render_cuda (renderer.cu:189-336) and the     */
/* render_triangle kernel (renderer.cu:83-187).                                                 */
/*   tris_dev   n_tris * 9 floats;  poses: n_poses row-major 4x4 (Model::mat4x4), host or device */
/*   out_depth_dev  n_poses * W' * H' int32 (W',H' = ROI size when a ROI is given), 0 = empty    */
/* Output equals render_cpu (renderer.cpp:259-298) bit for bit.                                 */
size_t pr_render_workspace_bytes(size_t n_poses, size_t n_tris, size_t width, size_t height);
int pr_render_batch(const float* tris_dev, size_t n_tris, const float* poses, int poses_on_device, size_t n_poses,
                    size_t width, size_t height, const float proj[16], pr_roi roi, int32_t* out_depth_dev,
                    void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* Indexed variant (no upstream counterpart: the reference renders a triangle soup).  Unique vertices are  */
/* projected once per pose, so a triangle's setup is three 16-byte loads instead of six transforms and six   */
/* divisions; the result is bit-identical to pr_render_batch.  pr_mesh_index (host) deduplicates a soup:      */
/* verts_out has room for 3*n_tris vertices (pass NULL to count), faces_out for 3*n_tris indices.            */
int pr_mesh_index(const float* tris_host, size_t n_tris, float* verts_out, int32_t* faces_out, size_t* n_verts);
size_t pr_render_indexed_workspace_bytes(size_t n_poses, size_t n_verts, size_t n_tris, size_t width, size_t height);
int pr_render_indexed_batch(const float* verts_dev, size_t n_verts, const int32_t* faces_dev, size_t n_tris,
                            const float* poses, int poses_on_device, size_t n_poses, size_t width, size_t height,
                            const float proj[16], pr_roi roi, int32_t* out_depth_dev,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* Fused render -> cloud: render_cuda_keep_in_gpu (renderer.cu:269-303) followed by depth2cloud_cuda           */
/* (icp.cu:256-286) for every pose, which is how the reference's own pipeline chains them (test.cpp:143-153).  */
/* The rasteriser's tile write-out counts the valid pixels, so the clouds are built from one more read of the   */
/* non-empty tiles only.  out_depth_dev is the same int32 batch pr_render_indexed_batch writes.  Cloud i is     */
/* out_pts_dev[3*offsets[i] .. 3*(offsets[i]+counts[i])): the same points depth2cloud_cuda produces, ordered     */
/* tile by tile (64x64-pixel screen tiles in row-major order, row-major inside a tile) instead of row-major      */
/* over the image.  counts / offsets / overflow / capacity_points / align_points as in pr_depth2cloud_count.     */
/* out_pts_dev == NULL: depth only (K, counts, offsets may then be NULL) -- pr_render_indexed_batch with clusters.  */
/* Optional acceleration structure for it (no upstream counterpart): pr_mesh_cluster (host) reorders the faces of an   */
/* indexed mesh along a Morton curve of their centroids, cuts them into clusters of 64 consecutive triangles and       */
/* lists every cluster's unique vertices.  With the lists on the device (pr_mesh_clusters) the rasteriser bins          */
/* clusters instead of triangles; the rendered depth is unchanged (depth is order independent).                         */
/*   faces_inout       n_tris * 3 indices, reordered in place                                                            */
/*   cluster_vert_off  room for (n_tris + 63) / 64 + 1 entries;  cluster_verts  room for 3 * n_tris entries              */
typedef struct pr_mesh_clusters {
    size_t n_clusters;
    const int32_t* vert_off_dev;   /* n_clusters + 1 */
    const int32_t* verts_dev;      /* vert_off[n_clusters] vertex ids */
} pr_mesh_clusters;
int pr_mesh_cluster(const float* verts_host, size_t n_verts, int32_t* faces_inout, size_t n_tris,
                    int32_t* cluster_vert_off, int32_t* cluster_verts, size_t* n_clusters);
size_t pr_render_cloud_workspace_bytes(size_t n_poses, size_t n_verts, size_t n_tris, size_t width, size_t height);
int pr_render_cloud_batch(const float* verts_dev, size_t n_verts, const int32_t* faces_dev, size_t n_tris,
                          const float* poses, int poses_on_device, size_t n_poses, size_t width, size_t height,
                          const float proj[16], const float K[9], int32_t* out_depth_dev,
                          float* out_pts_dev, size_t capacity_points, uint32_t align_points,
                          uint32_t* counts_dev, uint32_t* offsets_dev, uint32_t* overflow_dev,
                          const pr_mesh_clusters* clusters /* nullable; faces must then be pr_mesh_cluster's order */,
                          void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* PoseRenderer's outputs (pose_renderer.cpp:38-63: raw2depth_uint16 / raw2mask_uint8 / raw2depth_mask after render)   */
/* folded into the rasteriser's tile write-out: out_depth16 = uint16_t(depth) (truncation, renderer.cu:405), out_mask =   */
/* depth > 0 ? 255 : 0 (:406), n_poses * W' * H' each.  Any of the three outputs may be NULL (at least one is needed);     */
/* with out_depth_dev == NULL the int32 batch is never written.  Needs the tile path's workspace                          */
/* (pr_render_workspace_bytes).  Otherwise as pr_render_batch.                                                            */
int pr_render_outputs_batch(const float* tris_dev, size_t n_tris, const float* poses, int poses_on_device, size_t n_poses,
                            size_t width, size_t height, const float proj[16], pr_roi roi, int32_t* out_depth_dev,
                            uint16_t* out_depth16_dev, uint8_t* out_mask_dev,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* Parity entry point: the rasteriser divides with a refined reciprocal + residual correction instead of the       */
/* div.rn.f32 call (raster.cu, "IEEE division without the library call").  Compares n pseudo-random quotients      */
/* over the operand ranges the kernel guarantees with div.rn.f32; *mismatches_dev (uint64) must end up 0.          */
int pr_debug_div_check(uint64_t n, uint32_t seed, uint64_t* mismatches_dev, pr_stream_t stream);

/* raw2depth_uint16_cuda / raw2mask_uint8_cuda / raw2depth_mask_cuda (renderer.cu:338-439):      */
/* depth = uint16_t(raw), mask = raw > 0 ? 255 : 0.  Either output may be NULL.                  */
int pr_raw2depth_mask(const int32_t* raw_dev, size_t n, uint16_t* depth_dev, uint8_t* mask_dev, pr_stream_t stream);

/* ---------------------------------------------------------------------------------------- */
/* depth2cloud: replaces depth2cloud_cuda<T> (icp.cu:228-291) for a batch of n_images images.    */
/* Cloud i is out_pts_dev[3*offsets[i] .. 3*(offsets[i]+counts[i])), row-major valid-pixel order. */
/* Offsets are rounded up to a multiple of `align_points` (use 4 for 16-byte aligned clouds).    */
/* stride must be 1 (stride > 1 indexes out of bounds upstream, icp.cpp:77-82).                  */
/*   step 1 (count):  counts_dev[n_images], offsets_dev[n_images + 1] (last = total, padded).    */
/*                    capacity_points (0 = unlimited, in both steps; offsets are 32-bit, so 2^32-1 */
/*                    points bound a batch in any case): a cloud that would end beyond it is      */
/*                    emptied (counts[i] = 0) and *overflow_dev (nullable) is set to 1, so the    */
/*                    fill step and ICP stay in bounds without a host round trip.                 */
/*   step 2 (fill):   writes the points of every cloud that fits.                                 */
size_t pr_depth2cloud_workspace_bytes(size_t n_images, uint32_t width, uint32_t height);
int pr_depth2cloud_count(const void* depth_dev, int depth_is_int32, size_t n_images, uint32_t width, uint32_t height,
                         uint32_t stride, uint32_t align_points, size_t capacity_points, uint32_t* counts_dev,
                         uint32_t* offsets_dev, uint32_t* overflow_dev,
                         void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);
int pr_depth2cloud_fill(const void* depth_dev, int depth_is_int32, size_t n_images, uint32_t width, uint32_t height,
                        const float K[9], uint32_t stride, uint32_t tl_x, uint32_t tl_y,
                        const uint32_t* offsets_dev, float* out_pts_dev, size_t capacity_points,
                        const void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* ---------------------------------------------------------------------------------------- */
/* scene preparation (one-time per scene)                                                      */
/* Scene_projective::init_Scene_projective_cuda (depth_scene.cu:3-20 -> depth_scene.cpp:3-35):   */
/* organised cloud (dep2pcd, common.h:47-61) + LINEMOD-style normals (get_normal,               */
/* common.cpp:17-107), computed on the device from a DEVICE depth image (uint16 or int32, mm).  */
int pr_scene_projective_init(const void* depth_dev, int depth_is_int32, uint32_t width, uint32_t height,
                             const float K[9], float* pcd_dev, float* normal_dev, pr_stream_t stream);

/* Scene_nn::init_Scene_nn_cuda (pcd_scene.cu:3-20 -> pcd_scene.cpp:4-184): compacts the valid    */
/* pixels, builds the kd-tree on the HOST with the reference's level-by-level algorithm          */
/* (leaf <= max_leaf points) and returns host arrays the caller uploads.  Synchronous.           */
/*   depth_host: uint16 or int32; pcd/normal_host: capacity_points*3 floats; nodes_host:          */
/*   capacity_nodes nodes (2*n_points+1 always suffices).                                        */
int pr_scene_nn_build_host(const void* depth_host, int depth_is_int32, uint32_t width, uint32_t height,
                           const float K[9], int max_leaf, float* pcd_host, float* normal_host, size_t capacity_points,
                           pr_node_kdtree* nodes_host, size_t capacity_nodes, size_t* n_points, size_t* n_nodes);

/* The same build on the DEVICE (no upstream counterpart: init_Scene_nn_cuda builds on the host and uploads,       */
/* pcd_scene.cu:3-20; its README names a GPU build as future work).  depth_dev: uint16 or int32 image in device     */
/* memory.  Writes the leaf-ordered points / normals (capacity_points * 3 floats each) and the nodes                 */
/* (capacity_nodes; 2 * n_points + 1 always suffices) to DEVICE buffers -- bit-identical to pr_scene_nn_build_host.   */
/* Synchronous (one small device -> host read per tree level).  PR_ERR_CAPACITY when a buffer is too small            */
/* (*n_points / *n_nodes then hold the sizes needed so far).                                                          */
size_t pr_scene_nn_build_workspace_bytes(uint32_t width, uint32_t height);
int pr_scene_nn_build(const void* depth_dev, int depth_is_int32, uint32_t width, uint32_t height, const float K[9], int max_leaf,
                      float* pcd_dev, float* normal_dev, size_t capacity_points, pr_node_kdtree* nodes_dev, size_t capacity_nodes,
                      size_t* n_points, size_t* n_nodes, void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* ---------------------------------------------------------------------------------------- */
/* ICP: replaces ICP_Point2Plane_cuda<Scene> (icp.cu:156-223), thrust__pcd2Ab (icp.h:128-209),   */
/* Scene_*::query (depth_scene.h:30-48, pcd_scene.h:61-136), transform_pcd_cuda (icp.cu:142-153) */
/* and eigen_slover_666 (icp.cpp:29-45) for a ragged batch of hypotheses, entirely on device.    */
/*   pts_dev     packed Vec3f points; hypothesis h owns [offsets[h], offsets[h]+counts[h])        */
/*   results_dev n_hyp records; flags: PR_ICP_UPDATE_POINTS writes the refined points back        */
/*               (the reference mutates the model cloud in place, icp.cu:209).                    */
#define PR_ICP_UPDATE_POINTS 1
/* cross-check driver: one launch per pass, every operation in the reference's own order (IEEE divisions, the        */
/* reference's stackless kd-tree walk, Eigen's pivoted LDL^T).  Several times slower; same results within tolerance.  */
#define PR_ICP_REFERENCE_ARITHMETIC 2
/* scene_pixels: width*height of the projective scene the workspace will be used with; for Scene_nn pass   */
/* n_points + 2*n_nodes + 16 (room for the re-laid-out kd-tree; with less, the reference-layout walk is used). */
size_t pr_icp_workspace_bytes(size_t n_hyp, size_t capacity_points, size_t scene_pixels);
/* kd-tree scenes: the same plus (a) one int per model point, in which every pass leaves the index of its nearest       */
/* neighbour; the next pass starts its tree walk from that point (the pose moves little between passes, so the walk     */
/* only has to prove the candidate: 3-4x fewer node visits), and (b) room for a hash grid over the scene points (a      */
/* table of 2 n slots, 8 n points of 16 bytes, build scratch: ~200 bytes per scene point) that answers every query      */
/* nearer than half a grid cell to its neighbour without the tree.  The result is the same exact nearest neighbour      */
/* either way; pr_icp_nn_batch uses whichever of the two the workspace has room for (a workspace sized by               */
/* pr_icp_workspace_bytes still works, without either).  Asynchronous on `stream` (round 1 synchronised here).          */
size_t pr_icp_nn_workspace_bytes(size_t n_hyp, size_t capacity_points, size_t n_scene_points, size_t n_nodes);
int pr_icp_projective_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                            size_t capacity_points, const pr_scene_projective* scene, pr_icp_criteria criteria,
                            pr_registration_result* results_dev, int flags,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);
int pr_icp_nn_batch(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                    size_t capacity_points, const pr_scene_nn* scene, pr_icp_criteria criteria,
                    pr_registration_result* results_dev, int flags,
                    void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* The projective scene as the ICP kernel gathers it: one 32-byte record {q.xyz, n.xyz, 0, 0} per pixel (one L2 sector  */
/* per correspondence).  pr_icp_projective_batch re-packs the scene into its workspace on every call;                  */
/* a caller that refines many batches against one scene packs once (packed_dev: 32-byte aligned,                       */
/* pr_scene_projective_packed_bytes bytes) and calls the _packed variant (packed_dev == NULL: same as the plain call).  */
size_t pr_scene_projective_packed_bytes(uint32_t width, uint32_t height);
int pr_scene_projective_pack(const pr_scene_projective* scene, void* packed_dev, pr_stream_t stream);
int pr_icp_projective_batch_packed(float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                                   size_t capacity_points, const pr_scene_projective* scene, const void* packed_dev,
                                   pr_icp_criteria criteria, pr_registration_result* results_dev, int flags,
                                   void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* eigen_slover_666 (icp.cpp:29-45) on the host: A 6x6 symmetric (36 floats), b 6 -> row-major 4x4. */
int pr_solve_666(const float A[36], const float b[6], float T[16]);

/* one reduction pass of thrust__pcd2Ab over one cloud (icp.cu:170-172): out29_dev[29].  Debug /   */
/* parity entry point; synchronous-free.                                                         */
int pr_pcd2ab_projective(const float* pts_dev, size_t n, const pr_scene_projective* scene, float* out29_dev,
                         pr_stream_t stream);
int pr_pcd2ab_nn(const float* pts_dev, size_t n, const pr_scene_nn* scene, float* out29_dev, pr_stream_t stream);

/* Parity entry points onto the SHIPPED ICP kernel (the two above run the reference-arithmetic kernel).               */
/*   pr_pass_sums_*:  one evaluation pass (identity transform) of pr_icp_*_batch's own kernel over a ragged batch;     */
/*                    out32_dev[32*h .. 32*h+28] = the 29 sums of hypothesis h (icp.cu:170-172).  Arguments as         */
/*                    pr_icp_*_batch.                                                                                 */
/*   pr_correspondences_*: for every point, the scene index that kernel's search picks (projective: pixel u + v*W     */
/*                    after the depth gate, depth_scene.h:30-48; nn: index into the leaf-ordered scene points,        */
/*                    pcd_scene.h:61-136), -1 = no valid correspondence.  idx_dev: n int32.  workspace as pr_icp_*.    */
/*   pr_solve_666_device: n systems, one device thread each, S29_dev[29*i..] in thrust__pcd2Ab's order ->             */
/*                    E16_dev[16*i..]; fast = 1: the solver the kernel runs between passes, 0: Eigen's pivoted LDL^T.   */
int pr_pass_sums_projective(const float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                            size_t capacity_points, const pr_scene_projective* scene, float* out32_dev,
                            void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);
int pr_pass_sums_nn(const float* pts_dev, const uint32_t* offsets_dev, const uint32_t* counts_dev, size_t n_hyp,
                    size_t capacity_points, const pr_scene_nn* scene, float* out32_dev,
                    void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);
int pr_correspondences_projective(const float* pts_dev, size_t n, const pr_scene_projective* scene, int32_t* idx_dev,
                                  void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);
int pr_correspondences_nn(const float* pts_dev, size_t n, const pr_scene_nn* scene, int32_t* idx_dev,
                          void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);
int pr_solve_666_device(const float* S29_dev, size_t n, int fast, float* E16_dev, pr_stream_t stream);
/*   pr_nn_walk_stats: cost of the nearest-neighbour search over n queries, four uint64: stats_dev[0] = tree nodes fetched  */
/*                    (box tests) and [1] = leaf points tested by the unseeded tree walk; [2] = queries the hash grid      */
/*                    answers without the tree and [3] = points it tests for them.  For bench.py's C3 block.               */
int pr_nn_walk_stats(const float* pts_dev, size_t n, const pr_scene_nn* scene, uint64_t* stats2_dev,
                     void* workspace_dev, size_t workspace_bytes, pr_stream_t stream);

/* ---------------------------------------------------------------------------------------- */
/* pr_refiner: the whole path for one mesh + one scene behind a single call with HOST buffers -- */
/* what PoseRenderer (pose_renderer.h:9-32) plus the test.cpp:143-172 sequence do upstream:       */
/* render(poses) -> depth2cloud -> ICP_Point2Plane, for a batch of pose hypotheses.               */
typedef struct pr_refiner pr_refiner;

/* Uploads the mesh once (tris_host: n_tris*9 floats, mm). max_hyp bounds the batch size;          */
/* capacity_points bounds the total number of model points of a batch (0 = max_hyp * W*H/4).      */
/* A run whose clouds do not fit returns PR_ERR_CAPACITY (host variant) after emptying the clouds  */
/* that spilled.                                                                                 */
int pr_refiner_create(pr_refiner** out, const float* tris_host, size_t n_tris, uint32_t width, uint32_t height,
                      const float K[9], size_t max_hyp, size_t capacity_points);
void pr_refiner_destroy(pr_refiner* r);
/* Scene from a HOST depth image (uint16 or int32, mm), prepared on the device.  Synchronous (copy + preparation +   */
/* cudaStreamSynchronize on the default stream).  Scene buffers are allocated on first use and then kept.            */
int pr_refiner_set_scene_projective(pr_refiner* r, const void* depth_host, int depth_is_int32, float max_dist_diff);
int pr_refiner_set_scene_nn(pr_refiner* r, const void* depth_host, int depth_is_int32);
/* Scene from a DEVICE depth image (e.g. one that arrived by pr_broadcast_scene): no host copy.  The projective        */
/* variant is asynchronous on `stream`; the kd-tree variant synchronises (pr_scene_nn_build reads the level sizes).  */
int pr_refiner_set_scene_projective_device(pr_refiner* r, const void* depth_dev, int depth_is_int32, float max_dist_diff,
                                           pr_stream_t stream);
int pr_refiner_set_scene_nn_device(pr_refiner* r, const void* depth_dev, int depth_is_int32, pr_stream_t stream);
/* poses_host: n_hyp row-major 4x4 (model -> camera, mm).  results_host: n_hyp records.           */
/* Copies poses H2D, runs render -> cloud -> ICP on `stream`, copies results D2H, synchronises.    */
int pr_refiner_run(pr_refiner* r, const float* poses_host, size_t n_hyp, pr_icp_criteria criteria,
                   pr_registration_result* results_host, pr_stream_t stream);
/* Same with everything resident: poses_dev / results_dev on the device, no copies, no sync.     */
int pr_refiner_run_device(pr_refiner* r, const float* poses_dev, size_t n_hyp, pr_icp_criteria criteria,
                          pr_registration_result* results_dev, pr_stream_t stream);
/* Introspection for tests and bench: device pointers into the refiner's own buffers, valid until */
/* the next run: depth (n_hyp*W*H int32), points, offsets (n_hyp+1), counts (n_hyp).              */
int pr_refiner_buffers(pr_refiner* r, const int32_t** depth_dev, const float** pts_dev,
                       const uint32_t** offsets_dev, const uint32_t** counts_dev);
/* ... and into the prepared scene (organised cloud + normals, W*H Vec3f each; kd-tree scenes: the leaf-ordered points)  */
/* and the device copy of the results of the last pr_refiner_run (host variant).  Any pointer may be NULL.               */
int pr_refiner_scene_buffers(pr_refiner* r, const float** scene_pcd_dev, const float** scene_normal_dev,
                             const pr_registration_result** results_dev);
/* The device-resident run cannot return PR_ERR_CAPACITY (no host round trip): hypotheses whose clouds did not fit           */
/* capacity_points come back with fitness 0 and the identity transform, and *flag_dev (one uint32 on the device, valid after */
/* the run completes on its stream) is 1.  pr_refiner_run reads the same flag and returns PR_ERR_CAPACITY.                   */
int pr_refiner_overflow_flag(pr_refiner* r, const uint32_t** flag_dev);
/* Device time of the two stages of the runs since the last call (at most the last 256), from CUDA events the refiner      */
/* records on the caller's stream around the fused render -> cloud call and around the ICP call of every run: the mean per  */
/* run, in milliseconds, and how many runs it covers.  Synchronises with the last of them.  This is how bench.py measures   */
/* the ICP kernel INSIDE its timed steps.                                                                                    */
int pr_refiner_stage_ms(pr_refiner* r, float* render_cloud_ms, float* icp_ms, uint32_t* n_runs);
/* kernel launches issued by this library since load (all entry points), for bench accounting.    */
uint64_t pr_launch_count(void);

/* ---------------------------------------------------------------------------------------- */
/* Multi-GPU (no upstream counterpart: the reference is single-GPU, README.md:15 only suggests host threads).      */
/* The batch shards by hypothesis -- one process or host thread per GPU, no collective in the data path:            */
/*   pr_shard_plan       contiguous shards whose sizes differ by at most one: rank r owns [begin, begin + count)     */
/*   pr_broadcast_scene  the scene depth image (any device buffer) from `root` to every rank, in place               */
/*                       (ncclBroadcast on `stream`); each rank then prepares the scene on its own GPU               */
/*                       (pr_scene_projective_init / pr_refiner_set_scene_*_device), bit-identically                 */
/*   pr_gather_results   n_per_rank records of every rank -> all_dev[nranks * n_per_rank] on every rank              */
/*                       (ncclAllGather; pad the shards to the largest one).  Starts when the work queued on         */
/*                       `stream` so far is complete and runs on the communicator's own stream, so the caller's      */
/*                       next batch overlaps it; pr_gather_wait makes `stream` (host_sync = 0) or the host           */
/*                       (host_sync = 1) wait for it.                                                                */
/* NCCL is loaded at run time (dlopen libnccl.so.2); without it these return PR_ERR_COMM.  A communicator is made     */
/* from an ncclUniqueId (128 bytes, created on one rank by pr_nccl_unique_id and handed to the others by the host's   */
/* own means) or adopted from an existing ncclComm_t.                                                                 */
typedef struct pr_comm pr_comm;
int pr_device_count(int* count);
int pr_set_device(int device);
int pr_shard_plan(size_t n_hyp, int nranks, int rank, size_t* begin, size_t* count);
int pr_nccl_unique_id(void* id128);
int pr_comm_create(pr_comm** out, const void* id128, int nranks, int rank);     /* ncclCommInitRank on the current device */
int pr_comm_adopt(pr_comm** out, void* nccl_comm, int nranks, int rank);        /* an existing ncclComm_t; not destroyed  */
void pr_comm_destroy(pr_comm* comm);
int pr_broadcast_scene(pr_comm* comm, void* buf_dev, size_t bytes, int root, pr_stream_t stream);
int pr_gather_results(pr_comm* comm, const pr_registration_result* local_dev, size_t n_per_rank,
                      pr_registration_result* all_dev, pr_stream_t stream);
int pr_gather_wait(pr_comm* comm, pr_stream_t stream, int host_sync);

#ifdef __cplusplus
}
#endif
#endif /* POSE_REFINE_B200_H */
