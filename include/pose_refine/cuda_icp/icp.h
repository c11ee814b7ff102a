// icp.h -- the reference's ICP interface (cuda_icp/icp.h) as inline wrappers over the C ABI.
// Same names, argument meaning and in-place semantics; CUDA_ON is implied (there is no CPU path here).
#pragma once
#include "geometry.h"
#include "scene/depth_scene/depth_scene.h"
#include "scene/pcd_scene/pcd_scene.h"

namespace cuda_icp {

using V3f_holder = device_vector_holder<Vec3f>;

struct RegistrationResult {   // icp.h:26-36, 72 bytes
    RegistrationResult(const Mat4x4f& t = Mat4x4f::identity()) : transformation_(t), inlier_rmse_(0.f), fitness_(0.f) {}
    Mat4x4f transformation_;
    float inlier_rmse_;
    float fitness_;
};
static_assert(sizeof(RegistrationResult) == sizeof(pr_registration_result), "RegistrationResult layout");

struct ICPConvergenceCriteria {   // icp.h:38-50
    ICPConvergenceCriteria(float relative_fitness = 1e-5f, float relative_rmse = 1e-5f, int max_iteration = 30)
        : relative_fitness_(relative_fitness), relative_rmse_(relative_rmse), max_iteration_(max_iteration) {}
    float relative_fitness_, relative_rmse_;
    int max_iteration_;
};

// eigen_slover_666 (icp.cpp:29-45)
inline Mat4x4f eigen_slover_666(float* A, float* b) {
    float T[16];
    pose_refine::check(pr_solve_666(A, b, T), "pr_solve_666");
    return Mat4x4f(T);
}

// depth2cloud_cuda<T> (icp.cu:256-286): depth is a DEVICE pointer (int32_t or uint16_t)
template <class T>
device_vector_holder<Vec3f> depth2cloud_cuda(T* depth, uint32_t width, uint32_t height, Mat3x3f& K, uint32_t stride = 1,
                                             uint32_t tl_x = 0, uint32_t tl_y = 0) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 2, "int32_t or uint16_t depth");
    const int is_i32 = sizeof(T) == 4;
    const size_t ws_bytes = pr_depth2cloud_workspace_bytes(1, width, height);
    device_vector_holder<unsigned char> ws(ws_bytes + 16);
    device_vector_holder<uint32_t> meta(4);   // counts[1], offsets[2]
    pose_refine::check(pr_depth2cloud_count(depth, is_i32, 1, width, height, stride, 1, 0, meta.data(), meta.data() + 1, nullptr,
                                            ws.data(), ws_bytes, nullptr), "pr_depth2cloud_count");
    const std::vector<uint32_t> m = meta.download();        // the same small D2H read upstream does (icp.cu:272-274)
    device_vector_holder<Vec3f> cloud(m[0]);
    pose_refine::check(pr_depth2cloud_fill(depth, is_i32, 1, width, height, K.data(), stride, tl_x, tl_y, meta.data() + 1,
                                           reinterpret_cast<float*>(cloud.data()), m[0], ws.data(), ws_bytes, nullptr), "pr_depth2cloud_fill");
    pose_refine::check(pr_stream_synchronize(nullptr), "sync");
    return cloud;
}

namespace detail {
inline RegistrationResult run(device_vector_holder<Vec3f>& model, const pr_scene_projective* sp, const pr_scene_nn* sn,
                              const ICPConvergenceCriteria& c) {
    const size_t n = model.size();
    const size_t ws_bytes = pr_icp_workspace_bytes(1, n, sp ? sp->width * sp->height : sn->n_points + 2 * sn->n_nodes + 16);
    device_vector_holder<unsigned char> ws(ws_bytes);
    device_vector_holder<uint32_t> meta;
    meta.upload(std::vector<uint32_t>{0u, (uint32_t)n});    // offsets[0] = 0, counts[0] = n
    device_vector_holder<pr_registration_result> res(1);
    const pr_icp_criteria crit = {c.relative_fitness_, c.relative_rmse_, c.max_iteration_};
    const int rc = sp ? pr_icp_projective_batch(reinterpret_cast<float*>(model.data()), meta.data(), meta.data() + 1, 1, n, sp, crit, res.data(),
                                                PR_ICP_UPDATE_POINTS, ws.data(), ws_bytes, nullptr)
                      : pr_icp_nn_batch(reinterpret_cast<float*>(model.data()), meta.data(), meta.data() + 1, 1, n, sn, crit, res.data(),
                                        PR_ICP_UPDATE_POINTS, ws.data(), ws_bytes, nullptr);
    pose_refine::check(rc, "pr_icp_batch");
    const pr_registration_result r = res.download()[0];
    RegistrationResult out{Mat4x4f(r.transformation)};
    out.inlier_rmse_ = r.inlier_rmse; out.fitness_ = r.fitness;
    return out;
}
}  // namespace detail

// ICP_Point2Plane_cuda<Scene> (icp.cu:156-217).  model_pcd is transformed IN PLACE, as upstream.
inline RegistrationResult ICP_Point2Plane_cuda(device_vector_holder<Vec3f>& model_pcd, const Scene_projective scene,
                                               const ICPConvergenceCriteria criteria = ICPConvergenceCriteria()) {
    const pr_scene_projective s = scene.c_abi();
    return detail::run(model_pcd, &s, nullptr, criteria);
}
inline RegistrationResult ICP_Point2Plane_cuda(device_vector_holder<Vec3f>& model_pcd, const Scene_nn scene,
                                               const ICPConvergenceCriteria criteria = ICPConvergenceCriteria()) {
    const pr_scene_nn s = scene.c_abi();
    return detail::run(model_pcd, nullptr, &s, criteria);
}

// the compile-time dispatchers of icp.h:102-120 (CUDA_ON branch)
template <typename... Params> V3f_holder depth2cloud(Params&&... params) { return depth2cloud_cuda(std::forward<Params>(params)...); }
template <typename... Params> RegistrationResult ICP_Point2Plane(Params&&... params) { return ICP_Point2Plane_cuda(std::forward<Params>(params)...); }

}  // namespace cuda_icp
