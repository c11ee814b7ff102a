// PoseRenderer (pose_renderer.h:9-32, pose_renderer.cpp) over the C ABI: loads the model, uploads the
// triangles once, renders batches of poses; down_sample renders at width/ds x height/ds with the
// FULL-resolution projection matrix, exactly as pose_renderer.cpp:25-36 does.
#pragma once
#include "cuda_renderer/renderer.h"
#include "cuda_icp/icp.h"

class PoseRenderer {
public:
    float K[9];
    int width = 0, height = 0;
    cuda_renderer::Model model;
    cuda_renderer::device_vector_holder<cuda_renderer::Model::Triangle> tris;
    cuda_renderer::Model::mat4x4 proj_mat;

    explicit PoseRenderer(const std::string& model_path) : model(model_path) { tris.upload(model.tris); }
    void set_K_width_height(const float* K_, int width_, int height_) {
        std::memcpy(K, K_, 36); width = width_; height = height_;
        proj_mat = cuda_renderer::compute_proj(K, width, height);
    }
    // init_poses: row-major 4x4 each.  pose_renderer.cpp:25-63: render at width/ds x height/ds with the FULL-resolution
    // projection matrix, then raw2depth_uint16 / raw2mask_uint8 / raw2depth_mask.  Here the rasteriser writes the
    // uint16 depth and the mask itself (pr_render_outputs_batch); the int32 batch never exists.
    std::vector<std::vector<uint16_t>> render_depth(const std::vector<cuda_renderer::Model::mat4x4>& init_poses, float down_sample = 1) {
        auto dm = render_what(init_poses, down_sample, true, false);
        std::vector<std::vector<uint16_t>> out(dm.size());
        for (size_t i = 0; i < dm.size(); i++) out[i] = std::move(dm[i].depth);
        return out;
    }
    std::vector<std::vector<uint8_t>> render_mask(const std::vector<cuda_renderer::Model::mat4x4>& init_poses, float down_sample = 1) {
        auto dm = render_what(init_poses, down_sample, false, true);
        std::vector<std::vector<uint8_t>> out(dm.size());
        for (size_t i = 0; i < dm.size(); i++) out[i] = std::move(dm[i].mask);
        return out;
    }
    // pose -> {uint16 depth, uint8 mask} (upstream: std::vector<std::vector<cv::Mat>>, pose_renderer.cpp:56-63)
    std::vector<cuda_renderer::DepthMask> render_depth_mask(const std::vector<cuda_renderer::Model::mat4x4>& init_poses, float down_sample = 1) {
        return render_what(init_poses, down_sample, true, true);
    }
    std::vector<cuda_renderer::DepthMask> render_what(const std::vector<cuda_renderer::Model::mat4x4>& init_poses, float down_sample,
                                                      bool want_depth, bool want_mask) {
        const int w = int(width / down_sample), h = int(height / down_sample);
        return cuda_renderer::render_depth_mask_cuda(tris, init_poses, (size_t)w, (size_t)h, proj_mat, want_depth, want_mask);
    }
};

// One-call refinement of a batch of pose hypotheses: what the reference's test.cpp does
// per hypothesis -- render_cuda_keep_in_gpu -> depth2cloud_cuda -> ICP_Point2Plane_cuda (test.cpp:143-172) -- for all
// hypotheses at once, over pr_refiner_* (mesh uploaded once, all buffers preallocated, one kernel chain per batch).
class PoseRefiner {
    pr_refiner* r_ = nullptr;
public:
    // tris in model units (mm), K row-major 3x3; max_hyp bounds the batch size
    PoseRefiner(const std::vector<cuda_renderer::Model::Triangle>& tris, int width, int height, const float* K, size_t max_hyp) {
        pose_refine::check(pr_refiner_create(&r_, reinterpret_cast<const float*>(tris.data()), tris.size(), (uint32_t)width, (uint32_t)height,
                                             K, max_hyp, 0), "pr_refiner_create");
    }
    ~PoseRefiner() { pr_refiner_destroy(r_); }
    PoseRefiner(const PoseRefiner&) = delete;
    PoseRefiner& operator=(const PoseRefiner&) = delete;
    // Scene_projective::init_Scene_projective_cuda / Scene_nn::init_Scene_nn_cuda of a host depth image (mm)
    void set_scene_projective(const pose_refine::DepthImage& depth, float max_dist_diff = 0.1f) {
        pose_refine::check(pr_refiner_set_scene_projective(r_, depth.data, depth.is_int32, max_dist_diff), "pr_refiner_set_scene_projective");
    }
    void set_scene_nn(const pose_refine::DepthImage& depth) {
        pose_refine::check(pr_refiner_set_scene_nn(r_, depth.data, depth.is_int32), "pr_refiner_set_scene_nn");
    }
    // poses: model -> camera, row-major 4x4 (mm).  result[i].transformation_ is the ICP update of hypothesis i in
    // metres (apply it to the cloud of pose i, as ICP_Point2Plane_cuda's result is used upstream).
    std::vector<cuda_icp::RegistrationResult> refine(const std::vector<cuda_renderer::Model::mat4x4>& poses,
                                                     const cuda_icp::ICPConvergenceCriteria& c = cuda_icp::ICPConvergenceCriteria()) {
        std::vector<cuda_icp::RegistrationResult> out(poses.size());
        if (poses.empty()) return out;
        const pr_icp_criteria crit = {c.relative_fitness_, c.relative_rmse_, c.max_iteration_};
        pose_refine::check(pr_refiner_run(r_, reinterpret_cast<const float*>(poses.data()), poses.size(), crit,
                                          reinterpret_cast<pr_registration_result*>(out.data()), nullptr), "pr_refiner_run");
        return out;
    }
};
