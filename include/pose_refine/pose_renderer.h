// PoseRenderer (pose_renderer.h:9-32, pose_renderer.cpp) over the C ABI: loads the model, uploads the
// triangles once, renders batches of poses; down_sample renders at width/ds x height/ds with the
// FULL-resolution projection matrix, exactly as pose_renderer.cpp:25-36 does.
#pragma once
#include "cuda_renderer/renderer.h"

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
    // init_poses: row-major 4x4 each
    std::vector<std::vector<uint16_t>> render_depth(const std::vector<cuda_renderer::Model::mat4x4>& init_poses, float down_sample = 1) {
        const int w = int(width / down_sample), h = int(height / down_sample);
        auto raw = cuda_renderer::render(tris, init_poses, (size_t)w, (size_t)h, proj_mat);
        return cuda_renderer::raw2depth_uint16_cuda(raw, w, h, init_poses.size());
    }
    std::vector<std::vector<uint8_t>> render_mask(const std::vector<cuda_renderer::Model::mat4x4>& init_poses, float down_sample = 1) {
        const int w = int(width / down_sample), h = int(height / down_sample);
        auto raw = cuda_renderer::render(tris, init_poses, (size_t)w, (size_t)h, proj_mat);
        return cuda_renderer::raw2mask_uint8_cuda(raw, w, h, init_poses.size());
    }
};
