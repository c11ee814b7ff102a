// Scene_projective (cuda_icp/scene/depth_scene/depth_scene.h:7-48) over the C ABI.
#pragma once
#include <cstring>
#include "../common.h"

struct Scene_projective {
    size_t width = 640, height = 480;
    float max_dist_diff = 0.1f;   // m
    Mat3x3f K;
    Vec3f* pcd_ptr = nullptr;     // DEVICE pointers into caller-owned buffers, width*height Vec3f each
    Vec3f* normal_ptr = nullptr;

    // init_Scene_projective_cuda (depth_scene.cu:3-20): organised cloud + normals, computed on the device
    void init_Scene_projective_cuda(const pose_refine::DepthImage& scene_depth, Mat3x3f& scene_K,
                                    device_vector_holder<Vec3f>& pcd_buffer, device_vector_holder<Vec3f>& normal_buffer,
                                    size_t width_ = 640, size_t height_ = 480, float max_dist_diff_ = 0.1f) {
        assert((size_t)scene_depth.cols == width_ && (size_t)scene_depth.rows == height_);
        K = scene_K; width = width_; height = height_; max_dist_diff = max_dist_diff_;
        const size_t n = width * height;
        device_vector_holder<unsigned char> d(n * (scene_depth.is_int32 ? 4 : 2));
        pose_refine::check(pr_memcpy_h2d(d.data(), scene_depth.data, d.size(), nullptr), "h2d");
        pcd_buffer.__malloc(n); normal_buffer.__malloc(n);
        pose_refine::check(pr_scene_projective_init(d.data(), scene_depth.is_int32, (uint32_t)width, (uint32_t)height, K.data(),
                                                    reinterpret_cast<float*>(pcd_buffer.data()), reinterpret_cast<float*>(normal_buffer.data()), nullptr),
                           "pr_scene_projective_init");
        pose_refine::check(pr_stream_synchronize(nullptr), "sync");
        pcd_ptr = pcd_buffer.data(); normal_ptr = normal_buffer.data();
    }
    pr_scene_projective c_abi() const {
        pr_scene_projective s;
        s.width = width; s.height = height; s.max_dist_diff = max_dist_diff;
        std::memcpy(s.K, K.data(), 36);
        s.pcd_dev = reinterpret_cast<const float*>(pcd_ptr); s.normal_dev = reinterpret_cast<const float*>(normal_ptr);
        return s;
    }
};
