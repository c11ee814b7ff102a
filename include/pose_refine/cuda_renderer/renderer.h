// renderer.h -- the reference's renderer interface (cuda_renderer/renderer.h) as inline wrappers over
// the C ABI.  Same names and argument meaning; CUDA_ON is implied.  cv::Mat results become plain
// std::vector images (define POSE_REFINE_WITH_OPENCV before including to also get cv::Mat overloads).
#pragma once
#include <cstdint>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../../pose_refine_b200.h"

namespace cuda_renderer {

inline void check(int status, const char* where) {
    if (status != PR_OK) throw std::runtime_error(std::string(where) + ": " + pr_error_string(status));
}

class Model {   // renderer.h:27-155 (the parts the rendering path consumes)
public:
    struct int3 { int v0, v1, v2; };
    struct ROI { int x, y, width, height; };
    struct float3 { float x, y, z; };
    struct Triangle { float3 v0, v1, v2; };
    struct mat4x4 {   // row-major, renderer.h:71-141
        float a0 = 1, a1 = 0, a2 = 0, a3 = 0;
        float b0 = 0, b1 = 1, b2 = 0, b3 = 0;
        float c0 = 0, c1 = 0, c2 = 1, c3 = 0;
        float d0 = 0, d1 = 0, d2 = 0, d3 = 1;
        void t() {
            std::swap(a1, b0); std::swap(a2, c0); std::swap(a3, d0);
            std::swap(b2, c1); std::swap(b3, d1); std::swap(c3, d2);
        }
        void init_from_ptr(const float* m) { std::memcpy(this, m, 64); }
        void init_from_ptr(const float* R, const float* t) {
            a0 = R[0]; a1 = R[1]; a2 = R[2]; a3 = t[0];
            b0 = R[3]; b1 = R[4]; b2 = R[5]; b3 = t[1];
            c0 = R[6]; c1 = R[7]; c2 = R[8]; c3 = t[2];
        }
    };
    static_assert(sizeof(Triangle) == 36 && sizeof(mat4x4) == 64, "boundary layouts");

    std::vector<Triangle> tris;
    float3 bbox_min{0, 0, 0}, bbox_max{0, 0, 0};

    Model() {}
    explicit Model(const std::string& fileName) { LoadModel(fileName); }
    // LoadModel (renderer.cpp:16-58) without assimp: ASCII / binary little-endian PLY
    void LoadModel(const std::string& fileName) {
        size_t n = 0;
        check(pr_load_ply(fileName.c_str(), nullptr, 0, &n), "pr_load_ply");
        tris.resize(n);
        check(pr_load_ply(fileName.c_str(), reinterpret_cast<float*>(tris.data()), n, &n), "pr_load_ply");
        bbox_min = {1e10f, 1e10f, 1e10f}; bbox_max = {-1e10f, -1e10f, -1e10f};
        for (const Triangle& t : tris)
            for (const float3* v : {&t.v0, &t.v1, &t.v2}) {
                bbox_min.x = std::min(bbox_min.x, v->x); bbox_min.y = std::min(bbox_min.y, v->y); bbox_min.z = std::min(bbox_min.z, v->z);
                bbox_max.x = std::max(bbox_max.x, v->x); bbox_max.y = std::max(bbox_max.y, v->y); bbox_max.z = std::max(bbox_max.z, v->z);
            }
    }
};

// device_vector_holder (renderer.h:161-183), movable
template <typename T> class device_vector_holder {
public:
    T* __gpu_memory = nullptr;
    size_t __size = 0;
    bool valid = false;
    device_vector_holder() {}
    explicit device_vector_holder(size_t n) { __malloc(n); }
    device_vector_holder(const device_vector_holder&) = delete;
    device_vector_holder& operator=(const device_vector_holder&) = delete;
    device_vector_holder(device_vector_holder&& o) noexcept { *this = std::move(o); }
    device_vector_holder& operator=(device_vector_holder&& o) noexcept {
        if (this != &o) { __free(); __gpu_memory = o.__gpu_memory; __size = o.__size; valid = o.valid; o.__gpu_memory = nullptr; o.__size = 0; o.valid = false; }
        return *this;
    }
    ~device_vector_holder() { __free(); }
    T* data() { return __gpu_memory; }
    T* begin() { return __gpu_memory; }
    T* end() { return __gpu_memory + __size; }
    size_t size() const { return __size; }
    void __malloc(size_t n) {
        if (valid) __free();
        void* p = nullptr;
        check(pr_device_malloc(&p, n * sizeof(T)), "pr_device_malloc");
        __gpu_memory = static_cast<T*>(p); __size = n; valid = true;
    }
    void __free() { if (valid) { pr_device_free(__gpu_memory); valid = false; __size = 0; __gpu_memory = nullptr; } }
    void upload(const std::vector<T>& h) { __malloc(h.size()); check(pr_memcpy_h2d(__gpu_memory, h.data(), h.size() * sizeof(T), nullptr), "h2d"); }
    std::vector<T> download() const { std::vector<T> h(__size); check(pr_memcpy_d2h(h.data(), __gpu_memory, __size * sizeof(T), nullptr), "d2h"); return h; }
};
using Int_holder = device_vector_holder<int>;

// compute_proj (renderer.cpp:161-185); K: 9 floats row-major
inline Model::mat4x4 compute_proj(const float* K, int width, int height, float near_plane = 10, float far_plane = 10000) {
    Model::mat4x4 p;
    check(pr_compute_proj(K, width, height, near_plane, far_plane, reinterpret_cast<float*>(&p)), "pr_compute_proj");
    return p;
}

// render_cuda_keep_in_gpu (renderer.cu:269-336)
inline device_vector_holder<int> render_cuda_keep_in_gpu(device_vector_holder<Model::Triangle>& tris, const std::vector<Model::mat4x4>& poses,
                                                         size_t width, size_t height, const Model::mat4x4& proj_mat,
                                                         const Model::ROI roi = {0, 0, 0, 0}) {
    const bool has_roi = roi.width > 0 && roi.height > 0;
    const size_t rw = has_roi ? roi.width : width, rh = has_roi ? roi.height : height;
    device_vector_holder<int> depth(poses.size() * rw * rh);
    const size_t ws_bytes = pr_render_workspace_bytes(poses.size(), tris.size(), width, height);
    device_vector_holder<unsigned char> ws(ws_bytes);
    const pr_roi r = {roi.x, roi.y, roi.width, roi.height};
    check(pr_render_batch(reinterpret_cast<const float*>(tris.data()), tris.size(), reinterpret_cast<const float*>(poses.data()), 0, poses.size(),
                          width, height, reinterpret_cast<const float*>(&proj_mat), r, depth.data(), ws.data(), ws_bytes, nullptr), "pr_render_batch");
    check(pr_stream_synchronize(nullptr), "sync");   // synchronous on return, like upstream (renderer.cu:295)
    return depth;
}
inline device_vector_holder<int> render_cuda_keep_in_gpu(const std::vector<Model::Triangle>& tris, const std::vector<Model::mat4x4>& poses,
                                                         size_t width, size_t height, const Model::mat4x4& proj_mat,
                                                         const Model::ROI roi = {0, 0, 0, 0}) {
    device_vector_holder<Model::Triangle> d;
    d.upload(tris);
    return render_cuda_keep_in_gpu(d, poses, width, height, proj_mat, roi);
}
// render_cuda (renderer.cu:189-267): result on the host
template <class Tris>
std::vector<int32_t> render_cuda(Tris& tris, const std::vector<Model::mat4x4>& poses, size_t width, size_t height,
                                 const Model::mat4x4& proj_mat, const Model::ROI roi = {0, 0, 0, 0}) {
    return render_cuda_keep_in_gpu(tris, poses, width, height, proj_mat, roi).download();
}

// raw2depth_uint16_cuda / raw2mask_uint8_cuda / raw2depth_mask_cuda (renderer.cu:354-439): one image per pose
inline std::vector<std::vector<uint16_t>> raw2depth_uint16_cuda(device_vector_holder<int>& raw, size_t width, size_t height, size_t pose_size) {
    if (raw.size() != width * height * pose_size) throw std::invalid_argument("raw2depth_uint16_cuda: size mismatch");
    device_vector_holder<uint16_t> d(raw.size());
    check(pr_raw2depth_mask(raw.data(), raw.size(), d.data(), nullptr, nullptr), "pr_raw2depth_mask");
    const std::vector<uint16_t> all = d.download();
    std::vector<std::vector<uint16_t>> out(pose_size);
    for (size_t i = 0; i < pose_size; i++) out[i].assign(all.begin() + i * width * height, all.begin() + (i + 1) * width * height);
    return out;
}
inline std::vector<std::vector<uint8_t>> raw2mask_uint8_cuda(device_vector_holder<int>& raw, size_t width, size_t height, size_t pose_size) {
    if (raw.size() != width * height * pose_size) throw std::invalid_argument("raw2mask_uint8_cuda: size mismatch");
    device_vector_holder<uint8_t> m(raw.size());
    check(pr_raw2depth_mask(raw.data(), raw.size(), nullptr, m.data(), nullptr), "pr_raw2depth_mask");
    const std::vector<uint8_t> all = m.download();
    std::vector<std::vector<uint8_t>> out(pose_size);
    for (size_t i = 0; i < pose_size; i++) out[i].assign(all.begin() + i * width * height, all.begin() + (i + 1) * width * height);
    return out;
}

// raw2depth_mask_cuda (renderer.cu:409-439): upstream returns, per pose, a pair of cv::Mat {CV_16U depth, CV_8U mask};
// without OpenCV the pair is this struct.  depth = uint16_t(raw) (truncation, renderer.cu:405), mask = raw > 0 ? 255 : 0.
struct DepthMask {
    std::vector<uint16_t> depth;
    std::vector<uint8_t> mask;
};
inline std::vector<DepthMask> raw2depth_mask_cuda(device_vector_holder<int>& raw, size_t width, size_t height, size_t pose_size) {
    if (raw.size() != width * height * pose_size) throw std::invalid_argument("raw2depth_mask_cuda: size mismatch");
    device_vector_holder<uint16_t> d(raw.size());
    device_vector_holder<uint8_t> m(raw.size());
    check(pr_raw2depth_mask(raw.data(), raw.size(), d.data(), m.data(), nullptr), "pr_raw2depth_mask");
    const std::vector<uint16_t> dall = d.download();
    const std::vector<uint8_t> mall = m.download();
    std::vector<DepthMask> out(pose_size);
    const size_t step = width * height;
    for (size_t i = 0; i < pose_size; i++) {
        out[i].depth.assign(dall.begin() + i * step, dall.begin() + (i + 1) * step);
        out[i].mask.assign(mall.begin() + i * step, mall.begin() + (i + 1) * step);
    }
    return out;
}

// The same two images straight from the rasteriser (pr_render_outputs_batch): the tile write-out stores uint16 depth and
// the mask itself, the int32 batch is never materialised.  What PoseRenderer::render_* use.
inline std::vector<DepthMask> render_depth_mask_cuda(device_vector_holder<Model::Triangle>& tris, const std::vector<Model::mat4x4>& poses,
                                                     size_t width, size_t height, const Model::mat4x4& proj_mat,
                                                     bool want_depth = true, bool want_mask = true) {
    const size_t n = poses.size() * width * height;
    device_vector_holder<uint16_t> d;
    device_vector_holder<uint8_t> m;
    if (want_depth) d.__malloc(n);
    if (want_mask) m.__malloc(n);
    const size_t ws_bytes = pr_render_workspace_bytes(poses.size(), tris.size(), width, height);
    device_vector_holder<unsigned char> ws(ws_bytes);
    const pr_roi none = {0, 0, 0, 0};
    check(pr_render_outputs_batch(reinterpret_cast<const float*>(tris.data()), tris.size(), reinterpret_cast<const float*>(poses.data()), 0,
                                  poses.size(), width, height, reinterpret_cast<const float*>(&proj_mat), none, nullptr,
                                  want_depth ? d.data() : nullptr, want_mask ? m.data() : nullptr, ws.data(), ws_bytes, nullptr),
          "pr_render_outputs_batch");
    check(pr_stream_synchronize(nullptr), "sync");
    std::vector<DepthMask> out(poses.size());
    const size_t step = width * height;
    if (want_depth) { const std::vector<uint16_t> all = d.download(); for (size_t i = 0; i < out.size(); i++) out[i].depth.assign(all.begin() + i * step, all.begin() + (i + 1) * step); }
    if (want_mask) { const std::vector<uint8_t> all = m.download(); for (size_t i = 0; i < out.size(); i++) out[i].mask.assign(all.begin() + i * step, all.begin() + (i + 1) * step); }
    return out;
}

// the dispatchers of renderer.h:230-248 (CUDA_ON branch)
template <typename... Params> Int_holder render(Params&&... params) { return render_cuda_keep_in_gpu(std::forward<Params>(params)...); }
template <typename... Params> std::vector<int32_t> render_host(Params&&... params) { return render_cuda(std::forward<Params>(params)...); }

}  // namespace cuda_renderer
