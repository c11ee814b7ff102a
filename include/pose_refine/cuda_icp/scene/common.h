// scene/common.h -- device_vector_holder<T>, dep2pcd, pcd2dep (cuda_icp/scene/common.h).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../geometry.h"
#include "../../../pose_refine_b200.h"

namespace pose_refine {
inline void check(int status, const char* where) {
    if (status != PR_OK) throw std::runtime_error(std::string(where) + ": " + pr_error_string(status));
}
// A depth image without OpenCV: what the reference passes as cv::Mat (CV_16U or CV_32S, millimetres).
struct DepthImage {
    const void* data;
    int rows, cols;
    bool is_int32;
    DepthImage(const int32_t* d, int r, int c) : data(d), rows(r), cols(c), is_int32(true) {}
    DepthImage(const uint16_t* d, int r, int c) : data(d), rows(r), cols(c), is_int32(false) {}
};
}  // namespace pose_refine

// Owning device buffer (common.h:16-38).  Unlike upstream it is movable and non-copyable, so returning it
// by value does not depend on copy elision.
template <typename T> class device_vector_holder {
public:
    T* __gpu_memory = nullptr;
    size_t __size = 0;
    bool valid = false;
    device_vector_holder() {}
    explicit device_vector_holder(size_t n) { __malloc(n); }
    device_vector_holder(size_t n, T init) { __malloc(n); std::vector<T> h(n, init); pose_refine::check(pr_memcpy_h2d(__gpu_memory, h.data(), n * sizeof(T), nullptr), "h2d"); }
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
        pose_refine::check(pr_device_malloc(&p, n * sizeof(T)), "pr_device_malloc");
        __gpu_memory = static_cast<T*>(p); __size = n; valid = true;
    }
    void __free() { if (valid) { pr_device_free(__gpu_memory); valid = false; __size = 0; __gpu_memory = nullptr; } }
    void upload(const std::vector<T>& h) { __malloc(h.size()); pose_refine::check(pr_memcpy_h2d(__gpu_memory, h.data(), h.size() * sizeof(T), nullptr), "h2d"); }
    std::vector<T> download() const { std::vector<T> h(__size); pose_refine::check(pr_memcpy_d2h(h.data(), __gpu_memory, __size * sizeof(T), nullptr), "d2h"); return h; }
};

// dep2pcd / pcd2dep (common.h:47-73), host versions with the reference's operation order
template <class T> inline Vec3f dep2pcd(size_t x, size_t y, T dep, Mat3x3f& K, size_t tl_x = 0, size_t tl_y = 0) {
    if (dep == 0) return Vec3f(0, 0, 0);
    const float z = dep / 1000.0f;
    return Vec3f((x + tl_x - K[0][2]) / K[0][0] * z, (y + tl_y - K[1][2]) / K[1][1] * z, z);
}
inline Vec3i pcd2dep(const Vec3f& p, const Mat3x3f& K, size_t tl_x = 0, size_t tl_y = 0) {
    return Vec3i(int(p.x / p.z * K[0][0] + K[0][2] - tl_x + 0.5f), int(p.y / p.z * K[1][1] + K[1][2] - tl_y + 0.5f),
                 int(p.z * 1000.0f + 0.5f));
}
