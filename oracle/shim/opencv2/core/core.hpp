// Minimal stand-in for the slice of OpenCV's cv::Mat that the reference's hot-path
// sources touch (cuda_icp/scene/*.cpp, cuda_renderer/renderer.{h,cpp}).
// TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the reference's own .cpp
// files, where they lie under /root/reference, into oracle/_ref/ without OpenCV.
#pragma once
#include <cstdint>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cassert>
#include <cstring>
#include <numeric>
#include <algorithm>
#include <vector>
#include <memory>
#include <iostream>
#include <string>

typedef unsigned short ushort;
typedef unsigned char uchar;

#define CV_8U 0
#define CV_16U 2
#define CV_32S 4
#define CV_32F 5
#define CV_8UC1 CV_8U
#define CV_32SC1 CV_32S

namespace cv {

struct Scalar {
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};

inline size_t shim_elem_size(int type) {
    switch (type) {
    case CV_8U: return 1;
    case CV_16U: return 2;
    case CV_32S: return 4;
    case CV_32F: return 4;
    }
    assert(false && "shim cv::Mat: unsupported type");
    return 0;
}

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;

    Mat() {}
    Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data((uchar*)ext), type_(type) {}
    Mat(int r, int c, int type, const Scalar& s = Scalar()) { create(r, c, type); fill(s.v[0]); }

    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        owned_ = std::make_shared<std::vector<uchar>>(size_t(r) * c * shim_elem_size(type));
        data = owned_->data();
    }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    bool isContinuous() const { return true; }

    template <class T> T& at(int r, int c) { return ((T*)data)[size_t(r) * cols + c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)data)[size_t(r) * cols + c]; }
    template <class T> T* ptr(int r = 0) { return ((T*)data) + size_t(r) * cols; }
    template <class T> const T* ptr(int r = 0) const { return ((const T*)data) + size_t(r) * cols; }

    // only the conversions the reference performs: CV_32S -> CV_16U (saturate_cast<ushort>)
    void convertTo(Mat& dst, int rtype) const {
        Mat out(rows, cols, rtype, Scalar(0));
        const size_t n = size_t(rows) * cols;
        if (type_ == CV_32S && rtype == CV_16U) {
            const int32_t* s = (const int32_t*)data; uint16_t* d = (uint16_t*)out.data;
            for (size_t i = 0; i < n; i++) d[i] = (uint16_t)(s[i] < 0 ? 0 : (s[i] > 65535 ? 65535 : s[i]));
        } else if (type_ == rtype) {
            std::memcpy(out.data, data, n * shim_elem_size(rtype));
        } else {
            assert(false && "shim cv::Mat::convertTo: unsupported conversion");
        }
        dst = out;
    }

private:
    int type_ = CV_8U;
    std::shared_ptr<std::vector<uchar>> owned_;
    void fill(double v) {
        const size_t n = size_t(rows) * cols;
        switch (type_) {
        case CV_8U: std::fill((uchar*)data, (uchar*)data + n, (uchar)v); break;
        case CV_16U: std::fill((uint16_t*)data, (uint16_t*)data + n, (uint16_t)v); break;
        case CV_32S: std::fill((int32_t*)data, (int32_t*)data + n, (int32_t)v); break;
        case CV_32F: std::fill((float*)data, (float*)data + n, (float)v); break;
        }
    }
};

}  // namespace cv
