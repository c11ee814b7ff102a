// geometry.h -- the small fixed-size vector / matrix types of the reference's boundary
// (cuda_icp/geometry.h): Vec3f (3 packed floats), Vec3i, Mat3x3f, Mat4x4f (row-major), vec<29,float>.
// Host-only here: the device code of libpose_refine_b200 works on plain float arrays with the same
// byte layout, so these types can be handed across the C ABI by pointer.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <ostream>

template <size_t DIM, typename T> struct vec {
    T data_[DIM];
    vec() { for (size_t i = 0; i < DIM; i++) data_[i] = T(); }
    T& operator[](size_t i) { assert(i < DIM); return data_[i]; }
    const T& operator[](size_t i) const { assert(i < DIM); return data_[i]; }
    vec& operator+=(const vec& o) { for (size_t i = 0; i < DIM; i++) data_[i] += o.data_[i]; return *this; }
    vec operator+(const vec& o) const { vec r(*this); r += o; return r; }
    static vec Zero() { return vec(); }
};
template <typename T> struct vec<2, T> {
    T x, y;
    vec() : x(T()), y(T()) {}
    vec(T X, T Y) : x(X), y(Y) {}
    T& operator[](size_t i) { assert(i < 2); return i == 0 ? x : y; }
    const T& operator[](size_t i) const { assert(i < 2); return i == 0 ? x : y; }
};
template <typename T> struct vec<3, T> {
    T x, y, z;
    vec() : x(T()), y(T()), z(T()) {}
    vec(T X, T Y, T Z) : x(X), y(Y), z(Z) {}
    T& operator[](size_t i) { assert(i < 3); return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](size_t i) const { assert(i < 3); return i == 0 ? x : (i == 1 ? y : z); }
    float norm() const { return std::sqrt(float(x * x + y * y + z * z)); }
};
template <size_t DIM, typename T> vec<DIM, T> operator-(vec<DIM, T> a, const vec<DIM, T>& b) {
    for (size_t i = 0; i < DIM; i++) a[i] -= b[i];
    return a;
}
template <size_t DIM, typename T> T operator*(const vec<DIM, T>& a, const vec<DIM, T>& b) {   // dot, summed from the last index down
    T r = T();
    for (size_t i = DIM; i--;) r += a[i] * b[i];
    return r;
}

template <size_t R, size_t C, typename T> class mat {
    vec<C, T> rows_[R];
public:
    mat() {}
    explicit mat(const T* d) { for (size_t i = 0; i < R; i++) for (size_t j = 0; j < C; j++) rows_[i][j] = d[j + i * C]; }
    vec<C, T>& operator[](size_t i) { assert(i < R); return rows_[i]; }
    const vec<C, T>& operator[](size_t i) const { assert(i < R); return rows_[i]; }
    vec<R, T> col(size_t j) const { vec<R, T> r; for (size_t i = 0; i < R; i++) r[i] = rows_[i][j]; return r; }
    static mat identity() { mat m; for (size_t i = 0; i < R; i++) for (size_t j = 0; j < C; j++) m[i][j] = T(i == j); return m; }
    mat<C, R, T> transpose() const { mat<C, R, T> t; for (size_t i = 0; i < R; i++) for (size_t j = 0; j < C; j++) t[j][i] = rows_[i][j]; return t; }
    // contiguous row-major storage (vec<C,T> is C packed Ts for C = 2, 3 and for the generic vec)
    const T* data() const { return &rows_[0][0]; }
    T* data() { return &rows_[0][0]; }
};
template <size_t R, size_t C, typename T> vec<R, T> operator*(const mat<R, C, T>& m, const vec<C, T>& v) {
    vec<R, T> r;
    for (size_t i = 0; i < R; i++) r[i] = m[i] * v;
    return r;
}
template <size_t R1, size_t C1, size_t C2, typename T> mat<R1, C2, T> operator*(const mat<R1, C1, T>& a, const mat<C1, C2, T>& b) {
    mat<R1, C2, T> r;
    for (size_t i = 0; i < R1; i++) for (size_t j = 0; j < C2; j++) r[i][j] = a[i] * b.col(j);
    return r;
}
template <size_t R, size_t C, typename T> std::ostream& operator<<(std::ostream& o, const mat<R, C, T>& m) {
    for (size_t i = 0; i < R; i++) { for (size_t j = 0; j < C; j++) o << m[i][j] << " "; o << "\n"; }
    return o;
}

typedef vec<2, float> Vec2f;
typedef vec<2, int> Vec2i;
typedef vec<3, float> Vec3f;
typedef vec<3, int> Vec3i;
typedef vec<4, float> Vec4f;
typedef mat<4, 4, float> Mat4x4f;
typedef mat<3, 3, float> Mat3x3f;
static_assert(sizeof(Vec3f) == 12 && sizeof(Mat4x4f) == 64 && sizeof(Mat3x3f) == 36, "boundary layouts (SURVEY.md App. C)");
