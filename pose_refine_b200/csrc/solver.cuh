// solver.cuh -- the 6x6 damped normal-equation solve and pose update, host + device.
//
// Replaces eigen_slover_666 + TransformVector6dToMatrix4d (cuda_icp/icp.cpp:7-45), which the
// reference calls on the HOST from inside its CUDA iteration loop (icp.cu:207) -- the round trip
// that makes its loop latency-bound.  Here the same arithmetic runs on one device thread of the
// CTA that finishes a hypothesis' reduction.
//
// Arithmetic (all in double, cast to float at the end, as upstream):
//   (A + 0.01 I) x = b   by LDL^T with symmetric diagonal pivoting (Eigen's LDLT algorithm:
//   unblocked lower variant; the solve guards zero pivots), then
//   R = Rz(x2) Ry(x1) Rx(x0) composed through unit quaternions (Eigen's AngleAxis product ->
//   Quaternion::toRotationMatrix), t = (x3, x4, x5).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>

namespace prb {

template <class T> __host__ __device__ __forceinline__ void swap_vals(T& a, T& b) { T t = a; a = b; b = t; }

// x = (rx, ry, rz, tx, ty, tz) -> row-major 4x4 (TransformVector6dToMatrix4d, icp.cpp:7-17)
__host__ __device__ __forceinline__ void pose_from_x(const double* x, float* E) {
    // q = qz * qy * qx with q_axis(angle) = (cos(angle/2), sin(angle/2) * axis)
    const double cz = cos(0.5 * x[2]), sz = sin(0.5 * x[2]);
    const double cy = cos(0.5 * x[1]), sy = sin(0.5 * x[1]);
    const double cx = cos(0.5 * x[0]), sx = sin(0.5 * x[0]);
    // qz*qy (Hamilton product with the structural zeros kept as exact 0 terms dropped)
    const double aw = cz * cy, ax = -(sz * sy), ay = cz * sy, az = sz * cy;
    // (qz*qy) * qx, qx = (cx, sx, 0, 0)
    const double qw = aw * cx - ax * sx;
    const double qx = aw * sx + ax * cx;
    const double qy = ay * cx + az * sx;
    const double qz = az * cx - ay * sx;
    const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    E[0] = (float)(1.0 - (tyy + tzz)); E[1] = (float)(txy - twz); E[2] = (float)(txz + twy); E[3] = (float)x[3];
    E[4] = (float)(txy + twz); E[5] = (float)(1.0 - (txx + tzz)); E[6] = (float)(tyz - twx); E[7] = (float)x[4];
    E[8] = (float)(txz - twy); E[9] = (float)(tyz + twx); E[10] = (float)(1.0 - (txx + tyy)); E[11] = (float)x[5];
    E[12] = 0.f; E[13] = 0.f; E[14] = 0.f; E[15] = 1.f;
}

// A: 6x6 symmetric (any of row/column-major), b: 6.  E: row-major 4x4 (16 floats).
__host__ __device__ inline void solve_666(const float* A, const float* b, float* E) {
    const int n = 6;
    double a[6][6], x[6], tmp[6];
    int tr[6];
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) a[i][j] = (double)A[i + 6 * j] + (i == j ? 0.01 : 0.0);
        x[i] = (double)b[i];
    }
    for (int k = 0; k < n; k++) {
        int piv = k;
        double big = fabs(a[k][k]);
        for (int i = k + 1; i < n; i++) if (fabs(a[i][i]) > big) { big = fabs(a[i][i]); piv = i; }
        tr[k] = piv;
        if (piv != k) {
            for (int j = 0; j < k; j++) swap_vals(a[k][j], a[piv][j]);
            for (int i = piv + 1; i < n; i++) swap_vals(a[i][k], a[i][piv]);
            swap_vals(a[k][k], a[piv][piv]);
            for (int i = k + 1; i < piv; i++) swap_vals(a[i][k], a[piv][i]);
        }
        if (k > 0) {
            for (int j = 0; j < k; j++) tmp[j] = a[j][j] * a[k][j];
            double acc = 0.0;
            for (int j = 0; j < k; j++) acc += a[k][j] * tmp[j];
            a[k][k] -= acc;
            for (int i = k + 1; i < n; i++) {
                double acc2 = 0.0;
                for (int j = 0; j < k; j++) acc2 += a[i][j] * tmp[j];
                a[i][k] -= acc2;
            }
        }
        if (k + 1 < n && fabs(a[k][k]) > 0.0)
            for (int i = k + 1; i < n; i++) a[i][k] /= a[k][k];
    }
    for (int k = 0; k < n; k++) if (tr[k] != k) swap_vals(x[k], x[tr[k]]);
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) x[i] -= a[i][j] * x[j];
    for (int i = 0; i < n; i++) { if (fabs(a[i][i]) > DBL_MIN) x[i] /= a[i][i]; else x[i] = 0.0; }
    for (int i = n - 1; i >= 0; i--) for (int j = i + 1; j < n; j++) x[i] -= a[j][i] * x[j];
    for (int k = n - 1; k >= 0; k--) if (tr[k] != k) swap_vals(x[k], x[tr[k]]);

    pose_from_x(x, E);
}

// Same arithmetic as solve_666, operation for operation, with every array index a compile-time constant:
// the data-dependent pivot is applied as "if (piv == p) swap(...)" over the unrolled candidates, so the
// whole 6x6 factorisation lives in registers.  (solve_666's a[piv][j] indexing puts the matrix in local
// memory: measured ~21 us per solve on one B200 thread, 13 % of the ICP kernel's warp time.)
template <int K> struct PivotStep {
    __host__ __device__ static __forceinline__ void run(double (&a)[6][6], int (&tr)[6]) {
        int piv = K;
        double big = fabs(a[K][K]);
#pragma unroll
        for (int i = K + 1; i < 6; i++) if (fabs(a[i][i]) > big) { big = fabs(a[i][i]); piv = i; }
        tr[K] = piv;
#pragma unroll
        for (int p = K + 1; p < 6; p++) {
            if (piv == p) {
#pragma unroll
                for (int j = 0; j < K; j++) swap_vals(a[K][j], a[p][j]);
#pragma unroll
                for (int i = p + 1; i < 6; i++) swap_vals(a[i][K], a[i][p]);
                swap_vals(a[K][K], a[p][p]);
#pragma unroll
                for (int i = K + 1; i < p; i++) swap_vals(a[i][K], a[p][i]);
            }
        }
        if (K > 0) {
            double tmp[6];
#pragma unroll
            for (int j = 0; j < K; j++) tmp[j] = a[j][j] * a[K][j];
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < K; j++) acc += a[K][j] * tmp[j];
            a[K][K] -= acc;
#pragma unroll
            for (int i = K + 1; i < 6; i++) {
                double acc2 = 0.0;
#pragma unroll
                for (int j = 0; j < K; j++) acc2 += a[i][j] * tmp[j];
                a[i][K] -= acc2;
            }
        }
        if (K + 1 < 6 && fabs(a[K][K]) > 0.0) {
#pragma unroll
            for (int i = K + 1; i < 6; i++) a[i][K] /= a[K][K];
        }
    }
};

// S: the 29 sums in thrust__pcd2Ab's order (icp.h:165-206); unpacked as icp.cu:198-205 does.
__host__ __device__ __forceinline__ void solve_666_unrolled(const float (&S)[29], float (&E)[16]) {
    double a[6][6], x[6];
    int tr[6];
    {
        int shift = 0;
#pragma unroll
        for (int y = 0; y < 6; y++)
#pragma unroll
            for (int xx = y; xx < 6; xx++) {
                const double v = (double)S[shift++];
                a[xx][y] = v + (xx == y ? 0.01 : 0.0);
                a[y][xx] = a[xx][y];
            }
#pragma unroll
        for (int i = 0; i < 6; i++) x[i] = (double)S[21 + i];
    }
    PivotStep<0>::run(a, tr); PivotStep<1>::run(a, tr); PivotStep<2>::run(a, tr);
    PivotStep<3>::run(a, tr); PivotStep<4>::run(a, tr); PivotStep<5>::run(a, tr);
#pragma unroll
    for (int k = 0; k < 6; k++) {
#pragma unroll
        for (int p = k + 1; p < 6; p++) if (tr[k] == p) swap_vals(x[k], x[p]);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j < i; j++) x[i] -= a[i][j] * x[j];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) { if (fabs(a[i][i]) > DBL_MIN) x[i] /= a[i][i]; else x[i] = 0.0; }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
#pragma unroll
        for (int j = i + 1; j < 6; j++) x[i] -= a[j][i] * x[j];
    }
#pragma unroll
    for (int k = 5; k >= 0; k--) {
#pragma unroll
        for (int p = k + 1; p < 6; p++) if (tr[k] == p) swap_vals(x[k], x[p]);
    }
    pose_from_x(x, E);
}

// ---- the solver the hypothesis-resident ICP kernel runs between two passes ---------------------------
// The pass-to-pass chain of a hypothesis is serial (sums -> solve -> new 4x4 -> next pass), so what counts here
// is the LATENCY of one solve on one thread.  solve_666_unrolled reproduces Eigen's pivoted LDL^T operation for
// operation: 21 IEEE double divisions, each a call-sized sequence (~7 us per solve on a B200 thread).  (A + 0.01 I)
// is symmetric positive definite (A is a sum of J J^T), so LDL^T needs no pivoting for stability; this variant
// factors without it and multiplies by one reciprocal per column (hardware seed + three Newton steps).  The
// critical path is ~200 dependent double operations instead of ~1500.  The solution x differs from the pivoted
// one by O(cond * 1e-16), i.e. not at all after the cast to float except for an occasional last-bit flip
// (tests/test_gpu_parity.py::test_fast_solver_matches_exact pins the difference).
#ifdef __CUDACC__
__device__ __forceinline__ double rcp_newton(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));       // ~2^-23 relative
    double e = fma(-d, r, 1.0); r = fma(r, e, r);                // 2^-46
    e = fma(-d, r, 1.0); r = fma(r, e, r);                       // 2^-92 -> limited by double rounding
    e = fma(-d, r, 1.0); r = fma(r, e, r);                       // absorbs the rounding of the previous step
    return r;
}
__device__ __forceinline__ void solve_666_fast(const float (&S)[29], float (&E)[16]) {
    double a[6][6], x[6], dinv[6];
    {
        int shift = 0;
#pragma unroll
        for (int y = 0; y < 6; y++)
#pragma unroll
            for (int xx = y; xx < 6; xx++) a[xx][y] = (double)S[shift++] + (xx == y ? 0.01 : 0.0);   // lower triangle, icp.cu:198-205
#pragma unroll
        for (int i = 0; i < 6; i++) x[i] = (double)S[21 + i];
    }
    // A = L D L^T, L unit lower (stored in a[i][j], i > j), D in a[j][j]
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double v[6];
#pragma unroll
        for (int k = 0; k < j; k++) v[k] = a[j][k] * a[k][k];
        double d = a[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) d = fma(-a[j][k], v[k], d);
        a[j][j] = d;
        dinv[j] = rcp_newton(d);
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = a[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) s = fma(-a[i][k], v[k], s);
            a[i][j] = s * dinv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j < i; j++) x[i] = fma(-a[i][j], x[j], x[i]);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] *= dinv[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
#pragma unroll
        for (int j = i + 1; j < 6; j++) x[i] = fma(-a[j][i], x[j], x[i]);
    }
    pose_from_x(x, E);
}
#endif

}  // namespace prb
