// oracle/ref_cuda_glue.cu -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.
//
// extern "C" doorway onto the reference's OWN CUDA path, compiled verbatim for sm_100 from /root/reference by
// `make -C oracle refcuda` into oracle/_ref/libpose_refine_refcuda.so (renderer.cu, icp.cu, scene/*.cu and the
// host .cpp files with -DCUDA_ON, exactly the reference's USE_CUDA build; OpenCV / assimp / Eigen are the stand-ins
// of oracle/shim/).  It runs the hot path the way the reference's test.cpp does (test.cpp:143-172):
//     render_cuda_keep_in_gpu (one call, P poses)  ->  P x depth2cloud_cuda  ->  P x ICP_Point2Plane_cuda
// serially or from `threads` host threads (README.md:15: "call it from many host threads", per-thread default
// stream, cuda_icp/CMakeLists.txt:11).  This is "the reference CUDA build" north_star's 10x target refers to;
// scripts/time_ref_cuda.py times it on the GPU box next to our path.  Nothing in the product links this.
#include "cuda_renderer/renderer.h"
#include "cuda_icp/icp.h"

#include <chrono>
#include <cstring>
#include <vector>
#include <omp.h>
#include <cuda_runtime.h>

using cuda_renderer::Model;

namespace {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C" {

// tris: T*9 floats; poses: P*16 row-major; K: 9; proj: 16; scene_depth: W*H int32 (host).
// results: P*18 floats (4x4, rmse, fitness).  seconds[0..2] = render, depth2cloud, icp wall time (device synchronised).
// Returns 0, or a CUDA error code.
int refcuda_pipeline(const float* tris, size_t T, const float* poses, size_t P, int W, int H, const float* K,
                     const float* proj, const int32_t* scene_depth, float rel_fit, float rel_rmse, int max_iter,
                     int threads, float* results, double* seconds, long* n_points_total) {
    std::vector<Model::Triangle> tv(T);
    std::memcpy(tv.data(), tris, T * sizeof(Model::Triangle));
    std::vector<Model::mat4x4> pv(P);
    for (size_t i = 0; i < P; i++) pv[i].init_from_ptr(poses + 16 * i);
    Model::mat4x4 pm; pm.init_from_ptr(proj);
    Mat3x3f Km(K);

    // scene: Scene_projective::init_Scene_projective_cuda (depth_scene.cu:3-20) -- one-time, not timed
    cv::Mat depth_mat(H, W, CV_32S, const_cast<int32_t*>(scene_depth));
    device_vector_holder<Vec3f> pcd_buffer, normal_buffer;
    Scene_projective scene;
    scene.init_Scene_projective_cuda(depth_mat, Km, pcd_buffer, normal_buffer, (size_t)W, (size_t)H);
    cudaDeviceSynchronize();

    const cuda_icp::ICPConvergenceCriteria crit(rel_fit, rel_rmse, max_iter);
    double t0 = now_s();
    auto depth = cuda_renderer::render_cuda_keep_in_gpu(tv, pv, (size_t)W, (size_t)H, pm);     // renderer.cu:269-303
    cudaDeviceSynchronize();
    double t1 = now_s();

    long total = 0;
    double t_cloud = 0.0, t_icp = 0.0;
    if (threads <= 1) {
        for (size_t i = 0; i < P; i++) {
            double a = now_s();
            auto cloud = cuda_icp::depth2cloud_cuda(depth.data() + i * (size_t)W * H, (uint32_t)W, (uint32_t)H, Km);   // icp.cu:256-286
            cudaStreamSynchronize(cudaStreamPerThread);
            double b = now_s();
            auto res = cuda_icp::ICP_Point2Plane_cuda(cloud, scene, crit);                                               // icp.cu:156-217
            double c = now_s();
            t_cloud += b - a; t_icp += c - b;
            total += (long)cloud.size();
            for (int r = 0; r < 4; r++) for (int cc = 0; cc < 4; cc++) results[18 * i + 4 * r + cc] = res.transformation_[r][cc];
            results[18 * i + 16] = res.inlier_rmse_; results[18 * i + 17] = res.fitness_;
        }
    } else {
        double a = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : total)
        for (long i = 0; i < (long)P; i++) {
            auto cloud = cuda_icp::depth2cloud_cuda(depth.data() + (size_t)i * W * H, (uint32_t)W, (uint32_t)H, Km);
            auto res = cuda_icp::ICP_Point2Plane_cuda(cloud, scene, crit);
            total += (long)cloud.size();
            for (int r = 0; r < 4; r++) for (int cc = 0; cc < 4; cc++) results[18 * i + 4 * r + cc] = res.transformation_[r][cc];
            results[18 * i + 16] = res.inlier_rmse_; results[18 * i + 17] = res.fitness_;
        }
        cudaDeviceSynchronize();
        t_icp = now_s() - a;        // depth2cloud + ICP together
    }
    seconds[0] = t1 - t0; seconds[1] = t_cloud; seconds[2] = t_icp;
    if (n_points_total) *n_points_total = total;
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

}  // extern "C"
