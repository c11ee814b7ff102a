// The reference's end-to-end scenario (test.cpp:22-193, CUDA half) written against the drop-in headers:
// same calls, same names.  Prints the final transform; exit code 0 when fitness looks sane.
#include <cstdio>
#include <cmath>
#include "pose_refine/cuda_renderer/renderer.h"
#include "pose_refine/cuda_icp/icp.h"
#include "pose_refine/pose_renderer.h"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s model.ply [proj]\n", argv[0]); return 2; }
    const bool use_proj = argc > 2;
    const int width = 640, height = 480;
    cuda_renderer::Model model(argv[1]);
    float K[9] = {572.4114f, 0.f, 325.2611f, 0.f, 573.57043f, 242.04899f, 0.f, 0.f, 1.f};
    auto proj = cuda_renderer::compute_proj(K, width, height);
    const float R_ren[9] = {0.34768538f, 0.93761126f, 0.f, 0.70540612f, -0.26157897f, -0.65877056f, -0.61767070f, 0.22904489f, -0.75234390f};
    const float t_ren[3] = {0.f, 0.f, 300.f}, t_ren2[3] = {20.f, 20.f, 320.f};
    const float a = 10.0f / 180.0f * 3.14f, c = std::cos(a), s = std::sin(a);
    const float Rx[9] = {1, 0, 0, 0, c, -s, 0, s, c}, Ry[9] = {c, 0, s, 0, 1, 0, -s, 0, c}, Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    auto mul = [](const float* A, const float* B, float* C) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float v = 0; for (int k = 0; k < 3; k++) v += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = v; }
    };
    float t1[9], rot[9], R2[9];
    mul(Rz, Ry, t1); mul(t1, Rx, rot); mul(rot, R_ren, R2);
    cuda_renderer::Model::mat4x4 mat4, mat4_2;
    mat4.init_from_ptr(R_ren, t_ren);
    mat4_2.init_from_ptr(R2, t_ren2);
    std::vector<cuda_renderer::Model::mat4x4> mat4_v = {mat4, mat4_2};

    auto depth_cuda = cuda_renderer::render_cuda_keep_in_gpu(model.tris, mat4_v, width, height, proj);   // test.cpp:143
    Mat3x3f K_(K);
    auto pcd1_cuda = cuda_icp::depth2cloud_cuda(depth_cuda.data(), width, height, K_);                   // test.cpp:153
    std::vector<int32_t> depth_host = depth_cuda.download();
    pose_refine::DepthImage scene_depth(depth_host.data() + width * height, height, width);

    cuda_icp::RegistrationResult result;
    device_vector_holder<Vec3f> pcd_buffer_cuda, normal_buffer_cuda;
    KDTree_cuda kdtree_cuda;
    if (use_proj) {
        Scene_projective scene;
        scene.init_Scene_projective_cuda(scene_depth, K_, pcd_buffer_cuda, normal_buffer_cuda);          // test.cpp:163
        result = cuda_icp::ICP_Point2Plane_cuda(pcd1_cuda, scene);                                        // test.cpp:172
    } else {
        Scene_nn scene;
        scene.init_Scene_nn_cuda(scene_depth, K_, kdtree_cuda);                                           // test.cpp:166
        result = cuda_icp::ICP_Point2Plane_cuda(pcd1_cuda, scene);
    }
    std::printf("points %zu fitness %.6f rmse %.8f\n", pcd1_cuda.size(), result.fitness_, result.inlier_rmse_);
    for (int i = 0; i < 4; i++) std::printf("%.7f %.7f %.7f %.7f\n", result.transformation_[i][0], result.transformation_[i][1], result.transformation_[i][2], result.transformation_[i][3]);
    // the same hypothesis (and a copy of it) through the one-call batch refiner: must reproduce the single-call result
    PoseRefiner refiner(model.tris, width, height, K, 2);
    if (use_proj) refiner.set_scene_projective(scene_depth); else refiner.set_scene_nn(scene_depth);
    std::vector<cuda_renderer::Model::mat4x4> hyp = {mat4, mat4};
    auto batch = refiner.refine(hyp);
    float worst = 0.f;
    for (const auto& b : batch)
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) worst = std::fmax(worst, std::fabs(b.transformation_[i][j] - result.transformation_[i][j]));
    std::printf("batch refiner: max |T_batch - T_single| = %.3g, fitness %.6f\n", worst, batch[0].fitness_);
    // PoseRenderer (pose_renderer.h:9-32): both fixture poses, full size and down_sample = 2 (width/2 x height/2 with the
    // full-resolution projection, pose_renderer.cpp:25-36); printed sums are checked against the oracle by the test
    PoseRenderer pr(argv[1]);
    pr.set_K_width_height(K, width, height);
    for (float ds : {1.0f, 2.0f}) {
        auto dm = pr.render_depth_mask(mat4_v, ds);
        auto only_depth = pr.render_depth(mat4_v, ds);
        auto only_mask = pr.render_mask(mat4_v, ds);
        for (size_t i = 0; i < dm.size(); i++) {
            unsigned long long sum = 0, cnt = 0;
            for (uint16_t v : dm[i].depth) sum += v;
            for (uint8_t v : dm[i].mask) cnt += (v == 255);
            const bool same = only_depth[i] == dm[i].depth && only_mask[i] == dm[i].mask;
            std::printf("pose_renderer ds %.0f pose %zu size %zu depth_sum %llu mask_count %llu consistent %d\n", ds, i, dm[i].depth.size(), sum, cnt, (int)same);
        }
    }
    return (result.fitness_ > 0.9f && worst < 5e-4f) ? 0 : 1;
}
