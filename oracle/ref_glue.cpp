// oracle/ref_glue.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" doorway onto the reference's OWN CPU code, compiled verbatim from
// /root/reference by oracle/Makefile into oracle/_ref/libpose_refine_ref.so:
//   cuda_renderer/renderer.cpp, cuda_icp/icp.cpp, cuda_icp/scene/common.cpp,
//   cuda_icp/scene/depth_scene/depth_scene.cpp, cuda_icp/scene/pcd_scene/pcd_scene.cpp
// (OpenCV / assimp / Eigen replaced by the stand-ins in oracle/shim/).
// Nothing in the product (pose_refine_b200/, include/) links or loads this; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// The entry points mirror oracle/oracle.cpp's (prefix orc_ there, ref_ here) so the tests can
// check the restatement against the reference function by function.
#include "cuda_renderer/renderer.h"
#include "cuda_icp/icp.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <chrono>
#include <omp.h>

// ---------------------------------------------------------------------------------------------
// assimp stand-in: aiImportFile = ASCII PLY reader producing one mesh on the root node, which
// is all cuda_renderer::Model::LoadModel (renderer.cpp:16-58) walks.
// ---------------------------------------------------------------------------------------------
namespace {
struct OwnedScene {
    aiScene scene;
    aiNode root;
    aiMesh mesh;
    aiMesh* mesh_ptr;
    unsigned int mesh_index = 0;
    std::vector<aiVector3D> verts;
    std::vector<aiFace> faces;
    std::vector<unsigned int> indices;
};
}  // namespace

const aiScene* aiImportFile(const char* file, unsigned int) {
    std::ifstream in(file);
    if (!in) return nullptr;
    std::string line;
    size_t n_vert = 0, n_face = 0;
    int n_vert_props = 0;
    bool in_vertex = false, ascii = false;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tok;
        ss >> tok;
        if (tok == "format") { std::string f; ss >> f; ascii = (f == "ascii"); }
        else if (tok == "element") {
            std::string what; size_t n; ss >> what >> n;
            in_vertex = (what == "vertex");
            if (what == "vertex") n_vert = n;
            if (what == "face") n_face = n;
        } else if (tok == "property" && in_vertex) n_vert_props++;
        else if (tok == "end_header") break;
    }
    if (!ascii) return nullptr;
    auto* os = new OwnedScene();
    os->verts.resize(n_vert);
    for (size_t i = 0; i < n_vert; i++) {
        std::getline(in, line);
        const char* p = line.c_str(); char* e;
        os->verts[i].x = strtof(p, &e); p = e;
        os->verts[i].y = strtof(p, &e); p = e;
        os->verts[i].z = strtof(p, &e);
    }
    os->faces.resize(n_face);
    os->indices.reserve(n_face * 3);
    std::vector<size_t> starts(n_face);
    for (size_t i = 0; i < n_face; i++) {
        std::getline(in, line);
        std::istringstream ss(line);
        unsigned int k; ss >> k;
        os->faces[i].mNumIndices = k;
        starts[i] = os->indices.size();
        for (unsigned int j = 0; j < k; j++) { unsigned int v; ss >> v; os->indices.push_back(v); }
    }
    for (size_t i = 0; i < n_face; i++) os->faces[i].mIndices = os->indices.data() + starts[i];
    os->mesh.mNumVertices = (unsigned int)n_vert; os->mesh.mVertices = os->verts.data();
    os->mesh.mNumFaces = (unsigned int)n_face; os->mesh.mFaces = os->faces.data();
    os->mesh_ptr = &os->mesh;
    os->root.mNumMeshes = 1; os->root.mMeshes = &os->mesh_index;
    os->scene.mNumMeshes = 1; os->scene.mMeshes = &os->mesh_ptr; os->scene.mRootNode = &os->root;
    return &os->scene;
}
void aiReleaseImport(const aiScene* scene) {
    // the aiScene is the first member of OwnedScene
    delete reinterpret_cast<OwnedScene*>(const_cast<aiScene*>(scene));
}
void aiIdentityMatrix4(aiMatrix4x4* m) { *m = aiMatrix4x4(); }
void aiMultiplyMatrix4(aiMatrix4x4* dst, const aiMatrix4x4* src) {
    const float a[4][4] = {{dst->a1, dst->a2, dst->a3, dst->a4}, {dst->b1, dst->b2, dst->b3, dst->b4},
                           {dst->c1, dst->c2, dst->c3, dst->c4}, {dst->d1, dst->d2, dst->d3, dst->d4}};
    const float b[4][4] = {{src->a1, src->a2, src->a3, src->a4}, {src->b1, src->b2, src->b3, src->b4},
                           {src->c1, src->c2, src->c3, src->c4}, {src->d1, src->d2, src->d3, src->d4}};
    float r[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
        r[i][j] = 0;
        for (int k = 0; k < 4; k++) r[i][j] += a[i][k] * b[k][j];
    }
    dst->a1 = r[0][0]; dst->a2 = r[0][1]; dst->a3 = r[0][2]; dst->a4 = r[0][3];
    dst->b1 = r[1][0]; dst->b2 = r[1][1]; dst->b3 = r[1][2]; dst->b4 = r[1][3];
    dst->c1 = r[2][0]; dst->c2 = r[2][1]; dst->c3 = r[2][2]; dst->c4 = r[2][3];
    dst->d1 = r[3][0]; dst->d2 = r[3][1]; dst->d3 = r[3][2]; dst->d4 = r[3][3];
}
void aiTransformVecByMatrix4(aiVector3D* v, const aiMatrix4x4* m) {
    aiVector3D r;
    r.x = m->a1 * v->x + m->a2 * v->y + m->a3 * v->z + m->a4;
    r.y = m->b1 * v->x + m->b2 * v->y + m->b3 * v->z + m->b4;
    r.z = m->c1 * v->x + m->c2 * v->y + m->c3 * v->z + m->c4;
    *v = r;
}

// ---------------------------------------------------------------------------------------------
namespace {

using cuda_renderer::Model;

struct SceneHandle {
    int kind;  // 0 projective, 1 nn
    Scene_projective proj;
    std::vector<Vec3f> pcd, normal;
    Scene_nn nn;
    KDTree_cpu tree;
};

cv::Mat wrap_depth(const void* depth, int is_i32, int W, int H) {
    return cv::Mat(H, W, is_i32 ? CV_32S : CV_16U, const_cast<void*>(depth));
}
Mat3x3f wrap_K(const float* K) { return Mat3x3f(K); }

template <class Scene>
void query_many(const Scene& s, const float* pts, size_t n, float* dst, float* nrm, uint8_t* valid) {
    for (size_t i = 0; i < n; i++) {
        Vec3f p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), q, nn;
        bool v = false;
        s.query(p, q, nn, v);
        valid[i] = v;
        if (v) {
            dst[3 * i] = q.x; dst[3 * i + 1] = q.y; dst[3 * i + 2] = q.z;
            nrm[3 * i] = nn.x; nrm[3 * i + 1] = nn.y; nrm[3 * i + 2] = nn.z;
        }
    }
}
template <class Scene> void pcd2ab_sum(const Scene& s, const float* pts, size_t n, float* out29) {
    cuda_icp::thrust__pcd2Ab<Scene> f(s);
    cuda_icp::Vec29f acc = cuda_icp::Vec29f::Zero();
    for (size_t i = 0; i < n; i++) acc += f(Vec3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
    for (int k = 0; k < 29; k++) out29[k] = acc[k];
}
void pack_result(const cuda_icp::RegistrationResult& r, float* out18) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out18[4 * i + j] = r.transformation_[i][j];
    out18[16] = r.inlier_rmse_;
    out18[17] = r.fitness_;
}
template <class Scene>
void icp_run(const Scene& s, float* pts, size_t n, float rf, float rr, int mi, float* out18) {
    std::vector<Vec3f> cloud(n);
    for (size_t i = 0; i < n; i++) cloud[i] = Vec3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    auto res = cuda_icp::ICP_Point2Plane_cpu(cloud, s, cuda_icp::ICPConvergenceCriteria(rf, rr, mi));
    for (size_t i = 0; i < n; i++) { pts[3 * i] = cloud[i].x; pts[3 * i + 1] = cloud[i].y; pts[3 * i + 2] = cloud[i].z; }
    pack_result(res, out18);
}

}  // namespace

extern "C" {

const char* ref_kind() { return "reference"; }
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_max_threads(); }

// cuda_renderer::Model(fileName) -> tris (renderer.cpp:11-58). Returns #triangles; copies up to cap.
long ref_load_model(const char* path, float* tris_out, long cap) {
    FILE* save = stdout; (void)save;
    std::streambuf* old = std::cout.rdbuf(nullptr);  // LoadModel prints; keep test output clean
    Model model{std::string(path)};
    std::cout.rdbuf(old);
    long T = (long)model.tris.size();
    if (tris_out) std::memcpy(tris_out, model.tris.data(), sizeof(Model::Triangle) * (size_t)std::min(T, cap));
    return T;
}

// cuda_renderer::compute_proj (renderer.cpp:161-185)
void ref_compute_proj(const float* K, int W, int H, float near_, float far_, float* out16) {
    cv::Mat Km(3, 3, CV_32F, const_cast<float*>(K));
    Model::mat4x4 p = cuda_renderer::compute_proj(Km, W, H, near_, far_);
    std::memcpy(out16, &p, 64);
}

// cuda_renderer::render_cpu (renderer.cpp:259-298). out: P * W' * H' int32.
void ref_render(const float* tris, size_t T, const float* poses, size_t P, size_t W, size_t H,
                const float* proj, const int* roi, int32_t* out) {
    std::vector<Model::Triangle> tv(T);
    std::memcpy(tv.data(), tris, T * sizeof(Model::Triangle));
    std::vector<Model::mat4x4> pv(P);
    for (size_t i = 0; i < P; i++) pv[i].init_from_ptr(poses + 16 * i);
    Model::mat4x4 pm; pm.init_from_ptr(proj);
    Model::ROI r = {roi[0], roi[1], roi[2], roi[3]};
    std::vector<int32_t> res = cuda_renderer::render_cpu(tv, pv, W, H, pm, r);
    std::memcpy(out, res.data(), res.size() * sizeof(int32_t));
}

// cuda_icp::depth2cloud_cpu (icp.cpp:73-117). Returns #points; writes up to cap points.
long ref_depth2cloud(const void* depth, int is_i32, uint32_t W, uint32_t H, const float* K,
                     uint32_t stride, uint32_t tl_x, uint32_t tl_y, float* out_pts, long cap) {
    Mat3x3f Km = wrap_K(K);
    std::vector<Vec3f> c = is_i32
        ? cuda_icp::depth2cloud_cpu((int32_t*)depth, W, H, Km, stride, tl_x, tl_y)
        : cuda_icp::depth2cloud_cpu((uint16_t*)depth, W, H, Km, stride, tl_x, tl_y);
    long n = (long)c.size();
    for (long i = 0; i < std::min(n, cap); i++) { out_pts[3 * i] = c[i].x; out_pts[3 * i + 1] = c[i].y; out_pts[3 * i + 2] = c[i].z; }
    return n;
}

// get_normal (scene/common.cpp:17-107). normals: W*H*3 floats.
void ref_get_normal(const void* depth, int is_i32, int W, int H, const float* K, float* normals) {
    cv::Mat d = wrap_depth(depth, is_i32, W, H);
    std::vector<Vec3f> n = get_normal(d, wrap_K(K));
    for (size_t i = 0; i < n.size(); i++) { normals[3 * i] = n[i].x; normals[3 * i + 1] = n[i].y; normals[3 * i + 2] = n[i].z; }
}

// Scene_projective::init_Scene_projective_cpu (depth_scene.cpp:3-35)
void* ref_scene_projective_create(const void* depth, int is_i32, const float* K, size_t W, size_t H, float max_dist) {
    auto* h = new SceneHandle();
    h->kind = 0;
    cv::Mat d = wrap_depth(depth, is_i32, (int)W, (int)H);
    Mat3x3f Km = wrap_K(K);
    h->proj.init_Scene_projective_cpu(d, Km, h->pcd, h->normal, W, H, max_dist);
    return h;
}
// Scene_nn::init_Scene_nn_cpu + KDTree_cpu::build_tree (pcd_scene.cpp:4-184)
void* ref_scene_nn_create(const void* depth, int is_i32, const float* K, size_t W, size_t H) {
    auto* h = new SceneHandle();
    h->kind = 1;
    cv::Mat d = wrap_depth(depth, is_i32, (int)W, (int)H);
    Mat3x3f Km = wrap_K(K);
    h->nn.init_Scene_nn_cpu(d, Km, h->tree);
    return h;
}
void ref_scene_destroy(void* hv) { delete (SceneHandle*)hv; }
void ref_scene_sizes(void* hv, long* n_pts, long* n_nodes) {
    auto* h = (SceneHandle*)hv;
    if (h->kind == 0) { *n_pts = (long)h->pcd.size(); *n_nodes = 0; }
    else { *n_pts = (long)h->tree.pcd_buffer.size(); *n_nodes = (long)h->tree.nodes.size(); }
}
// copies out the scene arrays: pcd/normal n_pts*3 floats, nodes n_nodes*52 bytes (Node_kdtree, pcd_scene.h:5-25)
void ref_scene_get(void* hv, float* pcd, float* normal, void* nodes) {
    auto* h = (SceneHandle*)hv;
    const std::vector<Vec3f>& p = h->kind == 0 ? h->pcd : h->tree.pcd_buffer;
    const std::vector<Vec3f>& n = h->kind == 0 ? h->normal : h->tree.normal_buffer;
    for (size_t i = 0; i < p.size(); i++) {
        pcd[3 * i] = p[i].x; pcd[3 * i + 1] = p[i].y; pcd[3 * i + 2] = p[i].z;
        normal[3 * i] = n[i].x; normal[3 * i + 1] = n[i].y; normal[3 * i + 2] = n[i].z;
    }
    if (h->kind == 1 && nodes) std::memcpy(nodes, h->tree.nodes.data(), h->tree.nodes.size() * sizeof(Node_kdtree));
}

// Scene_*::query (depth_scene.h:30-48, pcd_scene.h:61-136) over n points
void ref_query(void* hv, const float* pts, size_t n, float* dst, float* nrm, uint8_t* valid) {
    auto* h = (SceneHandle*)hv;
    if (h->kind == 0) query_many(h->proj, pts, n, dst, nrm, valid);
    else query_many(h->nn, pts, n, dst, nrm, valid);
}
// sequential sum of thrust__pcd2Ab (icp.h:128-209) over n points
void ref_pcd2ab(void* hv, const float* pts, size_t n, float* out29) {
    auto* h = (SceneHandle*)hv;
    if (h->kind == 0) pcd2ab_sum(h->proj, pts, n, out29);
    else pcd2ab_sum(h->nn, pts, n, out29);
}
// eigen_slover_666 (icp.cpp:29-45)
void ref_solve_666(const float* A, const float* b, float* T16) {
    float Ac[36], bc[6];
    std::memcpy(Ac, A, sizeof(Ac)); std::memcpy(bc, b, sizeof(bc));
    Mat4x4f m = cuda_icp::eigen_slover_666(Ac, bc);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T16[4 * i + j] = m[i][j];
}
// ICP_Point2Plane_cpu (icp.cpp:125-188). pts mutated in place. out18 = T(16, row-major), rmse, fitness.
void ref_icp(void* hv, float* pts, size_t n, float rel_fit, float rel_rmse, int max_iter, float* out18) {
    auto* h = (SceneHandle*)hv;
    if (h->kind == 0) icp_run(h->proj, pts, n, rel_fit, rel_rmse, max_iter, out18);
    else icp_run(h->nn, pts, n, rel_fit, rel_rmse, max_iter, out18);
}

// The reference CPU pipeline for a batch of hypotheses, as BASELINE.md section 3 describes it:
// render_cpu (OpenMP over poses) -> per hypothesis depth2cloud_cpu -> ICP_Point2Plane_cpu.
// schedule 0: hypotheses serial, OpenMP inside each call (as shipped);
// schedule 1: omp parallel for over hypotheses, inner regions serialised.
// Returns wall seconds (steady_clock); results: P*18 floats; n_pts (optional): P longs.
double ref_pipeline(void* hv, const float* tris, size_t T, const float* poses, size_t P, size_t W, size_t H,
                    const float* proj, const float* K, float rel_fit, float rel_rmse, int max_iter,
                    int schedule, float* results, long* n_pts) {
    auto* h = (SceneHandle*)hv;
    std::vector<Model::Triangle> tv(T);
    std::memcpy(tv.data(), tris, T * sizeof(Model::Triangle));
    std::vector<Model::mat4x4> pv(P);
    for (size_t i = 0; i < P; i++) pv[i].init_from_ptr(poses + 16 * i);
    Model::mat4x4 pm; pm.init_from_ptr(proj);
    Mat3x3f Km = wrap_K(K);
    cuda_icp::ICPConvergenceCriteria crit(rel_fit, rel_rmse, max_iter);

    auto t0 = std::chrono::steady_clock::now();
    std::vector<int32_t> depth = cuda_renderer::render_cpu(tv, pv, W, H, pm);
    auto one = [&](size_t i) {
        std::vector<Vec3f> cloud = cuda_icp::depth2cloud_cpu(depth.data() + i * W * H, (uint32_t)W, (uint32_t)H, Km);
        if (n_pts) n_pts[i] = (long)cloud.size();
        cuda_icp::RegistrationResult r = h->kind == 0 ? cuda_icp::ICP_Point2Plane_cpu(cloud, h->proj, crit)
                                                       : cuda_icp::ICP_Point2Plane_cpu(cloud, h->nn, crit);
        pack_result(r, results + 18 * i);
    };
    if (schedule == 0) {
        for (size_t i = 0; i < P; i++) one(i);
    } else {
        omp_set_max_active_levels(1);
#pragma omp parallel for schedule(dynamic, 1)
        for (size_t i = 0; i < P; i++) one(i);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
