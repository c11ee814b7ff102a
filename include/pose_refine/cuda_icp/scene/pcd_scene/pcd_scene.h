// Scene_nn, KDTree_cuda, Node_kdtree (cuda_icp/scene/pcd_scene/pcd_scene.h) over the C ABI.
#pragma once
#include "../common.h"

typedef pr_node_kdtree Node_kdtree;   // 52 bytes, same field order as pcd_scene.h:5-25
static_assert(sizeof(Node_kdtree) == 52, "Node_kdtree layout");

class KDTree_cuda {   // pcd_scene.h:37-43
public:
    device_vector_holder<Vec3f> pcd_buffer;
    device_vector_holder<Vec3f> normal_buffer;
    device_vector_holder<Node_kdtree> nodes;
};

class Scene_nn {
    float max_dist_diff = 0.1f;   // m
    Vec3f* pcd_ptr = nullptr;
    Vec3f* normal_ptr = nullptr;
    Node_kdtree* node_ptr = nullptr;
    size_t n_points = 0, n_nodes = 0;
public:
    // init_Scene_nn_cuda (pcd_scene.cu:3-20): normals + compaction + kd-tree build (host, leaf <= 10), upload
    void init_Scene_nn_cuda(const pose_refine::DepthImage& scene_depth, Mat3x3f& scene_K, KDTree_cuda& kdtree) {
        const size_t cap = (size_t)scene_depth.rows * scene_depth.cols;
        std::vector<Vec3f> pcd(cap), nrm(cap);
        std::vector<Node_kdtree> nodes(2 * cap + 1);
        pose_refine::check(pr_scene_nn_build_host(scene_depth.data, scene_depth.is_int32, (uint32_t)scene_depth.cols, (uint32_t)scene_depth.rows,
                                                  scene_K.data(), 10, reinterpret_cast<float*>(pcd.data()), reinterpret_cast<float*>(nrm.data()), cap,
                                                  nodes.data(), nodes.size(), &n_points, &n_nodes), "pr_scene_nn_build_host");
        pcd.resize(n_points); nrm.resize(n_points); nodes.resize(n_nodes);
        kdtree.pcd_buffer.upload(pcd); kdtree.normal_buffer.upload(nrm); kdtree.nodes.upload(nodes);
        pcd_ptr = kdtree.pcd_buffer.data(); normal_ptr = kdtree.normal_buffer.data(); node_ptr = kdtree.nodes.data();
    }
    pr_scene_nn c_abi() const {
        pr_scene_nn s;
        s.max_dist_diff = max_dist_diff;
        s.pcd_dev = reinterpret_cast<const float*>(pcd_ptr); s.normal_dev = reinterpret_cast<const float*>(normal_ptr);
        s.nodes_dev = node_ptr; s.n_points = n_points; s.n_nodes = n_nodes;
        return s;
    }
};
