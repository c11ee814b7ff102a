// host_io.cpp -- host-only entry points: PLY ingestion and compute_proj.
//
// pr_load_ply replaces cuda_renderer::Model::LoadModel (renderer.cpp:16-58), which goes through
// assimp: what the renderer consumes is only Model::tris -- for every face the three vertex
// positions, transformed by the node matrix (identity for a PLY), in face order
// (renderer.cpp:69-108).  Faces with fewer than 3 indices are skipped as upstream does
// (renderer.cpp:79); polygons are fan-triangulated (assimp's Triangulate step).
// pr_compute_proj restates compute_proj (renderer.cpp:161-185).
#include "../../include/pose_refine_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <sstream>
#include <fstream>

namespace {

struct Prop { std::string name, type, list_count_type; bool is_list = false; };
struct Element { std::string name; size_t count = 0; std::vector<Prop> props; };

size_t type_size(const std::string& t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}

double read_binary_scalar(std::istream& in, const std::string& t) {
    char buf[8] = {0};
    const size_t n = type_size(t);
    in.read(buf, (std::streamsize)n);
    if (t == "char" || t == "int8") { int8_t v; memcpy(&v, buf, 1); return v; }
    if (t == "uchar" || t == "uint8") { uint8_t v; memcpy(&v, buf, 1); return v; }
    if (t == "short" || t == "int16") { int16_t v; memcpy(&v, buf, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, buf, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; memcpy(&v, buf, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, buf, 4); return v; }
    if (t == "float" || t == "float32") { float v; memcpy(&v, buf, 4); return v; }
    double v; memcpy(&v, buf, 8); return v;
}

}  // namespace

extern "C" {

int pr_load_ply(const char* path, float* tris_host, size_t capacity_tris, size_t* n_tris) {
    if (!path || !n_tris) return PR_ERR_INVALID_ARGUMENT;
    std::ifstream in(path, std::ios::binary);
    if (!in) return PR_ERR_IO;
    std::string line;
    if (!std::getline(in, line) || line.substr(0, 3) != "ply") return PR_ERR_IO;
    bool ascii = false, binary_le = false;
    std::vector<Element> elements;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string tok;
        ss >> tok;
        if (tok == "format") { std::string f; ss >> f; ascii = (f == "ascii"); binary_le = (f == "binary_little_endian"); }
        else if (tok == "element") { Element e; ss >> e.name >> e.count; elements.push_back(e); }
        else if (tok == "property" && !elements.empty()) {
            Prop p; std::string t; ss >> t;
            if (t == "list") { p.is_list = true; ss >> p.list_count_type >> p.type >> p.name; }
            else { p.type = t; ss >> p.name; }
            elements.back().props.push_back(p);
        } else if (tok == "end_header") break;
    }
    if (!ascii && !binary_le) return PR_ERR_UNSUPPORTED;

    std::vector<float> verts;
    size_t out = 0;
    for (const Element& el : elements) {
        int ix = -1, iy = -1, iz = -1;
        for (size_t k = 0; k < el.props.size(); k++) {
            if (el.props[k].name == "x") ix = (int)k;
            if (el.props[k].name == "y") iy = (int)k;
            if (el.props[k].name == "z") iz = (int)k;
        }
        const bool is_vertex = (el.name == "vertex");
        const bool is_face = (el.name == "face");
        if (is_vertex && (ix < 0 || iy < 0 || iz < 0)) return PR_ERR_IO;
        if (is_vertex) verts.resize(el.count * 3);
        std::vector<double> idx;
        for (size_t r = 0; r < el.count; r++) {
            std::istringstream ls;
            const char* cur = nullptr;
            if (ascii) {
                if (!std::getline(in, line)) return PR_ERR_IO;
                cur = line.c_str();
            }
            for (size_t k = 0; k < el.props.size(); k++) {
                const Prop& p = el.props[k];
                auto next = [&](const std::string& type) -> double {
                    if (!ascii) return read_binary_scalar(in, type);
                    char* endp = nullptr;
                    double v;
                    // vertex coordinates of a float property: parse as float (correctly rounded)
                    if (type == "float" || type == "float32") v = strtof(cur, &endp); else v = strtod(cur, &endp);
                    cur = endp;
                    return v;
                };
                if (p.is_list) {
                    const size_t cnt = (size_t)next(p.list_count_type);
                    idx.resize(cnt);
                    for (size_t j = 0; j < cnt; j++) idx[j] = next(p.type);
                    if (is_face && (p.name == "vertex_indices" || p.name == "vertex_index") && cnt >= 3) {
                        for (size_t j = 1; j + 1 < cnt; j++) {
                            const size_t tri[3] = {(size_t)idx[0], (size_t)idx[j], (size_t)idx[j + 1]};
                            if (tris_host && out < capacity_tris)
                                for (int c = 0; c < 3; c++) {
                                    if (tri[c] * 3 + 2 >= verts.size()) return PR_ERR_IO;
                                    memcpy(tris_host + out * 9 + 3 * c, verts.data() + tri[c] * 3, 12);
                                }
                            out++;
                        }
                    }
                } else {
                    const double v = next(p.type);
                    if (is_vertex) {
                        if ((int)k == ix) verts[r * 3 + 0] = (float)v;
                        if ((int)k == iy) verts[r * 3 + 1] = (float)v;
                        if ((int)k == iz) verts[r * 3 + 2] = (float)v;
                    }
                }
            }
            if (!ascii && !in) return PR_ERR_IO;
        }
    }
    *n_tris = out;
    if (tris_host && out > capacity_tris) return PR_ERR_CAPACITY;
    return PR_OK;
}

int pr_mesh_index(const float* tris_host, size_t n_tris, float* verts_out, int32_t* faces_out, size_t* n_verts) {
    if (!tris_host || !faces_out || !n_verts) return PR_ERR_INVALID_ARGUMENT;
    // exact-bit deduplication of the 3*n_tris corners (open addressing on the 96-bit pattern)
    const size_t n_corners = n_tris * 3;
    size_t cap = 16;
    while (cap < n_corners * 2) cap <<= 1;
    std::vector<int32_t> table(cap, -1);
    std::vector<float> verts;
    verts.reserve(n_corners);
    for (size_t c = 0; c < n_corners; c++) {
        uint32_t k[3];
        memcpy(k, tris_host + 3 * c, 12);
        uint64_t h = (k[0] * 0x9E3779B97F4A7C15ull) ^ (k[1] * 0xC2B2AE3D27D4EB4Full) ^ (k[2] * 0x165667B19E3779F9ull);
        size_t slot = (size_t)(h >> 20) & (cap - 1);
        for (;;) {
            const int32_t v = table[slot];
            if (v < 0) {
                table[slot] = (int32_t)(verts.size() / 3);
                verts.insert(verts.end(), tris_host + 3 * c, tris_host + 3 * c + 3);
                faces_out[c] = table[slot];
                break;
            }
            if (memcmp(verts.data() + 3 * (size_t)v, k, 12) == 0) { faces_out[c] = v; break; }
            slot = (slot + 1) & (cap - 1);
        }
    }
    *n_verts = verts.size() / 3;
    if (verts_out) memcpy(verts_out, verts.data(), verts.size() * 4);
    return PR_OK;
}

// Morton (z-order) code of a point quantised to 10 bits per axis
static inline uint32_t spread10(uint32_t v) {
    v &= 0x3FF;
    v = (v | (v << 16)) & 0x030000FF;
    v = (v | (v << 8)) & 0x0300F00F;
    v = (v | (v << 4)) & 0x030C30C3;
    v = (v | (v << 2)) & 0x09249249;
    return v;
}

int pr_mesh_cluster(const float* verts, size_t n_verts, int32_t* faces, size_t n_tris, int32_t* cluster_vert_off,
                    int32_t* cluster_verts, size_t* n_clusters) {
    if (!verts || !faces || !cluster_vert_off || !cluster_verts || !n_clusters || n_verts == 0) return PR_ERR_INVALID_ARGUMENT;
#ifndef PR_CLUSTER_TRIS
#define PR_CLUSTER_TRIS 64
#endif
    const size_t kTris = PR_CLUSTER_TRIS;
    for (size_t i = 0; i < 3 * n_tris; i++)
        if (faces[i] < 0 || (size_t)faces[i] >= n_verts) return PR_ERR_INVALID_ARGUMENT;
    float lo[3] = {verts[0], verts[1], verts[2]}, hi[3] = {verts[0], verts[1], verts[2]};
    for (size_t v = 0; v < n_verts; v++)
        for (int a = 0; a < 3; a++) {
            const float x = verts[3 * v + a];
            if (x < lo[a]) lo[a] = x;
            if (x > hi[a]) hi[a] = x;
        }
    std::vector<std::pair<uint32_t, uint32_t>> keys(n_tris);       // (morton code of the centroid, face)
    for (size_t t = 0; t < n_tris; t++) {
        uint32_t code = 0;
        for (int a = 0; a < 3; a++) {
            const float c = (verts[3 * (size_t)faces[3 * t] + a] + verts[3 * (size_t)faces[3 * t + 1] + a] + verts[3 * (size_t)faces[3 * t + 2] + a]) / 3.f;
            const float ext = hi[a] - lo[a];
            float q = ext > 0.f ? (c - lo[a]) / ext * 1023.f : 0.f;
            if (!(q >= 0.f)) q = 0.f;                              // also NaN
            if (q > 1023.f) q = 1023.f;
            code |= spread10((uint32_t)q) << a;
        }
        keys[t] = {code, (uint32_t)t};
    }
    std::stable_sort(keys.begin(), keys.end(), [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) { return a.first < b.first; });
    std::vector<int32_t> sorted(3 * n_tris);
    for (size_t t = 0; t < n_tris; t++)
        for (int k = 0; k < 3; k++) sorted[3 * t + k] = faces[3 * (size_t)keys[t].second + k];
    memcpy(faces, sorted.data(), sorted.size() * sizeof(int32_t));
    // unique vertices per cluster, in order of first use
    const size_t nc = (n_tris + kTris - 1) / kTris;
    std::vector<int32_t> stamp(n_verts, -1);
    size_t n = 0;
    for (size_t c = 0; c < nc; c++) {
        cluster_vert_off[c] = (int32_t)n;
        const size_t t1 = std::min(n_tris, (c + 1) * kTris);
        for (size_t i = 3 * c * kTris; i < 3 * t1; i++) {
            const int32_t v = faces[i];
            if (stamp[v] != (int32_t)c) { stamp[v] = (int32_t)c; cluster_verts[n++] = v; }
        }
    }
    cluster_vert_off[nc] = (int32_t)n;
    *n_clusters = nc;
    return PR_OK;
}

int pr_compute_proj(const float K[9], int width, int height, float near_plane, float far_plane, float p[16]) {
    if (!K || !p || width <= 0 || height <= 0) return PR_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 16; i++) p[i] = 0.f;
    // renderer.cpp:164-182; each "x = -x" flip of the original is folded into the sign here
    p[0] = 2 * K[0] / width;
    p[1] = -(-2 * K[1] / width);
    p[2] = -(-2 * K[2] / width + 1);
    p[5] = -(2 * K[4] / height);
    p[6] = -(2 * K[5] / height - 1);
    p[10] = -(-(far_plane + near_plane) / (far_plane - near_plane));
    p[11] = -2 * far_plane * near_plane / (far_plane - near_plane);
    p[14] = 1.f;
    return PR_OK;
}

}  // extern "C"
