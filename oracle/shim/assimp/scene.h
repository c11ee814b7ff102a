// Minimal stand-in for the assimp types named in cuda_renderer/renderer.cpp:16-150.
// TEST INFRASTRUCTURE ONLY (see oracle/Makefile). aiImportFile is implemented in
// oracle/ref_glue.cpp by a small ASCII-PLY reader.
#pragma once
#include <cstddef>
struct aiVector3D { float x = 0, y = 0, z = 0; };
struct aiMatrix4x4 {
    float a1 = 1, a2 = 0, a3 = 0, a4 = 0;
    float b1 = 0, b2 = 1, b3 = 0, b4 = 0;
    float c1 = 0, c2 = 0, c3 = 1, c4 = 0;
    float d1 = 0, d2 = 0, d3 = 0, d4 = 1;
};
struct aiFace { unsigned int mNumIndices = 0; unsigned int* mIndices = nullptr; };
struct aiMesh {
    unsigned int mNumVertices = 0; aiVector3D* mVertices = nullptr;
    unsigned int mNumFaces = 0; aiFace* mFaces = nullptr;
};
struct aiNode {
    aiMatrix4x4 mTransformation;
    unsigned int mNumMeshes = 0; unsigned int* mMeshes = nullptr;
    unsigned int mNumChildren = 0; aiNode** mChildren = nullptr;
};
struct aiScene {
    unsigned int mNumMeshes = 0; aiMesh** mMeshes = nullptr;
    aiNode* mRootNode = nullptr;
};
