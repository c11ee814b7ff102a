#pragma once
#include "scene.h"
const aiScene* aiImportFile(const char* file, unsigned int flags);
void aiReleaseImport(const aiScene* scene);
void aiMultiplyMatrix4(aiMatrix4x4* dst, const aiMatrix4x4* src);
void aiTransformVecByMatrix4(aiVector3D* vec, const aiMatrix4x4* mat);
void aiIdentityMatrix4(aiMatrix4x4* mat);
