// TEST INFRASTRUCTURE — not part of the product path.
//
// Minimal stand-in for the subset of GLM (OpenGL Mathematics, 0.9.9.8 — the
// `libglm-dev` of the reference's Ubuntu 22.04 docker image, see
// /root/reference/docker/Dockerfile:1 and CMakeLists.txt:44,107) that the
// reference's CUDA rasterizer uses (cuda_rasterizer/forward.cu:25-151,
// backward.cu:23-340).  GLM is a third-party dependency that is NOT vendored
// under /root/reference and is not installed in this image, so the reference
// kernels cannot be compiled without it.  This header restates GLM's published
// semantics for exactly those operations so that `oracle/Makefile` can compile
// the UNMODIFIED reference sources into oracle/_ref/:
//
//   * column-major mat3, m[col][row]; mat3(a..i) fills column by column;
//   * mat3*mat3 element (c,r) = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]
//     evaluated left to right (glm/detail/type_mat3x3.inl operator*);
//   * dot(vec3) = tmp = a*b; tmp.x + tmp.y + tmp.z (glm/detail/func_geometric.inl);
//   * length(v) = sqrt(dot(v, v)); scalar*vec / vec*scalar / vec/scalar componentwise.
//
// Evaluation order is kept as in GLM because nvcc's FMA contraction follows the
// expression shape and the integer outputs (radii, tiles_touched) depend on it.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

#define GLM_SHIM_FN __host__ __device__ __forceinline__

namespace glm {

struct vec3 {
    float x, y, z;
    GLM_SHIM_FN vec3() : x(0.f), y(0.f), z(0.f) {}
    GLM_SHIM_FN vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    GLM_SHIM_FN explicit vec3(float s) : x(s), y(s), z(s) {}
    GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
    GLM_SHIM_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLM_SHIM_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLM_SHIM_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    GLM_SHIM_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    GLM_SHIM_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    GLM_SHIM_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLM_SHIM_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_SHIM_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_SHIM_FN vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
GLM_SHIM_FN vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
GLM_SHIM_FN vec3 operator/(const vec3& v, float s) { return vec3(v.x / s, v.y / s, v.z / s); }

GLM_SHIM_FN float dot(const vec3& a, const vec3& b) {
    vec3 tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
GLM_SHIM_FN float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLM_SHIM_FN float length(const vec4& v) {
    return sqrtf((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
}
GLM_SHIM_FN vec3 max(const vec3& v, float s) {
    return vec3(fmaxf(v.x, s), fmaxf(v.y, s), fmaxf(v.z, s));
}

struct mat3 {
    vec3 c[3];
    GLM_SHIM_FN mat3() { c[0] = vec3(1.f, 0.f, 0.f); c[1] = vec3(0.f, 1.f, 0.f); c[2] = vec3(0.f, 0.f, 1.f); }
    GLM_SHIM_FN explicit mat3(float s) { c[0] = vec3(s, 0.f, 0.f); c[1] = vec3(0.f, s, 0.f); c[2] = vec3(0.f, 0.f, s); }
    GLM_SHIM_FN mat3(float x0, float y0, float z0,
                     float x1, float y1, float z1,
                     float x2, float y2, float z2) {
        c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2);
    }
    GLM_SHIM_FN mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    GLM_SHIM_FN vec3& operator[](int i) { return c[i]; }
    GLM_SHIM_FN const vec3& operator[](int i) const { return c[i]; }
};

GLM_SHIM_FN mat3 operator*(const mat3& m1, const mat3& m2) {
    const float A00 = m1[0][0], A01 = m1[0][1], A02 = m1[0][2];
    const float A10 = m1[1][0], A11 = m1[1][1], A12 = m1[1][2];
    const float A20 = m1[2][0], A21 = m1[2][1], A22 = m1[2][2];
    const float B00 = m2[0][0], B01 = m2[0][1], B02 = m2[0][2];
    const float B10 = m2[1][0], B11 = m2[1][1], B12 = m2[1][2];
    const float B20 = m2[2][0], B21 = m2[2][1], B22 = m2[2][2];
    mat3 r(0.f);
    r[0][0] = A00 * B00 + A10 * B01 + A20 * B02;
    r[0][1] = A01 * B00 + A11 * B01 + A21 * B02;
    r[0][2] = A02 * B00 + A12 * B01 + A22 * B02;
    r[1][0] = A00 * B10 + A10 * B11 + A20 * B12;
    r[1][1] = A01 * B10 + A11 * B11 + A21 * B12;
    r[1][2] = A02 * B10 + A12 * B11 + A22 * B12;
    r[2][0] = A00 * B20 + A10 * B21 + A20 * B22;
    r[2][1] = A01 * B20 + A11 * B21 + A21 * B22;
    r[2][2] = A02 * B20 + A12 * B21 + A22 * B22;
    return r;
}
GLM_SHIM_FN mat3 operator*(float s, const mat3& m) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
GLM_SHIM_FN mat3 operator*(const mat3& m, float s) { return mat3(m[0] * s, m[1] * s, m[2] * s); }

GLM_SHIM_FN mat3 transpose(const mat3& m) {
    return mat3(m[0][0], m[1][0], m[2][0],
                m[0][1], m[1][1], m[2][1],
                m[0][2], m[1][2], m[2][2]);
}

}  // namespace glm
