// Per-view constants of a keyframe, with the arithmetic of GaussianKeyframe::getWorld2View2,
// getProjectionMatrix and computeTransformTensors (/root/reference/src/gaussian_keyframe.cpp:151-184, 228-279) but
// without Eigen (absent from this image): the world->camera pose comes as a quaternion (w, x, y, z) + translation, the
// outputs are the three tensors the rasterizer reads — world_view_transform_ (transposed, i.e. the memory order
// m[4*col+row] the kernels index), full_proj_transform_ and camera_center_.  Host arithmetic on whatever device is
// asked for; checked against the reference's own dump (check_colmap.md) in tests/test_keyframe_golden_cpu.py.
#pragma once

#include <torch/torch.h>

#include <array>
#include <cmath>
#include <tuple>

namespace keyframe_transforms {

// GaussianKeyframe::getProjectionMatrix (:251-279), row-major P applied to column vectors
inline torch::Tensor getProjectionMatrix(float znear, float zfar, float fovX, float fovY, torch::Device device = torch::kCPU) {
    const float tanHalfFovY = std::tan(fovY / 2), tanHalfFovX = std::tan(fovX / 2);
    const float top = tanHalfFovY * znear, bottom = -top, right = tanHalfFovX * znear, left = -right;
    torch::Tensor P = torch::zeros({4, 4}, torch::TensorOptions().dtype(torch::kFloat32));
    auto a = P.accessor<float, 2>();
    a[0][0] = 2.0 * znear / (right - left);
    a[1][1] = 2.0 * znear / (top - bottom);
    a[0][2] = (right + left) / (right - left);
    a[1][2] = (top + bottom) / (top - bottom);
    a[3][2] = 1.0f;
    a[2][2] = zfar / (zfar - znear);
    a[2][3] = -(zfar * znear) / (zfar - znear);
    return P.to(device);
}

// rotation matrix of a unit quaternion (w, x, y, z) — Eigen::Quaternion::toRotationMatrix
inline std::array<float, 9> quaternionToRotation(double w, double x, double y, double z) {
    const double n = std::sqrt(w * w + x * x + y * y + z * z);
    w /= n; x /= n; y /= n; z /= n;
    return {float(1 - 2 * (y * y + z * z)), float(2 * (x * y - w * z)), float(2 * (x * z + w * y)),
            float(2 * (x * y + w * z)), float(1 - 2 * (x * x + z * z)), float(2 * (y * z - w * x)),
            float(2 * (x * z - w * y)), float(2 * (y * z + w * x)), float(1 - 2 * (x * x + y * y))};
}

// GaussianKeyframe::getWorld2View2 (:228-249): [R t; 0 1] with the camera centre shifted by `trans` and scaled
inline torch::Tensor getWorld2View2(const std::array<float, 9>& R, const std::array<float, 3>& t,
                                    const std::array<float, 3>& trans = {0.f, 0.f, 0.f}, float scale = 1.0f) {
    torch::Tensor Rt = torch::zeros({4, 4}, torch::TensorOptions().dtype(torch::kFloat32));
    auto a = Rt.accessor<float, 2>();
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) a[r][c] = R[3 * r + c];
        a[r][3] = t[r];
    }
    a[3][3] = 1.0f;
    torch::Tensor C2W = torch::linalg_inv(Rt);
    auto c = C2W.accessor<float, 2>();
    for (int r = 0; r < 3; ++r) c[r][3] = (c[r][3] + trans[r]) * scale;
    return torch::linalg_inv(C2W);
}

// GaussianKeyframe::computeTransformTensors (:151-184) -> (world_view_transform_, projection_matrix_,
// full_proj_transform_, camera_center_)
inline std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> computeTransformTensors(
    const std::array<float, 9>& R, const std::array<float, 3>& t, float FoVx, float FoVy, float znear = 0.01f,
    float zfar = 100.0f, torch::Device device = torch::kCPU) {
    torch::Tensor world_view_transform = getWorld2View2(R, t).to(device).transpose(0, 1);
    torch::Tensor projection_matrix = getProjectionMatrix(znear, zfar, FoVx, FoVy, device).transpose(0, 1);
    torch::Tensor full_proj_transform = world_view_transform.unsqueeze(0).bmm(projection_matrix.unsqueeze(0)).squeeze(0);
    torch::Tensor camera_center = torch::linalg_inv(world_view_transform).index({3, torch::indexing::Slice(0, 3)});
    return {world_view_transform, projection_matrix, full_proj_transform, camera_center};
}

}  // namespace keyframe_transforms
