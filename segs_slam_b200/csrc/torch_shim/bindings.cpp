// pybind11 surface of the LibTorch host layer, so the Python test-suite can drive the kernels
// through the SAME C++ entry points a SEGS-SLAM build links (rasterize_points.h) and through the
// C++ autograd function (gaussian_rasterizer.h).  Built in-tree as segs_slam_b200/_segs_torch.so.
#include <torch/extension.h>

#include "gaussian_rasterizer.h"
#include "rasterize_points.h"

namespace {

// GaussianRasterizer::forward of the L5 twin with the has_* flags derived from numel(), returning
// (color, radii); autograd flows through GaussianRasterizerFunction.
std::tuple<torch::Tensor, torch::Tensor> rasterizer_forward(
    int H, int W, double tanfovx, double tanfovy, torch::Tensor bg, double scale_modifier, torch::Tensor viewmatrix,
    torch::Tensor projmatrix, int sh_degree, torch::Tensor campos, bool prefiltered, torch::Tensor means3D,
    torch::Tensor means2D, torch::Tensor opacities, torch::Tensor shs, torch::Tensor colors_precomp,
    torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3D_precomp)
{
    GaussianRasterizationSettings settings(H, W, static_cast<float>(tanfovx), static_cast<float>(tanfovy), bg,
                                           static_cast<float>(scale_modifier), viewmatrix, projmatrix, sh_degree,
                                           campos, prefiltered);
    GaussianRasterizer rasterizer(settings);
    return rasterizer.forward(means3D, means2D, opacities, shs.numel() != 0, colors_precomp.numel() != 0,
                              scales.numel() != 0, rotations.numel() != 0, cov3D_precomp.numel() != 0, shs,
                              colors_precomp, scales, rotations, cov3D_precomp);
}

torch::Tensor mark_visible(torch::Tensor means3D, torch::Tensor viewmatrix, torch::Tensor projmatrix) {
    return markVisible(means3D, viewmatrix, projmatrix);
}

}  // namespace

PYBIND11_MODULE(_segs_torch, m) {
    m.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
    m.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
    m.def("markVisible", &mark_visible);
    m.def("RasterizeGaussiansfilterCUDA", &RasterizeGaussiansfilterCUDA);
    m.def("RasterizeGaussiansprojectCUDA", &RasterizeGaussiansprojectCUDA);
    m.def("distCUDA2", &distCUDA2);
    m.def("rasterizer_forward", &rasterizer_forward);
}
