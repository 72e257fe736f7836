// L5 twin: the autograd layer that sits on top of rasterize_points.h, with the interface of
// /root/reference/include/gaussian_rasterizer.h:25-160 (GaussianRasterizationSettings,
// GaussianRasterizerFunction, rasterizeGaussians, GaussianRasterizer).
//
// A SEGS-SLAM build keeps ITS OWN gaussian_rasterizer.{h,cpp} unchanged — they only call the six
// functions of rasterize_points.h.  This twin exists because the reference header pulls in
// gaussian_model.h (PCL, Sophus, torch_scatter ...: none of which are in this image), so the
// reference's L5 cannot be compiled here; the parity tests drive the kernels through this file
// so that the C++ autograd path (saved tensors, gradient order) is exercised end to end.
#pragma once
#include <torch/torch.h>

#include <tuple>

#include "rasterize_points.h"

struct GaussianRasterizationSettings {
    GaussianRasterizationSettings(int image_height, int image_width, float tanfovx, float tanfovy,
                                  torch::Tensor& bg, float scale_modifier, torch::Tensor& viewmatrix,
                                  torch::Tensor& projmatrix, int sh_degree, torch::Tensor& campos,
                                  bool prefiltered)
        : image_height_(image_height), image_width_(image_width), tanfovx_(tanfovx), tanfovy_(tanfovy),
          bg_(bg), scale_modifier_(scale_modifier), viewmatrix_(viewmatrix), projmatrix_(projmatrix),
          sh_degree_(sh_degree), campos_(campos), prefiltered_(prefiltered) {}

    int image_height_;
    int image_width_;
    float tanfovx_;
    float tanfovy_;
    torch::Tensor bg_;
    float scale_modifier_;
    torch::Tensor viewmatrix_;
    torch::Tensor projmatrix_;
    int sh_degree_;
    torch::Tensor campos_;
    bool prefiltered_;
};

class GaussianRasterizerFunction : public torch::autograd::Function<GaussianRasterizerFunction> {
public:
    // -> {color[3,H,W], radii[P]}
    static torch::autograd::tensor_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor means3D,
                                                torch::Tensor means2D, torch::Tensor sh,
                                                torch::Tensor colors_precomp, torch::Tensor opacities,
                                                torch::Tensor scales, torch::Tensor rotations,
                                                torch::Tensor cov3Ds_precomp,
                                                GaussianRasterizationSettings raster_settings);
    // -> grads of (means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
    //    cov3Ds_precomp, settings = undefined)   (src/gaussian_rasterizer.cpp:143-153)
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::tensor_list grad_outputs);
};

inline torch::autograd::tensor_list rasterizeGaussians(torch::Tensor& means3D, torch::Tensor& means2D,
                                                       torch::Tensor& sh, torch::Tensor& colors_precomp,
                                                       torch::Tensor& opacities, torch::Tensor& scales,
                                                       torch::Tensor& rotations, torch::Tensor& cov3Ds_precomp,
                                                       GaussianRasterizationSettings& raster_settings)
{
    return GaussianRasterizerFunction::apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                             cov3Ds_precomp, raster_settings);
}

class GaussianRasterizer : public torch::nn::Module {
public:
    explicit GaussianRasterizer(GaussianRasterizationSettings& raster_settings) : raster_settings_(raster_settings) {}

    torch::Tensor markVisibleGaussians(torch::Tensor& positions);

    std::tuple<torch::Tensor, torch::Tensor> forward(torch::Tensor means3D, torch::Tensor means2D,
                                                     torch::Tensor opacities, bool has_shs,
                                                     bool has_colors_precomp, bool has_scales,
                                                     bool has_rotations, bool has_cov3D_precomp,
                                                     torch::Tensor shs, torch::Tensor colors_precomp,
                                                     torch::Tensor scales, torch::Tensor rotations,
                                                     torch::Tensor cov3D_precomp);

    torch::Tensor visible_filter(torch::Tensor means3D, bool has_scales, bool has_rotations,
                                 bool has_cov3D_precomp, torch::Tensor scales, torch::Tensor rotations,
                                 torch::Tensor cov3D_precomp);

    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> project2_image(
        torch::Tensor means3D, torch::Tensor means2D, torch::Tensor opacities, bool has_shs,
        bool has_colors_precomp, bool has_scales, bool has_rotations, bool has_cov3D_precomp, torch::Tensor shs,
        torch::Tensor colors_precomp, torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3D_precomp);

public:
    GaussianRasterizationSettings raster_settings_;
};
