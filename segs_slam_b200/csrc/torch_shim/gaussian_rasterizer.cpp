// L5 twin — see gaussian_rasterizer.h.  Behaviour follows
// /root/reference/src/gaussian_rasterizer.cpp:18-330.
#include "gaussian_rasterizer.h"

#include <stdexcept>

namespace {

// which optional inputs a call carries; mirrors the has_* flags of GaussianRasterizer::forward
struct Presence {
    bool shs, colors, scales, rotations, cov3D;
};

void require_exactly_one_colour_source(const Presence& p) {
    if (p.shs == p.colors)   // both or neither (src/gaussian_rasterizer.cpp:172-174)
        throw std::runtime_error("Please provide excatly one of either SHs or precomputed colors!");
}

void require_exactly_one_covariance_source(const Presence& p) {
    const bool pair = p.scales && p.rotations, any = p.scales || p.rotations;
    if ((!pair && !p.cov3D) || (any && p.cov3D))   // src/gaussian_rasterizer.cpp:176-179
        throw std::runtime_error(
            "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
}

// an absent input travels as a 0-element CUDA tensor (src/gaussian_rasterizer.cpp:181-193)
void blank_if_absent(bool present, torch::Tensor& t) {
    if (!present) t = torch::tensor({}, torch::TensorOptions().device(torch::kCUDA));
}

}  // namespace

torch::Tensor GaussianRasterizer::markVisibleGaussians(torch::Tensor& positions)
{
    torch::NoGradGuard no_grad;
    return markVisible(positions, raster_settings_.viewmatrix_, raster_settings_.projmatrix_);
}

torch::autograd::tensor_list GaussianRasterizerFunction::forward(
    torch::autograd::AutogradContext* ctx, torch::Tensor means3D, torch::Tensor /*means2D*/, torch::Tensor sh,
    torch::Tensor colors_precomp, torch::Tensor opacities, torch::Tensor scales, torch::Tensor rotations,
    torch::Tensor cov3Ds_precomp, GaussianRasterizationSettings s)
{
    auto [num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer] = RasterizeGaussiansCUDA(
        s.bg_, means3D, colors_precomp, opacities, scales, rotations, s.scale_modifier_, cov3Ds_precomp,
        s.viewmatrix_, s.projmatrix_, s.tanfovx_, s.tanfovy_, s.image_height_, s.image_width_, sh, s.sh_degree_,
        s.campos_, s.prefiltered_);

    ctx->saved_data["num_rendered"] = num_rendered;
    ctx->saved_data["scale_modifier"] = s.scale_modifier_;
    ctx->saved_data["tanfovx"] = s.tanfovx_;
    ctx->saved_data["tanfovy"] = s.tanfovy_;
    ctx->saved_data["sh_degree"] = s.sh_degree_;
    // same 14 tensors, same order as src/gaussian_rasterizer.cpp:73-86
    ctx->save_for_backward({s.bg_, s.viewmatrix_, s.projmatrix_, s.campos_, colors_precomp, means3D, scales,
                            rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer});
    return {color, radii};
}

torch::autograd::tensor_list GaussianRasterizerFunction::backward(torch::autograd::AutogradContext* ctx,
                                                                  torch::autograd::tensor_list grad_outputs)
{
    const int num_rendered = static_cast<int>(ctx->saved_data["num_rendered"].toInt());
    const float scale_modifier = static_cast<float>(ctx->saved_data["scale_modifier"].toDouble());
    const float tanfovx = static_cast<float>(ctx->saved_data["tanfovx"].toDouble());
    const float tanfovy = static_cast<float>(ctx->saved_data["tanfovy"].toDouble());
    const int sh_degree = static_cast<int>(ctx->saved_data["sh_degree"].toInt());
    const auto v = ctx->get_saved_variables();
    enum { BG, VIEW, PROJ, CAMPOS, COLORS, MEANS3D, SCALES, ROTS, COV3D, RADII, SH, GEOM, BINNING, IMG };

    // only the colour gradient is consumed (src/gaussian_rasterizer.cpp:119)
    auto [dmeans2D, dcolors, dopacity, dmeans3D, dcov3D, dsh, dscales, drotations] =
        RasterizeGaussiansBackwardCUDA(v[BG], v[MEANS3D], v[RADII], v[COLORS], v[SCALES], v[ROTS], scale_modifier,
                                       v[COV3D], v[VIEW], v[PROJ], tanfovx, tanfovy, grad_outputs[0], v[SH],
                                       sh_degree, v[CAMPOS], v[GEOM], num_rendered, v[BINNING], v[IMG]);
    // forward argument order: means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
    // cov3Ds_precomp, raster_settings
    return {dmeans3D, dmeans2D, dsh, dcolors, dopacity, dscales, drotations, dcov3D, torch::Tensor()};
}

std::tuple<torch::Tensor, torch::Tensor> GaussianRasterizer::forward(
    torch::Tensor means3D, torch::Tensor means2D, torch::Tensor opacities, bool has_shs, bool has_colors_precomp,
    bool has_scales, bool has_rotations, bool has_cov3D_precomp, torch::Tensor shs, torch::Tensor colors_precomp,
    torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3D_precomp)
{
    const Presence p{has_shs, has_colors_precomp, has_scales, has_rotations, has_cov3D_precomp};
    require_exactly_one_colour_source(p);
    require_exactly_one_covariance_source(p);
    blank_if_absent(has_shs, shs);
    blank_if_absent(has_colors_precomp, colors_precomp);
    blank_if_absent(has_scales, scales);
    blank_if_absent(has_rotations, rotations);
    blank_if_absent(has_cov3D_precomp, cov3D_precomp);
    auto out = rasterizeGaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                  cov3D_precomp, raster_settings_);
    return std::make_tuple(out[0], out[1]);
}

torch::Tensor GaussianRasterizer::visible_filter(torch::Tensor means3D, bool has_scales, bool has_rotations,
                                                 bool has_cov3D_precomp, torch::Tensor scales,
                                                 torch::Tensor rotations, torch::Tensor cov3D_precomp)
{
    blank_if_absent(has_scales, scales);
    blank_if_absent(has_rotations, rotations);
    blank_if_absent(has_cov3D_precomp, cov3D_precomp);
    torch::NoGradGuard no_grad;
    const auto& s = raster_settings_;
    return RasterizeGaussiansfilterCUDA(means3D, scales, rotations, s.scale_modifier_, cov3D_precomp, s.viewmatrix_,
                                        s.projmatrix_, s.tanfovx_, s.tanfovy_, s.image_height_, s.image_width_,
                                        s.prefiltered_, false);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> GaussianRasterizer::project2_image(
    torch::Tensor means3D, torch::Tensor /*means2D*/, torch::Tensor opacities, bool has_shs,
    bool has_colors_precomp, bool has_scales, bool has_rotations, bool has_cov3D_precomp, torch::Tensor shs,
    torch::Tensor colors_precomp, torch::Tensor scales, torch::Tensor rotations, torch::Tensor cov3D_precomp)
{
    const Presence p{has_shs, has_colors_precomp, has_scales, has_rotations, has_cov3D_precomp};
    require_exactly_one_colour_source(p);
    require_exactly_one_covariance_source(p);
    blank_if_absent(has_shs, shs);
    blank_if_absent(has_colors_precomp, colors_precomp);
    blank_if_absent(has_scales, scales);
    blank_if_absent(has_rotations, rotations);
    blank_if_absent(has_cov3D_precomp, cov3D_precomp);
    const auto& s = raster_settings_;
    return RasterizeGaussiansprojectCUDA(s.bg_, means3D, colors_precomp, opacities, scales, rotations,
                                         s.scale_modifier_, cov3D_precomp, s.viewmatrix_, s.projmatrix_, s.tanfovx_,
                                         s.tanfovy_, s.image_height_, s.image_width_, shs, s.sh_degree_, s.campos_,
                                         s.prefiltered_);
}
