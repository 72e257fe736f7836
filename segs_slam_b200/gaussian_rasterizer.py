"""Autograd layer, mirroring the reference's L5 interface
(/root/reference/include/gaussian_rasterizer.h:25-151, src/gaussian_rasterizer.cpp:17-313):
GaussianRasterizationSettings, GaussianRasterizerFunction, GaussianRasterizer with
forward / visible_filter / project2_image / markVisibleGaussians — same names, argument
order, validation messages and gradient order.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import rasterize_points as rp


@dataclass
class GaussianRasterizationSettings:
    """include/gaussian_rasterizer.h:25-55."""
    image_height_: int
    image_width_: int
    tanfovx_: float
    tanfovy_: float
    bg_: torch.Tensor
    scale_modifier_: float
    viewmatrix_: torch.Tensor
    projmatrix_: torch.Tensor
    sh_degree_: int
    campos_: torch.Tensor
    prefiltered_: bool = False


class GaussianRasterizerFunction(torch.autograd.Function):
    """src/gaussian_rasterizer.cpp:27-154."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings: GaussianRasterizationSettings):
        s = raster_settings
        (num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer) = rp.RasterizeGaussiansCUDA(
            s.bg_, means3D, colors_precomp, opacities, scales, rotations, s.scale_modifier_,
            cov3Ds_precomp, s.viewmatrix_, s.projmatrix_, s.tanfovx_, s.tanfovy_, s.image_height_,
            s.image_width_, sh, s.sh_degree_, s.campos_, s.prefiltered_)
        ctx.num_rendered = num_rendered
        ctx.scale_modifier = s.scale_modifier_
        ctx.tanfovx = s.tanfovx_
        ctx.tanfovy = s.tanfovy_
        ctx.sh_degree = s.sh_degree_
        ctx.save_for_backward(s.bg_, s.viewmatrix_, s.projmatrix_, s.campos_, colors_precomp, means3D,
                              scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
                              imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii=None):
        (bg, viewmatrix, projmatrix, campos, colors_precomp, means3D, scales, rotations,
         cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
         dL_drotations) = rp.RasterizeGaussiansBackwardCUDA(
            bg, means3D, radii, colors_precomp, scales, rotations, ctx.scale_modifier, cov3Ds_precomp,
            viewmatrix, projmatrix, ctx.tanfovx, ctx.tanfovy, grad_out_color, sh, ctx.sh_degree, campos,
            geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer)
        # order of src/gaussian_rasterizer.cpp:143-153
        return (dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dopacity, dL_dscales, dL_drotations,
                dL_dcov3D, None)


def rasterizeGaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                       cov3Ds_precomp, raster_settings):
    return GaussianRasterizerFunction.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                            rotations, cov3Ds_precomp, raster_settings)


class GaussianRasterizer(torch.nn.Module):
    """src/gaussian_rasterizer.cpp:17-25, 157-313."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings_ = raster_settings

    def markVisibleGaussians(self, positions):
        with torch.no_grad():
            s = self.raster_settings_
            return rp.markVisible(positions, s.viewmatrix_, s.projmatrix_)

    @staticmethod
    def _empty(device):
        return torch.empty(0, dtype=torch.float32, device=device)

    def forward(self, means3D, means2D, opacities, has_shs, has_colors_precomp, has_scales,
                has_rotations, has_cov3D_precomp, shs, colors_precomp, scales, rotations, cov3D_precomp):
        if (not has_shs and not has_colors_precomp) or (has_shs and has_colors_precomp):
            raise RuntimeError("Please provide excatly one of either SHs or precomputed colors!")
        if ((not has_scales or not has_rotations) and not has_cov3D_precomp) or \
                ((has_scales or has_rotations) and has_cov3D_precomp):
            raise RuntimeError(
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        dev = means3D.device
        if not has_shs:
            shs = self._empty(dev)
        if not has_colors_precomp:
            colors_precomp = self._empty(dev)
        if not has_scales:
            scales = self._empty(dev)
        if not has_rotations:
            rotations = self._empty(dev)
        if not has_cov3D_precomp:
            cov3D_precomp = self._empty(dev)
        color, radii = rasterizeGaussians(means3D, means2D, shs, colors_precomp, opacities, scales,
                                          rotations, cov3D_precomp, self.raster_settings_)
        return color, radii

    def visible_filter(self, means3D, has_scales, has_rotations, has_cov3D_precomp, scales, rotations,
                       cov3D_precomp):
        # the reference wrapper does no XOR validation here (src/gaussian_rasterizer.cpp:210-245);
        # a missing covariance source is reported by the C-ABI status instead
        dev = means3D.device
        if not has_scales:
            scales = self._empty(dev)
        if not has_rotations:
            rotations = self._empty(dev)
        if not has_cov3D_precomp:
            cov3D_precomp = self._empty(dev)
        s = self.raster_settings_
        with torch.no_grad():
            return rp.RasterizeGaussiansfilterCUDA(
                means3D, scales, rotations, s.scale_modifier_, cov3D_precomp, s.viewmatrix_,
                s.projmatrix_, s.tanfovx_, s.tanfovy_, s.image_height_, s.image_width_, s.prefiltered_,
                False)

    def project2_image(self, means3D, means2D, opacities, has_shs, has_colors_precomp, has_scales,
                       has_rotations, has_cov3D_precomp, shs, colors_precomp, scales, rotations,
                       cov3D_precomp):
        if (not has_shs and not has_colors_precomp) or (has_shs and has_colors_precomp):
            raise RuntimeError("Please provide excatly one of either SHs or precomputed colors!")
        if ((not has_scales or not has_rotations) and not has_cov3D_precomp) or \
                ((has_scales or has_rotations) and has_cov3D_precomp):
            raise RuntimeError(
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        dev = means3D.device
        if not has_shs:
            shs = self._empty(dev)
        if not has_colors_precomp:
            colors_precomp = self._empty(dev)
        if not has_scales:
            scales = self._empty(dev)
        if not has_rotations:
            rotations = self._empty(dev)
        if not has_cov3D_precomp:
            cov3D_precomp = self._empty(dev)
        s = self.raster_settings_
        with torch.no_grad():
            return rp.RasterizeGaussiansprojectCUDA(
                s.bg_, means3D, colors_precomp, opacities, scales, rotations, s.scale_modifier_,
                cov3D_precomp, s.viewmatrix_, s.projmatrix_, s.tanfovx_, s.tanfovy_, s.image_height_,
                s.image_width_, shs, s.sh_degree_, s.campos_, s.prefiltered_)
