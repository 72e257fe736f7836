"""segs_slam_b200 — B200-native (sm_100a) drop-in for SEGS-SLAM's Gaussian rasterizer hot path.

Layout:
  csrc/                CUDA kernels + the C-ABI (include/segs_raster.h) -> libsegs_raster.so
  rasterize_points.py  L6 mirror: RasterizeGaussiansCUDA / ...BackwardCUDA / ...filterCUDA /
                       ...projectCUDA / markVisible / distCUDA2
  gaussian_rasterizer.py  L5 mirror: GaussianRasterizationSettings, GaussianRasterizer(+Function)
  gaussian_renderer.py L4 mirror: generate_neural_gaussians (fused anchor decode, autograd)
  synth.py             the synthetic scenes of BASELINE.md §3
"""
from .gaussian_rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                  GaussianRasterizerFunction, rasterizeGaussians)
from .rasterize_points import (RasterizeGaussiansBackwardCUDA, RasterizeGaussiansCUDA,  # noqa: F401
                               RasterizeGaussiansfilterCUDA, RasterizeGaussiansprojectCUDA, distCUDA2,
                               markVisible)
from .gaussian_renderer import generate_neural_gaussians  # noqa: F401
