"""segs_slam_b200 — B200-native (sm_100a) drop-in for SEGS-SLAM's Gaussian rasterizer hot path.

Layout:
  csrc/                CUDA kernels + the C-ABI (include/segs_raster.h) -> libsegs_raster.so
  rasterize_points.py  L6 mirror: RasterizeGaussiansCUDA / ...BackwardCUDA / ...filterCUDA /
                       ...projectCUDA / markVisible / distCUDA2
  gaussian_rasterizer.py  L5 mirror: GaussianRasterizationSettings, GaussianRasterizer(+Function)
  gaussian_renderer.py L4 mirror: generate_neural_gaussians (fused anchor decode, autograd)
  loss_utils.py        the reference's `loss_utils` namespace: l1_loss / ssim / psnr (fused kernels), the mapper's
                       fused L1+SSIM loss and scaling regulariser, the frequency terms (host compositions over cuFFT)
  optim.py             FusedAdam: one launch per step over the flat gradient bucket
  mapper.py            the keyframe-batched data-parallel mapping step: mapping_step (autograd composition),
                       FusedMapper (segs_mapper_views: views issued from C++ on concurrent lanes), RasterBatch
  checkpoint.py        the reference's anchor checkpoint format: tinyply-compatible PLY + MLP text files
  anchor_model.py      container with the reference GaussianModel's member names + the C3 / C4 synthetic configs
  synth.py             the synthetic scenes of BASELINE.md §3
"""
from .gaussian_rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                  GaussianRasterizerFunction, rasterizeGaussians)
from .rasterize_points import (RasterizeGaussiansBackwardCUDA, RasterizeGaussiansCUDA,  # noqa: F401
                               RasterizeGaussiansfilterCUDA, RasterizeGaussiansprojectCUDA, distCUDA2,
                               markVisible)
from .gaussian_renderer import generate_neural_gaussians  # noqa: F401
