// LibTorch host layer over the C-ABI (include/segs_raster.h): the drop-in replacement for the
// reference's tensor-level entry points.  A SEGS-SLAM build links THIS translation unit (plus
// libsegs_raster.so) instead of src/rasterize_points.cu + cuda_rasterizer/*.cu +
// third_party/simple-knn/*; GaussianRasterizer / GaussianRenderer / GaussianModel /
// GaussianMapper compile and run unchanged because names, argument lists, tuple orders and
// error behaviour are those of
//   /root/reference/include/rasterize_points.h:18-102   (five rasterizer entry points)
//   /root/reference/third_party/simple-knn/spatial.h:14 (distCUDA2)
#pragma once
#include <torch/torch.h>

#include <tuple>

// -> (rendered, out_color[3,H,W], radii[P] i32, geomBuffer u8, binningBuffer u8, imgBuffer u8)
//    (src/rasterize_points.cu:36-114)
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                       const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                       const float tan_fovx, const float tan_fovy, const int image_height,
                       const int image_width, const torch::Tensor& sh, const int degree,
                       const torch::Tensor& campos, const bool prefiltered);

// -> (dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dopacity[P,1], dL_dmeans3D[P,3], dL_dcov3D[P,6],
//     dL_dsh[P,M,3], dL_dscales[P,3], dL_drotations[P,4])   (src/rasterize_points.cu:116-193)
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                               const torch::Tensor& radii, const torch::Tensor& colors,
                               const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp,
                               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                               const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& sh,
                               const int degree, const torch::Tensor& campos,
                               const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer);

// -> bool[P]   (src/rasterize_points.cu:195-214)
torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix);

// -> radii[P] i32, > 0 where visible   (src/rasterize_points.cu:216-276)
torch::Tensor RasterizeGaussiansfilterCUDA(const torch::Tensor& means3D, const torch::Tensor& scales,
                                           const torch::Tensor& rotations, const float scale_modifier,
                                           const torch::Tensor& cov3D_precomp,
                                           const torch::Tensor& viewmatrix,
                                           const torch::Tensor& projmatrix, const float tan_fovx,
                                           const float tan_fovy, const int image_height,
                                           const int image_width, const bool prefiltered,
                                           const bool debug);

// -> (points_image[P,2], radii[P] i32, out_color[P,3])   (src/rasterize_points.cu:278-362)
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansprojectCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                              const torch::Tensor& colors, const torch::Tensor& opacity,
                              const torch::Tensor& scales, const torch::Tensor& rotations,
                              const float scale_modifier, const torch::Tensor& cov3D_precomp,
                              const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                              const float tan_fovx, const float tan_fovy, const int image_height,
                              const int image_width, const torch::Tensor& sh, const int degree,
                              const torch::Tensor& campos, const bool prefiltered);

// -> float[P]: mean squared distance to the 3 nearest neighbours (simple-knn/spatial.cu:16-25)
torch::Tensor distCUDA2(const torch::Tensor& points);
