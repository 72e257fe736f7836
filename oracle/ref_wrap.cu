// TEST INFRASTRUCTURE — not part of the product path.
//
// C-ABI veneer over the UNMODIFIED reference CUDA rasterizer and simple-knn, which
// oracle/Makefile compiles from the sources where they lie under /root/reference
// (cuda_rasterizer/{forward,backward,rasterizer_impl}.cu, third_party/simple-knn/
// simple_knn.cu) into oracle/_ref/libsegs_ref.so.  This file holds none of the
// reference's code: it only forwards to the reference's public C++ entry points
//   CudaRasterizer::Rasterizer::{forward,backward,visible_filter,markVisible}
//     (/root/reference/cuda_rasterizer/rasterizer.h:20-126)
//   SimpleKNN::knn (/root/reference/third_party/simple-knn/simple_knn.h:15-19)
// so that tests/ and bench.py can drive the reference from Python (ctypes) with the
// same allocation-callback contract as the product's C-ABI (include/segs_raster.h).
#include <cstddef>
#include <cstdint>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer.h"
#include "third_party/simple-knn/simple_knn.h"

typedef char* (*ref_alloc_fn)(void* ctx, size_t bytes);

static std::function<char*(size_t)> wrap(ref_alloc_fn fn, void* ctx) {
    return [fn, ctx](size_t n) { return fn(ctx, n); };
}

extern "C" {

int ref_raster_forward(
    ref_alloc_fn geom_alloc, void* geom_ctx,
    ref_alloc_fn binning_alloc, void* binning_ctx,
    ref_alloc_fn img_alloc, void* img_ctx,
    int P, int D, int M,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, int* radii)
{
    return CudaRasterizer::Rasterizer::forward(
        wrap(geom_alloc, geom_ctx), wrap(binning_alloc, binning_ctx), wrap(img_alloc, img_ctx),
        P, D, M, background, width, height, means3D, shs, colors_precomp, opacities,
        scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
        tan_fovx, tan_fovy, prefiltered != 0, out_color, radii);
}

void ref_raster_backward(
    int P, int D, int M, int R,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* campos, float tan_fovx, float tan_fovy, const int* radii,
    char* geom_buffer, char* binning_buffer, char* image_buffer,
    const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
    float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot)
{
    CudaRasterizer::Rasterizer::backward(
        P, D, M, R, background, width, height, means3D, shs, colors_precomp, scales,
        scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
        tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix,
        dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
        dL_dscale, dL_drot);
}

void ref_visible_filter(
    ref_alloc_fn geom_alloc, void* geom_ctx,
    ref_alloc_fn binning_alloc, void* binning_ctx,
    ref_alloc_fn img_alloc, void* img_ctx,
    int P, int M, int width, int height,
    const float* means3D, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix,
    float tan_fovx, float tan_fovy, int prefiltered, int* radii)
{
    CudaRasterizer::Rasterizer::visible_filter(
        wrap(geom_alloc, geom_ctx), wrap(binning_alloc, binning_ctx), wrap(img_alloc, img_ctx),
        P, M, width, height, means3D, scales, scale_modifier, rotations, cov3D_precomp,
        viewmatrix, projmatrix, tan_fovx, tan_fovy, prefiltered != 0, radii, false);
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

void ref_knn_mean_dist2(int P, float* points, float* mean_dists)
{
    SimpleKNN::knn(P, reinterpret_cast<float3*>(points), mean_dists);
}

int ref_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
