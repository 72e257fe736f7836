// LibTorch host layer over the C-ABI — see rasterize_points.h.  Replaces
// /root/reference/src/rasterize_points.cu and third_party/simple-knn/spatial.cu.
//
// torch is used for device memory and the current stream only; every kernel lives behind
// include/segs_raster.h.  Differences from the reference that a caller can observe:
//   * work is queued on at::cuda::getCurrentCUDAStream() (the reference uses the legacy default
//     stream, which is the same stream under LibTorch's defaults);
//   * CUDA failures surface here as c10::Error with the library's message instead of as a later
//     sticky error (the reference checks nothing);
//   * output tensors are torch::empty where the kernels write every element (the reference
//     zero-fills 9 gradient tensors + 2 outputs per call and then overwrites them).
#include "rasterize_points.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <stdexcept>
#include <string>

#include "../../../include/segs_raster.h"

namespace {

constexpr int kChannels = 3;   // NUM_CHANNELS, cuda_rasterizer/config.h:15

// "absent" optional input: the reference hands data_ptr() of a 0-element tensor to the kernels
// (src/gaussian_rasterizer.cpp:183-193); the C-ABI takes NULL.
const float* fptr(const torch::Tensor& t) { return t.numel() == 0 ? nullptr : t.data_ptr<float>(); }

torch::Tensor dense(const torch::Tensor& t) { return t.contiguous(); }

// std::function<char*(size_t)> resizeFunctional(t) of src/rasterize_points.cu:28-34 as a C callback
char* grow(void* user, size_t bytes) {
    auto* t = static_cast<torch::Tensor*>(user);
    t->resize_({static_cast<int64_t>(bytes)});
    return reinterpret_cast<char*>(t->data_ptr());
}

void* current_stream() { return static_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

void raise_if(int status) {
    if (status == SEGS_OK) return;
    const std::string msg = segs_last_error();
    // the reference throws std::runtime_error for the "exactly one of" checks
    // (rasterizer_impl.cu:241-244, gaussian_rasterizer.cpp:172-179)
    if (status == SEGS_ERR_INVALID_ARG && msg.rfind("Please provide", 0) == 0) throw std::runtime_error(msg);
    TORCH_CHECK(false, "segs_raster: ", msg, " (status ", status, ")");
}

void check_means(const torch::Tensor& means3D) {
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
        AT_ERROR("means3D must have dimensions (num_points, 3)");   // src/rasterize_points.cu:57-59
    }
    TORCH_CHECK(means3D.is_cuda(), "segs_raster has no CPU path: tensors must live on a CUDA device");
}

}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                       const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                       const float tan_fovx, const float tan_fovy, const int image_height,
                       const int image_width, const torch::Tensor& sh, const int degree,
                       const torch::Tensor& campos, const bool prefiltered)
{
    check_means(means3D);
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = static_cast<int>(means3D.size(0));
    const int H = image_height, W = image_width;
    const auto f32 = means3D.options().dtype(torch::kFloat32);
    const auto u8 = means3D.options().dtype(torch::kByte);

    // P == 0: the reference returns the zero image untouched (src/rasterize_points.cu:81)
    torch::Tensor out_color = P ? torch::empty({kChannels, H, W}, f32) : torch::zeros({kChannels, H, W}, f32);
    torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
    torch::Tensor geomBuffer = torch::empty({0}, u8);
    torch::Tensor binningBuffer = torch::empty({0}, u8);
    torch::Tensor imgBuffer = torch::empty({0}, u8);

    int rendered = 0;
    if (P != 0) {
        const int M = sh.numel() != 0 ? static_cast<int>(sh.size(1)) : 0;
        const auto bg = dense(background), m3 = dense(means3D), shc = dense(sh), col = dense(colors),
                   opa = dense(opacity), sca = dense(scales), rot = dense(rotations),
                   cov = dense(cov3D_precomp), view = dense(viewmatrix), proj = dense(projmatrix),
                   cam = dense(campos);
        raise_if(segs_raster_forward(grow, &geomBuffer, grow, &binningBuffer, grow, &imgBuffer, P, degree, M,
                                     fptr(bg), W, H, fptr(m3), fptr(shc), fptr(col), fptr(opa), fptr(sca),
                                     scale_modifier, fptr(rot), fptr(cov), fptr(view), fptr(proj), fptr(cam),
                                     tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color.data_ptr<float>(),
                                     radii.data_ptr<int>(), &rendered, current_stream()));
    }
    return std::make_tuple(rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer);
}

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
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer)
{
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = static_cast<int>(means3D.size(0));
    const int H = static_cast<int>(dL_dout_color.size(1));
    const int W = static_cast<int>(dL_dout_color.size(2));
    const int M = sh.numel() != 0 ? static_cast<int>(sh.size(1)) : 0;
    const auto f32 = means3D.options().dtype(torch::kFloat32);

    // every element is written by the kernels (zeros for Gaussians that were not rendered)
    torch::Tensor dL_dmeans3D = torch::empty({P, 3}, f32);
    torch::Tensor dL_dmeans2D = torch::empty({P, 3}, f32);
    torch::Tensor dL_dcolors = torch::empty({P, kChannels}, f32);
    torch::Tensor dL_dopacity = torch::empty({P, 1}, f32);
    torch::Tensor dL_dcov3D = torch::empty({P, 6}, f32);
    torch::Tensor dL_dsh = torch::empty({P, M, 3}, f32);
    torch::Tensor dL_dscales = torch::empty({P, 3}, f32);
    torch::Tensor dL_drotations = torch::empty({P, 4}, f32);

    if (P != 0) {
        const auto bg = dense(background), m3 = dense(means3D), shc = dense(sh), col = dense(colors),
                   sca = dense(scales), rot = dense(rotations), cov = dense(cov3D_precomp),
                   view = dense(viewmatrix), proj = dense(projmatrix), cam = dense(campos),
                   dpix = dense(dL_dout_color), rad = dense(radii);
        auto bytes = [](const torch::Tensor& t) {
            return t.numel() == 0 ? nullptr : reinterpret_cast<char*>(t.data_ptr());
        };
        raise_if(segs_raster_backward(
            P, degree, M, R, fptr(bg), W, H, fptr(m3), fptr(shc), fptr(col), fptr(sca), scale_modifier,
            fptr(rot), fptr(cov), fptr(view), fptr(proj), fptr(cam), tan_fovx, tan_fovy, rad.data_ptr<int>(),
            bytes(geomBuffer), bytes(binningBuffer), bytes(imageBuffer), fptr(dpix),
            dL_dmeans2D.data_ptr<float>(), /*dL_dconic (internal to the reference)*/ nullptr,
            dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(), dL_dmeans3D.data_ptr<float>(),
            dL_dcov3D.data_ptr<float>(), M ? dL_dsh.data_ptr<float>() : nullptr, dL_dscales.data_ptr<float>(),
            dL_drotations.data_ptr<float>(), current_stream()));
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
                           dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix)
{
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = static_cast<int>(means3D.size(0));
    torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
    if (P != 0) {
        const auto m3 = dense(means3D), view = dense(viewmatrix), proj = dense(projmatrix);
        raise_if(segs_mark_visible(P, fptr(m3), fptr(view), fptr(proj),
                                   reinterpret_cast<unsigned char*>(present.data_ptr<bool>()), current_stream()));
    }
    return present;
}

torch::Tensor RasterizeGaussiansfilterCUDA(const torch::Tensor& means3D, const torch::Tensor& scales,
                                           const torch::Tensor& rotations, const float scale_modifier,
                                           const torch::Tensor& cov3D_precomp,
                                           const torch::Tensor& viewmatrix,
                                           const torch::Tensor& projmatrix, const float tan_fovx,
                                           const float tan_fovy, const int image_height,
                                           const int image_width, const bool prefiltered,
                                           const bool /*debug*/)
{
    check_means(means3D);
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = static_cast<int>(means3D.size(0));
    torch::Tensor radii = torch::full({P}, 0, means3D.options().dtype(torch::kInt32));
    if (P != 0) {
        const auto m3 = dense(means3D), sca = dense(scales), rot = dense(rotations), cov = dense(cov3D_precomp),
                   view = dense(viewmatrix), proj = dense(projmatrix);
        raise_if(segs_visible_filter(P, 0, image_width, image_height, fptr(m3), fptr(sca), scale_modifier,
                                     fptr(rot), fptr(cov), fptr(view), fptr(proj), tan_fovx, tan_fovy,
                                     prefiltered ? 1 : 0, radii.data_ptr<int>(), current_stream()));
    }
    return radii;
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansprojectCUDA(const torch::Tensor& /*background*/, const torch::Tensor& means3D,
                              const torch::Tensor& colors, const torch::Tensor& opacity,
                              const torch::Tensor& scales, const torch::Tensor& rotations,
                              const float scale_modifier, const torch::Tensor& cov3D_precomp,
                              const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                              const float tan_fovx, const float tan_fovy, const int image_height,
                              const int image_width, const torch::Tensor& sh, const int degree,
                              const torch::Tensor& campos, const bool prefiltered)
{
    check_means(means3D);
    const c10::cuda::CUDAGuard guard(means3D.device());
    const int P = static_cast<int>(means3D.size(0));
    const auto f32 = means3D.options().dtype(torch::kFloat32);
    torch::Tensor out_color = torch::full({P, kChannels}, 0.0, f32);
    torch::Tensor radii = torch::full({P}, 0, means3D.options().dtype(torch::kInt32));
    torch::Tensor points_image = torch::full({P, 2}, 0.0, f32);
    if (P != 0) {
        const int M = sh.numel() != 0 ? static_cast<int>(sh.size(1)) : 0;
        const auto m3 = dense(means3D), shc = dense(sh), col = dense(colors), opa = dense(opacity),
                   sca = dense(scales), rot = dense(rotations), cov = dense(cov3D_precomp),
                   view = dense(viewmatrix), proj = dense(projmatrix), cam = dense(campos);
        raise_if(segs_project(P, degree, M, image_width, image_height, fptr(m3), fptr(shc), fptr(col), fptr(opa),
                              fptr(sca), scale_modifier, fptr(rot), fptr(cov), fptr(view), fptr(proj), fptr(cam),
                              tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color.data_ptr<float>(),
                              points_image.data_ptr<float>(), radii.data_ptr<int>(), current_stream()));
    }
    return std::make_tuple(points_image, radii, out_color);
}

torch::Tensor distCUDA2(const torch::Tensor& points)
{
    TORCH_CHECK(points.is_cuda(), "segs_raster has no CPU path: tensors must live on a CUDA device");
    const c10::cuda::CUDAGuard guard(points.device());
    const int P = static_cast<int>(points.size(0));
    torch::Tensor means = torch::full({P}, 0.0, points.options().dtype(torch::kFloat32));
    if (P != 0) {
        const auto pts = dense(points);
        // all temporaries come from one caller-owned scratch tensor (the reference cudaMallocs
        // and frees inside SimpleKNN::knn on every call, simple_knn.cu:187-220)
        torch::Tensor scratch = torch::empty({0}, points.options().dtype(torch::kByte));
        raise_if(segs_knn_mean_dist2(P, fptr(pts), means.data_ptr<float>(), grow, &scratch, current_stream()));
    }
    return means;
}
