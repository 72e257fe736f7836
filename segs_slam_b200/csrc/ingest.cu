// Keyframe image ingest for sm_100a.
//
// Replaces, per new keyframe, the OpenCV + ATen sequence of the reference (src/gaussian_mapper.cpp:557-645, 1230-1309):
//   camera.undistortImage(img)              cv::remap(src, dst, map1, map2, INTER_LINEAR)           include/camera.h:106-115
//   cvMat2TorchTensor_Float32(img, device)  from_blob([H,W,3]).clone().to(device).permute({2,0,1})   include/tensor_utils.h:40-69
//                                           .contiguous()
//   cv::cuda::resize(img, level size)       the Gaussian-pyramid levels                             gaussian_mapper.cpp:621-632
// as two kernels: (1) remap + interleaved-to-planar in ONE pass over the uploaded [H,W,C] image (the reference remaps on
// the host, uploads, permutes and copies again), (2) bilinear resize of the planar image to a pyramid level.
//
// Arithmetic follows OpenCV 4.x so that the result can be held to the real library (tests/test_ingest_*.py against cv2):
//   * remap, CV_32FC1 maps, INTER_LINEAR, BORDER_CONSTANT(0): the maps are quantised to 1/32 pixel
//     (sx = cvRound(map_x * 32) — round half to even; ix = sx >> 5, fx = sx & 31) and the four taps are weighted with the
//     FP32 table BilinearTab_f[fy][fx] = {(1-fy/32)(1-fx/32), (1-fy/32)(fx/32), (fy/32)(1-fx/32), (fy/32)(fx/32)}
//     (imgwarp.cpp: initInterTab2D / remapBilinear);
//   * resize, INTER_LINEAR, CV_32F: fx = (float)((dx + 0.5) * (src / dst) - 0.5) in double, sx = floor(fx), clamped at both
//     borders with the weight moved to the surviving tap; horizontal pass first, then vertical (resize.cpp: HResizeLinear /
//     VResizeLinear).
#include "common.cuh"

namespace segs {
namespace {

constexpr int IT = 256;
constexpr int INTER_BITS = 5, INTER_TAB_SIZE = 1 << INTER_BITS;

__global__ void __launch_bounds__(IT)
ingest_kernel(int H, int W, int C, int src_H, int src_W, const float* __restrict__ src, const float* __restrict__ map_x,
              const float* __restrict__ map_y, float* __restrict__ dst)
{
    const size_t e = size_t(blockIdx.x) * IT + threadIdx.x;
    if (e >= size_t(H) * W) return;
    const int x = int(e % W), y = int(e / W);
    const size_t plane = size_t(H) * W;
    if (map_x == nullptr) {                                   // cvMat2TorchTensor_Float32 alone: [H,W,C] -> [C,H,W]
        for (int c = 0; c < C; ++c) dst[c * plane + e] = __ldg(src + e * C + c);
        return;
    }
    const int sx = __float2int_rn(__fmul_rn(__ldg(map_x + e), (float)INTER_TAB_SIZE));      // cvRound: half to even
    const int sy = __float2int_rn(__fmul_rn(__ldg(map_y + e), (float)INTER_TAB_SIZE));
    const int ix = sx >> INTER_BITS, iy = sy >> INTER_BITS;
    const float fx = (float)(sx & (INTER_TAB_SIZE - 1)) * (1.f / INTER_TAB_SIZE), fy = (float)(sy & (INTER_TAB_SIZE - 1)) * (1.f / INTER_TAB_SIZE);
    // initInterTab2D: 2-D table = outer product of the 1-D FP32 tables {1 - f, f}
    const float w00 = __fmul_rn(1.f - fy, 1.f - fx), w01 = __fmul_rn(1.f - fy, fx), w10 = __fmul_rn(fy, 1.f - fx), w11 = __fmul_rn(fy, fx);
    const bool x0 = ix >= 0 && ix < src_W, x1 = ix + 1 >= 0 && ix + 1 < src_W;
    const bool y0 = iy >= 0 && iy < src_H, y1 = iy + 1 >= 0 && iy + 1 < src_H;
    for (int c = 0; c < C; ++c) {
        auto tap = [&](bool ok, int yy, int xx) { return ok ? __ldg(src + (size_t(yy) * src_W + xx) * C + c) : 0.f; };   // BORDER_CONSTANT 0
        const float v00 = tap(y0 && x0, iy, ix), v01 = tap(y0 && x1, iy, ix + 1);
        const float v10 = tap(y1 && x0, iy + 1, ix), v11 = tap(y1 && x1, iy + 1, ix + 1);
        // remapBilinear: S0[0]*w[0] + S0[cn]*w[1] + S1[0]*w[2] + S1[cn]*w[3]
        dst[c * plane + e] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, w00), __fmul_rn(v01, w01)), __fmul_rn(v10, w10)), __fmul_rn(v11, w11));
    }
}

__device__ __forceinline__ void resize_src(int d, double scale, int in_size, int& i0, int& i1, float& a0, float& a1) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= in_size - 1) { s = in_size - 1; f = 0.f; i0 = s; i1 = s; } else { i0 = s; i1 = s + 1; }
    a0 = 1.f - f; a1 = f;
}

__global__ void __launch_bounds__(IT)
resize_kernel(int C, int H, int W, int h, int w, const float* __restrict__ src, float* __restrict__ dst)
{
    const size_t e = size_t(blockIdx.x) * IT + threadIdx.x;
    if (e >= size_t(C) * h * w) return;
    const int x = int(e % w), y = int((e / w) % h), c = int(e / (size_t(w) * h));
    int x0, x1, y0, y1; float a0, a1, b0, b1;
    resize_src(x, (double)W / w, W, x0, x1, a0, a1);
    resize_src(y, (double)H / h, H, y0, y1, b0, b1);
    const float* p = src + size_t(c) * H * W;
    const float r0 = __fadd_rn(__fmul_rn(p[size_t(y0) * W + x0], a0), __fmul_rn(p[size_t(y0) * W + x1], a1));   // HResizeLinear
    const float r1 = __fadd_rn(__fmul_rn(p[size_t(y1) * W + x0], a0), __fmul_rn(p[size_t(y1) * W + x1], a1));
    dst[e] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));                                                   // VResizeLinear
}

}  // namespace
}  // namespace segs

using namespace segs;

extern "C" {

int segs_ingest_image(int H, int W, int C, int src_H, int src_W, const float* src_hwc, const float* map_x, const float* map_y,
                      float* dst_chw, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (H <= 0 || W <= 0 || C <= 0 || C > 4 || src_H <= 0 || src_W <= 0 || !src_hwc || !dst_chw || ((map_x == nullptr) != (map_y == nullptr))) {
        set_error("ingest: invalid argument"); return SEGS_ERR_INVALID_ARG;
    }
    if (!map_x && (src_H != H || src_W != W)) { set_error("ingest: without maps the source must have the output size"); return SEGS_ERR_INVALID_ARG; }
    ingest_kernel<<<int((size_t(H) * W + IT - 1) / IT), IT, 0, stream>>>(H, W, C, src_H, src_W, src_hwc, map_x, map_y, dst_chw);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int segs_resize_bilinear(int C, int H, int W, const float* src_chw, int h, int w, float* dst_chw, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (C <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0 || !src_chw || !dst_chw) { set_error("resize: invalid argument"); return SEGS_ERR_INVALID_ARG; }
    resize_kernel<<<int((size_t(C) * h * w + IT - 1) / IT), IT, 0, stream>>>(C, H, W, h, w, src_chw, dst_chw);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // extern "C"
