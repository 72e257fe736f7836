// Per-Gaussian backward preprocess for sm_100a.
//
// Replaces computeCov2DCUDA + preprocessCUDA(backward) + computeCov3D(backward) +
// computeColorFromSH(backward) of the reference (cuda_rasterizer/backward.cu:20-139,
// 144-274, 278-341, 347-396) and the nine torch::zeros gradient fills of
// RasterizeGaussiansBackwardCUDA (src/rasterize_points.cu:146-154).
//
// Design (B200): ONE pass over P instead of two kernels + nine memsets.  Each thread reads
// its Gaussian's packed 48-byte gradient accumulator (written by the blend backward),
// runs the conic -> cov2D -> cov3D/mean -> scale/rotation chain in registers and writes
// every output gradient exactly once (zeros for Gaussians that were not rendered), so the
// caller can hand in uninitialised memory.  dL_dcov3D / dL_dmean3D never make the extra
// HBM round trip they make between the reference's two kernels.  (The accumulator is cleared
// by segs_raster_backward before the blend pass, so repeated backwards over one forward state
// stay correct.)
//
// Floating-point outputs only (tolerance 1e-4 relative), so expressions are written
// plainly, in the reference's evaluation order.
#include "common.cuh"

namespace segs {

namespace {

constexpr int BWD_THREADS = 256;

struct M3 {   // column-major like glm::mat3: m[c][r]
    float m[3][3];
};
__device__ __forceinline__ M3 mul(const M3& A, const M3& B) {
    M3 R;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            R.m[c][r] = A.m[0][r] * B.m[c][0] + A.m[1][r] * B.m[c][1] + A.m[2][r] * B.m[c][2];
    return R;
}
__device__ __forceinline__ M3 transpose(const M3& A) {
    M3 R;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) R.m[c][r] = A.m[r][c];
    return R;
}

__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {   // auxiliary.h:108-118
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

// SH backward (backward.cu:20-139): writes dL_dsh[0..M) (zeros above the active degree) and
// returns the mean gradient through the view direction.
__device__ float3 sh_backward(int deg, int M, const float* __restrict__ sh, float3 pos, float3 campos,
                              const uint8_t* __restrict__ clamped, float3 dL_dcolor,
                              float* __restrict__ dL_dsh)
{
    constexpr float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
    constexpr float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
    constexpr float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                             0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                             -0.5900435899266435f};
    const float3 dir_orig = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    float g[3] = {dL_dcolor.x * (clamped[0] ? 0.f : 1.f), dL_dcolor.y * (clamped[1] ? 0.f : 1.f),
                  dL_dcolor.z * (clamped[2] ? 0.f : 1.f)};
    float w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = 0.f;
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;   // dL/ddir
    w[0] = C0;
    float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
    auto S = [&](int k, int ch) { return sh[3 * k + ch]; };
    if (deg > 0) {
        w[1] = -C1 * y; w[2] = C1 * z; w[3] = -C1 * x;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            dRGBdx[ch] = -C1 * S(3, ch);
            dRGBdy[ch] = -C1 * S(1, ch);
            dRGBdz[ch] = C1 * S(2, ch);
        }
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            w[4] = C2[0] * xy; w[5] = C2[1] * yz; w[6] = C2[2] * (2.f * zz - xx - yy);
            w[7] = C2[3] * xz; w[8] = C2[4] * (xx - yy);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                dRGBdx[ch] += C2[0] * y * S(4, ch) + C2[2] * 2.f * -x * S(6, ch) + C2[3] * z * S(7, ch) + C2[4] * 2.f * x * S(8, ch);
                dRGBdy[ch] += C2[0] * x * S(4, ch) + C2[1] * z * S(5, ch) + C2[2] * 2.f * -y * S(6, ch) + C2[4] * 2.f * -y * S(8, ch);
                dRGBdz[ch] += C2[1] * y * S(5, ch) + C2[2] * 2.f * 2.f * z * S(6, ch) + C2[3] * x * S(7, ch);
            }
            if (deg > 2) {
                w[9] = C3[0] * y * (3.f * xx - yy);
                w[10] = C3[1] * xy * z;
                w[11] = C3[2] * y * (4.f * zz - xx - yy);
                w[12] = C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                w[13] = C3[4] * x * (4.f * zz - xx - yy);
                w[14] = C3[5] * z * (xx - yy);
                w[15] = C3[6] * x * (xx - 3.f * yy);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    dRGBdx[ch] += (C3[0] * S(9, ch) * 3.f * 2.f * xy + C3[1] * S(10, ch) * yz +
                                   C3[2] * S(11, ch) * -2.f * xy + C3[3] * S(12, ch) * -3.f * 2.f * xz +
                                   C3[4] * S(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                                   C3[5] * S(14, ch) * 2.f * xz + C3[6] * S(15, ch) * 3.f * (xx - yy));
                    dRGBdy[ch] += (C3[0] * S(9, ch) * 3.f * (xx - yy) + C3[1] * S(10, ch) * xz +
                                   C3[2] * S(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                                   C3[3] * S(12, ch) * -3.f * 2.f * yz + C3[4] * S(13, ch) * -2.f * xy +
                                   C3[5] * S(14, ch) * -2.f * yz + C3[6] * S(15, ch) * -3.f * 2.f * xy);
                    dRGBdz[ch] += (C3[1] * S(10, ch) * xy + C3[2] * S(11, ch) * 4.f * 2.f * yz +
                                   C3[3] * S(12, ch) * 3.f * (2.f * zz - xx - yy) +
                                   C3[4] * S(13, ch) * 4.f * 2.f * xz + C3[5] * S(14, ch) * (xx - yy));
                }
            }
        }
    }
    const int ncoef = (deg + 1) * (deg + 1);
    for (int k = 0; k < M; ++k) {
        const float wk = (k < ncoef && k < 16) ? w[k] : 0.f;
        dL_dsh[3 * k + 0] = wk * g[0];
        dL_dsh[3 * k + 1] = wk * g[1];
        dL_dsh[3 * k + 2] = wk * g[2];
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        ddx += dRGBdx[ch] * g[ch];
        ddy += dRGBdy[ch] * g[ch];
        ddz += dRGBdz[ch] * g[ch];
    }
    return dnormvdv3(dir_orig, make_float3(ddx, ddy, ddz));
}

__global__ void __launch_bounds__(BWD_THREADS, 4)
preprocess_backward_kernel(int P, int D, int M,
                           const float* __restrict__ means3D, const float* __restrict__ scales,
                           const float* __restrict__ rotations, const float* __restrict__ shs,
                           const float* __restrict__ cov3D_precomp,
                           const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix,
                           const float* __restrict__ campos, const ViewParams vp,
                           const uint32_t* __restrict__ tiles_touched,
                           const float* __restrict__ cov3D_planes, const uint8_t* __restrict__ clamped,
                           float4* __restrict__ acc,
                           float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic,
                           float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor,
                           float* __restrict__ dL_dmean3D, float* __restrict__ dL_dcov3D,
                           float* __restrict__ dL_dsh, float* __restrict__ dL_dscale,
                           float* __restrict__ dL_drot)
{
    __shared__ float s_view[16], s_proj[16];
    // AoS outputs ([P,3] x4, [P,6]) are staged here and written back with coalesced stores
    __shared__ float s_out[BWD_THREADS * 18];
    if (threadIdx.x < 16) s_view[threadIdx.x] = __ldg(viewmatrix + threadIdx.x);
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = __ldg(projmatrix + threadIdx.x - 16);
    __syncthreads();
    const size_t first = size_t(blockIdx.x) * BWD_THREADS;
    const size_t idx = first + threadIdx.x;
    const int n_blk = (int)min(size_t(BWD_THREADS), size_t(P) - first);
    const bool in_range = idx < (size_t)P;

    float d_mean2D[2] = {0.f, 0.f};
    float d_conic[3] = {0.f, 0.f, 0.f};     // x, y, w
    float d_opacity = 0.f;
    float d_color[3] = {0.f, 0.f, 0.f};
    float d_mean3D[3] = {0.f, 0.f, 0.f};
    float d_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float d_scale[3] = {0.f, 0.f, 0.f};
    float d_rot[4] = {0.f, 0.f, 0.f, 0.f};
    bool sh_written = false;

    // Every global load of this CTA is issued up front, before the first dependent use, so the
    // kernel pays ONE memory round trip.  The AoS inputs (48-B accumulators, [P,3] means and scales)
    // are fetched with fully coalesced loads into the staging buffer that later carries the
    // outputs: a thread-strided 12/48-B access pattern would touch every sector three times.
    const size_t li = in_range ? idx : 0;
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0, t2 = t0;
    float m0[3], sc0[3];
    {
        const float4* acc_blk = acc + 3 * first;
        const int n4 = 3 * n_blk;
        if ((int)threadIdx.x < n4) t0 = acc_blk[threadIdx.x];
        if ((int)threadIdx.x + BWD_THREADS < n4) t1 = acc_blk[threadIdx.x + BWD_THREADS];
        if ((int)threadIdx.x + 2 * BWD_THREADS < n4) t2 = acc_blk[threadIdx.x + 2 * BWD_THREADS];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = threadIdx.x + k * BWD_THREADS;
            m0[k] = j < 3 * n_blk ? __ldg(means3D + 3 * first + j) : 0.f;
            sc0[k] = (scales != nullptr && j < 3 * n_blk) ? __ldg(scales + 3 * first + j) : 1.f;
        }
    }
    float c3[6];
#pragma unroll
    for (int k = 0; k < 6; ++k)
        c3[k] = cov3D_precomp ? __ldg(cov3D_precomp + 6 * li + k) : __ldg(cov3D_planes + size_t(k) * P + li);
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    if (scales != nullptr) q = __ldg(reinterpret_cast<const float4*>(rotations) + li);
    const uint32_t touched = __ldg(tiles_touched + li);
    {
        // (stores after the last load: the SM issues in order, so nothing that waits on a load may
        // sit in front of another load)
        float4* s_acc = reinterpret_cast<float4*>(s_out);                    // [256][3] float4
        float* s_mean = s_out + BWD_THREADS * 12;                            // [256][3]
        float* s_scale = s_mean + BWD_THREADS * 3;                           // [256][3]
        s_acc[threadIdx.x] = t0; s_acc[threadIdx.x + BWD_THREADS] = t1; s_acc[threadIdx.x + 2 * BWD_THREADS] = t2;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s_mean[threadIdx.x + k * BWD_THREADS] = m0[k];
            s_scale[threadIdx.x + k * BWD_THREADS] = sc0[k];
        }
    }
    __syncthreads();
    const int tl = in_range ? (int)threadIdx.x : 0;
    const float4 a0 = reinterpret_cast<const float4*>(s_out)[3 * tl + 0];
    const float4 a1 = reinterpret_cast<const float4*>(s_out)[3 * tl + 1];
    const float4 a2 = reinterpret_cast<const float4*>(s_out)[3 * tl + 2];
    const float3 mean = make_float3(s_out[BWD_THREADS * 12 + 3 * tl], s_out[BWD_THREADS * 12 + 3 * tl + 1],
                                    s_out[BWD_THREADS * 12 + 3 * tl + 2]);
    const float sc3[3] = {s_out[BWD_THREADS * 15 + 3 * tl], s_out[BWD_THREADS * 15 + 3 * tl + 1],
                          s_out[BWD_THREADS * 15 + 3 * tl + 2]};
    __syncthreads();       // everyone holds its inputs in registers: the buffer is free for the outputs

    // The chain below runs for EVERY lane (no branch for the loads to sink into; ~11% of the
    // Gaussians are not rendered and compute on don't-care values) and the results are masked
    // by `vis` at the end.
    const bool vis = in_range && touched > 0;   // == radii[idx] > 0 of backward.cu:156,367
    {
        d_mean2D[0] = a0.x; d_mean2D[1] = a0.y;
        d_conic[0] = a0.z; d_conic[1] = a0.w; d_conic[2] = a1.x;
        d_opacity = a1.y;
        d_color[0] = a1.z; d_color[1] = a1.w; d_color[2] = a2.x;

        const float* V = s_view;
        const float* Pm = s_proj;

        // ---- computeCov2DCUDA (backward.cu:144-274) --------------------------------------
        float3 t;
        t.x = V[0] * mean.x + V[4] * mean.y + V[8] * mean.z + V[12];
        t.y = V[1] * mean.x + V[5] * mean.y + V[9] * mean.z + V[13];
        t.z = V[2] * mean.x + V[6] * mean.y + V[10] * mean.z + V[14];
        const float limx = 1.3f * vp.tan_fovx, limy = 1.3f * vp.tan_fovy;
        const float txtz = t.x / t.z, tytz = t.y / t.z;
        t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
        t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float h_x = vp.focal_x, h_y = vp.focal_y;

        M3 J, Wm, Vrk;
        J.m[0][0] = h_x / t.z; J.m[0][1] = 0.f; J.m[0][2] = -(h_x * t.x) / (t.z * t.z);
        J.m[1][0] = 0.f; J.m[1][1] = h_y / t.z; J.m[1][2] = -(h_y * t.y) / (t.z * t.z);
        J.m[2][0] = 0.f; J.m[2][1] = 0.f; J.m[2][2] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int r = 0; r < 3; ++r) Wm.m[k][r] = V[k + 4 * r];
        Vrk.m[0][0] = c3[0]; Vrk.m[0][1] = c3[1]; Vrk.m[0][2] = c3[2];
        Vrk.m[1][0] = c3[1]; Vrk.m[1][1] = c3[3]; Vrk.m[1][2] = c3[4];
        Vrk.m[2][0] = c3[2]; Vrk.m[2][1] = c3[4]; Vrk.m[2][2] = c3[5];
        const M3 Tm = mul(Wm, J);
        const M3 cov2D = mul(mul(transpose(Tm), transpose(Vrk)), Tm);
        const float a = cov2D.m[0][0] + 0.3f;
        const float b = cov2D.m[0][1];
        const float c = cov2D.m[1][1] + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        const float dcx = d_conic[0], dcy = d_conic[1], dcz = d_conic[2];
        const float (*T)[3] = Tm.m;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            d_cov[0] = (T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc);
            d_cov[3] = (T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc);
            d_cov[5] = (T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc);
            d_cov[1] = 2 * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][1] * dL_dc;
            d_cov[2] = 2 * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][2] * dL_dc;
            d_cov[4] = 2 * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db + 2 * T[1][1] * T[1][2] * dL_dc;
        }
        const float (*K)[3] = Vrk.m;
        const float dL_dT00 = 2 * (T[0][0] * K[0][0] + T[0][1] * K[0][1] + T[0][2] * K[0][2]) * dL_da +
                              (T[1][0] * K[0][0] + T[1][1] * K[0][1] + T[1][2] * K[0][2]) * dL_db;
        const float dL_dT01 = 2 * (T[0][0] * K[1][0] + T[0][1] * K[1][1] + T[0][2] * K[1][2]) * dL_da +
                              (T[1][0] * K[1][0] + T[1][1] * K[1][1] + T[1][2] * K[1][2]) * dL_db;
        const float dL_dT02 = 2 * (T[0][0] * K[2][0] + T[0][1] * K[2][1] + T[0][2] * K[2][2]) * dL_da +
                              (T[1][0] * K[2][0] + T[1][1] * K[2][1] + T[1][2] * K[2][2]) * dL_db;
        const float dL_dT10 = 2 * (T[1][0] * K[0][0] + T[1][1] * K[0][1] + T[1][2] * K[0][2]) * dL_dc +
                              (T[0][0] * K[0][0] + T[0][1] * K[0][1] + T[0][2] * K[0][2]) * dL_db;
        const float dL_dT11 = 2 * (T[1][0] * K[1][0] + T[1][1] * K[1][1] + T[1][2] * K[1][2]) * dL_dc +
                              (T[0][0] * K[1][0] + T[0][1] * K[1][1] + T[0][2] * K[1][2]) * dL_db;
        const float dL_dT12 = 2 * (T[1][0] * K[2][0] + T[1][1] * K[2][1] + T[1][2] * K[2][2]) * dL_dc +
                              (T[0][0] * K[2][0] + T[0][1] * K[2][1] + T[0][2] * K[2][2]) * dL_db;
        const float (*Wp)[3] = Wm.m;
        const float dL_dJ00 = Wp[0][0] * dL_dT00 + Wp[0][1] * dL_dT01 + Wp[0][2] * dL_dT02;
        const float dL_dJ02 = Wp[2][0] * dL_dT00 + Wp[2][1] * dL_dT01 + Wp[2][2] * dL_dT02;
        const float dL_dJ11 = Wp[1][0] * dL_dT10 + Wp[1][1] * dL_dT11 + Wp[1][2] * dL_dT12;
        const float dL_dJ12 = Wp[2][0] * dL_dT10 + Wp[2][1] * dL_dT11 + Wp[2][2] * dL_dT12;
        const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                             (2 * h_y * t.y) * tz3 * dL_dJ12;
        // transformVec4x3Transpose (auxiliary.h:90-98); assignment, backward.cu:273
        d_mean3D[0] = V[0] * dL_dtx + V[1] * dL_dty + V[2] * dL_dtz;
        d_mean3D[1] = V[4] * dL_dtx + V[5] * dL_dty + V[6] * dL_dtz;
        d_mean3D[2] = V[8] * dL_dtx + V[9] * dL_dty + V[10] * dL_dtz;

        // ---- preprocessCUDA backward (backward.cu:347-396) -------------------------------
        const float m_homw = Pm[3] * mean.x + Pm[7] * mean.y + Pm[11] * mean.z + Pm[15];
        const float m_w = 1.0f / (m_homw + 0.0000001f);
        const float mul1 = (Pm[0] * mean.x + Pm[4] * mean.y + Pm[8] * mean.z + Pm[12]) * m_w * m_w;
        const float mul2 = (Pm[1] * mean.x + Pm[5] * mean.y + Pm[9] * mean.z + Pm[13]) * m_w * m_w;
        d_mean3D[0] += (Pm[0] * m_w - Pm[3] * mul1) * d_mean2D[0] + (Pm[1] * m_w - Pm[3] * mul2) * d_mean2D[1];
        d_mean3D[1] += (Pm[4] * m_w - Pm[7] * mul1) * d_mean2D[0] + (Pm[5] * m_w - Pm[7] * mul2) * d_mean2D[1];
        d_mean3D[2] += (Pm[8] * m_w - Pm[11] * mul1) * d_mean2D[0] + (Pm[9] * m_w - Pm[11] * mul2) * d_mean2D[1];

        if (shs != nullptr && vis) {
            const float3 cp = make_float3(__ldg(campos), __ldg(campos + 1), __ldg(campos + 2));
            const float3 dm = sh_backward(D, M, shs + idx * M * 3, mean, cp, clamped + 3 * idx,
                                          make_float3(d_color[0], d_color[1], d_color[2]),
                                          dL_dsh + idx * M * 3);
            d_mean3D[0] += dm.x; d_mean3D[1] += dm.y; d_mean3D[2] += dm.z;
            sh_written = true;
        }

        if (scales != nullptr) {
            // ---- computeCov3D backward (backward.cu:278-341) -----------------------------
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            M3 R;
            R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
            R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
            R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
            const float s[3] = {vp.scale_modifier * sc3[0], vp.scale_modifier * sc3[1], vp.scale_modifier * sc3[2]};
            M3 S;
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
                for (int r2 = 0; r2 < 3; ++r2) S.m[c2][r2] = (c2 == r2) ? s[c2] : 0.f;
            const M3 Mm = mul(S, R);
            M3 dS;
            dS.m[0][0] = d_cov[0]; dS.m[0][1] = 0.5f * d_cov[1]; dS.m[0][2] = 0.5f * d_cov[2];
            dS.m[1][0] = 0.5f * d_cov[1]; dS.m[1][1] = d_cov[3]; dS.m[1][2] = 0.5f * d_cov[4];
            dS.m[2][0] = 0.5f * d_cov[2]; dS.m[2][1] = 0.5f * d_cov[4]; dS.m[2][2] = d_cov[5];
            M3 M2;
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2)
#pragma unroll
                for (int r2 = 0; r2 < 3; ++r2) M2.m[c2][r2] = Mm.m[c2][r2] * 2.0f;
            const M3 dL_dM = mul(M2, dS);
            const M3 Rt = transpose(R);
            M3 dMt = transpose(dL_dM);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                d_scale[k] = Rt.m[k][0] * dMt.m[k][0] + Rt.m[k][1] * dMt.m[k][1] + Rt.m[k][2] * dMt.m[k][2];
                dMt.m[k][0] *= s[k]; dMt.m[k][1] *= s[k]; dMt.m[k][2] *= s[k];
            }
            const float (*d)[3] = dMt.m;
            d_rot[0] = 2 * z * (d[0][1] - d[1][0]) + 2 * y * (d[2][0] - d[0][2]) + 2 * x * (d[1][2] - d[2][1]);
            d_rot[1] = 2 * y * (d[1][0] + d[0][1]) + 2 * z * (d[2][0] + d[0][2]) + 2 * r * (d[1][2] - d[2][1]) - 4 * x * (d[2][2] + d[1][1]);
            d_rot[2] = 2 * x * (d[1][0] + d[0][1]) + 2 * r * (d[2][0] - d[0][2]) + 2 * z * (d[1][2] + d[2][1]) - 4 * y * (d[2][2] + d[0][0]);
            d_rot[3] = 2 * r * (d[0][1] - d[1][0]) + 2 * x * (d[2][0] + d[0][2]) + 2 * y * (d[1][2] + d[2][1]) - 4 * z * (d[1][1] + d[0][0]);
        }
    }
    if (!vis) {
        d_mean2D[0] = d_mean2D[1] = 0.f;
        d_conic[0] = d_conic[1] = d_conic[2] = 0.f;
        d_opacity = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { d_color[k] = 0.f; d_mean3D[k] = 0.f; d_scale[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < 6; ++k) d_cov[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) d_rot[k] = 0.f;
    }

    {
        const int t = threadIdx.x;
        float* o2 = s_out;                          // dL_dmean2D [256,3]
        float* oc = s_out + BWD_THREADS * 3;        // dL_dcolor  [256,3]
        float* o3 = s_out + BWD_THREADS * 6;        // dL_dmean3D [256,3]
        float* os = s_out + BWD_THREADS * 9;        // dL_dscale  [256,3]
        float* ov = s_out + BWD_THREADS * 12;       // dL_dcov3D  [256,6]
        o2[3 * t] = d_mean2D[0]; o2[3 * t + 1] = d_mean2D[1]; o2[3 * t + 2] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            oc[3 * t + k] = d_color[k];
            o3[3 * t + k] = d_mean3D[k];
            os[3 * t + k] = d_scale[k];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) ov[6 * t + k] = d_cov[k];
    }
    if (in_range) {
        if (dL_dconic != nullptr)
            *reinterpret_cast<float4*>(dL_dconic + 4 * idx) = make_float4(d_conic[0], d_conic[1], 0.f, d_conic[2]);
        dL_dopacity[idx] = d_opacity;
        *reinterpret_cast<float4*>(dL_drot + 4 * idx) = make_float4(d_rot[0], d_rot[1], d_rot[2], d_rot[3]);
        if (dL_dsh != nullptr && M > 0 && !sh_written)
            for (int k = 0; k < 3 * M; ++k) dL_dsh[idx * M * 3 + k] = 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 3 * n_blk; j += BWD_THREADS) {
        dL_dmean2D[3 * first + j] = s_out[j];
        dL_dcolor[3 * first + j] = s_out[BWD_THREADS * 3 + j];
        dL_dmean3D[3 * first + j] = s_out[BWD_THREADS * 6 + j];
        dL_dscale[3 * first + j] = s_out[BWD_THREADS * 9 + j];
    }
    for (int j = threadIdx.x; j < 6 * n_blk; j += BWD_THREADS) dL_dcov3D[6 * first + j] = s_out[BWD_THREADS * 12 + j];
}

}  // namespace

int launch_preprocess_backward(int P, int D, int M, const float* means3D, const float* scales,
                               const float* rotations, const float* shs,
                               const float* cov3D_precomp, const float* viewmatrix,
                               const float* projmatrix, const float* campos,
                               const ViewParams& vp, const int* radii, GeomState& g,
                               float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                               float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                               float* dL_dsh, float* dL_dscale, float* dL_drot,
                               cudaStream_t stream)
{
    (void)radii;   // visibility is taken from this library's own forward state
    preprocess_backward_kernel<<<(P + BWD_THREADS - 1) / BWD_THREADS, BWD_THREADS, 0, stream>>>(
        P, D, M, means3D, scales, rotations, shs, cov3D_precomp, viewmatrix, projmatrix, campos, vp,
        g.tiles_touched, g.cov3D, g.clamped, g.acc, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
        dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // namespace segs
