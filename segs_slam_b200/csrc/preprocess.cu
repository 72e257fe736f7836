// Per-Gaussian forward preprocess for sm_100a.
//
// Replaces FORWARD::preprocess / filter_preprocess / project and checkFrustum of the
// reference (cuda_rasterizer/forward.cu:156-256, 260-334, 573-673; rasterizer_impl.cu:54-66).
//
// Design (B200): one thread per Gaussian.  The AoS inputs ([P,3] means, scales, colours) are read
// directly: the three 4-byte loads of a warp cover one contiguous 384-byte run, so the second and
// third are L1 hits; every output is written to this library's own SoA / 16-byte-record layout so
// that the binning and blend stages read coalesced or gather whole 16-byte sectors (the three
// 16-byte stores of a thread's 48-byte record are merged in L2 before they reach DRAM).  Round 1
// staged inputs and records through shared memory (two extra barriers, 21 KB per CTA); without the
// staging the kernel has no barrier between load and store, so more warps are in flight
// (PRE_MIN_CTAS) and the IEEE divisions / square roots of one warp hide behind the loads of others.
// The pass also emits the depth-sort keys, which removes a separate pass over P.
//
// Bit-exactness: radii / tiles_touched / rect / depth / mean2D / conic decide integer
// outputs downstream (keys, order, ranges, n_contrib), so the arithmetic below follows the
// reference's expressions operation by operation INCLUDING nvcc's FMA contraction of them
// (default -fmad=true).  The contraction is spelled out with __fmaf_rn/__fmul_rn/__fadd_rn
// so it does not depend on how this file happens to be optimised:  for a sum of three
// products  a0*b0 + a1*b1 + a2*b2  nvcc emits  fma(a2,b2, fma(a0,b0, mul(a1,b1)))  — see
// dot3c() — and products with a structural zero are kept (0*x is not folded without
// fast-math).  The pattern was read from the SASS of the reference build (NVVM contracts
// some pairs in PTX, ptxas fuses more of the remaining mul/add pairs), see DESIGN.md
// §"bit-exact preprocess".
#include <algorithm>
#include "common.cuh"

namespace segs {

namespace {

constexpr int PRE_THREADS = 128;
#ifndef PRE_MIN_CTAS
#define PRE_MIN_CTAS 8         // 1024 threads per SM at 64 registers (48 and 40 spill and are slower; 80 is no faster)
#endif

__device__ __forceinline__ float dot3c(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}

// row `r` of transformPoint4x4/4x3 (auxiliary.h:59-78): m[r]*x + m[4+r]*y + m[8+r]*z + m[12+r]
__device__ __forceinline__ float affine_row(const float* __restrict__ m, int r, float x, float y, float z) {
    return __fadd_rn(m[12 + r], dot3c(m[r], x, m[4 + r], y, m[8 + r], z));
}

struct Projected {
    float depth;      // p_view.z
    float px, py;     // pixel-space mean (ndc2Pix)
    float cov_x, cov_y, cov_z;
    float det;
    int radius;
    unsigned x0, y0, x1, y1;
    unsigned tiles;
};

// Stage 1: cull + 3D covariance (forward.cu:118-152).  Returns false if culled.
__device__ __forceinline__ void cov3d_from_scale_rot(float mod, float s0, float s1, float s2,
                                                     float4 q, float* c)
{
    const float r = q.x, x = q.y, y = q.z, z = q.w;   // (r,x,y,z), NOT normalised (forward.cu:127)
    const float sx = __fmul_rn(mod, s0), sy = __fmul_rn(mod, s1), sz = __fmul_rn(mod, s2);
    // Final contraction as it appears in the reference's SASS (ptxas fuses one product of each
    // a*b +- c*d pair that NVVM left as mul/add): xy, yz and ry are never rounded on their own.
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float rz = __fmul_rn(r, z), xz = __fmul_rn(x, z), rx = __fmul_rn(r, x);
    const float yy_zz = __fadd_rn(yy, zz);
    const float xx_zz = __fmaf_rn(x, x, zz);
    const float xx_yy = __fmaf_rn(x, x, yy);
    const float xy_m_rz = __fmaf_rn(x, y, -rz), xy_p_rz = __fmaf_rn(x, y, rz);
    const float xz_p_ry = __fmaf_rn(r, y, xz), xz_m_ry = __fmaf_rn(-r, y, xz);
    const float yz_m_rx = __fmaf_rn(y, z, -rx), yz_p_rx = __fmaf_rn(y, z, rx);
    // glm::mat3 R(...) column-major: R[c][r]
    const float R00 = __fsub_rn(1.f, __fadd_rn(yy_zz, yy_zz));
    const float R01 = __fadd_rn(xy_m_rz, xy_m_rz);
    const float R02 = __fadd_rn(xz_p_ry, xz_p_ry);
    const float R10 = __fadd_rn(xy_p_rz, xy_p_rz);
    const float R11 = __fsub_rn(1.f, __fadd_rn(xx_zz, xx_zz));
    const float R12 = __fadd_rn(yz_m_rx, yz_m_rx);
    const float R20 = __fadd_rn(xz_m_ry, xz_m_ry);
    const float R21 = __fadd_rn(yz_p_rx, yz_p_rx);
    const float R22 = __fsub_rn(1.f, __fadd_rn(xx_yy, xx_yy));
    // M = S * R with S = diag(sx,sy,sz): M[c][r] = S[0][r]*R[c][0] + S[1][r]*R[c][1] + S[2][r]*R[c][2]
    const float M00 = dot3c(sx, R00, 0.f, R01, 0.f, R02);
    const float M01 = dot3c(0.f, R00, sy, R01, 0.f, R02);
    const float M02 = dot3c(0.f, R00, 0.f, R01, sz, R02);
    const float M10 = dot3c(sx, R10, 0.f, R11, 0.f, R12);
    const float M11 = dot3c(0.f, R10, sy, R11, 0.f, R12);
    const float M12 = dot3c(0.f, R10, 0.f, R11, sz, R12);
    const float M20 = dot3c(sx, R20, 0.f, R21, 0.f, R22);
    const float M21 = dot3c(0.f, R20, sy, R21, 0.f, R22);
    const float M22 = dot3c(0.f, R20, 0.f, R21, sz, R22);
    // Sigma = transpose(M) * M: Sigma[c][r] = M[r][0]*M[c][0] + M[r][1]*M[c][1] + M[r][2]*M[c][2]
    c[0] = dot3c(M00, M00, M01, M01, M02, M02);
    c[1] = dot3c(M00, M10, M01, M11, M02, M12);
    c[2] = dot3c(M00, M20, M01, M21, M02, M22);
    c[3] = dot3c(M10, M10, M11, M11, M12, M12);
    c[4] = dot3c(M10, M20, M11, M21, M12, M22);
    c[5] = dot3c(M20, M20, M21, M21, M22, M22);
}

// ndc2Pix (auxiliary.h:41-45) is evaluated in double: ((v + 1.0) * S - 1.0) * 0.5, with the
// multiply-subtract contracted to one DFMA.
__device__ __forceinline__ float ndc_to_pix(float v, int S) {
    const double t = __fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0);
    return (float)__dmul_rn(t, 0.5);
}

// Everything after the cull: projection, EWA 2D covariance (forward.cu:74-113), extent,
// tile rectangle (auxiliary.h:47-57).  Returns false for det == 0 or an empty rectangle.
__device__ __forceinline__ bool project_gaussian(float px, float py, float pz, const float* c3,
                                                 const float* __restrict__ V,
                                                 const float* __restrict__ Pm,
                                                 const ViewParams& vp, Projected& o)
{
    // p_hom / p_proj (forward.cu:197-200)
    const float hx = affine_row(Pm, 0, px, py, pz);
    const float hy = affine_row(Pm, 1, px, py, pz);
    const float hw = affine_row(Pm, 3, px, py, pz);
    const float p_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
    const float projx = __fmul_rn(hx, p_w);
    const float projy = __fmul_rn(hy, p_w);

    // computeCov2D
    const float tx0 = affine_row(V, 0, px, py, pz);
    const float ty0 = affine_row(V, 1, px, py, pz);
    const float tz = affine_row(V, 2, px, py, pz);
    const float limx = __fmul_rn(1.3f, vp.tan_fovx);
    const float limy = __fmul_rn(1.3f, vp.tan_fovy);
    const float txtz = __fdiv_rn(tx0, tz);
    const float tytz = __fdiv_rn(ty0, tz);
    const float tx = __fmul_rn(fminf(limx, fmaxf(-limx, txtz)), tz);
    const float ty = __fmul_rn(fminf(limy, fmaxf(-limy, tytz)), tz);
    const float tz2 = __fmul_rn(tz, tz);
    const float J00 = __fdiv_rn(vp.focal_x, tz);
    const float J02 = __fdiv_rn(-__fmul_rn(vp.focal_x, tx), tz2);
    const float J11 = __fdiv_rn(vp.focal_y, tz);
    const float J12 = __fdiv_rn(-__fmul_rn(vp.focal_y, ty), tz2);
    // W[k][r] = V[k + 4r];  T = W * J:  T[c][r] = W[0][r]*J[c][0] + W[1][r]*J[c][1] + W[2][r]*J[c][2]
    const float T00 = dot3c(V[0], J00, V[1], 0.f, V[2], J02);
    const float T01 = dot3c(V[4], J00, V[5], 0.f, V[6], J02);
    const float T02 = dot3c(V[8], J00, V[9], 0.f, V[10], J02);
    const float T10 = dot3c(V[0], 0.f, V[1], J11, V[2], J12);
    const float T11 = dot3c(V[4], 0.f, V[5], J11, V[6], J12);
    const float T12 = dot3c(V[8], 0.f, V[9], J11, V[10], J12);
    // A = transpose(T) * transpose(Vrk):  A[c][r] = T[r][0]*Vrk[0][c] + T[r][1]*Vrk[1][c] + T[r][2]*Vrk[2][c]
    const float A00 = dot3c(T00, c3[0], T01, c3[1], T02, c3[2]);
    const float A01 = dot3c(T10, c3[0], T11, c3[1], T12, c3[2]);
    const float A10 = dot3c(T00, c3[1], T01, c3[3], T02, c3[4]);
    const float A11 = dot3c(T10, c3[1], T11, c3[3], T12, c3[4]);
    const float A20 = dot3c(T00, c3[2], T01, c3[4], T02, c3[5]);
    const float A21 = dot3c(T10, c3[2], T11, c3[4], T12, c3[5]);
    // cov = A * T:  cov[c][r] = A[0][r]*T[c][0] + A[1][r]*T[c][1] + A[2][r]*T[c][2]
    const float cov00 = dot3c(A00, T00, A10, T01, A20, T02);
    const float cov01 = dot3c(A01, T00, A11, T01, A21, T02);
    const float cov11 = dot3c(A01, T10, A11, T11, A21, T12);
    o.cov_x = __fadd_rn(cov00, 0.3f);
    o.cov_y = cov01;
    o.cov_z = __fadd_rn(cov11, 0.3f);

    // forward.cu:219-232; in SASS: det = fma(a, c, -(b*b)), discriminant = fma(mid, mid, -det)
    o.det = __fmaf_rn(o.cov_x, o.cov_z, -__fmul_rn(o.cov_y, o.cov_y));
    if (o.det == 0.0f) return false;
    const float mid = __fmul_rn(0.5f, __fadd_rn(o.cov_x, o.cov_z));
    const float disc = __fsqrt_rn(fmaxf(0.1f, __fmaf_rn(mid, mid, -o.det)));
    const float lambda1 = __fadd_rn(mid, disc);
    const float lambda2 = __fsub_rn(mid, disc);
    const float my_radius = ceilf(__fmul_rn(3.f, __fsqrt_rn(fmaxf(lambda1, lambda2))));
    o.px = ndc_to_pix(projx, vp.W);
    o.py = ndc_to_pix(projy, vp.H);

    // getRect with max_radius = (int)my_radius converted back to float
    o.radius = (int)my_radius;
    const float rad = (float)o.radius;
    const float inv16 = 0.0625f;  // "/ BLOCK_X" is compiled to an exact multiply
    const int ix0 = (int)__fmul_rn(__fsub_rn(o.px, rad), inv16);
    const int iy0 = (int)__fmul_rn(__fsub_rn(o.py, rad), inv16);
    const int ix1 = (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(o.px, rad), 16.f), -1.f), inv16);
    const int iy1 = (int)__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(o.py, rad), 16.f), -1.f), inv16);
    o.x0 = min((unsigned)vp.grid_x, (unsigned)max(0, ix0));
    o.y0 = min((unsigned)vp.grid_y, (unsigned)max(0, iy0));
    o.x1 = min((unsigned)vp.grid_x, (unsigned)max(0, ix1));
    o.y1 = min((unsigned)vp.grid_y, (unsigned)max(0, iy1));
    o.tiles = (o.x1 - o.x0) * (o.y1 - o.y0);
    return o.tiles != 0;
}

// SH -> RGB (forward.cu:20-71).  Floating-point output only (tolerance 1e-5), so written
// plainly.  `sh` points at this Gaussian's [M,3] coefficients.
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float* __restrict__ sh, float3 pos,
                                            float3 campos, bool* clamped)
{
    constexpr float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
    constexpr float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                             -1.0925484305920792f, 0.5462742152960396f};
    constexpr float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                             0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                             -0.5900435899266435f};
    float dx = pos.x - campos.x, dy = pos.y - campos.y, dz = pos.z - campos.z;
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= len; dy /= len; dz /= len;
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        auto S = [&](int k) { return sh[3 * k + ch]; };
        float r = C0 * S(0);
        if (deg > 0) {
            const float x = dx, y = dy, z = dz;
            r = r - C1 * y * S(1) + C1 * z * S(2) - C1 * x * S(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + C2[0] * xy * S(4) + C2[1] * yz * S(5) + C2[2] * (2.0f * zz - xx - yy) * S(6) +
                    C2[3] * xz * S(7) + C2[4] * (xx - yy) * S(8);
                if (deg > 2) {
                    r = r + C3[0] * y * (3.0f * xx - yy) * S(9) + C3[1] * xy * z * S(10) +
                        C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
                        C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                        C3[4] * x * (4.0f * zz - xx - yy) * S(13) + C3[5] * z * (xx - yy) * S(14) +
                        C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
        r += 0.5f;
        clamped[ch] = (r < 0.f);
        res[ch] = fmaxf(r, 0.f);
    }
    return make_float3(res[0], res[1], res[2]);
}

// Conservative half-extent of the region where this Gaussian can pass the blend's
// alpha >= 1/255 test (forward.cu:421-423): power >= -ln(255*o)  <=>  d^T conic d <= 2 ln(255 o),
// whose axis-aligned bound is sqrt(2 ln(255 o) * cov_xx) (cov = conic^-1).  A small relative
// and absolute margin absorbs the rounding of conic and of expf; anything non-finite or
// non-positive-definite disables culling (extent = +inf).
__device__ __forceinline__ float2 cull_extent(float opacity, float cov_x, float cov_z, float det) {
    const float o255 = 255.f * opacity;
    if (!(o255 >= 1.f)) {
        // alpha <= opacity < 1/255 for every pixel: never blended.  NaN opacity: never cull.
        const float e = (opacity == opacity) ? -1e30f : __int_as_float(0x7f800000);
        return make_float2(e, e);
    }
    if (!(det > 0.f && cov_x > 0.f && cov_z > 0.f)) {
        const float inf = __int_as_float(0x7f800000);
        return make_float2(inf, inf);
    }
    const float two_tau = 2.f * logf(o255);
    return make_float2(sqrtf(two_tau * cov_x) * 1.0005f + 0.01f,
                       sqrtf(two_tau * cov_z) * 1.0005f + 0.01f);
}

// Coalesced staging of `n` consecutive [.,3] rows starting at row `first` into smem.
__device__ __forceinline__ void stage_rows3(const float* __restrict__ src, float* dst, size_t first, int n) {
    const float* s = src + 3 * first;
    for (int i = threadIdx.x; i < 3 * n; i += PRE_THREADS) dst[i] = __ldg(s + i);
}

enum class Mode { Render, Filter, Project };

// one Gaussian (`item`) of one thread; returns through tiles_acc / cand_acc the WARP totals of tiles_touched and
// super-tile candidates (same value in every lane)
template <Mode MODE>
__device__ __forceinline__ void preprocess_one(size_t item, const float* s_view, const float* s_proj, uint32_t& tiles_acc, uint32_t& cand_acc,
                  int P, int D, int M,
                  const float* __restrict__ means3D, const float* __restrict__ scales,
                  const float* __restrict__ rotations, const float* __restrict__ opacities,
                  const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                  const float* __restrict__ colors_precomp,
                  const float* __restrict__ cam_pos, const ViewParams& vp, const int prefiltered,
                  int* __restrict__ radii,
                  float* __restrict__ depths, uint32_t* __restrict__ tiles_touched,
                  ushort4* __restrict__ rect, float4* __restrict__ rec, float* __restrict__ cov3D,
                  uint8_t* __restrict__ clamped, uint32_t* __restrict__ sort_key,
                  uint32_t* __restrict__ err_flag, float* __restrict__ out_rgb, float* __restrict__ points_image)
{
    const bool in_range = item < size_t(P);
    const size_t idx = in_range ? item : size_t(P) - 1;      // tail lanes re-read the last Gaussian and write nothing
    // all independent loads first, the matrices' barrier behind them
    const float px = __ldg(means3D + 3 * idx), py = __ldg(means3D + 3 * idx + 1), pz = __ldg(means3D + 3 * idx + 2);
    float sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    float c3[6];
    if (cov3D_precomp != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; ++k) c3[k] = __ldg(cov3D_precomp + 6 * idx + k);
    } else {
        sc0 = __ldg(scales + 3 * idx); sc1 = __ldg(scales + 3 * idx + 1); sc2 = __ldg(scales + 3 * idx + 2);
        q = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
    }
    float3 rgb = make_float3(0.f, 0.f, 0.f);
    float opacity = 0.f;
    if (MODE == Mode::Render) {
        if (colors_precomp != nullptr)
            rgb = make_float3(__ldg(colors_precomp + 3 * idx), __ldg(colors_precomp + 3 * idx + 1), __ldg(colors_precomp + 3 * idx + 2));
        opacity = __ldg(opacities + idx);
    }
    bool visible = false;
    Projected pr;

    // in_frustum (auxiliary.h:140-166): only the near-plane test is live.
    const float depth = affine_row(s_view, 2, px, py, pz);
    if (!in_range) {
        // tail lanes of the last CTA: nothing to do, but stay for the warp-wide sums below
    } else if (depth <= 0.2f) {
        if (prefiltered && err_flag != nullptr) atomicExch(err_flag, 1u);
    } else {
        if (cov3D_precomp == nullptr) cov3d_from_scale_rot(vp.scale_modifier, sc0, sc1, sc2, q, c3);
        visible = project_gaussian(px, py, pz, c3, s_view, s_proj, vp, pr);
    }

    if (MODE == Mode::Render) {
        // num_rendered = sum of tiles_touched (what the reference gets from its inclusive scan,
        // rasterizer_impl.cu:276-281): one atomic per warp
        const uint32_t warp_tiles = __reduce_add_sync(0xFFFFFFFFu, visible ? pr.tiles : 0u);
        // ... and the number of (super-tile, Gaussian) candidates of the two-level binning
        uint32_t cand = 0;
        if (visible) {
            const int sh = vp.sshift;
            cand = (((pr.x1 - 1) >> sh) - (pr.x0 >> sh) + 1) * (((pr.y1 - 1) >> sh) - (pr.y0 >> sh) + 1);
        }
        const uint32_t warp_cand = __reduce_add_sync(0xFFFFFFFFu, cand);
        tiles_acc += warp_tiles;          // warp totals; one pair of REDs per CTA at the end of the kernel
        cand_acc += warp_cand;
    }
    if (MODE != Mode::Render && !in_range) return;

    if (MODE == Mode::Filter) {
        radii[idx] = visible ? pr.radius : 0;
        return;
    }

    bool cl[3] = {false, false, false};
    if (visible && colors_precomp == nullptr) {
        const float3 cp = make_float3(__ldg(cam_pos), __ldg(cam_pos + 1), __ldg(cam_pos + 2));
        rgb = sh_to_rgb(D, shs + size_t(idx) * M * 3, make_float3(px, py, pz), cp, cl);
    }

    if (MODE == Mode::Project) {
        // project2_image (rasterizer_impl.cu:494-585): pixel means, radii and the SH colours.
        radii[idx] = visible ? pr.radius : 0;
        points_image[2 * idx] = visible ? pr.px : 0.f;
        points_image[2 * idx + 1] = visible ? pr.py : 0.f;
        const bool has_rgb = visible && colors_precomp == nullptr;
        out_rgb[3 * idx] = has_rgb ? rgb.x : 0.f;
        out_rgb[3 * idx + 1] = has_rgb ? rgb.y : 0.f;
        out_rgb[3 * idx + 2] = has_rgb ? rgb.z : 0.f;
        return;
    }

    // ---- Mode::Render -------------------------------------------------------------
    if (in_range) {
        if (radii != nullptr) radii[idx] = visible ? pr.radius : 0;
        tiles_touched[idx] = visible ? pr.tiles : 0u;
        if (!visible) {
            sort_key[idx] = 0xFFFFFFFFu;   // sorts behind every real depth (depth > 0.2 => sign bit 0)
            depths[idx] = 0.f;
            rect[idx] = make_ushort4(0, 0, 0, 0);
            // the record of a culled Gaussian is never read
        } else {
            depths[idx] = depth;
            sort_key[idx] = __float_as_uint(depth);
            rect[idx] = make_ushort4((unsigned short)pr.x0, (unsigned short)pr.y0,
                                     (unsigned short)pr.x1, (unsigned short)pr.y1);
            const float det_inv = __frcp_rn(pr.det);
            const float conic_x = __fmul_rn(pr.cov_z, det_inv);
            const float conic_y = __fmul_rn(det_inv, -pr.cov_y);
            const float conic_z = __fmul_rn(pr.cov_x, det_inv);
            const float2 ext = cull_extent(opacity, pr.cov_x, pr.cov_z, pr.det);
            float4* r = rec + 3 * idx;
            r[0] = make_float4(pr.px, pr.py, ext.x, ext.y);
            r[1] = make_float4(conic_x, conic_y, conic_z, opacity);
            r[2] = make_float4(rgb.x, rgb.y, rgb.z, depth);
            if (cov3D_precomp == nullptr) {
#pragma unroll
                for (int k = 0; k < 6; ++k) cov3D[size_t(k) * P + idx] = c3[k];
            }
            if (colors_precomp == nullptr) {
                clamped[3 * idx + 0] = cl[0];
                clamped[3 * idx + 1] = cl[1];
                clamped[3 * idx + 2] = cl[2];
            }
        }
    }
}

// Persistent grid (148 SMs x PRE_MIN_CTAS CTAs), each CTA strides over chunks of PRE_THREADS Gaussians: 7800 short-lived
// 128-thread CTAs were bound by CTA launch rate, not by memory or issue slots.
template <Mode MODE>
__global__ void __launch_bounds__(PRE_THREADS, PRE_MIN_CTAS)
preprocess_kernel(int P, int D, int M,
                  const float* __restrict__ means3D, const float* __restrict__ scales,
                  const float* __restrict__ rotations, const float* __restrict__ opacities,
                  const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                  const float* __restrict__ colors_precomp,
                  const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix,
                  const float* __restrict__ cam_pos, const ViewParams vp, const int prefiltered,
                  int* __restrict__ radii,
                  // Render outputs
                  float* __restrict__ depths, uint32_t* __restrict__ tiles_touched,
                  ushort4* __restrict__ rect, float4* __restrict__ rec, float* __restrict__ cov3D,
                  float4* __restrict__ acc, uint8_t* __restrict__ clamped,
                  uint32_t* __restrict__ sort_key, uint32_t* __restrict__ sort_val,
                  uint32_t* __restrict__ err_flag, uint32_t* __restrict__ num_rendered,
                  volatile uint32_t* __restrict__ host_counters,
                  // Project outputs
                  float* __restrict__ out_rgb, float* __restrict__ points_image)
{
    __shared__ float s_view[16], s_proj[16];
    __shared__ uint32_t s_part[2 * PRE_THREADS / 32];
    if (threadIdx.x < 16) s_view[threadIdx.x] = __ldg(viewmatrix + threadIdx.x);
    else if (threadIdx.x < 32) s_proj[threadIdx.x - 16] = __ldg(projmatrix + threadIdx.x - 16);
    __syncthreads();
    (void)acc;   // the gradient accumulators are cleared by the backward (forward-only renders never touch them)
    (void)sort_val;
    uint32_t tiles_acc = 0, cand_acc = 0;
    const size_t chunks = (size_t(P) + PRE_THREADS - 1) / PRE_THREADS;
    for (size_t chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x)
        preprocess_one<MODE>(chunk * PRE_THREADS + threadIdx.x, s_view, s_proj, tiles_acc, cand_acc, P, D, M, means3D, scales, rotations,
                             opacities, shs, cov3D_precomp, colors_precomp, cam_pos, vp, prefiltered, radii, depths, tiles_touched, rect,
                             rec, cov3D, clamped, sort_key, err_flag, out_rgb, points_image);
    if (MODE != Mode::Render) return;
    if ((threadIdx.x & 31) == 0) {
        s_part[threadIdx.x >> 5] = tiles_acc;
        s_part[PRE_THREADS / 32 + (threadIdx.x >> 5)] = cand_acc;
    }
    // The host needs the totals to size the binning buffer (rasterizer_impl.cu:279-285).  The
    // last CTA to get here stores them straight into mapped pinned host memory: no copy
    // engine is involved, so the read-back never queues behind a caller's bulk D2H copies.
    if (host_counters != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tiles_cta = 0, cand_cta = 0;
#pragma unroll
            for (int w = 0; w < PRE_THREADS / 32; ++w) { tiles_cta += s_part[w]; cand_cta += s_part[PRE_THREADS / 32 + w]; }
            if (tiles_cta) {
                uint32_t* slot = num_rendered + COUNTER_BASE + 2 * (blockIdx.x % COUNTER_SLOTS);
                atomicAdd(slot, tiles_cta);
                atomicAdd(slot + 1, cand_cta);
            }
            __threadfence();
            if (atomicAdd(num_rendered + 4, 1u) == gridDim.x - 1) {
                __threadfence();
                uint32_t tiles = 0, cands = 0;
                for (int k = 0; k < COUNTER_SLOTS; ++k) {
                    tiles += atomicAdd(num_rendered + COUNTER_BASE + 2 * k, 0u);
                    cands += atomicAdd(num_rendered + COUNTER_BASE + 2 * k + 1, 0u);
                }
                num_rendered[0] = tiles;
                num_rendered[2] = cands;
                host_counters[0] = tiles;
                host_counters[1] = atomicAdd(num_rendered + 1, 0u);
                host_counters[2] = cands;
                __threadfence_system();
            }
        }
    }
}

__global__ void __launch_bounds__(PRE_THREADS)
mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                    unsigned char* __restrict__ present)
{
    __shared__ float s_mean[3 * PRE_THREADS];
    const size_t first = size_t(blockIdx.x) * PRE_THREADS;
    const int n = (int)min(size_t(PRE_THREADS), size_t(P) - first);
    stage_rows3(means3D, s_mean, first, n);
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= n) return;
    const float depth = affine_row(viewmatrix, 2, s_mean[3 * t], s_mean[3 * t + 1], s_mean[3 * t + 2]);
    present[first + t] = !(depth <= 0.2f);
}

inline int grid_for(int P) { return std::min((P + PRE_THREADS - 1) / PRE_THREADS, SM_COUNT * PRE_MIN_CTAS); }
inline int grid_small(int P) { return (P + PRE_THREADS - 1) / PRE_THREADS; }

}  // namespace

int launch_preprocess(int P, int D, int M, const float* means3D, const float* scales,
                      const float* rotations, const float* opacities, const float* shs,
                      const float* cov3D_precomp, const float* colors_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                      const ViewParams& vp, bool prefiltered, int* radii, GeomState& g,
                      uint32_t* host_counters, cudaStream_t stream)
{
    preprocess_kernel<Mode::Render><<<grid_for(P), PRE_THREADS, 0, stream>>>(
        P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp,
        viewmatrix, projmatrix, cam_pos, vp, prefiltered ? 1 : 0, radii, g.depths,
        g.tiles_touched, g.rect, g.rec, g.cov3D, g.acc, g.clamped, g.key_a, g.val_a,
        g.counters + 1, g.counters, host_counters, nullptr, nullptr);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int launch_filter(int P, const float* means3D, const float* scales, const float* rotations,
                  const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                  const ViewParams& vp, bool prefiltered, int* radii, uint32_t* err_flag,
                  cudaStream_t stream)
{
    preprocess_kernel<Mode::Filter><<<grid_for(P), PRE_THREADS, 0, stream>>>(
        P, 0, 0, means3D, scales, rotations, nullptr, nullptr, cov3D_precomp, nullptr,
        viewmatrix, projmatrix, nullptr, vp, prefiltered ? 1 : 0, radii, nullptr, nullptr,
        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, err_flag, nullptr, nullptr, nullptr, nullptr);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int launch_project(int P, int D, int M, const float* means3D, const float* scales,
                   const float* rotations, const float* opacities, const float* shs,
                   const float* cov3D_precomp, const float* colors_precomp,
                   const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                   const ViewParams& vp, bool prefiltered, float* out_rgb, float* points_image,
                   int* radii, cudaStream_t stream)
{
    preprocess_kernel<Mode::Project><<<grid_for(P), PRE_THREADS, 0, stream>>>(
        P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp,
        viewmatrix, projmatrix, cam_pos, vp, prefiltered ? 1 : 0, radii, nullptr, nullptr,
        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out_rgb,
        points_image);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix,
                        unsigned char* present, cudaStream_t stream)
{
    mark_visible_kernel<<<grid_small(P), PRE_THREADS, 0, stream>>>(P, means3D, viewmatrix, present);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // namespace segs
