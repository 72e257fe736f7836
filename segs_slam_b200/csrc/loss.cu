// Fused photometric loss of the mapper for sm_100a:
//     loss = w_l1 * mean|x - y| + w_ssim * mean(SSIM(x, y)) + bias
// with the forward pass producing the three per-pixel derivative maps of the SSIM index and the
// backward pass turning them into dL/dx — the `dL_dout_color` the rasterizer backward consumes —
// in ONE kernel each.
//
// Replaces loss_utils::l1_loss / loss_utils::ssim / loss_utils::_ssim
// (/root/reference/include/loss_utils.h:29-32, 50-127) as they are combined by the mapper
// (/root/reference/src/gaussian_mapper.cpp:917-925:
//   loss = (1 - lambda) * l1(rendered * mask, gt * mask) + lambda * (1 - ssim(rendered * mask, gt))
//  i.e. w_l1 = 1 - lambda, w_ssim = -lambda, bias = lambda).
// The reference runs 5 grouped 11x11 conv2d (121 taps each) + ~25 elementwise ATen kernels forward
// and the same again backward, with ~12 image-sized temporaries; here the 11x11 Gaussian window
// (sigma 1.5, zero padding 5: loss_utils.h:50-75, 85-87) is applied separably out of shared
// memory (horizontal pass: 4 outputs per thread sliding over 14 registers fed by LDS.128; vertical pass: two
// rows per thread) and the only temporaries are the three derivative maps.
//
// SSIM algebra (per pixel, per channel; mu = G*x etc., G = the window):
//   a = 2 mu1 mu2 + C1, b = 2 s12 + C2, c = mu1^2 + mu2^2 + C1, d = s1 + s2 + C2   (s1 = G*x^2 - mu1^2 ...)
//   ssim = a b / (c d)
//   d ssim / d mu1     = 2 mu2 (b - a)/(c d) - 2 mu1 ssim (d - c)/(c d)      (holding G*x^2, G*xy fixed)
//   d ssim / d (G*x^2) = -ssim / d
//   d ssim / d (G*xy)  = 2 a / (c d)
//   dL/dx = G*(g m1) + 2 x G*(g m2) + y G*(g m3),   g = w_ssim / (C H W)
// Both kernels are HBM-streaming: forward reads 2 and writes 3 floats per pixel-channel, backward
// reads 5 and writes 1.
#include "common.cuh"

namespace segs {

namespace {

constexpr int TX = 32, TY = 16;        // output tile of one 256-thread CTA (two rows of one column per thread)
constexpr int LTHREADS = 256;
constexpr int HALO = 5;                // window 11
constexpr int LWX = TX + 2 * HALO;     // 42 staged columns
constexpr int LWY = TY + 2 * HALO;     // 26 staged rows
constexpr int WIN = 11;
constexpr int LS = 44;                 // shared row stride in floats: 16-byte aligned rows for LDS.128
constexpr int RUN = 4;                 // outputs per thread in the horizontal pass (14 inputs -> 4 x LDS.128)
constexpr int HRUNS = LWY * (TX / RUN);   // 208 row segments
static_assert(HRUNS <= LTHREADS, "one horizontal run per thread");

struct Window { float w[WIN]; };

// loss_utils::gaussian (loss_utils.h:50-65): exp(-(x-5)^2 / (2 sigma^2)), normalised, sigma = 1.5
Window make_window() {
    Window g;
    float v[WIN];
    float sum = 0.f;
    for (int x = 0; x < WIN; ++x) {
        const int t = x - WIN / 2;
        v[x] = expf(float(-t * t) / (2.0f * 1.5f * 1.5f));
        sum += v[x];
    }
    for (int x = 0; x < WIN; ++x) g.w[x] = v[x] / sum;
    return g;
}

struct LossState {
    float* m1;        // [C*H*W] d ssim / d mu1
    float* m2;        // [C*H*W] d ssim / d (G*x^2)
    float* m3;        // [C*H*W] d ssim / d (G*xy)
    double* partial;  // [2 * blocks]
    uint32_t* ticket; // [1]
    static LossState carve(char* base, size_t n, size_t blocks, size_t* bytes) {
        Carver c(base);
        LossState s;
        s.m1 = c.take<float>(n);
        s.m2 = c.take<float>(n);
        s.m3 = c.take<float>(n);
        s.partial = c.take<double>(2 * blocks);
        s.ticket = c.take<uint32_t>(1);
        if (bytes) *bytes = c.used(base) + 128;
        return s;
    }
};

__device__ __forceinline__ float masked_load(const float* __restrict__ img, const float* __restrict__ row_mask,
                                             int ch, int y, int x, int H, int W) {
    if (x < 0 || x >= W || y < 0 || y >= H) return 0.f;
    float v = __ldg(img + (size_t(ch) * H + y) * W + x);
    if (row_mask) v *= __ldg(row_mask + ch * H + y);
    return v;
}

// horizontal 11-tap pass of one row segment: 4 outputs from 14 staged inputs (read as 4 x float4)
__device__ __forceinline__ void load_run(const float* row, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = reinterpret_cast<const float4*>(row)[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

__global__ void __launch_bounds__(LTHREADS)
ssim_l1_forward_kernel(int C, int H, int W, const float* __restrict__ image, const float* __restrict__ gt,
                       const float* __restrict__ row_mask, const Window win, float w_l1, float w_ssim, float bias,
                       LossState st, float* __restrict__ loss_out)
{
    __shared__ __align__(16) float sx[LWY][LS];
    __shared__ __align__(16) float sy[LWY][LS];
    __shared__ __align__(16) float hb[5][LWY][TX];
    __shared__ double s_red[2][LTHREADS / 32];
    __shared__ bool s_last;

    const int tid = threadIdx.x;
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;

    // staging: thread = (column c of 64, row phase r0 of 4); no divisions, one pointer bump per row
    {
        const int c = tid & 63, r0 = tid >> 6;
        const int gx = x0 + c - HALO;
        const bool col_ok = c < LWX && gx >= 0 && gx < W;
        if (c < LS) {
#pragma unroll
            for (int it = 0; it < (LWY + 3) / 4; ++it) {
                const int r = r0 + 4 * it;
                if (r < LWY) {
                    const int gy = y0 + r - HALO;
                    float vx = 0.f, vy = 0.f;
                    if (col_ok && gy >= 0 && gy < H) {
                        // gaussian_mapper.cpp:915: gt * mask == gt (masked rows of gt are all-zero by construction), but
                        // the product is applied anyway so arbitrary masks behave like the reference
                        const float m = row_mask ? __ldg(row_mask + ch * H + gy) : 1.f;
                        const size_t o = (size_t(ch) * H + gy) * W + gx;
                        vx = __ldg(image + o) * m;
                        vy = __ldg(gt + o) * m;
                    }
                    sx[r][c] = vx;
                    sy[r][c] = vy;
                }
            }
        }
    }
    __syncthreads();
    // horizontal pass: 26 rows x 8 segments of 4 outputs x 5 quantities, sliding over registers
    if (tid < HRUNS) {
        const int r = tid / (TX / RUN), c0 = (tid % (TX / RUN)) * RUN;
        float x[16], y[16];
        load_run(&sx[r][c0], x);
        load_run(&sy[r][c0], y);
        float a[5][RUN];
#pragma unroll
        for (int o = 0; o < RUN; ++o) { a[0][o] = a[1][o] = a[2][o] = a[3][o] = a[4][o] = 0.f; }
#pragma unroll
        for (int k = 0; k < RUN + WIN - 1; ++k) {
            const float xx = x[k] * x[k], yy = y[k] * y[k], xy = x[k] * y[k];
#pragma unroll
            for (int o = 0; o < RUN; ++o) {
                const int tap = k - o;
                if (tap >= 0 && tap < WIN) {
                    const float w = win.w[tap];
                    a[0][o] = fmaf(w, x[k], a[0][o]);
                    a[1][o] = fmaf(w, y[k], a[1][o]);
                    a[2][o] = fmaf(w, xx, a[2][o]);
                    a[3][o] = fmaf(w, yy, a[3][o]);
                    a[4][o] = fmaf(w, xy, a[4][o]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q)
            *reinterpret_cast<float4*>(&hb[q][r][c0]) = make_float4(a[q][0], a[q][1], a[q][2], a[q][3]);
    }
    __syncthreads();
    // vertical pass: thread = column tx, output rows 2*ty and 2*ty + 1 (12 staged rows feed both)
    const int tx = tid & (TX - 1), ty2 = (tid >> 5) * 2;
    float m[5][2];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        float v[WIN + 1];
#pragma unroll
        for (int k = 0; k < WIN + 1; ++k) v[k] = hb[q][ty2 + k][tx];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) { s0 = fmaf(win.w[k], v[k], s0); s1 = fmaf(win.w[k], v[k + 1], s1); }
        m[q][0] = s0; m[q][1] = s1;
    }
    double l1 = 0.0, ss = 0.0;
    const int px = x0 + tx;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int py = y0 + ty2 + j;
        if (px < W && py < H) {
            const float mu1 = m[0][j], mu2 = m[1][j], e11 = m[2][j], e22 = m[3][j], e12 = m[4][j];
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;          // loss_utils.h:101-102
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
            const float a = 2.f * mu12 + C1, b = 2.f * s12 + C2;
            const float c = mu1_sq + mu2_sq + C1, d = s1 + s2 + C2;
            const float rcd = 1.f / (c * d);
            const float ssim = a * b * rcd;                              // loss_utils.h:104
            const size_t o = (size_t(ch) * H + py) * W + px;
            st.m1[o] = 2.f * mu2 * (b - a) * rcd - 2.f * mu1 * ssim * (d - c) * rcd;
            st.m2[o] = -ssim / d;
            st.m3[o] = 2.f * a * rcd;
            l1 += (double)fabsf(sx[ty2 + j + HALO][tx + HALO] - sy[ty2 + j + HALO][tx + HALO]);
            ss += (double)ssim;
        }
    }
    // deterministic reduction: lanes -> warps -> CTA partial -> (last CTA) fixed-order total
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, o);
        ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = l1; s_red[1][tid >> 5] = ss; }
    __syncthreads();
    const unsigned blocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < LTHREADS / 32; ++w) { t1 += s_red[0][w]; t2 += s_red[1][w]; }
        st.partial[2 * bid] = t1;
        st.partial[2 * bid + 1] = t2;
        __threadfence();
        s_last = atomicAdd(st.ticket, 1u) == blocks - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double t1 = 0.0, t2 = 0.0;
    for (unsigned i = tid; i < blocks; i += LTHREADS) {
        t1 += __ldcg(st.partial + 2 * i);
        t2 += __ldcg(st.partial + 2 * i + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t1 += __shfl_xor_sync(0xFFFFFFFFu, t1, o);
        t2 += __shfl_xor_sync(0xFFFFFFFFu, t2, o);
    }
    __syncthreads();
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = t1; s_red[1][tid >> 5] = t2; }
    __syncthreads();
    if (tid == 0) {
        t1 = 0.0; t2 = 0.0;
        for (int w = 0; w < LTHREADS / 32; ++w) { t1 += s_red[0][w]; t2 += s_red[1][w]; }
        const double n = double(C) * H * W;
        const float l1_mean = float(t1 / n), ssim_mean = float(t2 / n);
        loss_out[0] = l1_mean;
        loss_out[1] = ssim_mean;
        loss_out[2] = w_l1 * l1_mean + w_ssim * ssim_mean + bias;
    }
}

__device__ __forceinline__ float map_load(const float* __restrict__ m, int ch, int y, int x, int H, int W) {
    if (x < 0 || x >= W || y < 0 || y >= H) return 0.f;
    return __ldg(m + (size_t(ch) * H + y) * W + x);
}

__global__ void __launch_bounds__(LTHREADS)
ssim_l1_backward_kernel(int C, int H, int W, const float* __restrict__ image, const float* __restrict__ gt,
                        const float* __restrict__ row_mask, const Window win, float w_l1, float w_ssim,
                        const float* __restrict__ dL_dloss, LossState st, float* __restrict__ dL_dimage)
{
    __shared__ __align__(16) float sm[3][LWY][LS];
    __shared__ __align__(16) float hb[3][LWY][TX];

    const int tid = threadIdx.x;
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;

    {
        const int c = tid & 63, r0 = tid >> 6;
        const int gx = x0 + c - HALO;
        const bool col_ok = c < LWX && gx >= 0 && gx < W;
        if (c < LS) {
#pragma unroll
            for (int it = 0; it < (LWY + 3) / 4; ++it) {
                const int r = r0 + 4 * it;
                if (r < LWY) {
                    const int gy = y0 + r - HALO;
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                    if (col_ok && gy >= 0 && gy < H) {
                        const size_t o = (size_t(ch) * H + gy) * W + gx;
                        v0 = __ldg(st.m1 + o); v1 = __ldg(st.m2 + o); v2 = __ldg(st.m3 + o);
                    }
                    sm[0][r][c] = v0; sm[1][r][c] = v1; sm[2][r][c] = v2;
                }
            }
        }
    }
    __syncthreads();
    if (tid < HRUNS) {
        const int r = tid / (TX / RUN), c0 = (tid % (TX / RUN)) * RUN;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float v[16];
            load_run(&sm[q][r][c0], v);
            float a[RUN] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < RUN + WIN - 1; ++k) {
#pragma unroll
                for (int o = 0; o < RUN; ++o) {
                    const int tap = k - o;
                    if (tap >= 0 && tap < WIN) a[o] = fmaf(win.w[tap], v[k], a[o]);
                }
            }
            *reinterpret_cast<float4*>(&hb[q][r][c0]) = make_float4(a[0], a[1], a[2], a[3]);
        }
    }
    __syncthreads();
    const int tx = tid & (TX - 1), ty2 = (tid >> 5) * 2;
    float cq[3][2];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        float v[WIN + 1];
#pragma unroll
        for (int k = 0; k < WIN + 1; ++k) v[k] = hb[q][ty2 + k][tx];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) { s0 = fmaf(win.w[k], v[k], s0); s1 = fmaf(win.w[k], v[k + 1], s1); }
        cq[q][0] = s0; cq[q][1] = s1;
    }
    const float up = dL_dloss ? __ldg(dL_dloss) : 1.f;
    const float inv_n = 1.f / (float(C) * float(H) * float(W));
    const int px = x0 + tx;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int py = y0 + ty2 + j;
        if (px >= W || py >= H) continue;
        const size_t o = (size_t(ch) * H + py) * W + px;
        const float m = row_mask ? __ldg(row_mask + ch * H + py) : 1.f;
        const float x = __ldg(image + o) * m, y = __ldg(gt + o) * m;
        const float diff = x - y;
        const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);   // abs backward: sign(0) = 0
        const float g = (w_ssim * inv_n) * (cq[0][j] + 2.f * x * cq[1][j] + y * cq[2][j]) + (w_l1 * inv_n) * sgn;
        dL_dimage[o] = up * m * g;
    }
}

// 0.01 * scaling.prod(1).mean() of gaussian_mapper.cpp:922-925, forward value + gradient added in place
__global__ void __launch_bounds__(256)
scaling_reg_kernel(int n, const float* __restrict__ scaling, float weight, const float* __restrict__ dL_dloss,
                   float* __restrict__ dL_dscaling, float* __restrict__ reg_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float p = 0.f;
    if (i < n) {
        const float s0 = scaling[3 * i], s1 = scaling[3 * i + 1], s2 = scaling[3 * i + 2];
        p = s0 * s1 * s2;
        if (dL_dscaling) {
            const float g = (dL_dloss ? __ldg(dL_dloss) : 1.f) * weight / float(n);
            dL_dscaling[3 * i] += g * s1 * s2;
            dL_dscaling[3 * i + 1] += g * s0 * s2;
            dL_dscaling[3 * i + 2] += g * s0 * s1;
        }
    }
    if (reg_out) {
        // one atomic per CTA (a single address: per-warp atomics serialise in L2)
        __shared__ float s_p[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xFFFFFFFFu, p, o);
        if ((threadIdx.x & 31) == 0) s_p[threadIdx.x >> 5] = p;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += s_p[w];
            if (t != 0.f) atomicAdd(reg_out, t * (weight / float(n)));
        }
    }
}

}  // namespace

}  // namespace segs

using namespace segs;

extern "C" {

size_t segs_loss_state_bytes(int C, int H, int W)
{
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    size_t bytes = 0;
    const size_t blocks = size_t((W + TX - 1) / TX) * ((H + TY - 1) / TY) * C;
    LossState::carve(nullptr, size_t(C) * H * W, blocks, &bytes);
    return bytes;
}

int segs_loss_l1_ssim_forward(int C, int H, int W, const float* image, const float* gt, const float* row_mask,
                              float w_l1, float w_ssim, float bias, float* loss_out, char* state, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) { set_error("loss: invalid sizes C=%d H=%d W=%d", C, H, W); return SEGS_ERR_INVALID_ARG; }
    if (!image || !gt || !loss_out || !state) { set_error("loss: NULL required pointer"); return SEGS_ERR_INVALID_ARG; }
    const dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, C);
    if (grid.y > 65535) { set_error("loss: image too tall"); return SEGS_ERR_INVALID_ARG; }
    const size_t blocks = size_t(grid.x) * grid.y * grid.z;
    LossState st = LossState::carve(state, size_t(C) * H * W, blocks, nullptr);
    SEGS_CUDA_CHECK(cudaMemsetAsync(st.ticket, 0, sizeof(uint32_t), stream));
    static const Window win = make_window();
    ssim_l1_forward_kernel<<<grid, LTHREADS, 0, stream>>>(C, H, W, image, gt, row_mask, win, w_l1, w_ssim, bias, st, loss_out);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int segs_loss_l1_ssim_backward(int C, int H, int W, const float* image, const float* gt, const float* row_mask,
                               float w_l1, float w_ssim, const float* dL_dloss, char* state, float* dL_dimage,
                               void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) { set_error("loss: invalid sizes C=%d H=%d W=%d", C, H, W); return SEGS_ERR_INVALID_ARG; }
    if (!image || !gt || !dL_dimage || !state) { set_error("loss: NULL required pointer"); return SEGS_ERR_INVALID_ARG; }
    const dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, C);
    if (grid.y > 65535) { set_error("loss: image too tall"); return SEGS_ERR_INVALID_ARG; }
    const size_t blocks = size_t(grid.x) * grid.y * grid.z;
    LossState st = LossState::carve(state, size_t(C) * H * W, blocks, nullptr);
    static const Window win = make_window();
    ssim_l1_backward_kernel<<<grid, LTHREADS, 0, stream>>>(C, H, W, image, gt, row_mask, win, w_l1, w_ssim, dL_dloss, st, dL_dimage);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int segs_scaling_reg(int n, const float* scaling, float weight, const float* dL_dloss, float* dL_dscaling,
                     float* reg_out, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n < 0 || (n > 0 && !scaling)) { set_error("scaling_reg: invalid argument"); return SEGS_ERR_INVALID_ARG; }
    if (n == 0) return SEGS_OK;
    scaling_reg_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, scaling, weight, dL_dloss, dL_dscaling, reg_out);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // extern "C"
