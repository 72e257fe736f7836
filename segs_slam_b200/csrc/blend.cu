// Tile alpha-blending forward and backward for sm_100a.
//
// Replaces FORWARD::render / BACKWARD::render (renderCUDA) of the reference
// (cuda_rasterizer/forward.cu:339-478, backward.cu:399-557, 624-657).
//
// Design (B200):
//  * One 256-thread CTA per 16x16 tile, but each WARP owns an 8x4-pixel sub-tile (not a
//    16x2 strip): the per-warp footprint is compact, so most of a tile's Gaussians miss it.
//  * A tile's sorted Gaussian list is streamed in batches of 256 records.  Each thread
//    gathers one 48-byte record (mean+extent | conic+opacity | rgb+depth) with three 16-byte
//    cp.async (LDGSTS) copies straight into shared memory, double-buffered, with the record
//    ids prefetched one batch further ahead, so the gather latency of batch b+1 hides behind
//    the blending of batch b.  Colours ride in the record: the reference's per-blended-pair
//    global colour loads (forward.cu:433) are gone.
//  * Per 32 staged Gaussians each lane tests ONE Gaussian's conservative extent against the
//    warp's sub-tile; a ballot turns that into a bitmask and the warp only runs the per-pixel
//    maths for the set bits.  Culled Gaussians cannot pass the reference's alpha >= 1/255
//    test for any pixel of the sub-tile, so the blended result, final_T and n_contrib are
//    unchanged (n_contrib counts list positions, which are tracked arithmetically).
//  * Early termination is per warp (all 32 pixels saturated) on top of the reference's
//    per-CTA vote.
//  * Backward: gradients of one Gaussian are reduced across the warp with a 9-shuffle
//    transpose-reduce and sent as ONE reduction per value per (Gaussian, 8x4 sub-tile) into a
//    packed 48-byte accumulator (8 lanes hit 8 consecutive floats) — instead of the
//    reference's 9 global atomics per (Gaussian, pixel).  (Shared-memory float atomics
//    compile to CAS loops on sm_100a, so a per-tile shared accumulator costs more
//    instructions than it saves; measured in profiles/r1.)
//
// The per-pair arithmetic (power, alpha, T update, colour accumulation) follows the
// reference's operations and FMA contraction one by one so that n_contrib / final_T decisions
// are bit-identical.
#include "common.cuh"

namespace segs {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int BATCH = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// power of forward.cu:413 / backward.cu:494 with nvcc's contraction:
//   -0.5f * (con.x*d.x*d.x + con.z*d.y*d.y) - con.y*d.x*d.y
__device__ __forceinline__ float gauss_power(float dx, float dy, float cx, float cy, float cz) {
    // SASS of the reference: FFMA(a, -0.5, -(dy*(dx*cy)))
    const float a = __fmaf_rn(dx, __fmul_rn(dx, cx), __fmul_rn(dy, __fmul_rn(dy, cz)));
    return __fmaf_rn(a, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, cy)));
}

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// conservative sub-tile test; written so that NaNs never cull
__device__ __forceinline__ bool extent_hits(float4 g0, float wx0, float wx1, float wy0, float wy1) {
    return !(g0.x + g0.z < wx0 || g0.x - g0.z > wx1 || g0.y + g0.w < wy0 || g0.y - g0.w > wy1);
}

struct SubTile {
    int tile_x, tile_y;
    int px, py;          // this lane's pixel
    float wx0, wx1, wy0, wy1;
};

__device__ __forceinline__ SubTile make_subtile(int tile, int grid_x) {
    SubTile s;
    s.tile_x = tile % grid_x;
    s.tile_y = tile / grid_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sx = s.tile_x * TILE_X + (warp & 1) * 8;
    const int sy = s.tile_y * TILE_Y + (warp >> 1) * 4;
    s.px = sx + (lane & 7);
    s.py = sy + (lane >> 3);
    s.wx0 = (float)sx; s.wx1 = (float)(sx + 7);
    s.wy0 = (float)sy; s.wy1 = (float)(sy + 3);
    return s;
}

// =======================================================================================
// forward
// =======================================================================================
__global__ void __launch_bounds__(TILE_PIX)
blend_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                     int W, int H, int grid_x, const float4* __restrict__ rec,
                     const float* __restrict__ bg_color, float* __restrict__ final_T,
                     uint32_t* __restrict__ n_contrib, float* __restrict__ out_color)
{
    __shared__ float4 s_rec[2][3][BATCH];

    const int tile = blockIdx.x;
    const SubTile st = make_subtile(tile, grid_x);
    const bool inside = st.px < W && st.py < H;
    const float pixfx = (float)st.px, pixfy = (float)st.py;
    const int tid = threadIdx.x;

    const uint2 range = ranges[tile];
    const int len = (int)(range.y - range.x);
    const int nbatches = (len + BATCH - 1) / BATCH;
    const uint32_t* list = point_list + range.x;

    bool done = !inside;
    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last_contributor = 0;

    // software pipeline: ids two batches ahead, records one batch ahead
    uint32_t id_next = (tid < len) ? __ldg(list + tid) : 0u;             // ids of batch 0
    {
        if (tid < len) {
            const float4* src = rec + 3 * size_t(id_next);
            cp_async16(&s_rec[0][0][tid], src);
            cp_async16(&s_rec[0][1][tid], src + 1);
            cp_async16(&s_rec[0][2][tid], src + 2);
        }
        cp_async_commit();
        id_next = (BATCH + tid < len) ? __ldg(list + BATCH + tid) : 0u;  // ids of batch 1
    }

    for (int b = 0; b < nbatches; ++b) {
        const int buf = b & 1;
        // issue gather of batch b+1 (its ids were loaded an iteration ago)
        if (b + 1 < nbatches) {
            const int p = (b + 1) * BATCH + tid;
            if (p < len) {
                const float4* src = rec + 3 * size_t(id_next);
                cp_async16(&s_rec[buf ^ 1][0][tid], src);
                cp_async16(&s_rec[buf ^ 1][1][tid], src + 1);
                cp_async16(&s_rec[buf ^ 1][2][tid], src + 2);
            }
            cp_async_commit();
            const int p2 = (b + 2) * BATCH + tid;
            id_next = (p2 < len) ? __ldg(list + p2) : 0u;
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        // batch b visible to all; also the CTA-wide "everyone is done" vote of forward.cu:387
        if (__syncthreads_and(done)) break;

        if (!__all_sync(FULL, done)) {
            const int n_in = min(BATCH, len - b * BATCH);
            const int base = b * BATCH;
            for (int c = 0; c * 32 < n_in; ++c) {
                const int slot_l = c * 32 + (tid & 31);
                bool hit = false;
                if (slot_l < n_in) hit = extent_hits(s_rec[buf][0][slot_l], st.wx0, st.wx1, st.wy0, st.wy1);
                unsigned mask = __ballot_sync(FULL, hit);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int slot = c * 32 + j;
                    // Straight-line body (no divergent branches): every lane evaluates the pair,
                    // lanes that are done / miss the reference's tests simply do not commit.
                    const float4 g0 = s_rec[buf][0][slot];
                    const float4 g1 = s_rec[buf][1][slot];
                    const float4 g2 = s_rec[buf][2][slot];
                    const float dx = __fsub_rn(g0.x, pixfx);
                    const float dy = __fsub_rn(g0.y, pixfy);
                    const float power = gauss_power(dx, dy, g1.x, g1.y, g1.z);
                    const float alpha = fminf(0.99f, __fmul_rn(g1.w, expf(power)));
                    const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
                    // forward.cu:414-429: power > 0 / alpha < 1/255 skip; T < 1e-4 stops the pixel
                    const bool contrib = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                    const bool stop = contrib && (test_T < 0.0001f);
                    const bool blend = contrib && !stop;
                    done = done || stop;
                    if (blend) {
                        C0 = __fmaf_rn(T, __fmul_rn(alpha, g2.x), C0);
                        C1 = __fmaf_rn(T, __fmul_rn(alpha, g2.y), C1);
                        C2 = __fmaf_rn(T, __fmul_rn(alpha, g2.z), C2);
                        T = test_T;
                        last_contributor = (uint32_t)(base + slot + 1);
                    }
                }
                if (__all_sync(FULL, done)) break;
            }
        }
        __syncthreads();   // everyone finished reading s_rec[buf] before it is refilled
    }
    cp_async_wait<0>();

    if (inside) {
        const size_t pix = size_t(st.py) * W + st.px;
        const size_t plane = size_t(H) * W;
        final_T[pix] = T;
        n_contrib[pix] = last_contributor;
        out_color[pix] = __fmaf_rn(T, __ldg(bg_color + 0), C0);
        out_color[plane + pix] = __fmaf_rn(T, __ldg(bg_color + 1), C1);
        out_color[2 * plane + pix] = __fmaf_rn(T, __ldg(bg_color + 2), C2);
    }
}

// =======================================================================================
// backward
// =======================================================================================
// Sum v[0..7] over the warp with 9 shuffles.  Afterwards lane l holds the total of value
// ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1) (all four lanes of a quad hold the same total).
__device__ __forceinline__ float transpose_reduce8(float (&v)[8], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h16 ? v[i] : v[i + 4];
        const float keep = h16 ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h8 ? v[i] : v[i + 2];
        const float keep = h8 ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    {
        const float send = h4 ? v[0] : v[1];
        const float keep = h4 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 2);
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
    return v[0];
}

__global__ void __launch_bounds__(TILE_PIX)
blend_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      int W, int H, int grid_x, const float4* __restrict__ rec,
                      const float* __restrict__ bg_color, const float* __restrict__ final_T,
                      const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                      float4* __restrict__ acc)
{
    __shared__ float4 s_rec[2][3][BATCH];
    __shared__ uint32_t s_id[2][BATCH];
    __shared__ int s_bmax;

    const int tile = blockIdx.x;
    const SubTile st = make_subtile(tile, grid_x);
    const bool inside = st.px < W && st.py < H;
    const float pixfx = (float)st.px, pixfy = (float)st.py;
    const int tid = threadIdx.x, lane = tid & 31;

    const uint2 range = ranges[tile];
    const uint32_t* list = point_list + range.x;

    const size_t pix = size_t(st.py) * W + st.px;
    const size_t plane = size_t(H) * W;
    const float T_final = inside ? final_T[pix] : 0.f;
    const int my_last = inside ? (int)n_contrib[pix] : 0;
    float dpix0 = 0.f, dpix1 = 0.f, dpix2 = 0.f;
    if (inside) {
        dpix0 = dL_dpix[pix];
        dpix1 = dL_dpix[plane + pix];
        dpix2 = dL_dpix[2 * plane + pix];
    }
    const float bg_dot_dpixel = __ldg(bg_color) * dpix0 + __ldg(bg_color + 1) * dpix1 + __ldg(bg_color + 2) * dpix2;
    const float ddelx_dx = (float)(0.5 * W);
    const float ddely_dy = (float)(0.5 * H);

    // last list position any pixel of the warp / CTA needs
    const int wmax = __reduce_max_sync(FULL, my_last);
    if (tid == 0) s_bmax = 0;
    __syncthreads();
    if (lane == 0 && wmax > 0) atomicMax(&s_bmax, wmax);
    __syncthreads();
    const int bmax = s_bmax;
    if (bmax == 0) return;
    const int nbatches = (bmax + BATCH - 1) / BATCH;

    float T = T_final;
    float accum0 = 0.f, accum1 = 0.f, accum2 = 0.f;
    float last_alpha = 0.f, lastc0 = 0.f, lastc1 = 0.f, lastc2 = 0.f;

    // batch b (b = 0 is the BACK of the list) covers positions [lo_b, hi_b), slot s <-> lo_b + s
    auto batch_lo = [&](int b) { return max(0, bmax - (b + 1) * BATCH); };
    auto batch_hi = [&](int b) { return bmax - b * BATCH; };

    uint32_t id_next;
    {
        const int lo = batch_lo(0), n0 = batch_hi(0) - lo;
        id_next = (tid < n0) ? __ldg(list + lo + tid) : 0u;
        if (tid < n0) {
            const float4* src = rec + 3 * size_t(id_next);
            cp_async16(&s_rec[0][0][tid], src);
            cp_async16(&s_rec[0][1][tid], src + 1);
            cp_async16(&s_rec[0][2][tid], src + 2);
            s_id[0][tid] = id_next;
        }
        cp_async_commit();
        if (nbatches > 1) {
            const int lo1 = batch_lo(1), n1 = batch_hi(1) - lo1;
            id_next = (tid < n1) ? __ldg(list + lo1 + tid) : 0u;
        }
    }

    for (int b = 0; b < nbatches; ++b) {
        const int buf = b & 1;
        const int lo = batch_lo(b);
        const int n_in = batch_hi(b) - lo;
        if (b + 1 < nbatches) {
            const int n1 = batch_hi(b + 1) - batch_lo(b + 1);
            if (tid < n1) {
                const float4* src = rec + 3 * size_t(id_next);
                cp_async16(&s_rec[buf ^ 1][0][tid], src);
                cp_async16(&s_rec[buf ^ 1][1][tid], src + 1);
                cp_async16(&s_rec[buf ^ 1][2][tid], src + 2);
                s_id[buf ^ 1][tid] = id_next;
            }
            cp_async_commit();
            if (b + 2 < nbatches) {
                const int lo2 = batch_lo(b + 2), n2 = batch_hi(b + 2) - lo2;
                id_next = (tid < n2) ? __ldg(list + lo2 + tid) : 0u;
            }
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (lo < wmax) {
            for (int c = (n_in - 1) >> 5; c >= 0; --c) {
                const int slot_l = c * 32 + lane;
                bool hit = false;
                if (slot_l < n_in && lo + slot_l < wmax)
                    hit = extent_hits(s_rec[buf][0][slot_l], st.wx0, st.wx1, st.wy0, st.wy1);
                unsigned mask = __ballot_sync(FULL, hit);
                while (mask) {
                    const int j = 31 - __clz(mask);
                    mask &= ~(1u << j);
                    const int slot = c * 32 + j;
                    const int pos = lo + slot;
                    const float4 g0 = s_rec[buf][0][slot];
                    const float4 g1 = s_rec[buf][1][slot];
                    const float dx = __fsub_rn(g0.x, pixfx);
                    const float dy = __fsub_rn(g0.y, pixfy);
                    const float power = gauss_power(dx, dy, g1.x, g1.y, g1.z);
                    const float G_raw = expf(power);
                    const float alpha_raw = fminf(0.99f, __fmul_rn(g1.w, G_raw));
                    // backward.cu:486-501: behind the last contributor / power > 0 / alpha < 1/255
                    const bool act = (pos < my_last) && !(power > 0.0f) && !(alpha_raw < 1.0f / 255.0f);
                    if (!__any_sync(FULL, act)) continue;

                    // Straight-line maths: lanes that do not contribute run with G = alpha = 0, which
                    // makes every gradient term exactly 0 and T unchanged; only the recurrence state
                    // (accum_rec / last_color / last_alpha) needs explicit selects.
                    const float G = act ? G_raw : 0.f;
                    const float alpha = act ? alpha_raw : 0.f;
                    const float4 g2 = s_rec[buf][2][slot];
                    const float rcp_oma = rcp_approx(1.f - alpha);           // 1 - alpha in [0.01, 1]: one MUFU.RCP
                    T = T * rcp_oma;                                         // T / (1 - alpha)
                    const float dchannel_dcolor = alpha * T;
                    const float na0 = last_alpha * lastc0 + (1.f - last_alpha) * accum0;
                    const float na1 = last_alpha * lastc1 + (1.f - last_alpha) * accum1;
                    const float na2 = last_alpha * lastc2 + (1.f - last_alpha) * accum2;
                    accum0 = act ? na0 : accum0;
                    accum1 = act ? na1 : accum1;
                    accum2 = act ? na2 : accum2;
                    lastc0 = act ? g2.x : lastc0;
                    lastc1 = act ? g2.y : lastc1;
                    lastc2 = act ? g2.z : lastc2;
                    last_alpha = act ? alpha : last_alpha;
                    float dL_dalpha = (g2.x - accum0) * dpix0;
                    dL_dalpha += (g2.y - accum1) * dpix1;
                    dL_dalpha += (g2.z - accum2) * dpix2;
                    dL_dalpha *= T;
                    dL_dalpha += (-T_final * rcp_oma) * bg_dot_dpixel;

                    const float dL_dG = g1.w * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * g1.x - gdy * g1.y;
                    const float dG_ddely = -gdy * g1.z - gdx * g1.y;
                    float v[8];
                    v[0] = dL_dG * dG_ddelx * ddelx_dx;        // dL_dmean2D.x
                    v[1] = dL_dG * dG_ddely * ddely_dy;        // dL_dmean2D.y
                    v[2] = -0.5f * gdx * dx * dL_dG;           // dL_dconic.x
                    v[3] = -0.5f * gdx * dy * dL_dG;           // dL_dconic.y
                    v[4] = -0.5f * gdy * dy * dL_dG;           // dL_dconic.w
                    v[5] = G * dL_dalpha;                      // dL_dopacity
                    v[6] = dchannel_dcolor * dpix0;            // dL_dcolor.r
                    v[7] = dchannel_dcolor * dpix1;            // dL_dcolor.g
                    float v8 = dchannel_dcolor * dpix2;        // dL_dcolor.b
                    const float tot = transpose_reduce8(v, lane);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(FULL, v8, o);
                    // one reduction per (Gaussian, warp sub-tile) and value, straight into the packed
                    // 48-byte accumulator: 8 lanes hit 8 consecutive floats, lane 1 the ninth
                    float* dst = reinterpret_cast<float*>(acc + 3 * size_t(s_id[buf][slot]));
                    if ((lane & 3) == 0) {
                        const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        atomicAdd(dst + k, tot);
                    } else if (lane == 1) {
                        atomicAdd(dst + 8, v8);
                    }
                }
            }
        }
        __syncthreads();   // all warps finished with s_rec[buf] / s_id[buf] before they are refilled
    }
}

}  // namespace

int launch_blend_forward(const ViewParams& vp, const GeomState& g, const BinningState& b,
                         ImageState& img, const float* background, float* out_color,
                         cudaStream_t stream)
{
    const int T = vp.grid_x * vp.grid_y;
    const uint32_t* point_list = b.point_list;
    blend_forward_kernel<<<T, TILE_PIX, 0, stream>>>(img.ranges, point_list, vp.W, vp.H, vp.grid_x, g.rec,
                                                    background, img.final_T, img.n_contrib, out_color);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int launch_blend_backward(const ViewParams& vp, const GeomState& g, const BinningState& b,
                          const ImageState& img, const float* background,
                          const float* dL_dpix, cudaStream_t stream)
{
    const int T = vp.grid_x * vp.grid_y;
    const uint32_t* point_list = b.point_list;
    blend_backward_kernel<<<T, TILE_PIX, 0, stream>>>(img.ranges, point_list, vp.W, vp.H, vp.grid_x, g.rec,
                                                     background, img.final_T, img.n_contrib, dL_dpix, g.acc);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // namespace segs
