// Tile alpha-blending forward and backward for sm_100a.
//
// Replaces FORWARD::render / BACKWARD::render (renderCUDA) of the reference
// (cuda_rasterizer/forward.cu:339-478, backward.cu:399-557, 624-657).
//
// Design (B200):
//  * One 128-thread CTA per 16x16 tile; each WARP owns a compact 8x8-pixel sub-tile (most of a
//    tile's Gaussians miss it) and each lane two pixels of one column, which share dx and all
//    per-Gaussian operands and halve the per-pixel share of the per-hit overheads.
//  * A tile's sorted Gaussian list is streamed in batches of 256 records.  Each thread
//    gathers two 48-byte records (mean+extent | conic+opacity | rgb+depth) with three 16-byte
//    cp.async (LDGSTS) copies straight into shared memory, double-buffered, with the record
//    ids prefetched one batch further ahead, so the gather latency of batch b+1 hides behind
//    the blending of batch b.  Colours ride in the record: the reference's per-blended-pair
//    global colour loads (forward.cu:433) are gone.
//  * Every staged Gaussian's conservative extent is tested ONCE per CTA against the tile's four
//    sub-tiles (a 4-bit mask in shared memory); per 32 staged Gaussians a warp ballots its own bit
//    and only runs the per-pixel maths for the set bits.  Culled Gaussians cannot pass the reference's alpha >= 1/255
//    test for any pixel of the sub-tile, so the blended result, final_T and n_contrib are
//    unchanged (n_contrib counts list positions, which are tracked arithmetically).
//  * Early termination is per warp (all 32 pixels saturated) on top of the reference's
//    per-CTA vote.
//  * Backward: gradients of one Gaussian are reduced across the warp with a 9-shuffle
//    transpose-reduce and sent as ONE reduction per value per (Gaussian, 8x8 sub-tile) into a
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

// dev aid (make EXTRA=-DSEGS_BLEND_STATS): work counters of the two kernels, read with segs_debug_blend_stats
#ifdef SEGS_BLEND_STATS
__device__ unsigned long long g_blend_stats[8];
#define BSTAT(i, v) do { const unsigned long long bstat_v = (unsigned long long)(v); if ((threadIdx.x & 31) == 0) atomicAdd(&g_blend_stats[i], bstat_v); } while (0)
#else
#define BSTAT(i, v) do { (void)sizeof(v); } while (0)
#endif

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

// Each warp owns an 8x8-pixel sub-tile and each lane TWO pixels of one column, (x, y) and (x, y + 4):
// the two pixels share dx and every per-Gaussian operand, give the scheduler two independent dependency
// chains, and halve the per-(Gaussian, pixel) share of everything that is paid once per warp and hit —
// the cull test, the loop control, and in the backward the 9-value warp reduction and its 9 RED.ADDs.
constexpr int BLEND_THREADS = 128;      // 4 warps x (8x8 pixels) = one 16x16 tile
constexpr int PER_THREAD = BATCH / BLEND_THREADS;

struct SubTile {
    int px, py0, py1;    // this lane's two pixels: (px, py0) and (px, py1 = py0 + 4)
    float wx0, wx1, wy0, wy1;
};

__device__ __forceinline__ SubTile make_subtile(int tile, int grid_x) {
    SubTile s;
    const int tile_x = tile % grid_x, tile_y = tile / grid_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sx = tile_x * TILE_X + (warp & 1) * 8;
    const int sy = tile_y * TILE_Y + (warp >> 1) * 8;
    s.px = sx + (lane & 7);
    s.py0 = sy + (lane >> 3);
    s.py1 = s.py0 + 4;
    s.wx0 = (float)sx; s.wx1 = (float)(sx + 7);
    s.wy0 = (float)sy; s.wy1 = (float)(sy + 7);
    return s;
}

// 4-bit mask of the tile's 8x8 sub-tiles (bit = warp index: x half | y half << 1) the conservative extent of one
// staged Gaussian reaches; computed ONCE per CTA and record (each warp used to run its own test on every record).
// Written so that NaNs never cull.
__device__ __forceinline__ uint32_t subtile_mask(float4 g0, float tx0, float ty0) {
    const float xlo = g0.x - g0.z, xhi = g0.x + g0.z, ylo = g0.y - g0.w, yhi = g0.y + g0.w;
    const bool c0 = !(xhi < tx0 || xlo > tx0 + 7.f), c1 = !(xhi < tx0 + 8.f || xlo > tx0 + 15.f);
    const bool r0 = !(yhi < ty0 || ylo > ty0 + 7.f), r1 = !(yhi < ty0 + 8.f || ylo > ty0 + 15.f);
    return (uint32_t)(c0 && r0) | ((uint32_t)(c1 && r0) << 1) | ((uint32_t)(c0 && r1) << 2) | ((uint32_t)(c1 && r1) << 3);
}

__device__ __forceinline__ void compute_masks(uint8_t* dst, const float4* g0s, int n, float tx0, float ty0) {
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int s = k * BLEND_THREADS + threadIdx.x;
        if (s < n) dst[s] = (uint8_t)subtile_mask(g0s[s], tx0, ty0);
    }
}

// gather `n` records (list positions first .. first + n - 1 of `list`) into one shared-memory stage
__device__ __forceinline__ void stage_records(float4 (*dst)[BATCH], uint32_t* dst_id, const float4* __restrict__ rec,
                                              const uint32_t (&ids)[PER_THREAD], int n) {
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int s = k * BLEND_THREADS + threadIdx.x;
        if (s < n) {
            const float4* src = rec + 3 * size_t(ids[k]);
            cp_async16(&dst[0][s], src);
            cp_async16(&dst[1][s], src + 1);
            cp_async16(&dst[2][s], src + 2);
            if (dst_id) dst_id[s] = ids[k];
        }
    }
}

__device__ __forceinline__ void load_ids(uint32_t (&ids)[PER_THREAD], const uint32_t* __restrict__ list, int first, int n) {
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int s = k * BLEND_THREADS + threadIdx.x;
        ids[k] = (s < n) ? __ldg(list + first + s) : 0u;
    }
}

// =======================================================================================
// forward
// =======================================================================================
__global__ void __launch_bounds__(BLEND_THREADS)
blend_forward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                     int W, int H, int grid_x, const float4* __restrict__ rec,
                     const float* __restrict__ bg_color, float* __restrict__ final_T,
                     uint32_t* __restrict__ n_contrib, float* __restrict__ out_color)
{
    __shared__ float4 s_rec[2][3][BATCH];
    __shared__ uint8_t s_mask[BATCH];

    const int tile = blockIdx.x;
    const SubTile st = make_subtile(tile, grid_x);
    const bool inside0 = st.px < W && st.py0 < H, inside1 = st.px < W && st.py1 < H;
    const float pixfx = (float)st.px, pixfy0 = (float)st.py0, pixfy1 = (float)st.py1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float tile_x0 = (float)((tile % grid_x) * TILE_X), tile_y0 = (float)((tile / grid_x) * TILE_Y);

    const uint2 range = ranges[tile];
    const int len = (int)(range.y - range.x);
    const int nbatches = (len + BATCH - 1) / BATCH;
    const uint32_t* list = point_list + range.x;

    bool done0 = !inside0, done1 = !inside1;
    float T0 = 1.0f, T1 = 1.0f;
    float C00 = 0.f, C01 = 0.f, C02 = 0.f, C10 = 0.f, C11 = 0.f, C12 = 0.f;
    uint32_t last0 = 0, last1 = 0;

    // software pipeline: ids two batches ahead, records one batch ahead
    uint32_t ids[PER_THREAD];
    load_ids(ids, list, 0, min(BATCH, len));
    stage_records(s_rec[0], nullptr, rec, ids, min(BATCH, len));
    cp_async_commit();
    load_ids(ids, list, BATCH, min(BATCH, len - BATCH));

    for (int b = 0; b < nbatches; ++b) {
        const int buf = b & 1;
        if (b + 1 < nbatches) {
            stage_records(s_rec[buf ^ 1], nullptr, rec, ids, min(BATCH, len - (b + 1) * BATCH));
            cp_async_commit();
            load_ids(ids, list, (b + 2) * BATCH, min(BATCH, len - (b + 2) * BATCH));
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        // batch b visible to all; also the CTA-wide "everyone is done" vote of forward.cu:387
        if (__syncthreads_and(done0 && done1)) break;
        const int n_in = min(BATCH, len - b * BATCH);
        compute_masks(s_mask, s_rec[buf][0], n_in, tile_x0, tile_y0);
        __syncthreads();
        if (threadIdx.x == 0) BSTAT(0, n_in);                 // records staged

        if (!__all_sync(FULL, done0 && done1)) {
            const int base = b * BATCH;
            for (int c = 0; c * 32 < n_in; ++c) {
                const int slot_l = c * 32 + lane;
                const bool hit = slot_l < n_in && ((s_mask[slot_l] >> warp) & 1u);
                unsigned mask = __ballot_sync(FULL, hit);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int slot = c * 32 + j;
                    BSTAT(1, 1);                                  // (Gaussian, sub-tile) pairs evaluated
                    // Straight-line body (no divergent branches): every lane evaluates both of its pixels,
                    // pixels that are done / miss the reference's tests simply do not commit.
                    const float4 g0 = s_rec[buf][0][slot];
                    const float4 g1 = s_rec[buf][1][slot];
                    const float dx = __fsub_rn(g0.x, pixfx);
                    const float dy0 = __fsub_rn(g0.y, pixfy0), dy1 = __fsub_rn(g0.y, pixfy1);
                    const float power0 = gauss_power(dx, dy0, g1.x, g1.y, g1.z);
                    const float power1 = gauss_power(dx, dy1, g1.x, g1.y, g1.z);
                    const float alpha0 = fminf(0.99f, __fmul_rn(g1.w, expf(power0)));
                    const float alpha1 = fminf(0.99f, __fmul_rn(g1.w, expf(power1)));
                    // forward.cu:414-429: power > 0 / alpha < 1/255 skip; T < 1e-4 stops the pixel
                    const bool contrib0 = !done0 && !(power0 > 0.0f) && !(alpha0 < 1.0f / 255.0f);
                    const bool contrib1 = !done1 && !(power1 > 0.0f) && !(alpha1 < 1.0f / 255.0f);
                    // the extent test is a bounding box: a fifth of the hits reach no pixel of the sub-tile at all
                    if (!__any_sync(FULL, contrib0 || contrib1)) continue;
                    BSTAT(2, 1);                                  // ... that reach at least one pixel
                    BSTAT(3, __popc(__ballot_sync(FULL, contrib0)) + __popc(__ballot_sync(FULL, contrib1)));   // blended (Gaussian, pixel) pairs
                    const float4 g2 = s_rec[buf][2][slot];
                    const float test_T0 = __fmul_rn(T0, __fsub_rn(1.f, alpha0));
                    const float test_T1 = __fmul_rn(T1, __fsub_rn(1.f, alpha1));
                    const bool stop0 = contrib0 && (test_T0 < 0.0001f), stop1 = contrib1 && (test_T1 < 0.0001f);
                    done0 = done0 || stop0;
                    done1 = done1 || stop1;
                    if (contrib0 && !stop0) {
                        C00 = __fmaf_rn(T0, __fmul_rn(alpha0, g2.x), C00);
                        C01 = __fmaf_rn(T0, __fmul_rn(alpha0, g2.y), C01);
                        C02 = __fmaf_rn(T0, __fmul_rn(alpha0, g2.z), C02);
                        T0 = test_T0;
                        last0 = (uint32_t)(base + slot + 1);
                    }
                    if (contrib1 && !stop1) {
                        C10 = __fmaf_rn(T1, __fmul_rn(alpha1, g2.x), C10);
                        C11 = __fmaf_rn(T1, __fmul_rn(alpha1, g2.y), C11);
                        C12 = __fmaf_rn(T1, __fmul_rn(alpha1, g2.z), C12);
                        T1 = test_T1;
                        last1 = (uint32_t)(base + slot + 1);
                    }
                }
                if (__all_sync(FULL, done0 && done1)) break;
            }
        }
        __syncthreads();   // everyone finished reading s_rec[buf] before it is refilled
    }
    cp_async_wait<0>();

    const size_t plane = size_t(H) * W;
    const float bg0 = __ldg(bg_color + 0), bg1 = __ldg(bg_color + 1), bg2 = __ldg(bg_color + 2);
    if (inside0) {
        const size_t pix = size_t(st.py0) * W + st.px;
        final_T[pix] = T0;
        n_contrib[pix] = last0;
        out_color[pix] = __fmaf_rn(T0, bg0, C00);
        out_color[plane + pix] = __fmaf_rn(T0, bg1, C01);
        out_color[2 * plane + pix] = __fmaf_rn(T0, bg2, C02);
    }
    if (inside1) {
        const size_t pix = size_t(st.py1) * W + st.px;
        final_T[pix] = T1;
        n_contrib[pix] = last1;
        out_color[pix] = __fmaf_rn(T1, bg0, C10);
        out_color[plane + pix] = __fmaf_rn(T1, bg1, C11);
        out_color[2 * plane + pix] = __fmaf_rn(T1, bg2, C12);
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

// per-pixel state of the back-to-front replay
struct PixB {
    float T, Tf;              // running transmittance, final transmittance
    float S0, S1, S2;         // colour accumulated BEHIND the current Gaussian (accum_rec of backward.cu:507-511,
                              // advanced after each contributor: S <- alpha c + (1 - alpha) S)
    float d0, d1, d2;         // dL_dpixel
    float bgdot;
    int last;                 // n_contrib
};

__global__ void __launch_bounds__(BLEND_THREADS)
blend_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                      int W, int H, int grid_x, const float4* __restrict__ rec,
                      const float* __restrict__ bg_color, const float* __restrict__ final_T,
                      const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                      float4* __restrict__ acc)
{
    __shared__ float4 s_rec[2][3][BATCH];
    __shared__ uint32_t s_id[2][BATCH];
    __shared__ uint8_t s_mask[BATCH];
    __shared__ int s_bmax;

    const int tile = blockIdx.x;
    const SubTile st = make_subtile(tile, grid_x);
    const float pixfx = (float)st.px, pixfy0 = (float)st.py0, pixfy1 = (float)st.py1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float tile_x0 = (float)((tile % grid_x) * TILE_X), tile_y0 = (float)((tile / grid_x) * TILE_Y);

    const uint2 range = ranges[tile];
    const uint32_t* list = point_list + range.x;
    const size_t plane = size_t(H) * W;
    const float bg0 = __ldg(bg_color), bg1 = __ldg(bg_color + 1), bg2 = __ldg(bg_color + 2);

    auto load_pixel = [&](int py) {
        PixB p;
        p.T = p.Tf = 0.f; p.S0 = p.S1 = p.S2 = 0.f; p.d0 = p.d1 = p.d2 = 0.f; p.last = 0;
        if (st.px < W && py < H) {
            const size_t pix = size_t(py) * W + st.px;
            p.T = p.Tf = final_T[pix];
            p.last = (int)n_contrib[pix];
            p.d0 = dL_dpix[pix]; p.d1 = dL_dpix[plane + pix]; p.d2 = dL_dpix[2 * plane + pix];
        }
        p.bgdot = bg0 * p.d0 + bg1 * p.d1 + bg2 * p.d2;
        return p;
    };
    PixB p0 = load_pixel(st.py0), p1 = load_pixel(st.py1);

    // the nine sums leave the warp scaled by these per-lane constants (lane -> value index, see transpose_reduce8):
    //   0 dL_dmean2D.x * 0.5 W   1 dL_dmean2D.y * 0.5 H   2,3,4 dL_dconic * -0.5   5 dL_dopacity   6,7,(8) dL_dcolor
    const int kval = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const float kscale = kval == 0 ? (float)(0.5 * W) : kval == 1 ? (float)(0.5 * H) : (kval <= 4 ? -0.5f : 1.f);

    // last list position any pixel of the warp / CTA needs
    const int wmax = __reduce_max_sync(FULL, max(p0.last, p1.last));
    if (tid == 0) s_bmax = 0;
    __syncthreads();
    if (lane == 0 && wmax > 0) atomicMax(&s_bmax, wmax);
    __syncthreads();
    const int bmax = s_bmax;
    if (bmax == 0) return;
    const int nbatches = (bmax + BATCH - 1) / BATCH;

    // batch b (b = 0 is the BACK of the list) covers positions [lo_b, hi_b), slot s <-> lo_b + s
    auto batch_lo = [&](int b) { return max(0, bmax - (b + 1) * BATCH); };
    auto batch_hi = [&](int b) { return bmax - b * BATCH; };

    uint32_t ids[PER_THREAD];
    load_ids(ids, list, batch_lo(0), batch_hi(0) - batch_lo(0));
    stage_records(s_rec[0], s_id[0], rec, ids, batch_hi(0) - batch_lo(0));
    cp_async_commit();
    if (nbatches > 1) load_ids(ids, list, batch_lo(1), batch_hi(1) - batch_lo(1));

    for (int b = 0; b < nbatches; ++b) {
        const int buf = b & 1;
        const int lo = batch_lo(b);
        const int n_in = batch_hi(b) - lo;
        if (b + 1 < nbatches) {
            stage_records(s_rec[buf ^ 1], s_id[buf ^ 1], rec, ids, batch_hi(b + 1) - batch_lo(b + 1));
            cp_async_commit();
            if (b + 2 < nbatches) load_ids(ids, list, batch_lo(b + 2), batch_hi(b + 2) - batch_lo(b + 2));
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        compute_masks(s_mask, s_rec[buf][0], n_in, tile_x0, tile_y0);
        __syncthreads();
        if (threadIdx.x == 0) BSTAT(4, n_in);

        if (lo < wmax) {
            for (int c = (n_in - 1) >> 5; c >= 0; --c) {
                const int slot_l = c * 32 + lane;
                const bool hit = slot_l < n_in && lo + slot_l < wmax && ((s_mask[slot_l] >> warp) & 1u);
                unsigned mask = __ballot_sync(FULL, hit);
                while (mask) {
                    const int j = 31 - __clz(mask);
                    mask &= ~(1u << j);
                    const int slot = c * 32 + j;
                    const int pos = lo + slot;
                    BSTAT(5, 1);
                    const float4 g0 = s_rec[buf][0][slot];
                    const float4 g1 = s_rec[buf][1][slot];
                    const float dx = __fsub_rn(g0.x, pixfx);
                    const float dy0 = __fsub_rn(g0.y, pixfy0), dy1 = __fsub_rn(g0.y, pixfy1);
                    const float power0 = gauss_power(dx, dy0, g1.x, g1.y, g1.z);
                    const float power1 = gauss_power(dx, dy1, g1.x, g1.y, g1.z);
                    const float Gr0 = expf(power0), Gr1 = expf(power1);
                    const float ar0 = fminf(0.99f, __fmul_rn(g1.w, Gr0)), ar1 = fminf(0.99f, __fmul_rn(g1.w, Gr1));
                    // backward.cu:486-501: behind the last contributor / power > 0 / alpha < 1/255
                    const bool act0 = (pos < p0.last) && !(power0 > 0.0f) && !(ar0 < 1.0f / 255.0f);
                    const bool act1 = (pos < p1.last) && !(power1 > 0.0f) && !(ar1 < 1.0f / 255.0f);
                    if (!__any_sync(FULL, act0 || act1)) continue;
                    BSTAT(6, 1);
                    BSTAT(7, __popc(__ballot_sync(FULL, act0)) + __popc(__ballot_sync(FULL, act1)));

                    // Straight-line maths: a pixel that does not contribute runs with G = alpha = 0, which makes
                    // every gradient term exactly 0 and leaves T and S unchanged (1/(1-0) = 1, 0 c + 1 S = S).
                    const float4 g2 = s_rec[buf][2][slot];
                    float q[2], qy[2], gda[2], dcol[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        PixB& p = k ? p1 : p0;
                        const float G = (k ? act1 : act0) ? (k ? Gr1 : Gr0) : 0.f;
                        const float alpha = (k ? act1 : act0) ? (k ? ar1 : ar0) : 0.f;
                        const float dy = k ? dy1 : dy0;
                        const float oma = 1.f - alpha;
                        const float rcp_oma = rcp_approx(oma);               // 1 - alpha in [0.01, 1]: one MUFU.RCP
                        p.T = p.T * rcp_oma;                                 // T / (1 - alpha)
                        dcol[k] = alpha * p.T;                               // dchannel_dcolor
                        float dL_dalpha = (g2.x - p.S0) * p.d0;
                        dL_dalpha = fmaf(g2.y - p.S1, p.d1, dL_dalpha);
                        dL_dalpha = fmaf(g2.z - p.S2, p.d2, dL_dalpha);
                        dL_dalpha *= p.T;
                        dL_dalpha = fmaf(-p.Tf * rcp_oma, p.bgdot, dL_dalpha);
                        p.S0 = fmaf(alpha, g2.x, oma * p.S0);
                        p.S1 = fmaf(alpha, g2.y, oma * p.S1);
                        p.S2 = fmaf(alpha, g2.z, oma * p.S2);
                        gda[k] = G * dL_dalpha;                              // -> dL_dopacity
                        q[k] = gda[k] * g1.w;                                // G * dL_dG
                        qy[k] = q[k] * dy;
                    }
                    // sums over the lane's two pixels (they share dx), common factors pulled out:
                    //   dL_dG dG_ddelx = -(q dx cx + q dy cy), dL_dconic.x ~ q dx dx, .y ~ q dx dy, .w ~ q dy dy
                    const float sq = q[0] + q[1];
                    const float sqy = qy[0] + qy[1];
                    const float A = dx * sq;
                    float v[8];
                    v[0] = -fmaf(g1.x, A, g1.y * sqy);                       // dL_dmean2D.x / (0.5 W)
                    v[1] = -fmaf(g1.z, sqy, g1.y * A);                       // dL_dmean2D.y / (0.5 H)
                    v[2] = dx * A;                                           // dL_dconic.x / -0.5
                    v[3] = dx * sqy;                                         // dL_dconic.y / -0.5
                    v[4] = fmaf(qy[1], dy1, qy[0] * dy0);                    // dL_dconic.w / -0.5
                    v[5] = gda[0] + gda[1];                                  // dL_dopacity
                    v[6] = fmaf(dcol[1], p1.d0, dcol[0] * p0.d0);            // dL_dcolor.r
                    v[7] = fmaf(dcol[1], p1.d1, dcol[0] * p0.d1);            // dL_dcolor.g
                    float v8 = fmaf(dcol[1], p1.d2, dcol[0] * p0.d2);        // dL_dcolor.b
                    const float tot = transpose_reduce8(v, lane) * kscale;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(FULL, v8, o);
                    // one reduction per (Gaussian, warp sub-tile) and value, straight into the packed
                    // 48-byte accumulator: 8 lanes hit 8 consecutive floats, lane 1 the ninth
                    float* dst = reinterpret_cast<float*>(acc + 3 * size_t(s_id[buf][slot]));
                    if ((lane & 3) == 0) atomicAdd(dst + kval, tot);
                    else if (lane == 1) atomicAdd(dst + 8, v8);
                }
            }
        }
        __syncthreads();   // all warps finished with s_rec[buf] / s_id[buf] before they are refilled
    }
}

}  // namespace

int debug_blend_stats(unsigned long long* out, bool reset)
{
#ifdef SEGS_BLEND_STATS
    SEGS_CUDA_CHECK(cudaDeviceSynchronize());
    SEGS_CUDA_CHECK(cudaMemcpyFromSymbol(out, g_blend_stats, sizeof(g_blend_stats)));
    if (reset) { unsigned long long z[8] = {0}; SEGS_CUDA_CHECK(cudaMemcpyToSymbol(g_blend_stats, z, sizeof(z))); }
    return SEGS_OK;
#else
    (void)reset;
    for (int i = 0; i < 8; ++i) out[i] = 0;
    return SEGS_OK;
#endif
}

int launch_blend_forward(const ViewParams& vp, const GeomState& g, const BinningState& b,
                         ImageState& img, const float* background, float* out_color,
                         cudaStream_t stream)
{
    const int T = vp.grid_x * vp.grid_y;
    const uint32_t* point_list = b.point_list;
    blend_forward_kernel<<<T, BLEND_THREADS, 0, stream>>>(img.ranges, point_list, vp.W, vp.H, vp.grid_x, g.rec,
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
    blend_backward_kernel<<<T, BLEND_THREADS, 0, stream>>>(img.ranges, point_list, vp.W, vp.H, vp.grid_x, g.rec,
                                                     background, img.final_T, img.n_contrib, dL_dpix, g.acc);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // namespace segs
