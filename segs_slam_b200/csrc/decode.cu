// Structure-enhanced anchor decode for sm_100a.
//
// Replaces GaussianRenderer::generate_neural_gaussians of the reference
// (src/gaussian_renderer.cpp:214-334; module shapes src/gaussian_model.cpp:60-98): boolean-mask
// gathers of the anchor tensors, the feature-bank mix, the pose-appearance Linear, three small
// MLPs, the opacity > 0 mask, repeat/cat into a [A*10, 22] temporary and a second boolean gather —
// about 40 ATen kernels and ~1 GB of temporaries at 200k anchors.
//
// Two kernel sets live behind segs_decode_forward / segs_decode_backward (segs_decode_set_variant):
//
// Variant 2 (default):
//   forward   decode_compact_kernel lists the visible anchors; decode_forward_v2_kernel: persistent CTAs of 512 threads,
//             tile = 128 visible anchors, BOTH layers of the three MLPs on tcgen05 (3xTF32, 51 MMAs per tile), rows
//             assembled by thread = output row in the reference's (anchor, offset) order; the layer activations stay in
//             the state buffer for the backward;
//   backward  decode_backward_kernel<true> (thread = visible anchor) back-propagates the row gradients to d_anchor /
//             d_offset / d_anchor_feat / d_scaling from those activations and stores the per-anchor factors of the weight
//             gradients feature-major in scratch; decode_wgrad_tc_kernel turns them into dW on tcgen05 (K = anchor
//             index, accumulators in TMEM over all stages of a persistent CTA).
// Variant 1 (round 1):
//   forward   decode_forward_kernel: thread = visible anchor, CTA = 128 consecutive anchors, first layers on tcgen05,
//             second layers as FP32 FFMA chains out of shared memory (weights padded to float4: one broadcast LDS.128
//             feeds four FFMAs); both compactions (visible anchors, surviving offsets) in-kernel with CTA-local ballot
//             scans plus a decoupled look-back;
//   backward  decode_backward_kernel<false> recomputes the forward; decode_wgrad_kernel: persistent CTAs,
//             dW[p][q] = sum_k U[k][p] V[k][q] with q on the 32 lanes, FP32 FFMA, flushed once with atomics.
// In both, the two totals (visible anchors, emitted rows) are stored to mapped host memory by the last tile (the host needs
// them to shape the outputs, as the reference's boolean indexing does).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include "common.cuh"

namespace segs {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int FEAT = 32;        // Model.feat_dim
constexpr int NOFF = 10;        // Model.n_offsets
constexpr int XDIM = 36;        // feat(32) + view(3) + dist(1)
constexpr int W2S = 36;         // shared-memory row stride of the second-layer weights
constexpr int DEC_THREADS = 128;
constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;

struct Pose7 { float v[7]; };

// shared-memory copy of the weights.  SWF: everything except the first-layer matrices (the forward
// feeds those to the tensor cores from its own operand tiles); SW adds them for the backward.
struct SWF {
    float b1[3][FEAT];           // 0 opacity, 1 cov, 2 colour (colour: b1 + W1[:, appearance columns] * appearance)
    float w2o[NOFF][W2S];        // second layers: rows padded to W2S floats (conflict-free when lanes
    float w2s[7 * NOFF][W2S];    // of a warp read the rows of different offsets)
    float w2c[3 * NOFF][W2S];
    float b2o[12];
    float b2s[72];
    float b2c[32];
    float wb1[FEAT][4];
    float bb1[FEAT];
    float wb2[3][FEAT];
    float bb2[4];
    float app[32];               // appearance vector of this view
};
struct SW : SWF {
    float w1[3][FEAT][XDIM];     // first 35(+dist) input columns, zero padded
};

constexpr int O2_W = NOFF + 7 * NOFF + 3 * NOFF;   // second-layer outputs per anchor: opacity(10) | scale_rot(70) | colour(30)

// opaque state shared by forward and backward
struct DecodeState {
    uint32_t* counters;       // [8]: 0 ticket
    uint32_t* look_vis;       // [tiles]
    uint32_t* look_row;       // [tiles]
    uint32_t* anchor_index;   // [A] anchor id of every visible ordinal
    uint32_t* row_start;      // [A] first output row of every visible ordinal
    uint32_t* mask_bits;      // [A] surviving-offset bits of every visible ordinal
    float* hidden;            // [A][96]  relu(layer 1) of the three MLPs per visible ordinal   } written by forward variant 2
    float* out2;              // [A][110] tanh(opacity) | scale_rot | colour pre-activations     } (counters[3] = 1), read by backward
    size_t zero_bytes;        // counters + look words
    static DecodeState carve(char* base, size_t A, size_t* bytes) {
        Carver c(base);
        DecodeState s;
        const size_t tiles = (A + DEC_THREADS - 1) / DEC_THREADS;
        s.counters = c.take<uint32_t>(8 + 2 * tiles);
        s.look_vis = s.counters + 8;
        s.look_row = s.look_vis + tiles;
        s.zero_bytes = (8 + 2 * tiles) * sizeof(uint32_t);
        s.anchor_index = c.take<uint32_t>(A);
        s.row_start = c.take<uint32_t>(A);
        s.mask_bits = c.take<uint32_t>(A);
        s.hidden = c.take<float>(A * size_t(3 * FEAT));
        s.out2 = c.take<float>(A * size_t(O2_W));
        if (bytes) *bytes = c.used(base) + 128;
        return s;
    }
};

template <class SWT>
__device__ __forceinline__ void stage_weights(SWT& s, const segs_decode_params& p, const Pose7& pose)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int in_o = 35 + (p.add_opacity_dist ? 1 : 0), in_s = 35 + (p.add_cov_dist ? 1 : 0);
    const int in_c = 35 + (p.add_color_dist ? 1 : 0);
    const int ld_c = in_c + p.appearance_dim;
    if constexpr (sizeof(SWT) == sizeof(SW)) {
        for (int e = tid; e < FEAT * XDIM; e += nt) {
            const int j = e / XDIM, i = e % XDIM;
            s.w1[0][j][i] = i < in_o ? __ldg(p.opacity_w1 + j * in_o + i) : 0.f;
            s.w1[1][j][i] = i < in_s ? __ldg(p.cov_w1 + j * in_s + i) : 0.f;
            s.w1[2][j][i] = i < in_c ? __ldg(p.color_w1 + j * ld_c + i) : 0.f;
        }
    } else {
        (void)in_o; (void)in_s;
    }
    for (int e = tid; e < NOFF * FEAT; e += nt) s.w2o[e / FEAT][e % FEAT] = __ldg(p.opacity_w2 + e);
    for (int e = tid; e < 7 * NOFF * FEAT; e += nt) s.w2s[e / FEAT][e % FEAT] = __ldg(p.cov_w2 + e);
    for (int e = tid; e < 3 * NOFF * FEAT; e += nt) s.w2c[e / FEAT][e % FEAT] = __ldg(p.color_w2 + e);
    for (int e = tid; e < 72; e += nt) {
        if (e < 12) s.b2o[e] = e < NOFF ? __ldg(p.opacity_b2 + e) : 0.f;
        s.b2s[e] = e < 7 * NOFF ? __ldg(p.cov_b2 + e) : 0.f;
        if (e < 32) s.b2c[e] = e < 3 * NOFF ? __ldg(p.color_b2 + e) : 0.f;
    }
    if (tid < FEAT) {
        s.b1[0][tid] = __ldg(p.opacity_b1 + tid);
        s.b1[1][tid] = __ldg(p.cov_b1 + tid);
        // appearance = Linear(7 -> app)(pose)   (gaussian_renderer.cpp:256-270)
        float a = 0.f;
        if (tid < p.appearance_dim) {
            a = __ldg(p.app_b + tid);
#pragma unroll
            for (int q = 0; q < 7; ++q) a = fmaf(__ldg(p.app_w + tid * 7 + q), pose.v[q], a);
        }
        s.app[tid] = a;
        if (p.use_feat_bank) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s.wb1[tid][i] = __ldg(p.bank_w1 + tid * 4 + i);
            s.bb1[tid] = __ldg(p.bank_b1 + tid);
#pragma unroll
            for (int m = 0; m < 3; ++m) s.wb2[m][tid] = __ldg(p.bank_w2 + m * FEAT + tid);
            if (tid < 4) s.bb2[tid] = tid < 3 ? __ldg(p.bank_b2 + tid) : 0.f;
        }
    }
    __syncthreads();
    if (tid < FEAT) {
        // the appearance input is the same for every anchor of the view: fold its columns into the bias
        float b = __ldg(p.color_b1 + tid);
        for (int k = 0; k < p.appearance_dim; ++k) b = fmaf(__ldg(p.color_w1 + tid * ld_c + in_c + k), s.app[k], b);
        s.b1[2][tid] = b;
    }
    __syncthreads();
}

// h = relu(W1 x + b1)
__device__ __forceinline__ void layer1(const float (*w)[XDIM], const float* b, const float (&x)[XDIM], float (&h)[FEAT])
{
#pragma unroll
    for (int j = 0; j < FEAT; ++j) {
        float acc = b[j];
        const float4* row = reinterpret_cast<const float4*>(w[j]);
#pragma unroll
        for (int q = 0; q < XDIM / 4; ++q) {
            const float4 w4 = row[q];
            acc = fmaf(w4.x, x[4 * q], acc);
            acc = fmaf(w4.y, x[4 * q + 1], acc);
            acc = fmaf(w4.z, x[4 * q + 2], acc);
            acc = fmaf(w4.w, x[4 * q + 3], acc);
        }
        h[j] = fmaxf(acc, 0.f);
    }
}

__device__ __forceinline__ float dot32(const float* wrow, float bias, const float (&h)[FEAT])
{
    float acc = bias;
    const float4* row = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int q = 0; q < FEAT / 4; ++q) {
        const float4 w4 = row[q];
        acc = fmaf(w4.x, h[4 * q], acc);
        acc = fmaf(w4.y, h[4 * q + 1], acc);
        acc = fmaf(w4.z, h[4 * q + 2], acc);
        acc = fmaf(w4.w, h[4 * q + 3], acc);
    }
    return acc;
}

// v[j] += s * wrow[j]
__device__ __forceinline__ void axpy32(const float* wrow, float s, float (&v)[FEAT])
{
    const float4* row = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int q = 0; q < FEAT / 4; ++q) {
        const float4 w4 = row[q];
        v[4 * q] = fmaf(w4.x, s, v[4 * q]);
        v[4 * q + 1] = fmaf(w4.y, s, v[4 * q + 1]);
        v[4 * q + 2] = fmaf(w4.z, s, v[4 * q + 2]);
        v[4 * q + 3] = fmaf(w4.w, s, v[4 * q + 3]);
    }
}

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

struct AnchorIn {
    float ax, ay, az;      // anchor
    float vx, vy, vz;      // anchor - camera centre
    float dist;
    float s[6];            // exp(_scaling)
};

// D1 + D2: inputs of the three MLPs for one anchor.  x = [feat'(32), ob_view(3), ob_dist]
__device__ __forceinline__ void build_input(const SWF& sw, bool use_bank, const float* __restrict__ feat_row,
                                            const AnchorIn& in, float (&x)[XDIM], float (&bankw)[3])
{
    const float ux = in.vx / in.dist, uy = in.vy / in.dist, uz = in.vz / in.dist;
    float f[FEAT];
#pragma unroll
    for (int q = 0; q < FEAT / 4; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(feat_row) + q);
        f[4 * q] = t.x; f[4 * q + 1] = t.y; f[4 * q + 2] = t.z; f[4 * q + 3] = t.w;
    }
    bankw[0] = 0.f; bankw[1] = 0.f; bankw[2] = 1.f;
    if (use_bank) {
        // bank weights = softmax(W2 relu(W1 [view, dist] + b1) + b2)   (gaussian_renderer.cpp:236-239)
        float l0 = sw.bb2[0], l1 = sw.bb2[1], l2 = sw.bb2[2];
#pragma unroll
        for (int j = 0; j < FEAT; ++j) {
            const float4 w4 = *reinterpret_cast<const float4*>(sw.wb1[j]);
            float hb = sw.bb1[j];
            hb = fmaf(w4.x, ux, hb); hb = fmaf(w4.y, uy, hb); hb = fmaf(w4.z, uz, hb); hb = fmaf(w4.w, in.dist, hb);
            hb = fmaxf(hb, 0.f);
            l0 = fmaf(sw.wb2[0][j], hb, l0); l1 = fmaf(sw.wb2[1][j], hb, l1); l2 = fmaf(sw.wb2[2][j], hb, l2);
        }
        const float mx = fmaxf(l0, fmaxf(l1, l2));
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx);
        const float den = e0 + e1 + e2;
        bankw[0] = e0 / den; bankw[1] = e1 / den; bankw[2] = e2 / den;
        // feat'[j] = feat[4 (j mod 8)] w0 + feat[2 (j mod 16)] w1 + feat[j] w2   (:241-248; repeat = tiling)
#pragma unroll
        for (int j = 0; j < FEAT; ++j)
            x[j] = f[4 * (j & 7)] * bankw[0] + f[2 * (j & 15)] * bankw[1] + f[j] * bankw[2];
    } else {
#pragma unroll
        for (int j = 0; j < FEAT; ++j) x[j] = f[j];
    }
    x[32] = ux; x[33] = uy; x[34] = uz; x[35] = in.dist;
}

__device__ __forceinline__ AnchorIn load_anchor(const float* __restrict__ anchor, const float* __restrict__ scaling,
                                                const float* __restrict__ cam, size_t a)
{
    AnchorIn in;
    in.ax = __ldg(anchor + 3 * a); in.ay = __ldg(anchor + 3 * a + 1); in.az = __ldg(anchor + 3 * a + 2);
    in.vx = in.ax - __ldg(cam); in.vy = in.ay - __ldg(cam + 1); in.vz = in.az - __ldg(cam + 2);
    in.dist = sqrtf(in.vx * in.vx + in.vy * in.vy + in.vz * in.vz);
#pragma unroll
    for (int k = 0; k < 6; ++k) in.s[k] = __ldg(scaling + 6 * a + k);
    return in;
}

// decoupled look-back of one running total (executed by warp 0); returns the exclusive prefix
__device__ __forceinline__ uint32_t lookback(volatile uint32_t* look, uint32_t tile, uint32_t total, int lane)
{
    if (lane == 0) look[tile] = (tile == 0 ? FLAG_PREFIX : FLAG_AGG) | total;
    uint32_t excl = 0;
    if (tile != 0) {
        for (int t = (int)tile - 1;; t -= 32) {
            uint32_t w = FLAG_PREFIX;
            if (t - lane >= 0) {
                do { w = look[t - lane]; } while ((w & FLAG_MASK) == 0u);
            }
            const unsigned pref = __ballot_sync(FULL, (w & FLAG_MASK) == FLAG_PREFIX);
            const int stop = __ffs(pref) - 1;
            const uint32_t v = (stop < 0 || lane <= stop) ? (w & ~FLAG_MASK) : 0u;
            excl += __reduce_add_sync(FULL, v);
            if (stop >= 0) break;
        }
        if (lane == 0) look[tile] = FLAG_PREFIX | (excl + total);
    }
    return excl;
}

// CTA-wide exclusive scan of one value per thread (DEC_THREADS threads); *total = CTA sum
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();                 // s_warp may still be read from a previous scan
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < DEC_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < warp) woff += c;
        tot += c;
    }
    *total = tot;
    return woff + inc - v;
}

// =======================================================================================
// forward
// =======================================================================================
// The first layers of the three MLPs are ONE dense contraction per tile of 128 anchors,
//   H[128 x 96] = X[128 x 40] * W1cat^T[40 x 96]      (X = [feat'(32), view(3), dist, 0-pad]),
// and run on the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 128, N = 96, K = 8 per
// instruction, both operands K-major in shared memory (no-swizzle core-matrix layout), the
// accumulator in TMEM.  A single TF32 product (10-bit mantissa) is not enough — the opacity > 0
// mask decides which Gaussians exist, and parity is 1e-5 — so every operand is split into
// hi = tf32(v) and lo = v - hi and three MMAs (hi*hi + hi*lo + lo*hi) accumulate per K-step
// (3xTF32, error ~2^-21, the level of FP32 summation-order noise).  The epilogue reads each
// anchor's 96 pre-activations back with tcgen05.ld (thread = TMEM lane = anchor).
namespace tc {

constexpr int TM = 128;                  // anchors per tile = MMA M = TMEM lanes
constexpr int TN = 3 * FEAT;             // 96 hidden units of the three MLPs
constexpr int TK = 40;                   // XDIM padded to a multiple of 8
constexpr int KB = TK / 4;               // 16-byte K-blocks per row
constexpr int CORE = 128;                // bytes of one 8 x 16 B core matrix
constexpr uint32_t SBO = KB * CORE;      // byte stride between 8-row groups
constexpr uint32_t LBO = CORE;           // byte stride between K-blocks
constexpr int A_BYTES = TM * TK * 4;     // 20480
constexpr int B_BYTES = TN * TK * 4;     // 15360
constexpr int TMEM_COLS = 128;           // power of two >= TN
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(TN >> 3) << 17) | (uint32_t(TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) in the K-major no-swizzle canonical layout
__device__ __forceinline__ uint32_t canon_off(int row, int k) {
    return uint32_t(((row >> 3) * KB + (k >> 2)) * CORE + (row & 7) * 16 + (k & 3) * 4);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), no swizzle, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(LBO >> 4) << 16) | (uint64_t(SBO >> 4) << 32) | (uint64_t(1) << 46);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_%=;\n"
        "bra WAIT_LOOP_%=;\n"
        "WAIT_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate) : "memory");
}
// arrive on the mbarrier when every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive columns of this thread's TMEM lane (warp-collective)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// round-to-nearest TF32 (the tensor core ignores the 13 low mantissa bits of its operands, i.e. truncates: rounding
// both parts of a hi / lo split here halves the representation error of the split and removes its bias)
// (cvt.rna.tf32.f32 does exactly this, but on the conversion pipe at a fraction of the ALU rate: with 56 - 64 conversions
// per thread and tile it showed up as `math` throttle stalls; add-half-and-mask is two full-rate integer instructions.
// Round to nearest, ties away from zero, as .rna.)
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
struct Split4 { float4 hi, lo; };
__device__ __forceinline__ Split4 split4(float v0, float v1, float v2, float v3) {
    Split4 s;
    s.hi = make_float4(tf32_rn(v0), tf32_rn(v1), tf32_rn(v2), tf32_rn(v3));
    s.lo = make_float4(tf32_rn(v0 - s.hi.x), tf32_rn(v1 - s.hi.y), tf32_rn(v2 - s.hi.z), tf32_rn(v3 - s.hi.w));
    return s;
}

}  // namespace tc

constexpr int H_W = 2 * FEAT + 4;      // floats per anchor in s_h (68: conflict-free float4 reads across anchors)
constexpr int REC_W = 20;
constexpr size_t align128(size_t x) { return (x + 127) & ~size_t(127); }
constexpr size_t OFF_B_HI = align128(sizeof(SWF));
constexpr size_t OFF_B_LO = OFF_B_HI + tc::B_BYTES;
constexpr size_t OFF_A_HI = OFF_B_LO + tc::B_BYTES;          // A operand tiles; re-used as s_h once the MMAs are done
constexpr size_t OFF_A_LO = OFF_A_HI + tc::A_BYTES;
constexpr size_t OFF_REC = OFF_A_HI + 2 * tc::A_BYTES;
constexpr size_t OFF_ROW = OFF_REC + sizeof(float) * DEC_THREADS * REC_W;
constexpr size_t FWD_SMEM = OFF_ROW + sizeof(uint32_t) * (DEC_THREADS + 4);
static_assert(sizeof(float) * DEC_THREADS * H_W <= 2 * tc::A_BYTES, "s_h must fit in the A operand tiles");
static_assert(DEC_THREADS == tc::TM, "one thread per TMEM lane");

__global__ void __launch_bounds__(DEC_THREADS, 2)
decode_forward_kernel(int A, const unsigned char* __restrict__ visible_mask, const float* __restrict__ anchor,
                      const float* __restrict__ anchor_feat, const float* __restrict__ offset,
                      const float* __restrict__ scaling, const float* __restrict__ cam, const Pose7 pose,
                      const segs_decode_params p, float* __restrict__ out_xyz, float* __restrict__ out_color,
                      float* __restrict__ out_opacity, float* __restrict__ out_scaling, float* __restrict__ out_rot,
                      float* __restrict__ neural_opacity, unsigned char* __restrict__ out_mask, DecodeState st,
                      volatile uint32_t* __restrict__ host_counts)
{
    extern __shared__ __align__(1024) unsigned char s_dec[];
    SWF& sw = *reinterpret_cast<SWF*>(s_dec);
    unsigned char* sB_hi = s_dec + OFF_B_HI;
    unsigned char* sB_lo = s_dec + OFF_B_LO;
    unsigned char* sA_hi = s_dec + OFF_A_HI;
    unsigned char* sA_lo = s_dec + OFF_A_LO;
    float* s_h = reinterpret_cast<float*>(s_dec + OFF_A_HI);               // [128][H_W]: h_cov | h_colour per anchor
    float* s_rec = reinterpret_cast<float*>(s_dec + OFF_REC);              // [128][REC_W]: anchor, scaling, opacities, mask
    uint32_t* s_row = reinterpret_cast<uint32_t*>(s_dec + OFF_ROW);        // [129] first row of every anchor
    __shared__ uint32_t s_aid[DEC_THREADS];
    __shared__ uint32_t s_warp[DEC_THREADS / 32];
    __shared__ uint32_t s_tile, s_vis_base, s_row_base, s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ntiles = (uint32_t)((A + DEC_THREADS - 1) / DEC_THREADS);

    // ---- once per (persistent) CTA: TMEM, mbarrier, weights ----
    if (warp == 0) tc::tmem_alloc(&s_tmem, tc::TMEM_COLS);
    if (tid == 0) tc::mbar_init(&s_bar, 1);
    {
        // B operand: W1cat[n][k], n = 32 * mlp + hidden unit, k = input column; hi/lo TF32 split
        const int in_o = 35 + (p.add_opacity_dist ? 1 : 0), in_s = 35 + (p.add_cov_dist ? 1 : 0);
        const int in_c = 35 + (p.add_color_dist ? 1 : 0), ld_c = in_c + p.appearance_dim;
        for (int e = tid; e < tc::TN * tc::TK; e += DEC_THREADS) {
            const int n = e / tc::TK, k = e % tc::TK, j = n & 31;
            float w = 0.f;
            if (n < 32) { if (k < in_o) w = __ldg(p.opacity_w1 + j * in_o + k); }
            else if (n < 64) { if (k < in_s) w = __ldg(p.cov_w1 + j * in_s + k); }
            else { if (k < in_c) w = __ldg(p.color_w1 + j * ld_c + k); }
            const float hi = tc::tf32_hi(w);
            const uint32_t off = tc::canon_off(n, k);
            *reinterpret_cast<float*>(sB_hi + off) = hi;
            *reinterpret_cast<float*>(sB_lo + off) = w - hi;
        }
    }
    stage_weights(sw, p, pose);          // contains __syncthreads
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    uint32_t phase = 0;

    for (;;) {
        __syncthreads();                 // previous tile is done with s_tile / s_aid / the bases / s_h
        if (tid == 0) s_tile = atomicAdd(st.counters, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) break;

        // ---- D1: compact the visible anchors of this tile (ascending anchor index) ----
        const size_t a0 = size_t(tile) * DEC_THREADS + tid;
        const bool vis = a0 < (size_t)A && (visible_mask == nullptr || visible_mask[a0] != 0);
        uint32_t n_vis;
        const uint32_t ord = cta_exclusive_scan(vis ? 1u : 0u, s_warp, &n_vis);
        if (vis) s_aid[ord] = (uint32_t)a0;
        __syncthreads();
        const bool active = (uint32_t)tid < n_vis;
        const size_t a = active ? s_aid[tid] : 0;

        // ---- D2/D3: MLP input row of this anchor -> A operand tiles (hi / lo) ----
        AnchorIn in;
        if (active) {
            in = load_anchor(anchor, scaling, cam, a);
            float x[XDIM], bankw[3];
            build_input(sw, p.use_feat_bank != 0, anchor_feat + a * FEAT, in, x, bankw);
#pragma unroll
            for (int kb = 0; kb < tc::KB; ++kb) {
                float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
                if (kb < XDIM / 4) {
                    hi = make_float4(tc::tf32_hi(x[4 * kb]), tc::tf32_hi(x[4 * kb + 1]), tc::tf32_hi(x[4 * kb + 2]), tc::tf32_hi(x[4 * kb + 3]));
                    lo = make_float4(x[4 * kb] - hi.x, x[4 * kb + 1] - hi.y, x[4 * kb + 2] - hi.z, x[4 * kb + 3] - hi.w);
                }
                const uint32_t off = tc::canon_off(tid, 4 * kb);
                *reinterpret_cast<float4*>(sA_hi + off) = hi;
                *reinterpret_cast<float4*>(sA_lo + off) = lo;
            }
        }
        tc::fence_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
        tc::fence_before_sync();
        __syncthreads();

        // ---- D4 first layers on the tensor cores: 5 K-steps x 3 split products ----
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t a_hi = tc::smem_u32(sA_hi), a_lo = tc::smem_u32(sA_lo);
            const uint32_t b_hi = tc::smem_u32(sB_hi), b_lo = tc::smem_u32(sB_lo);
#pragma unroll
            for (int j = 0; j < tc::TK / 8; ++j) {
                const uint32_t ko = 2 * j * tc::CORE;        // two K-blocks per instruction
                tc::umma_tf32(tmem, tc::make_desc(a_hi + ko), tc::make_desc(b_hi + ko), j > 0 ? 1u : 0u);
                tc::umma_tf32(tmem, tc::make_desc(a_hi + ko), tc::make_desc(b_lo + ko), 1u);
                tc::umma_tf32(tmem, tc::make_desc(a_lo + ko), tc::make_desc(b_hi + ko), 1u);
            }
            tc::umma_commit(&s_bar);
        }
        tc::mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();

        // ---- epilogue: thread = TMEM lane = anchor.  (tcgen05.ld is warp-collective: every
        //      thread loads, only active anchors use the values.)  The A tiles are free now: s_h
        //      lives there. ----
        const uint32_t lane_base = tmem + (uint32_t(warp * 32) << 16);
        float op[NOFF];
        uint32_t m = 0;
        {
            float h[FEAT];
            tc::tmem_ld32(lane_base + 0, h);
            if (active) {
#pragma unroll
                for (int j = 0; j < FEAT; ++j) h[j] = fmaxf(h[j] + sw.b1[0][j], 0.f);
#pragma unroll
                for (int o = 0; o < NOFF; ++o) {
                    op[o] = tanhf(dot32(sw.w2o[o], sw.b2o[o], h));
                    if (op[o] > 0.0f) m |= 1u << o;               // mask = neural_opacity > 0   (:278-279)
                }
            }
        }
        // (the MMAs — the last readers of the A tiles — have completed: s_h may overwrite them)
#pragma unroll
        for (int mlp = 1; mlp < 3; ++mlp) {
            float h[FEAT];
            tc::tmem_ld32(lane_base + mlp * FEAT, h);
            if (active && m != 0) {
#pragma unroll
                for (int q = 0; q < FEAT / 4; ++q)
                    *reinterpret_cast<float4*>(s_h + tid * H_W + (mlp - 1) * FEAT + 4 * q) =
                        make_float4(fmaxf(h[4 * q] + sw.b1[mlp][4 * q], 0.f), fmaxf(h[4 * q + 1] + sw.b1[mlp][4 * q + 1], 0.f),
                                    fmaxf(h[4 * q + 2] + sw.b1[mlp][4 * q + 2], 0.f), fmaxf(h[4 * q + 3] + sw.b1[mlp][4 * q + 3], 0.f));
            }
        }
        tc::fence_before_sync();         // TMEM reads are ordered before the next tile's MMAs (across the barriers below)

        uint32_t n_rows;
        const uint32_t row_off = cta_exclusive_scan(__popc(m), s_warp, &n_rows);

        // ---- order across CTAs: exclusive prefixes of visible anchors and surviving rows ----
        if (tid < 32) {
            const uint32_t vb = lookback(st.look_vis, tile, n_vis, lane);
            const uint32_t rb = lookback(st.look_row, tile, n_rows, lane);
            if (lane == 0) {
                s_vis_base = vb;
                s_row_base = rb;
                if (tile == ntiles - 1 && host_counts != nullptr) {
                    host_counts[0] = vb + n_vis;
                    host_counts[1] = rb + n_rows;
                    __threadfence_system();
                }
            }
        }
        if (active) {
            s_row[tid] = row_off;
            float* rec = s_rec + tid * REC_W;
            rec[0] = in.ax; rec[1] = in.ay; rec[2] = in.az;
#pragma unroll
            for (int k = 0; k < 6; ++k) rec[3 + k] = in.s[k];
#pragma unroll
            for (int o = 0; o < NOFF; ++o) rec[9 + o] = op[o];
            rec[19] = __uint_as_float(m);
        }
        if (tid == 0) s_row[n_vis] = n_rows;                  // sentinel for the search below
        __syncthreads();
        const size_t vis_base = s_vis_base, row_base = s_row_base;
        if (active) {
            const size_t ordinal = vis_base + tid;
            st.anchor_index[ordinal] = (uint32_t)a;
            st.row_start[ordinal] = (uint32_t)(row_base + row_off);
            st.mask_bits[ordinal] = m;
#pragma unroll
            for (int o = 0; o < NOFF; ++o) {
                neural_opacity[ordinal * NOFF + o] = op[o];
                out_mask[ordinal * NOFF + o] = (m >> o) & 1u;
            }
        }

        // ---- D4 second layers + D5: thread = output row; consecutive threads write consecutive
        //      rows, and no lane idles on a masked-out offset ----
        for (uint32_t r = tid; r < n_rows; r += DEC_THREADS) {
            int lo = 0, hi = (int)n_vis;                 // s_row[lo] <= r < s_row[hi]
#pragma unroll
            for (int step = 0; step < 7; ++step) {
                const int mid = (lo + hi) >> 1;
                if (mid > lo && s_row[mid] <= r) lo = mid; else if (mid > lo) hi = mid;
            }
            const float* rec = s_rec + lo * REC_W;
            const uint32_t mm = __float_as_uint(rec[19]);
            const int o = __fns(mm, 0, (int)(r - s_row[lo]) + 1);     // the (r - first)-th surviving offset
            const size_t aid = s_aid[lo];
            float h[FEAT];
#pragma unroll
            for (int q = 0; q < FEAT / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(s_h + lo * H_W + 4 * q);
                h[4 * q] = t.x; h[4 * q + 1] = t.y; h[4 * q + 2] = t.z; h[4 * q + 3] = t.w;
            }
            float sr[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) sr[k] = dot32(sw.w2s[7 * o + k], sw.b2s[7 * o + k], h);
            const size_t row = row_base + r;
            const float* off = offset + (aid * NOFF + o) * 3;
            // xyz = anchor + offset * scaling[:3]; scaling = scaling[3:] * sigmoid(sr[:3]); rot = normalize(sr[3:7])
            out_xyz[3 * row] = rec[0] + __ldg(off) * rec[3];
            out_xyz[3 * row + 1] = rec[1] + __ldg(off + 1) * rec[4];
            out_xyz[3 * row + 2] = rec[2] + __ldg(off + 2) * rec[5];
            out_scaling[3 * row] = rec[6] * sigmoidf_(sr[0]);
            out_scaling[3 * row + 1] = rec[7] * sigmoidf_(sr[1]);
            out_scaling[3 * row + 2] = rec[8] * sigmoidf_(sr[2]);
            const float nrm = fmaxf(sqrtf(sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6]), 1e-12f);
            *reinterpret_cast<float4*>(out_rot + 4 * row) = make_float4(sr[3] / nrm, sr[4] / nrm, sr[5] / nrm, sr[6] / nrm);
            out_opacity[row] = rec[9 + o];
#pragma unroll
            for (int q = 0; q < FEAT / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(s_h + lo * H_W + FEAT + 4 * q);
                h[4 * q] = t.x; h[4 * q + 1] = t.y; h[4 * q + 2] = t.z; h[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) out_color[3 * row + k] = sigmoidf_(dot32(sw.w2c[3 * o + k], sw.b2c[3 * o + k], h));
        }
    }

    // ---- teardown: the allocating warp frees TMEM once every warp is done with it ----
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tc::TMEM_COLS);
}

// =======================================================================================
// forward, variant 2 (round 2): BOTH layers of the three MLPs on the tensor cores, tile = 128 VISIBLE anchors
// =======================================================================================
// Variant 1 above walks the anchors in index order (a tile of 128 anchors holds ~25 visible ones in a mapping view),
// runs the second layers as scalar FFMA chains with thread = anchor / thread = row, and keeps 128 threads busy with
// ~8000 dependent instructions per tile: 43 us per tile, 12 % occupancy, issue slots 19 % busy.  Variant 2:
//   * `decode_compact_kernel` lists the visible anchors first (one scan + look-back pass over the mask bytes), so a
//     tile is 128 consecutive visible ORDINALS: 5x fewer tiles in a mapping view, and the visible prefix of a tile is
//     its index (one look-back per tile instead of two);
//   * CTA = 512 threads, one per SM (all operands of a tile stay in its 192 KB of shared memory); the input rows are
//     built by 4 threads per anchor (feature-bank hidden units split 4 ways, logits reduced with two shuffles);
//   * layer 1:  H[128 x 96]  = X[128 x 40] W1cat^T            15 tcgen05.mma (M 128, N 96, K 8, 3xTF32) -> TMEM cols 0..95
//     epilogue 1 (12 warps, one MLP's 32 hidden units of 32 anchors each): bias + ReLU, hi/lo split, written straight
//     back to shared memory as the K-major A operands of
//   * layer 2:  O_m[128 x N_m] = relu(H_m)[128 x 32] W2_m^T    3 x 12 tcgen05.mma, N = 16 / 80 / 32 (opacity / cov /
//     colour, zero-padded rows) -> TMEM cols 128.. / 160.. / 96..
//     epilogue 2 (16 warps x 32 columns): + bias, into a [128][113]-float tile in shared memory (it overlays the dead
//     A operands); tanh + mask of the opacity columns then run one (anchor, offset) pair per thread;
//   * the compacted rows are then assembled by thread = output row exactly as in variant 1 (activation, quaternion
//     normalisation, coalesced stores), now with no dot products left in the loop.
// 51 UTCHMMA per tile; every scalar FMA chain of the MLPs is gone from the forward.
namespace tc2 {

constexpr int THREADS = 512;
constexpr int KB2 = FEAT / 4;                          // 16-byte K-blocks per row of a layer-2 operand (K = 32)
constexpr uint32_t SBO1 = tc::KB * tc::CORE;           // layer-1 operands (K = 40)
constexpr uint32_t SBO2 = KB2 * tc::CORE;              // layer-2 operands
constexpr int N_OP = 16, N_COV = 80, N_COL = 32;       // MMA N of the second layers (10 / 70 / 30 outputs, zero-padded)
constexpr int B2_OP = 0;                               // byte offsets of the three W2 tiles inside the B2 buffer
constexpr int B2_COV = B2_OP + N_OP * FEAT * 4;        // 2048
constexpr int B2_COL = B2_COV + N_COV * FEAT * 4;      // 12288
constexpr int B2_BYTES = B2_COL + N_COL * FEAT * 4;    // 16384
constexpr int A2_BYTES = tc::TM * FEAT * 4;            // 16384 per MLP and per hi / lo
constexpr int COL_D2C = 96, COL_D2O = 128, COL_D2S = 160;   // TMEM columns of the layer-2 accumulators (layer 1: 0..95)
constexpr int TMEM_COLS = 256;
constexpr int OUT_W = 113;                             // floats per anchor in s_out: op(10) | scale_rot(70) | colour(30); odd stride
constexpr int O_OP = 0, O_COV = NOFF, O_COL = NOFF + 7 * NOFF;
constexpr int U_BYTES = 6 * A2_BYTES;                  // union: A1 hi|lo  /  A2 hi[3] | lo[3]  /  s_out
static_assert(2 * tc::A_BYTES <= U_BYTES && tc::TM * OUT_W * 4 <= U_BYTES, "operand union");

__host__ __device__ constexpr uint32_t idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(tc::TM >> 4) << 24);
}
// byte offset of element (row, k) of a K-major no-swizzle operand with KB2 K-blocks per row
__device__ __forceinline__ uint32_t canon2(int row, int k) {
    return uint32_t(((row >> 3) * KB2 + (k >> 2)) * tc::CORE + (row & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(tc::LBO >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t id, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(id), "r"(accumulate) : "memory");
}

constexpr size_t OFF_B1H = align128(sizeof(SWF));
constexpr size_t OFF_B1L = OFF_B1H + tc::B_BYTES;
constexpr size_t OFF_B2H = OFF_B1L + tc::B_BYTES;
constexpr size_t OFF_B2L = OFF_B2H + B2_BYTES;
constexpr size_t OFF_U = (OFF_B2L + B2_BYTES + 1023) & ~size_t(1023);
constexpr size_t OFF_REC = OFF_U + U_BYTES;
constexpr size_t OFF_ROW = OFF_REC + sizeof(float) * tc::TM * REC_W;          // uint32 [TM + 4]
constexpr size_t OFF_AID = OFF_ROW + sizeof(uint32_t) * (tc::TM + 4);          // uint32 [TM]
constexpr size_t OFF_MSK = OFF_AID + sizeof(uint32_t) * tc::TM;                // uint32 [TM]
constexpr size_t SMEM = OFF_MSK + sizeof(uint32_t) * tc::TM;
static_assert(SMEM <= 227 * 1024, "decode forward v2: shared memory");

}  // namespace tc2

constexpr int CMP_PER = 8;                             // anchors per thread of the compaction pass
constexpr int CMP_TILE = DEC_THREADS * CMP_PER;

// Visible-anchor list: anchor_index[ordinal] = a for the visible anchors in ascending order, counters[1] = their number.
__global__ void __launch_bounds__(DEC_THREADS)
decode_compact_kernel(int A, const unsigned char* __restrict__ visible_mask, DecodeState st)
{
    __shared__ uint32_t s_warp[DEC_THREADS / 32];
    __shared__ uint32_t s_tile, s_base;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t ntiles = (uint32_t)((A + CMP_TILE - 1) / CMP_TILE);
    if (tid == 0) s_tile = atomicAdd(st.counters + 2, 1u);          // ticket: a tile only waits for tiles that run
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= ntiles) return;
    const size_t a0 = size_t(tile) * CMP_TILE + size_t(tid) * CMP_PER;
    uint32_t bits = 0;
    if ((reinterpret_cast<uintptr_t>(visible_mask) & 7u) == 0 && a0 + CMP_PER <= (size_t)A) {
        const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long*>(visible_mask + a0));
#pragma unroll
        for (int k = 0; k < CMP_PER; ++k) if ((v >> (8 * k)) & 0xFFull) bits |= 1u << k;
    } else {
#pragma unroll
        for (int k = 0; k < CMP_PER; ++k) if (a0 + k < (size_t)A && visible_mask[a0 + k] != 0) bits |= 1u << k;
    }
    uint32_t total;
    const uint32_t ord = cta_exclusive_scan(__popc(bits), s_warp, &total);
    if (tid < 32) {
        const uint32_t b = lookback(st.look_vis, tile, total, lane);
        if (lane == 0) {
            s_base = b;
            if (tile == ntiles - 1) st.counters[1] = b + total;
        }
    }
    __syncthreads();
    uint32_t o = s_base + ord;
#pragma unroll
    for (int k = 0; k < CMP_PER; ++k) if ((bits >> k) & 1u) st.anchor_index[o++] = (uint32_t)(a0 + k);
}

__global__ void __launch_bounds__(tc2::THREADS, 1)
decode_forward_v2_kernel(int A, const int identity, const float* __restrict__ anchor,
                         const float* __restrict__ anchor_feat, const float* __restrict__ offset,
                         const float* __restrict__ scaling, const float* __restrict__ cam, const Pose7 pose,
                         const segs_decode_params p, float* __restrict__ out_xyz, float* __restrict__ out_color,
                         float* __restrict__ out_opacity, float* __restrict__ out_scaling, float* __restrict__ out_rot,
                         float* __restrict__ neural_opacity, unsigned char* __restrict__ out_mask, DecodeState st,
                         volatile uint32_t* __restrict__ host_counts)
{
    extern __shared__ __align__(1024) unsigned char s_dec[];
    SWF& sw = *reinterpret_cast<SWF*>(s_dec);
    unsigned char* sB1_hi = s_dec + tc2::OFF_B1H;
    unsigned char* sB1_lo = s_dec + tc2::OFF_B1L;
    unsigned char* sB2_hi = s_dec + tc2::OFF_B2H;
    unsigned char* sB2_lo = s_dec + tc2::OFF_B2L;
    unsigned char* sU = s_dec + tc2::OFF_U;
    unsigned char* sA1_hi = sU;                                            // layer-1 A operand, hi / lo
    unsigned char* sA1_lo = sU + tc::A_BYTES;
    unsigned char* sA2_hi = sU;                                            // layer-2 A operands [3], hi / lo
    unsigned char* sA2_lo = sU + 3 * tc2::A2_BYTES;
    float* s_out = reinterpret_cast<float*>(sU);                           // [128][OUT_W] layer-2 outputs (+ bias)
    float* s_rec = reinterpret_cast<float*>(s_dec + tc2::OFF_REC);         // [128][REC_W]: anchor, scaling, opacities
    uint32_t* s_row = reinterpret_cast<uint32_t*>(s_dec + tc2::OFF_ROW);   // [129] first row of every anchor of the tile
    uint32_t* s_aid = reinterpret_cast<uint32_t*>(s_dec + tc2::OFF_AID);   // [128] anchor ids
    uint32_t* s_msk = reinterpret_cast<uint32_t*>(s_dec + tc2::OFF_MSK);   // [128] surviving-offset bits
    __shared__ uint32_t s_warp[4];
    __shared__ uint32_t s_tile[2], s_look[tc2::THREADS / 32], s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = warp & 3, grp = warp >> 2;          // TMEM lane quarter this warp may read; column group
    const uint32_t n_vis = identity ? (uint32_t)A : st.counters[1];
    const uint32_t ntiles = (n_vis + tc::TM - 1) / tc::TM;

    // ---- once per (persistent) CTA: TMEM, mbarrier, weights ----
    if (warp == 0) tc::tmem_alloc(&s_tmem, tc2::TMEM_COLS);
    if (tid == 0) {
        tc::mbar_init(&s_bar, 1);
        s_tile[0] = atomicAdd(st.counters, 1u);
        if (blockIdx.x == 0) st.counters[3] = 1u;                          // hidden / out2 are valid for the backward
        if (ntiles == 0 && blockIdx.x == 0 && host_counts != nullptr) {    // nothing visible: nobody else reports
            host_counts[0] = 0;
            host_counts[1] = 0;
            __threadfence_system();
        }
    }
    __syncthreads();                                   // s_tile[0]
    const bool use_bank = p.use_feat_bank != 0;
    // raw inputs of this thread's (anchor slot, quarter) for one tile: fetched one tile ahead, so their global-memory
    // latency hides behind the previous tile's second layer and row assembly (the first tile's: behind the weight staging)
    struct Raw { size_t a; float ax, ay, az; float4 f0, f1; float fa[8], fb[8]; };
    auto fetch = [&](uint32_t t) {
        Raw r;
        const int i = tid >> 2, q = tid & 3;
        const uint32_t o0 = t * tc::TM, na = min((uint32_t)tc::TM, n_vis - o0);
        const uint32_t oi = o0 + ((uint32_t)i < na ? (uint32_t)i : 0u);     // idle slots read the tile's first anchor
        r.a = identity ? size_t(oi) : size_t(st.anchor_index[oi]);
        r.ax = __ldg(anchor + 3 * r.a); r.ay = __ldg(anchor + 3 * r.a + 1); r.az = __ldg(anchor + 3 * r.a + 2);
        const float* frow = anchor_feat + r.a * FEAT;
        r.f0 = __ldg(reinterpret_cast<const float4*>(frow) + 2 * q);
        r.f1 = __ldg(reinterpret_cast<const float4*>(frow) + 2 * q + 1);
        if (use_bank) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                r.fa[jj] = __ldg(frow + 4 * jj);                      // feat[4 (j mod 8)],  j = 8q + jj
                r.fb[jj] = __ldg(frow + 16 * (q & 1) + 2 * jj);       // feat[2 (j mod 16)]
            }
        }
        return r;
    };
    Raw raw;
    bool have_raw = s_tile[0] < ntiles;
    if (have_raw) raw = fetch(s_tile[0]);
    {
        // B operands.  Layer 1: W1cat[n][k], n = 32 * mlp + hidden unit, k = input column (as in variant 1);
        // layer 2: one tile per MLP, row = output unit (zero rows pad N to 16 / 80 / 32), k = hidden unit.
        const int in_o = 35 + (p.add_opacity_dist ? 1 : 0), in_s = 35 + (p.add_cov_dist ? 1 : 0);
        const int in_c = 35 + (p.add_color_dist ? 1 : 0), ld_c = in_c + p.appearance_dim;
        for (int e = tid; e < tc::TN * tc::TK; e += tc2::THREADS) {
            const int n = e / tc::TK, k = e % tc::TK, j = n & 31;
            float w = 0.f;
            if (n < 32) { if (k < in_o) w = __ldg(p.opacity_w1 + j * in_o + k); }
            else if (n < 64) { if (k < in_s) w = __ldg(p.cov_w1 + j * in_s + k); }
            else { if (k < in_c) w = __ldg(p.color_w1 + j * ld_c + k); }
            const float hi = tc::tf32_rn(w);
            const uint32_t off = tc::canon_off(n, k);
            *reinterpret_cast<float*>(sB1_hi + off) = hi;
            *reinterpret_cast<float*>(sB1_lo + off) = tc::tf32_rn(w - hi);
        }
        for (int e = tid; e < 128 * FEAT; e += tc2::THREADS) {
            const int n = e / FEAT, k = e % FEAT;
            float w = 0.f;
            uint32_t off;
            if (n < tc2::N_OP) {
                if (n < NOFF) w = __ldg(p.opacity_w2 + n * FEAT + k);
                off = tc2::B2_OP + tc2::canon2(n, k);
            } else if (n < tc2::N_OP + tc2::N_COV) {
                const int r = n - tc2::N_OP;
                if (r < 7 * NOFF) w = __ldg(p.cov_w2 + r * FEAT + k);
                off = tc2::B2_COV + tc2::canon2(r, k);
            } else {
                const int r = n - tc2::N_OP - tc2::N_COV;
                if (r < 3 * NOFF) w = __ldg(p.color_w2 + r * FEAT + k);
                off = tc2::B2_COL + tc2::canon2(r, k);
            }
            const float hi = tc::tf32_rn(w);
            *reinterpret_cast<float*>(sB2_hi + off) = hi;
            *reinterpret_cast<float*>(sB2_lo + off) = tc::tf32_rn(w - hi);
        }
    }
    stage_weights(sw, p, pose);          // biases, feature-bank weights, appearance fold; contains __syncthreads
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    uint32_t phase = 0;
    // descriptors of the operand tiles (the start address sits in the low bits in units of 16 bytes: stepping along K or to
    // another tile is an add)
    const uint64_t dA1_hi = tc2::make_desc(tc::smem_u32(sA1_hi), tc2::SBO1), dA1_lo = tc2::make_desc(tc::smem_u32(sA1_lo), tc2::SBO1);
    const uint64_t dB1_hi = tc2::make_desc(tc::smem_u32(sB1_hi), tc2::SBO1), dB1_lo = tc2::make_desc(tc::smem_u32(sB1_lo), tc2::SBO1);
    const uint64_t dA2_hi = tc2::make_desc(tc::smem_u32(sA2_hi), tc2::SBO2), dA2_lo = tc2::make_desc(tc::smem_u32(sA2_lo), tc2::SBO2);
    const uint64_t dB2_hi = tc2::make_desc(tc::smem_u32(sB2_hi), tc2::SBO2), dB2_lo = tc2::make_desc(tc::smem_u32(sB2_lo), tc2::SBO2);

    for (int it = 0;; ++it) {
        const uint32_t tile = s_tile[it & 1];
        if (tile >= ntiles) break;
        if (tid == 0) s_tile[(it + 1) & 1] = atomicAdd(st.counters, 1u);   // next ticket (read after the barriers below)
        const uint32_t ord0 = tile * tc::TM;                               // visible ordinal of row 0
        const uint32_t n_act = min((uint32_t)tc::TM, n_vis - ord0);
        if (!have_raw) raw = fetch(tile);

        // ---- D1-D3: MLP input rows -> layer-1 A operand (hi / lo).  4 threads per anchor: thread q owns the
        //      feature columns / bank hidden units [8q, 8q + 8) ----
        if (tid < tc::TM) s_msk[tid] = 0u;               // filled with atomicOr after the second layer
        {
            const int i = tid >> 2, q = tid & 3;
            const bool act = (uint32_t)i < n_act;        // idle slots compute on the tile's first anchor (the shuffles
            {                                            // below need every lane) and store nothing
                const size_t a = raw.a;
                const float ax = raw.ax, ay = raw.ay, az = raw.az;
                const float vx = ax - __ldg(cam), vy = ay - __ldg(cam + 1), vz = az - __ldg(cam + 2);
                const float dist = sqrtf(vx * vx + vy * vy + vz * vz);
                const float ux = vx / dist, uy = vy / dist, uz = vz / dist;
                float x[8];
                {
                    const float4 t0 = raw.f0, t1 = raw.f1;
                    x[0] = t0.x; x[1] = t0.y; x[2] = t0.z; x[3] = t0.w; x[4] = t1.x; x[5] = t1.y; x[6] = t1.z; x[7] = t1.w;
                }
                if (use_bank) {
                    // bank weights = softmax(W2 relu(W1 [view, dist] + b1) + b2)   (gaussian_renderer.cpp:236-239)
                    float l0 = 0.f, l1 = 0.f, l2 = 0.f;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = 8 * q + jj;
                        const float4 w4 = *reinterpret_cast<const float4*>(sw.wb1[j]);
                        float hb = sw.bb1[j];
                        hb = fmaf(w4.x, ux, hb); hb = fmaf(w4.y, uy, hb); hb = fmaf(w4.z, uz, hb); hb = fmaf(w4.w, dist, hb);
                        hb = fmaxf(hb, 0.f);
                        l0 = fmaf(sw.wb2[0][j], hb, l0); l1 = fmaf(sw.wb2[1][j], hb, l1); l2 = fmaf(sw.wb2[2][j], hb, l2);
                    }
                    // the four threads of an anchor are adjacent lanes: butterfly sums are identical in all four
                    l0 += __shfl_xor_sync(FULL, l0, 1); l1 += __shfl_xor_sync(FULL, l1, 1); l2 += __shfl_xor_sync(FULL, l2, 1);
                    l0 += __shfl_xor_sync(FULL, l0, 2); l1 += __shfl_xor_sync(FULL, l1, 2); l2 += __shfl_xor_sync(FULL, l2, 2);
                    l0 += sw.bb2[0]; l1 += sw.bb2[1]; l2 += sw.bb2[2];
                    const float mx = fmaxf(l0, fmaxf(l1, l2));
                    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx);
                    const float den = e0 + e1 + e2;
                    const float w0 = e0 / den, w1 = e1 / den, w2 = e2 / den;
                    // feat'[j] = feat[4 (j mod 8)] w0 + feat[2 (j mod 16)] w1 + feat[j] w2   (:241-248; repeat = tiling)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) x[jj] = raw.fa[jj] * w0 + raw.fb[jj] * w1 + x[jj] * w2;
                }
                auto put4 = [&](int kb, float v0, float v1, float v2, float v3) {
                    if (!act) return;
                    const tc::Split4 sp = tc::split4(v0, v1, v2, v3);
                    const uint32_t off = tc::canon_off(i, 4 * kb);
                    *reinterpret_cast<float4*>(sA1_hi + off) = sp.hi;
                    *reinterpret_cast<float4*>(sA1_lo + off) = sp.lo;
                };
                put4(2 * q, x[0], x[1], x[2], x[3]);
                put4(2 * q + 1, x[4], x[5], x[6], x[7]);
                if (q == 0 && act) {
                    put4(8, ux, uy, uz, dist);
                    float* rec = s_rec + i * REC_W;
                    rec[0] = ax; rec[1] = ay; rec[2] = az;
#pragma unroll
                    for (int k = 0; k < 6; ++k) rec[3 + k] = __ldg(scaling + 6 * a + k);
                    s_aid[i] = (uint32_t)a;
                } else if (q == 1) {
                    put4(9, 0.f, 0.f, 0.f, 0.f);          // K padding: the weights there are zero, the operand must be finite
                }
            }
        }
        tc::fence_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
        tc::fence_before_sync();
        __syncthreads();

        // ---- D4, first layers: 5 K-steps x 3 split products ----
        if (tid == 0) {
            tc::fence_after_sync();
            constexpr uint32_t id = tc2::idesc(tc::TN);
#pragma unroll
            for (int j = 0; j < tc::TK / 8; ++j) {
                const uint64_t ko = (2 * j * tc::CORE) >> 4;        // two K-blocks per instruction
                tc2::umma_tf32(tmem, dA1_hi + ko, dB1_hi + ko, id, j > 0 ? 1u : 0u);
                tc2::umma_tf32(tmem, dA1_hi + ko, dB1_lo + ko, id, 1u);
                tc2::umma_tf32(tmem, dA1_lo + ko, dB1_hi + ko, id, 1u);
            }
            tc::umma_commit(&s_bar);
        }
        tc::mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();

        // ---- epilogue 1: warp (lq, grp < 3) = 32 anchors x the 32 hidden units of MLP grp.  bias + ReLU, hi / lo
        //      split, straight into the layer-2 A operand (the layer-1 operand it overlays has been consumed) ----
        const int row = lq * 32 + lane;                  // TMEM lane = anchor slot of this thread in both epilogues
        const uint32_t lane_base = tmem + (uint32_t(lq * 32) << 16);
        if (grp < 3) {
            float h[FEAT];
            tc::tmem_ld32(lane_base + grp * FEAT, h);
            unsigned char* a_hi = sA2_hi + grp * tc2::A2_BYTES;
            unsigned char* a_lo = sA2_lo + grp * tc2::A2_BYTES;
#pragma unroll
            for (int kb = 0; kb < tc2::KB2; ++kb) {
                const float p0 = h[4 * kb] + sw.b1[grp][4 * kb], p1 = h[4 * kb + 1] + sw.b1[grp][4 * kb + 1];
                const float p2 = h[4 * kb + 2] + sw.b1[grp][4 * kb + 2], p3 = h[4 * kb + 3] + sw.b1[grp][4 * kb + 3];
                const float v0 = fmaxf(p0, 0.f), v1 = fmaxf(p1, 0.f), v2 = fmaxf(p2, 0.f), v3 = fmaxf(p3, 0.f);
                const tc::Split4 sp = tc::split4(v0, v1, v2, v3);
                const uint32_t off = tc2::canon2(row, 4 * kb);
                *reinterpret_cast<float4*>(a_hi + off) = sp.hi;
                *reinterpret_cast<float4*>(a_lo + off) = sp.lo;
                // the backward reads the PRE-activations back instead of recomputing the layer (128 contiguous bytes per
                // thread); it redoes a unit in FP32 only where the sign of the ReLU argument is within the 3xTF32 error
                if ((uint32_t)row < n_act)
                    *reinterpret_cast<float4*>(st.hidden + size_t(ord0 + row) * (3 * FEAT) + grp * FEAT + 4 * kb) = make_float4(p0, p1, p2, p3);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();

        // ---- D4, second layers: per MLP 4 K-steps x 3 split products ----
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int mlp = 0; mlp < 3; ++mlp) {
                const uint32_t boff = mlp == 0 ? tc2::B2_OP : mlp == 1 ? tc2::B2_COV : tc2::B2_COL;
                const uint32_t id = mlp == 0 ? tc2::idesc(tc2::N_OP) : mlp == 1 ? tc2::idesc(tc2::N_COV) : tc2::idesc(tc2::N_COL);
                const uint32_t d = tmem + (mlp == 0 ? tc2::COL_D2O : mlp == 1 ? tc2::COL_D2S : tc2::COL_D2C);
#pragma unroll
                for (int j = 0; j < FEAT / 8; ++j) {
                    const uint64_t ka = (mlp * tc2::A2_BYTES + 2 * j * tc::CORE) >> 4, kb = (boff + 2 * j * tc::CORE) >> 4;
                    tc2::umma_tf32(d, dA2_hi + ka, dB2_hi + kb, id, j > 0 ? 1u : 0u);
                    tc2::umma_tf32(d, dA2_hi + ka, dB2_lo + kb, id, 1u);
                    tc2::umma_tf32(d, dA2_lo + ka, dB2_hi + kb, id, 1u);
                }
            }
            tc::umma_commit(&s_bar);
        }
        {
            // the next tile's raw inputs go out now: behind the last proxy fence of this tile (fence.proxy.async is
            // MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and would wait for them), ahead of everything that remains
            const uint32_t nxt = s_tile[(it + 1) & 1];
            have_raw = nxt < ntiles;
            if (have_raw) raw = fetch(nxt);
        }
        tc::mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();

        // ---- epilogue 2: 16 warps x 32 accumulator columns -> s_out (+ bias); the opacity warps apply tanh and
        //      take the mask.  s_out overlays the layer-2 A operands, which the MMAs have consumed. ----
        {
            float v[32];
            float* so = s_out + row * tc2::OUT_W;
            if (grp == 0) {                                  // colour pre-activations
                tc::tmem_ld32(lane_base + tc2::COL_D2C, v);
#pragma unroll
                for (int c = 0; c < 3 * NOFF; ++c) so[tc2::O_COL + c] = v[c] + sw.b2c[c];
            } else if (grp == 1) {                           // opacity pre-activations
                tc::tmem_ld32(lane_base + tc2::COL_D2O, v);
#pragma unroll
                for (int o = 0; o < NOFF; ++o) so[tc2::O_OP + o] = v[o] + sw.b2o[o];
                tc::tmem_ld32(lane_base + tc2::COL_D2S + 64, v);     // scale_rot units 64..69
#pragma unroll
                for (int c = 0; c < 7 * NOFF - 64; ++c) so[tc2::O_COV + 64 + c] = v[c] + sw.b2s[64 + c];
            } else {                                         // scale_rot units 0..31 / 32..63
                const int c0 = (grp - 2) * 32;
                tc::tmem_ld32(lane_base + tc2::COL_D2S + c0, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) so[tc2::O_COV + c0 + c] = v[c] + sw.b2s[c0 + c];
            }
        }
        tc::fence_before_sync();         // TMEM reads are ordered before the next tile's MMAs (across the barriers below)
        __syncthreads();

        // ---- opacity = tanh, mask = neural_opacity > 0 (:278-279): one (anchor, offset) pair per thread and step ----
        for (uint32_t e = tid; e < n_act * NOFF; e += tc2::THREADS) {
            const uint32_t k = e / NOFF, o = e - k * NOFF;
            const float t = tanhf(s_out[k * tc2::OUT_W + tc2::O_OP + o]);
            s_out[k * tc2::OUT_W + tc2::O_OP + o] = t;
            s_rec[k * REC_W + 9 + o] = t;
            if (t > 0.0f) atomicOr(&s_msk[k], 1u << o);
        }
        __syncthreads();

        // ---- rows of the tile: exclusive scan of the surviving offsets per anchor (warps 0..3) ----
        uint32_t cnt = 0, inc = 0;
        if (tid < tc::TM) {
            cnt = __popc(s_msk[tid]);
            inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane == 31) s_warp[warp] = inc;
        }
        __syncthreads();
        const uint32_t n_rows = s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
        if (tid < tc::TM) {
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) if (w < warp) woff += s_warp[w];
            s_row[tid] = woff + inc - cnt;
        }
        if (tid == 0) s_row[tc::TM] = n_rows;                 // sentinel for the search below
        // ---- order across CTAs: exclusive prefix of the surviving rows (the visible prefix is ord0).  The tiles of a
        //      wave publish their aggregates at about the same time, so a look-back walks ~148 flags to the previous
        //      wave's prefix: the 16 warps read 16 windows of 32 flags AT ONCE (one global-memory round trip instead of
        //      five in a row) and every thread combines them ----
        volatile uint32_t* look = st.look_row;
        if (tid == 0) look[tile] = (tile == 0 ? FLAG_PREFIX : FLAG_AGG) | n_rows;
        // per-anchor outputs, coalesced over (ordinal, offset) — in flight while the flags are polled
        for (uint32_t e = tid; e < n_act * NOFF; e += tc2::THREADS) {
            const uint32_t k = e / NOFF, o = e - k * NOFF;
            neural_opacity[size_t(ord0) * NOFF + e] = s_rec[k * REC_W + 9 + o];
            out_mask[size_t(ord0) * NOFF + e] = (unsigned char)((s_msk[k] >> o) & 1u);
        }
        // second-layer outputs for the backward (which then has no dot product of the forward left to redo)
        for (uint32_t e = tid; e < n_act * O2_W; e += tc2::THREADS) {
            const uint32_t k = e / O2_W, c = e - k * O2_W;
            st.out2[size_t(ord0) * O2_W + e] = s_out[k * tc2::OUT_W + c];
        }
        uint32_t excl = 0;
        if (tile != 0) {
            for (int newest = (int)tile - 1;; newest -= tc2::THREADS) {
                const int t = newest - 32 * warp - lane;
                uint32_t w = FLAG_PREFIX;                            // before tile 0: prefix 0
                if (t >= 0) {
                    do { w = look[t]; } while ((w & FLAG_MASK) == 0u);
                }
                const unsigned pref = __ballot_sync(FULL, (w & FLAG_MASK) == FLAG_PREFIX);
                const int stop = __ffs(pref) - 1;
                const uint32_t v = (stop < 0 || lane <= stop) ? (w & ~FLAG_MASK) : 0u;
                const uint32_t sum = __reduce_add_sync(FULL, v);
                if (lane == 0) s_look[warp] = sum | (stop >= 0 ? 0x80000000u : 0u);
                __syncthreads();
                bool done = false;
#pragma unroll
                for (int q = 0; q < tc2::THREADS / 32; ++q) {
                    if (!done) {
                        const uint32_t x = s_look[q];
                        excl += x & 0x7FFFFFFFu;
                        done = (x & 0x80000000u) != 0u;
                    }
                }
                if (done) break;
                __syncthreads();                                     // s_look is rewritten in the next round
            }
            if (tid == 0) look[tile] = FLAG_PREFIX | (excl + n_rows);
        }
        if (tid == 0 && tile == ntiles - 1 && host_counts != nullptr) {
            host_counts[0] = n_vis;
            host_counts[1] = excl + n_rows;
            __threadfence_system();
        }
        const size_t row_base = excl;
        __syncthreads();                 // s_row / s_msk of this tile are complete for every thread (tile 0 polls nothing)
        if ((uint32_t)tid < n_act) {
            if (identity) st.anchor_index[ord0 + tid] = s_aid[tid];
            st.row_start[ord0 + tid] = (uint32_t)(row_base + s_row[tid]);
            st.mask_bits[ord0 + tid] = s_msk[tid];
        }

        // ---- D5: thread = output row; consecutive threads write consecutive rows.  Two rows per thread are located
        //      (binary search, offset load) before either is assembled, so their latencies overlap ----
        struct RowRef { int lo, o; float ox, oy, oz; };
        auto locate = [&](uint32_t r) {
            RowRef rr;
            int lo = 0, hi = tc::TM;                     // s_row[lo] <= r < s_row[hi]
#pragma unroll
            for (int step = 0; step < 7; ++step) {
                const int mid = (lo + hi) >> 1;
                if (mid > lo && s_row[mid] <= r) lo = mid; else if (mid > lo) hi = mid;
            }
            rr.lo = lo;
            rr.o = __fns(s_msk[lo], 0, (int)(r - s_row[lo]) + 1);     // the (r - first)-th surviving offset
            const float* off = offset + (size_t(s_aid[lo]) * NOFF + rr.o) * 3;
            rr.ox = __ldg(off); rr.oy = __ldg(off + 1); rr.oz = __ldg(off + 2);
            return rr;
        };
        auto emit = [&](uint32_t r, const RowRef& rr) {
            const float* rec = s_rec + rr.lo * REC_W;
            const float* sr = s_out + rr.lo * tc2::OUT_W + tc2::O_COV + 7 * rr.o;
            const float* sc = s_out + rr.lo * tc2::OUT_W + tc2::O_COL + 3 * rr.o;
            const size_t orow = row_base + r;
            // xyz = anchor + offset * scaling[:3]; scaling = scaling[3:] * sigmoid(sr[:3]); rot = normalize(sr[3:7])
            out_xyz[3 * orow] = rec[0] + rr.ox * rec[3];
            out_xyz[3 * orow + 1] = rec[1] + rr.oy * rec[4];
            out_xyz[3 * orow + 2] = rec[2] + rr.oz * rec[5];
            out_scaling[3 * orow] = rec[6] * sigmoidf_(sr[0]);
            out_scaling[3 * orow + 1] = rec[7] * sigmoidf_(sr[1]);
            out_scaling[3 * orow + 2] = rec[8] * sigmoidf_(sr[2]);
            const float q0 = sr[3], q1 = sr[4], q2 = sr[5], q3 = sr[6];
            const float nrm = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
            *reinterpret_cast<float4*>(out_rot + 4 * orow) = make_float4(q0 / nrm, q1 / nrm, q2 / nrm, q3 / nrm);
            out_opacity[orow] = rec[9 + rr.o];
#pragma unroll
            for (int k = 0; k < 3; ++k) out_color[3 * orow + k] = sigmoidf_(sc[k]);
        };
        for (uint32_t r0 = tid; r0 < n_rows; r0 += 2 * tc2::THREADS) {
            const uint32_t r1 = r0 + tc2::THREADS;
            const bool two = r1 < n_rows;
            const RowRef a0 = locate(r0);
            const RowRef a1 = locate(two ? r1 : r0);
            emit(r0, a0);
            if (two) emit(r1, a1);
        }
        __syncthreads();                 // the tile is done with s_out / s_rec / s_row / s_aid / s_msk; next ticket is visible
    }

    // ---- teardown: the allocating warp frees TMEM once every warp is done with it ----
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tc2::TMEM_COLS);
}

// =======================================================================================
// backward, kernel 1: per-anchor back-propagation + factors of the weight gradients
// =======================================================================================
// scratch: [tile of 64 ordinals][FACT_ROWS][64] floats (feature-major inside a tile)
constexpr int FT = 64;                 // ordinals per scratch tile
constexpr int F_X = 0;                 // x[36]
constexpr int F_H = F_X + XDIM;        // h of opacity / cov / colour [3][32]
constexpr int F_DPRE = F_H + 3 * FEAT; // pre-activation gradients [3][32]
constexpr int F_D2O = F_DPRE + 3 * FEAT;   // d(pre-tanh) [10]
constexpr int F_D2S = F_D2O + NOFF;        // d(scale_rot) [70]
constexpr int F_D2C = F_D2S + 7 * NOFF;    // d(pre-sigmoid colour) [30]
constexpr int F_HB = F_D2C + 3 * NOFF;     // bank hidden [32]
constexpr int F_DPREB = F_HB + FEAT;       // bank pre-activation gradient [32]
constexpr int F_DLOG = F_DPREB + FEAT;     // bank logit gradients [3]
constexpr int F_CAT = F_DLOG + 3;          // bank input [view, dist] [4]
constexpr int FACT_ROWS = F_CAT + 4;       // 409

// CACHED: the forward left the layer activations in the state (variant 2).  Both instantiations are launched; the one
// that does not match the flag the forward wrote returns at once (the host cannot read the flag without a sync).  The
// split keeps each instantiation's code small: with both paths in one kernel the SASS was 820 KB and a quarter of the
// warp-stall samples were instruction-cache misses.
template <bool CACHED>
__global__ void __launch_bounds__(DEC_THREADS, 4)
decode_backward_kernel(int n_vis, const float* __restrict__ anchor, const float* __restrict__ anchor_feat,
                       const float* __restrict__ offset, const float* __restrict__ scaling,
                       const float* __restrict__ cam, const Pose7 pose, const segs_decode_params p, DecodeState st,
                       const float* __restrict__ g_xyz, const float* __restrict__ g_color,
                       const float* __restrict__ g_opacity, const float* __restrict__ g_scaling,
                       const float* __restrict__ g_rot, const float* __restrict__ g_nop,
                       float* __restrict__ d_anchor, float* __restrict__ d_feat, float* __restrict__ d_offset,
                       float* __restrict__ d_scaling, float* __restrict__ fact, const int flags)
{
    if ((st.counters[3] != 0u) != CACHED) return;
    constexpr bool cached = CACHED;
    __shared__ __align__(16) SW sw;
    stage_weights(sw, p, pose);
    // SEGS_DECODE_ACCUMULATE: add to the caller's arrays.  Every anchor row is owned by exactly one thread of one view, so
    // the sums do not depend on atomicity; RED.ADD is used either way because it does not wait for the old value
    // (`*dst += v` stalled on a cold global load per element: 17 % of the kernel's stall samples in a mapping view).
    // SEGS_DECODE_ATOMIC (several views accumulate concurrently) therefore needs nothing extra here.
    const bool acc_mode = (flags & SEGS_DECODE_ACCUMULATE) != 0;
    auto accum = [&](float* dst, float v) { atomicAdd(dst, v); };
  for (size_t ordinal = size_t(blockIdx.x) * DEC_THREADS + threadIdx.x; ordinal < (size_t)n_vis;
       ordinal += size_t(gridDim.x) * DEC_THREADS) {
    const size_t a = st.anchor_index[ordinal];
    const uint32_t m = st.mask_bits[ordinal];
    const size_t row0 = st.row_start[ordinal];
    float* F = fact + (ordinal / FT) * size_t(FACT_ROWS) * FT + (ordinal % FT);
    auto put = [&](int row, float v) { F[size_t(row) * FT] = v; };

    const AnchorIn in = load_anchor(anchor, scaling, cam, a);
    float x[XDIM], bankw[3];
    build_input(sw, p.use_feat_bank != 0, anchor_feat + a * FEAT, in, x, bankw);
#pragma unroll
    for (int i = 0; i < XDIM; ++i) put(F_X + i, x[i]);
    // forward variant 2 left relu(layer 1) and the second-layer outputs of every visible ordinal: read them back instead
    // of redoing 3 x 32 x 36 + ~2600 FMAs per anchor
    const float* hid = st.hidden + ordinal * size_t(3 * FEAT);
    const float* o2 = st.out2 + ordinal * size_t(O2_W);
    auto hidden_of = [&](int mlp, float (&h)[FEAT]) {
        if (cached) {
#pragma unroll
            for (int q = 0; q < FEAT / 4; ++q) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(hid + mlp * FEAT) + q);
                h[4 * q] = t.x; h[4 * q + 1] = t.y; h[4 * q + 2] = t.z; h[4 * q + 3] = t.w;
            }
            // ReLU'(0): a pre-activation within the tensor-core split's error of zero (~3e-6; a few units in 1e5 are
            // inside this band) is redone as the FP32 FMA chain, so the derivative mask is the one an FP32 forward takes
            uint32_t amb = 0;
#pragma unroll
            for (int j = 0; j < FEAT; ++j) if (fabsf(h[j]) < 2e-5f) amb |= 1u << j;
            while (amb != 0u) {                              // rare: one iteration per ambiguous unit
                const int j = __ffs(amb) - 1;
                amb &= amb - 1u;
                float acc = sw.b1[mlp][j];
                const float4* wrow = reinterpret_cast<const float4*>(sw.w1[mlp][j]);
#pragma unroll
                for (int q = 0; q < XDIM / 4; ++q) {
                    const float4 w4 = wrow[q];
                    acc = fmaf(w4.x, x[4 * q], acc); acc = fmaf(w4.y, x[4 * q + 1], acc);
                    acc = fmaf(w4.z, x[4 * q + 2], acc); acc = fmaf(w4.w, x[4 * q + 3], acc);
                }
#pragma unroll
                for (int jj = 0; jj < FEAT; ++jj) h[jj] = jj == j ? acc : h[jj];
            }
#pragma unroll
            for (int j = 0; j < FEAT; ++j) h[j] = fmaxf(h[j], 0.f);
        } else {
            layer1(sw.w1[mlp], sw.b1[mlp], x, h);
        }
    };

    float dx[XDIM];
#pragma unroll
    for (int i = 0; i < XDIM; ++i) dx[i] = 0.f;
    float ds[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dax = 0.f, day = 0.f, daz = 0.f;

    // back-propagate one MLP's hidden gradient through ReLU and its first layer
    auto finish_mlp = [&](int mlp, const float (&h)[FEAT], float (&dh)[FEAT]) {
#pragma unroll
        for (int j = 0; j < FEAT; ++j) {
            const float dpre = h[j] > 0.f ? dh[j] : 0.f;
            put(F_H + mlp * FEAT + j, h[j]);
            put(F_DPRE + mlp * FEAT + j, dpre);
            const float4* row = reinterpret_cast<const float4*>(sw.w1[mlp][j]);
#pragma unroll
            for (int q = 0; q < XDIM / 4; ++q) {
                const float4 w4 = row[q];
                dx[4 * q] = fmaf(w4.x, dpre, dx[4 * q]);
                dx[4 * q + 1] = fmaf(w4.y, dpre, dx[4 * q + 1]);
                dx[4 * q + 2] = fmaf(w4.z, dpre, dx[4 * q + 2]);
                dx[4 * q + 3] = fmaf(w4.w, dpre, dx[4 * q + 3]);
            }
        }
    };

    // ---- the three MLPs, one after the other in ONE loop body: the code they share (activations, ReLU backward, the
    //      first layer's transposed product) exists once in the instruction stream ----
#pragma unroll 1
    for (int mlp = 0; mlp < 3; ++mlp) {
        float h[FEAT], dh[FEAT];
        hidden_of(mlp, h);
#pragma unroll
        for (int j = 0; j < FEAT; ++j) dh[j] = 0.f;
        if (mlp == 0) {
            // opacity MLP
            size_t r = row0;
#pragma unroll 1
            for (int o = 0; o < NOFF; ++o) {
                const float t = cached ? __ldg(o2 + o) : tanhf(dot32(sw.w2o[o], sw.b2o[o], h));
                float g = g_nop ? __ldg(g_nop + ordinal * NOFF + o) : 0.f;
                if ((m >> o) & 1u) { g += __ldg(g_opacity + r); ++r; }
                const float dz = g * (1.f - t * t);
                put(F_D2O + o, dz);
                axpy32(sw.w2o[o], dz, dh);
            }
        } else if (mlp == 1) {
            // covariance MLP + geometry assembly
            size_t r = row0;
#pragma unroll 1
            for (int o = 0; o < NOFF; ++o) {
                float* dof = d_offset + (a * NOFF + o) * 3;
                if (!((m >> o) & 1u)) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) put(F_D2S + 7 * o + k, 0.f);
                    if (!acc_mode) { dof[0] = 0.f; dof[1] = 0.f; dof[2] = 0.f; }
                    continue;
                }
                float sr[7], dsr[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) sr[k] = cached ? __ldg(o2 + NOFF + 7 * o + k) : dot32(sw.w2s[7 * o + k], sw.b2s[7 * o + k], h);
                const float gx = __ldg(g_xyz + 3 * r), gy = __ldg(g_xyz + 3 * r + 1), gz = __ldg(g_xyz + 3 * r + 2);
                const float ox = __ldg(offset + (a * NOFF + o) * 3), oy = __ldg(offset + (a * NOFF + o) * 3 + 1),
                            oz = __ldg(offset + (a * NOFF + o) * 3 + 2);
                // xyz = anchor + offset * s[:3]
                dax += gx; day += gy; daz += gz;
                if (acc_mode) { accum(dof, gx * in.s[0]); accum(dof + 1, gy * in.s[1]); accum(dof + 2, gz * in.s[2]); }
                else { dof[0] = gx * in.s[0]; dof[1] = gy * in.s[1]; dof[2] = gz * in.s[2]; }
                ds[0] = fmaf(gx, ox, ds[0]); ds[1] = fmaf(gy, oy, ds[1]); ds[2] = fmaf(gz, oz, ds[2]);
                // scaling = s[3:] * sigmoid(sr[:3])
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float sg = sigmoidf_(sr[k]);
                    const float g = __ldg(g_scaling + 3 * r + k);
                    ds[3 + k] = fmaf(g, sg, ds[3 + k]);
                    dsr[k] = g * in.s[3 + k] * sg * (1.f - sg);
                }
                // rot = v / max(|v|, 1e-12)
                {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(g_rot + 4 * r));
                    const float n2 = sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6];
                    const float nrm = sqrtf(n2);
                    if (nrm > 1e-12f) {
                        const float inv = 1.f / nrm;
                        const float u0 = sr[3] * inv, u1 = sr[4] * inv, u2 = sr[5] * inv, u3 = sr[6] * inv;
                        const float ug = u0 * g.x + u1 * g.y + u2 * g.z + u3 * g.w;
                        dsr[3] = (g.x - u0 * ug) * inv; dsr[4] = (g.y - u1 * ug) * inv;
                        dsr[5] = (g.z - u2 * ug) * inv; dsr[6] = (g.w - u3 * ug) * inv;
                    } else {
                        dsr[3] = g.x * 1e12f; dsr[4] = g.y * 1e12f; dsr[5] = g.z * 1e12f; dsr[6] = g.w * 1e12f;
                    }
                }
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    put(F_D2S + 7 * o + k, dsr[k]);
                    axpy32(sw.w2s[7 * o + k], dsr[k], dh);
                }
                ++r;
            }
        } else {
            // colour MLP
            size_t r = row0;
#pragma unroll 1
            for (int o = 0; o < NOFF; ++o) {
                if (!((m >> o) & 1u)) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) put(F_D2C + 3 * o + k, 0.f);
                    continue;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float c = sigmoidf_(cached ? __ldg(o2 + 8 * NOFF + 3 * o + k) : dot32(sw.w2c[3 * o + k], sw.b2c[3 * o + k], h));
                    const float dz = __ldg(g_color + 3 * r + k) * c * (1.f - c);
                    put(F_D2C + 3 * o + k, dz);
                    axpy32(sw.w2c[3 * o + k], dz, dh);
                }
                ++r;
            }
        }
        finish_mlp(mlp, h, dh);
    }

    // ---- inputs: dx = [d feat'(32), d view(3), d dist] ----
    const float ux = x[32], uy = x[33], uz = x[34];
    float dux = dx[32], duy = dx[33], duz = dx[34], ddist = dx[35];
    float df[FEAT];
    if (p.use_feat_bank) {
        float f[FEAT];
#pragma unroll
        for (int q = 0; q < FEAT / 4; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(anchor_feat + a * FEAT) + q);
            f[4 * q] = t.x; f[4 * q + 1] = t.y; f[4 * q + 2] = t.z; f[4 * q + 3] = t.w;
        }
        float dw0 = 0.f, dw1 = 0.f, dw2 = 0.f;
#pragma unroll
        for (int j = 0; j < FEAT; ++j) df[j] = 0.f;
#pragma unroll
        for (int j = 0; j < FEAT; ++j) {
            dw0 = fmaf(dx[j], f[4 * (j & 7)], dw0);
            dw1 = fmaf(dx[j], f[2 * (j & 15)], dw1);
            dw2 = fmaf(dx[j], f[j], dw2);
            df[4 * (j & 7)] = fmaf(dx[j], bankw[0], df[4 * (j & 7)]);
            df[2 * (j & 15)] = fmaf(dx[j], bankw[1], df[2 * (j & 15)]);
            df[j] = fmaf(dx[j], bankw[2], df[j]);
        }
        // softmax backward
        const float dot = bankw[0] * dw0 + bankw[1] * dw1 + bankw[2] * dw2;
        const float dl[3] = {bankw[0] * (dw0 - dot), bankw[1] * (dw1 - dot), bankw[2] * (dw2 - dot)};
        put(F_DLOG + 0, dl[0]); put(F_DLOG + 1, dl[1]); put(F_DLOG + 2, dl[2]);
        put(F_CAT + 0, ux); put(F_CAT + 1, uy); put(F_CAT + 2, uz); put(F_CAT + 3, in.dist);
        float dc0 = 0.f, dc1 = 0.f, dc2 = 0.f, dc3 = 0.f;
#pragma unroll
        for (int j = 0; j < FEAT; ++j) {
            const float4 w4 = *reinterpret_cast<const float4*>(sw.wb1[j]);
            float hb = sw.bb1[j];
            hb = fmaf(w4.x, ux, hb); hb = fmaf(w4.y, uy, hb); hb = fmaf(w4.z, uz, hb); hb = fmaf(w4.w, in.dist, hb);
            hb = fmaxf(hb, 0.f);
            const float dhb = sw.wb2[0][j] * dl[0] + sw.wb2[1][j] * dl[1] + sw.wb2[2][j] * dl[2];
            const float dpre = hb > 0.f ? dhb : 0.f;
            put(F_HB + j, hb);
            put(F_DPREB + j, dpre);
            dc0 = fmaf(w4.x, dpre, dc0); dc1 = fmaf(w4.y, dpre, dc1); dc2 = fmaf(w4.z, dpre, dc2); dc3 = fmaf(w4.w, dpre, dc3);
        }
        dux += dc0; duy += dc1; duz += dc2; ddist += dc3;
    } else {
#pragma unroll
        for (int j = 0; j < FEAT; ++j) df[j] = dx[j];
    }
#pragma unroll
    for (int q = 0; q < FEAT / 4; ++q) {
        float4* dst = reinterpret_cast<float4*>(d_feat + a * FEAT) + q;
        float4 v = make_float4(df[4 * q], df[4 * q + 1], df[4 * q + 2], df[4 * q + 3]);
        if (acc_mode) {
            float* d1 = reinterpret_cast<float*>(dst);
            atomicAdd(d1, v.x); atomicAdd(d1 + 1, v.y); atomicAdd(d1 + 2, v.z); atomicAdd(d1 + 3, v.w);
            continue;
        }
        *dst = v;
    }

    // ob_view = v / |v|, ob_dist = |v|, v = anchor - camera
    {
        const float inv = 1.f / in.dist;
        const float ud = ux * dux + uy * duy + uz * duz;
        dax += (dux - ux * ud) * inv + ddist * ux;
        day += (duy - uy * ud) * inv + ddist * uy;
        daz += (duz - uz * ud) * inv + ddist * uz;
    }
    // SEGS_DECODE_LOG_SCALING: the caller's parameter is _scaling = log(scaling) (GaussianModel::get_scaling,
    // gaussian_model.cpp:186-189), so chain through the exp: d/d_log = d/d_scaling * scaling
    if (flags & SEGS_DECODE_LOG_SCALING) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ds[k] *= in.s[k];
    }
    if (acc_mode) {
        accum(d_anchor + 3 * a, dax); accum(d_anchor + 3 * a + 1, day); accum(d_anchor + 3 * a + 2, daz);
#pragma unroll
        for (int k = 0; k < 6; ++k) accum(d_scaling + 6 * a + k, ds[k]);
    } else {
        d_anchor[3 * a] = dax; d_anchor[3 * a + 1] = day; d_anchor[3 * a + 2] = daz;
#pragma unroll
        for (int k = 0; k < 6; ++k) d_scaling[6 * a + k] = ds[k];
    }
  }
}

// =======================================================================================
// backward, kernel 2: weight gradients  dW[p][q] = sum_k U[k][p] V[k][q]
// =======================================================================================
constexpr int WG_THREADS = 256;
constexpr int WG_STRIDE = FT + 4;      // floats per factor row in shared memory (conflict-free LDS.128)

struct WJob {            // one outer-product sum over the anchors k:
    int u0, np, v0;      //   out[p][q] += U[k][u0 + p] * V[k][v0 + q]   (p < np rows, q < nq lanes)
    float* out;          //   v0 < 0: V == 1 (bias sums): lanes index the U rows, out[p*32 + lane]
    int ld, nq;          //   tr != 0: transposed store, out[q * ld + p]
    int tr, warp;        //   warp that owns the job (its rows share the V operand, cached in registers)
};
constexpr int MAX_JOBS = 24;
struct WJobs { WJob j[MAX_JOBS]; int n; };
constexpr int WG_ROWS_MAX = 36;        // rows (accumulators) per warp

__global__ void __launch_bounds__(WG_THREADS, 2)
decode_wgrad_kernel(const float* __restrict__ fact, int n_vis, const WJobs jobs, const int nrows_used)
{
    extern __shared__ __align__(16) float s_f[];          // [FACT_ROWS][WG_STRIDE]
    __shared__ short s_job[WG_THREADS / 32][WG_ROWS_MAX], s_p[WG_THREADS / 32][WG_ROWS_MAX];
    __shared__ short s_u[WG_THREADS / 32][WG_ROWS_MAX], s_v[WG_THREADS / 32][WG_ROWS_MAX];
    __shared__ int s_nrows[WG_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntiles = (n_vis + FT - 1) / FT;

    // row table of this warp: its jobs in list order, one entry per accumulator
    if (lane == 0) {
        int r = 0;
        for (int j = 0; j < jobs.n; ++j) {
            if (jobs.j[j].warp != warp) continue;
            const int rows = jobs.j[j].v0 >= 0 ? jobs.j[j].np : (jobs.j[j].np + 31) / 32;
            for (int pp = 0; pp < rows && r < WG_ROWS_MAX; ++pp, ++r) {
                s_job[warp][r] = (short)j;
                s_p[warp][r] = (short)pp;
                s_u[warp][r] = (short)(jobs.j[j].v0 >= 0 ? jobs.j[j].u0 + pp : jobs.j[j].u0 + min(pp * 32 + lane, jobs.j[j].np - 1));
                s_v[warp][r] = (short)(jobs.j[j].v0 >= 0 ? jobs.j[j].v0 : -1 - j);     // distinct negative tag per bias job
            }
        }
        s_nrows[warp] = r;
    }
    __syncthreads();
    const int nrows = s_nrows[warp];
    float acc[WG_ROWS_MAX];
#pragma unroll
    for (int i = 0; i < WG_ROWS_MAX; ++i) acc[i] = 0.f;

    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int nk = min(FT, n_vis - t * FT);
        __syncthreads();
        const float* src = fact + size_t(t) * FACT_ROWS * FT;
        // 16-byte asynchronous copies, one factor row = 16 of them (rows land WG_STRIDE apart)
        for (int e = threadIdx.x; e < nrows_used * (FT / 4); e += WG_THREADS) {
            const int row = e / (FT / 4), k4 = e % (FT / 4);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_f + row * WG_STRIDE + 4 * k4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + row * FT + 4 * k4));
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        if (nk < FT) {
            // tail tile: the unused ordinals hold whatever the scratch held — zero them
            for (int e = threadIdx.x; e < nrows_used * FT; e += WG_THREADS) {
                const int row = e / FT, k = e % FT;
                if (k >= nk) s_f[row * WG_STRIDE + k] = 0.f;
            }
            __syncthreads();
        }
        int cur_v = -1000;
        float4 V[FT / 4];
#pragma unroll
        for (int i = 0; i < WG_ROWS_MAX; ++i) {
            if (i >= nrows) break;
            const int vtag = s_v[warp][i];            // warp-uniform
            if (vtag >= 0) {
                if (vtag != cur_v) {                   // new V operand: this lane's column, all 64 anchors, into registers
                    const int nq = jobs.j[s_job[warp][i]].nq;
                    const float4* v = reinterpret_cast<const float4*>(s_f + (vtag + (lane < nq ? lane : 0)) * WG_STRIDE);
#pragma unroll
                    for (int k4 = 0; k4 < FT / 4; ++k4) V[k4] = v[k4];
                    cur_v = vtag;
                }
                const float4* u = reinterpret_cast<const float4*>(s_f + s_u[warp][i] * WG_STRIDE);   // broadcast
                float a = acc[i];
#pragma unroll
                for (int k4 = 0; k4 < FT / 4; ++k4) {
                    const float4 uu = u[k4];
                    a = fmaf(uu.x, V[k4].x, a); a = fmaf(uu.y, V[k4].y, a); a = fmaf(uu.z, V[k4].z, a); a = fmaf(uu.w, V[k4].w, a);
                }
                acc[i] = a;
            } else {
                // bias sums: lane <-> U row
                const WJob& job = jobs.j[s_job[warp][i]];
                const int r = s_p[warp][i] * 32 + lane;
                if (r < job.np) {
                    const float4* u = reinterpret_cast<const float4*>(s_f + (job.u0 + r) * WG_STRIDE);
                    float a = acc[i];
#pragma unroll
                    for (int k4 = 0; k4 < FT / 4; ++k4) {
                        const float4 uu = u[k4];
                        a += uu.x + uu.y + uu.z + uu.w;
                    }
                    acc[i] = a;
                }
            }
        }
    }
    // flush: one atomic per accumulator
#pragma unroll
    for (int i = 0; i < WG_ROWS_MAX; ++i) {
        if (i >= nrows) break;
        const WJob& job = jobs.j[s_job[warp][i]];
        const int pidx = s_p[warp][i];
        if (job.v0 >= 0) {
            if (lane < job.nq) atomicAdd(job.tr ? job.out + lane * job.ld + pidx : job.out + pidx * job.ld + lane, acc[i]);
        } else {
            const int r = pidx * 32 + lane;
            if (r < job.np) atomicAdd(job.out + r, acc[i]);
        }
    }
}

// =======================================================================================
// backward, kernel 2, variant 2 (round 2): the weight gradients on the tensor cores
// =======================================================================================
// dW[p][q] = sum over anchors k of U[k][p] V[k][q] is a contraction over the ANCHOR index, and kernel 1 already stores
// every factor feature-major ([row][64 ordinals]) — K-major operands as they lie.  Per stage of 32 ordinals the rows
// are loaded once (coalesced 16-byte loads, 100 % sector use, two stages in flight in registers), split hi / lo (3xTF32)
// and stored as the no-swizzle core-matrix tiles of two products that accumulate in TMEM over ALL stages of the CTA:
//   D1[128 x 48]  = [dpre_opacity | dpre_cov | dpre_colour (96) | dpre_bank (32)]  x  [x (36) | 1 | view,dist (4)]
//                   rows < 96: dW1 of the three MLPs and (the 1-column) db1; rows 96..127: feature-bank dW1 / db1
//   D2[128 x 144] = [d2_opacity (10) | d2_scale_rot (70) | d2_colour (30) | dlogit (3)]  x  [h_opacity | h_cov | h_colour | 1 | h_bank]
//                   rows < 110: the diagonal blocks are dW2 of the three MLPs, the 1-column db2; rows 110..112: bank dW2 / db2
// (a row of ones in the B operand turns the bias sums into one more output column; products nobody needs are computed
// and ignored — the tensor pipe is far from busy).  The operand tiles are double-buffered (2 x 112 KB of shared memory):
// the split + store of stage s+1 does not wait for the MMAs of stage s.
// 24 tcgen05.mma per stage replace ~2300 FFMA + 580 LDS.128 per thread.  One flush per CTA: tcgen05.ld + one atomic
// per needed element, as in variant 1.
namespace wg2 {

constexpr int THREADS = 512;
constexpr int KS = 32;                          // ordinals per stage = MMA K per stage (4 instructions of K = 8)
constexpr int KBS = KS / 4;                     // 16-byte K-blocks per operand row
constexpr uint32_t SBO = KBS * tc::CORE;        // 1024
constexpr int N1 = 48, N2 = 144;
constexpr int A_T = 128 * KS * 4;               // bytes of a 128-row operand tile (hi or lo)
constexpr int OFF_A1 = 0;
constexpr int OFF_B1 = OFF_A1 + A_T;
constexpr int OFF_A2 = OFF_B1 + N1 * KS * 4;
constexpr int OFF_B2 = OFF_A2 + A_T;
constexpr int HALF = OFF_B2 + N2 * KS * 4;      // 57344: the lo tiles follow the hi tiles
constexpr int BUF = 2 * HALF;                   // one operand buffer (hi + lo)
constexpr int SMEM = 2 * BUF;                   // two buffers
constexpr int COL_D1 = 0, COL_D2 = 64;
constexpr int TMEM_COLS = 256;
// tile rows: B1 = x (0..35) | ones (36) | view,dist (37..40);  A1 = dpre (0..95) | dpre_bank (96..127)
//            B2 = h (0..95) | ones (96) | h_bank (97..128);    A2 = d2 (0..109) | dlogit (110..112)
constexpr int ONES1 = XDIM, CAT1 = XDIM + 1, ONES2 = 3 * FEAT, HB2 = 3 * FEAT + 1, BANK_A1 = 3 * FEAT, LOG_A2 = O2_W;
constexpr int MAX_LD = ((FACT_ROWS + 7) / 8 * 8 * KBS + THREADS - 1) / THREADS;   // 16-byte loads per thread and stage (7)
static_assert(HALF % 1024 == 0 && SMEM + 1024 <= 227 * 1024, "decode wgrad v2: shared memory");
static_assert(CAT1 + 4 <= N1 && HB2 + FEAT <= N2 && LOG_A2 + 3 <= 128, "decode wgrad v2: tile rows");

__device__ __forceinline__ uint32_t canon(int row, int k) {
    return uint32_t(((row >> 3) * KBS + (k >> 2)) * tc::CORE + (row & 7) * 16 + (k & 3) * 4);
}
// factor row -> byte offset of its K-block 0 inside the hi half of a buffer (tile base + row placement)
__device__ __forceinline__ uint32_t place(int r) {
    int base, trow;
    if (r < F_H) { base = OFF_B1; trow = r - F_X; }
    else if (r < F_DPRE) { base = OFF_B2; trow = r - F_H; }
    else if (r < F_D2O) { base = OFF_A1; trow = r - F_DPRE; }
    else if (r < F_HB) { base = OFF_A2; trow = r - F_D2O; }
    else if (r < F_DPREB) { base = OFF_B2; trow = HB2 + (r - F_HB); }
    else if (r < F_DLOG) { base = OFF_A1; trow = BANK_A1 + (r - F_DPREB); }
    else if (r < F_CAT) { base = OFF_A2; trow = LOG_A2 + (r - F_DLOG); }
    else { base = OFF_B1; trow = CAT1 + (r - F_CAT); }
    return uint32_t(base) + uint32_t(((trow >> 3) * KBS) * tc::CORE + (trow & 7) * 16);
}

}  // namespace wg2

struct WGradOut {
    float* w1[3]; float* b1[3]; int ld1[3]; int in1[3];     // first layers: dW1[j * ld + i], i < in
    float* w2[3]; float* b2[3];                             // second layers: opacity [10,32], cov [70,32], colour [30,32]
    float* bank_w1; float* bank_b1; float* bank_w2; float* bank_b2;   // NULL without the feature bank
};

__global__ void __launch_bounds__(wg2::THREADS, 1)
decode_wgrad_tc_kernel(const float* __restrict__ fact, int n_vis, const WGradOut out, const int nrows_used)
{
    extern __shared__ __align__(1024) unsigned char s_op[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nstages = (n_vis + wg2::KS - 1) / wg2::KS;
    const bool bank = out.bank_w1 != nullptr;

    if (warp == 0) tc::tmem_alloc(&s_tmem, wg2::TMEM_COLS);
    if (tid == 0) { tc::mbar_init(&s_bar[0], 1); tc::mbar_init(&s_bar[1], 1); }
    // padding rows stay zero for the whole kernel; the rows of ones are written once (hi halves of both buffers)
    for (int e = tid; e < wg2::SMEM / 16; e += wg2::THREADS) reinterpret_cast<float4*>(s_op)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (tid < 4 * wg2::KS) {
        const int which = tid / wg2::KS, k = tid % wg2::KS;
        const int base = (which >> 1) * wg2::BUF + ((which & 1) ? wg2::OFF_B2 : wg2::OFF_B1);
        *reinterpret_cast<float*>(s_op + base + wg2::canon((which & 1) ? wg2::ONES2 : wg2::ONES1, k)) = 1.0f;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;

    // this thread's 16-byte pieces of a stage: piece e -> factor row (e >> 6) * 8 + (e & 7), K-block ((e >> 3) & 3) + 4 * ((e >> 5) & 1)
    // (8 rows x 4 K-blocks per warp instruction: 64 contiguous bytes per row from global memory, 128 contiguous bytes per
    // quarter-warp in shared memory).  Where a piece goes and where it comes from is the same for every stage.
    const int npieces = (nrows_used + 7) / 8 * 8 * wg2::KBS;
    uint32_t dst_off[wg2::MAX_LD], src_off[wg2::MAX_LD];
#pragma unroll
    for (int it = 0; it < wg2::MAX_LD; ++it) {
        const int e = tid + it * wg2::THREADS;
        const int r3 = e & 7, c2 = (e >> 3) & 3, rest = e >> 5;
        const int kb = c2 + 4 * (rest & 1), row = ((rest >> 1) << 3) + r3;
        const bool ok = e < npieces && row < nrows_used;
        dst_off[it] = ok ? wg2::place(row) + kb * tc::CORE : 0xFFFFFFFFu;
        src_off[it] = ok ? uint32_t(row * FT + 4 * kb) : 0u;
    }
    float4 buf0[wg2::MAX_LD], buf1[wg2::MAX_LD];      // two stages of loads in flight
    auto load_stage = [&](float4 (&buf)[wg2::MAX_LD], int s) {
        const float* src = fact + size_t(s >> 1) * FACT_ROWS * FT + (s & 1) * wg2::KS;
        const int nk = min(wg2::KS, n_vis - s * wg2::KS);
#pragma unroll
        for (int it = 0; it < wg2::MAX_LD; ++it) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dst_off[it] != 0xFFFFFFFFu) {
                v = __ldg(reinterpret_cast<const float4*>(src + src_off[it]));
                if (nk < wg2::KS) {
                    // tail stage: the ordinals past n_vis hold whatever the scratch held
                    const int k0 = int(src_off[it]) & (FT - 1);
                    if (k0 + 0 >= nk) v.x = 0.f;
                    if (k0 + 1 >= nk) v.y = 0.f;
                    if (k0 + 2 >= nk) v.z = 0.f;
                    if (k0 + 3 >= nk) v.w = 0.f;
                }
            }
            buf[it] = v;
        }
    };
    auto store_stage = [&](const float4 (&buf)[wg2::MAX_LD], unsigned char* dst) {
#pragma unroll
        for (int it = 0; it < wg2::MAX_LD; ++it) {
            if (dst_off[it] != 0xFFFFFFFFu) {
                const float4 v = buf[it];
                const tc::Split4 sp = tc::split4(v.x, v.y, v.z, v.w);
                *reinterpret_cast<float4*>(dst + dst_off[it]) = sp.hi;
                *reinterpret_cast<float4*>(dst + wg2::HALF + dst_off[it]) = sp.lo;
            }
        }
    };

    const uint64_t d_base = tc2::make_desc(tc::smem_u32(s_op), wg2::SBO);
    uint32_t ph0 = 0, ph1 = 0;
    const int G = gridDim.x;
    int nth = 0;                                          // stages this CTA has issued
    // one stage: (the MMAs that read this operand buffer two stages ago have finished — checked, rarely waited for)
    // split + store the rows, refill the register buffer with the stage two ahead, hand the tiles to the tensor core
    auto stage = [&](float4 (&buf)[wg2::MAX_LD], int s, const int b) {
        if (nth >= 2) {
            if (b == 0) { tc::mbar_wait(&s_bar[0], ph0); ph0 ^= 1u; } else { tc::mbar_wait(&s_bar[1], ph1); ph1 ^= 1u; }
        }
        store_stage(buf, s_op + b * wg2::BUF);
        if (s + 2 * G < nstages) load_stage(buf, s + 2 * G);
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t acc0 = nth > 0 ? 1u : 0u;
            const uint64_t d_hi = d_base + uint64_t((b * wg2::BUF) >> 4), d_lo = d_hi + uint64_t(wg2::HALF >> 4);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint32_t a = g == 0 ? wg2::OFF_A1 : wg2::OFF_A2;
                const uint32_t bb = g == 0 ? wg2::OFF_B1 : wg2::OFF_B2;
                const uint32_t d = tmem + (g == 0 ? wg2::COL_D1 : wg2::COL_D2);
                const uint32_t id = g == 0 ? tc2::idesc(wg2::N1) : tc2::idesc(wg2::N2);
#pragma unroll
                for (int j = 0; j < wg2::KS / 8; ++j) {
                    const uint64_t ka = (a + 2 * j * tc::CORE) >> 4, kb = (bb + 2 * j * tc::CORE) >> 4;
                    tc2::umma_tf32(d, d_hi + ka, d_hi + kb, id, j > 0 ? 1u : acc0);
                    tc2::umma_tf32(d, d_hi + ka, d_lo + kb, id, 1u);
                    tc2::umma_tf32(d, d_lo + ka, d_hi + kb, id, 1u);
                }
            }
            tc::umma_commit(&s_bar[b]);
        }
        ++nth;
    };
    {
        const int s0 = blockIdx.x;
        if (s0 < nstages) load_stage(buf0, s0);
        if (s0 + G < nstages) load_stage(buf1, s0 + G);
        for (int s = s0; s < nstages; s += 2 * G) {
            stage(buf0, s, 0);
            if (s + G < nstages) stage(buf1, s + G, 1);
        }
    }
    if (nth == 0) {                                   // no stage for this CTA (grid <= stages, so this does not happen)
        tc::fence_before_sync();
        __syncthreads();
        if (warp == 0) tc::tmem_dealloc(tmem, wg2::TMEM_COLS);
        return;
    }
    // the commit of the last stage covers every MMA issued before it
    if ((nth - 1) & 1) tc::mbar_wait(&s_bar[1], ph1); else tc::mbar_wait(&s_bar[0], ph0);
    tc::fence_after_sync();

    // ---- flush: thread = accumulator row (TMEM lane); one atomic per needed element ----
    const int lq = warp & 3, part = warp >> 2;       // 16 warps: D1 | D2 | - | -
    const int row = lq * 32 + lane;
    const uint32_t lane_base = tmem + (uint32_t(lq * 32) << 16);
    float v[32];
    if (part == 0) {
        // D1 rows < 96: row = 32 * mlp + hidden unit j, columns = input column i (36: the bias sum);
        //    rows 96..127: bank hidden unit j = row - 96, columns 37..40 = [view, dist], 36 = the bias sum
        const int m = min(row >> 5, 2), j = row & 31;
        tc::tmem_ld32(lane_base + wg2::COL_D1, v);
        if (row < 3 * FEAT) {
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(out.w1[m] + j * out.ld1[m] + i, v[i]);
        }
        tc::tmem_ld32(lane_base + wg2::COL_D1 + 32, v);       // columns 32..47 are D1's, the rest is not used
        if (row < 3 * FEAT) {
#pragma unroll
            for (int i = 32; i < XDIM; ++i) if (i < out.in1[m]) atomicAdd(out.w1[m] + j * out.ld1[m] + i, v[i - 32]);
            atomicAdd(out.b1[m] + j, v[wg2::ONES1 - 32]);
        } else if (bank) {
#pragma unroll
            for (int i = 0; i < 4; ++i) atomicAdd(out.bank_w1 + j * 4 + i, v[wg2::CAT1 - 32 + i]);
            atomicAdd(out.bank_b1 + j, v[wg2::ONES1 - 32]);
        }
    } else if (part == 1) {
        // D2 rows < 110 = second-layer output units (opacity 0..9, scale_rot 10..79, colour 80..109): the block of 32
        //    columns of the same MLP's hidden units is the weight gradient, column 96 the bias sum;
        //    rows 110..112 = bank logit m: columns 97..128 = bank hidden units, column 96 the bias sum
        const int m = row < NOFF ? 0 : row < NOFF + 7 * NOFF ? 1 : 2;
        const int n = m == 0 ? row : m == 1 ? row - NOFF : row - NOFF - 7 * NOFF;
        const bool used = row < O2_W;
        const bool logit = bank && row >= wg2::LOG_A2 && row < wg2::LOG_A2 + 3;
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
            tc::tmem_ld32(lane_base + wg2::COL_D2 + 32 * blk, v);
            if (used && blk == m) {
#pragma unroll
                for (int i = 0; i < 32; ++i) atomicAdd(out.w2[m] + n * FEAT + i, v[i]);
            }
        }
        tc::tmem_ld32(lane_base + wg2::COL_D2 + wg2::ONES2, v);        // columns 96..127: bias sum, bank hidden 0..30
        if (used) atomicAdd(out.b2[m] + n, v[0]);
        if (logit) {
            atomicAdd(out.bank_b2 + (row - wg2::LOG_A2), v[0]);
#pragma unroll
            for (int i = 0; i < FEAT - 1; ++i) atomicAdd(out.bank_w2 + (row - wg2::LOG_A2) * FEAT + i, v[1 + i]);
        }
        tc::tmem_ld32(lane_base + wg2::COL_D2 + wg2::ONES2 + 32, v);   // column 128: bank hidden 31
        if (logit) atomicAdd(out.bank_w2 + (row - wg2::LOG_A2) * FEAT + FEAT - 1, v[0]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, wg2::TMEM_COLS);
}

// appearance path: the appearance vector is the same for all anchors, so its gradients follow from
// db1_colour:  d_app = W1c[:, app cols]^T db1c ;  dW1c[:, app cols] = db1c (x) app ;
// d app_w = d_app (x) pose ; d app_b = d_app
__global__ void __launch_bounds__(32)
decode_appgrad_kernel(const segs_decode_params p, const segs_decode_grads d, const Pose7 pose)
{
    __shared__ float s_app[32], s_dapp[32];
    const int t = threadIdx.x;
    const int in_c = 35 + (p.add_color_dist ? 1 : 0), ld_c = in_c + p.appearance_dim;
    float a = 0.f;
    if (t < p.appearance_dim) {
        a = p.app_b[t];
        for (int q = 0; q < 7; ++q) a = fmaf(p.app_w[t * 7 + q], pose.v[q], a);
    }
    s_app[t] = a;
    float da = 0.f;
    if (t < p.appearance_dim)
        for (int j = 0; j < FEAT; ++j) da = fmaf(p.color_w1[j * ld_c + in_c + t], d.color_b1[j], da);
    s_dapp[t] = da;
    __syncwarp();
    for (int k = 0; k < p.appearance_dim; ++k) d.color_w1[t * ld_c + in_c + k] = d.color_b1[t] * s_app[k];
    if (t < p.appearance_dim) {
        d.app_b[t] = da;
        for (int q = 0; q < 7; ++q) d.app_w[t * 7 + q] = da * pose.v[q];
    }
}

// zero up to 24 arrays in ONE launch (22 separate memsets cost more in launch gaps than in bytes)
struct ZeroList { float* p[24]; unsigned long long n[24]; int count; };
__global__ void __launch_bounds__(256)
zero_many_kernel(const ZeroList z)
{
    for (int a = 0; a < z.count; ++a) {
        float* ptr = z.p[a];
        const unsigned long long n = z.n[a];
        const unsigned long long n4 = ((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0) ? n / 4 : 0;
        float4* p4 = reinterpret_cast<float4*>(ptr);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (unsigned long long i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) p4[i] = z4;
        for (unsigned long long i = 4 * n4 + size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) ptr[i] = 0.f;
    }
}

struct HostCounts { uint32_t* host = nullptr; uint32_t* dev = nullptr; cudaEvent_t ev = nullptr; };
HostCounts decode_host_counts()
{
    // the mapped words are portable (one address on every device under UVA); the event is per device (capi.cu)
    static thread_local HostCounts h;
    if (!h.host) {
        void *hp = nullptr, *dp = nullptr;
        if (cudaHostAlloc(&hp, 4 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
            cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
            h.host = static_cast<uint32_t*>(hp);
            h.dev = static_cast<uint32_t*>(dp);
        }
    }
    HostCounts out = h;
    out.ev = readback_event();
    if (!out.ev) out.host = nullptr;
    return out;
}

// which forward kernel runs: 1 = thread-per-anchor kernel with the first layers on tcgen05, 2 = both layers on
// tcgen05 over tiles of visible anchors.  SEGS_DECODE_VARIANT in the environment sets the initial value.
std::atomic<int> g_decode_variant{0};
int decode_variant()
{
    int v = g_decode_variant.load(std::memory_order_relaxed);
    if (v == 0) {
        const char* e = getenv("SEGS_DECODE_VARIANT");
        v = (e && e[0] == '1') ? 1 : 2;
        g_decode_variant.store(v, std::memory_order_relaxed);
    }
    return v;
}

// the weight-gradient kernel follows the forward variant unless SEGS_DECODE_WGRAD (1 | 2) says otherwise (development)
int decode_wgrad_variant()
{
    static const int env = [] { const char* e = getenv("SEGS_DECODE_WGRAD"); return e ? (e[0] == '1' ? 1 : e[0] == '2' ? 2 : 0) : 0; }();
    return env ? env : decode_variant();
}

int check_params(const segs_decode_params* p)
{
    if (!p) { set_error("decode: params must not be NULL"); return SEGS_ERR_INVALID_ARG; }
    if (p->appearance_dim < 0 || p->appearance_dim > 32) { set_error("decode: appearance_dim must be in [0, 32]"); return SEGS_ERR_INVALID_ARG; }
    if (!p->opacity_w1 || !p->opacity_b1 || !p->opacity_w2 || !p->opacity_b2 || !p->cov_w1 || !p->cov_b1 || !p->cov_w2 ||
        !p->cov_b2 || !p->color_w1 || !p->color_b1 || !p->color_w2 || !p->color_b2) {
        set_error("decode: NULL MLP weight"); return SEGS_ERR_INVALID_ARG;
    }
    if (p->appearance_dim > 0 && (!p->app_w || !p->app_b)) { set_error("decode: NULL appearance weight"); return SEGS_ERR_INVALID_ARG; }
    if (p->use_feat_bank && (!p->bank_w1 || !p->bank_b1 || !p->bank_w2 || !p->bank_b2)) { set_error("decode: NULL feature-bank weight"); return SEGS_ERR_INVALID_ARG; }
    return SEGS_OK;
}

}  // namespace
}  // namespace segs

using namespace segs;

extern "C" int segs_decode_set_variant(int variant)
{
    if (variant != 1 && variant != 2) { set_error("decode: variant must be 1 or 2"); return SEGS_ERR_INVALID_ARG; }
    g_decode_variant.store(variant, std::memory_order_relaxed);
    return SEGS_OK;
}

extern "C" int segs_decode_get_variant(void) { return decode_variant(); }

extern "C" size_t segs_decode_state_bytes(int A)
{
    size_t bytes = 0;
    DecodeState::carve(nullptr, A > 0 ? A : 0, &bytes);
    return bytes;
}

extern "C" int segs_decode_forward(
    int A, const unsigned char* visible_mask, const float* anchor, const float* anchor_feat, const float* offset,
    const float* scaling, const float* camera_center, const float* pose, const segs_decode_params* params,
    float* xyz, float* color, float* opacity, float* out_scaling, float* rot, float* neural_opacity,
    unsigned char* mask, char* state, int* counts, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (counts) { counts[0] = 0; counts[1] = 0; }
    if (A == 0) return SEGS_OK;
    int rc = check_params(params);
    if (rc) return rc;
    if (A < 0 || !anchor || !anchor_feat || !offset || !scaling || !camera_center || !pose || !xyz || !color || !opacity ||
        !out_scaling || !rot || !neural_opacity || !mask || !state || !counts) {
        set_error("decode: invalid argument"); return SEGS_ERR_INVALID_ARG;
    }
    const HostCounts hc = decode_host_counts();
    if (!hc.host) { set_error("decode: mapped host memory allocation failed"); return SEGS_ERR_CUDA; }
    DecodeState st = DecodeState::carve(state, A, nullptr);
    SEGS_CUDA_CHECK(cudaMemsetAsync(st.counters, 0, st.zero_bytes, stream));
    Pose7 p7;
    for (int i = 0; i < 7; ++i) p7.v[i] = pose[i];
    const int tiles = (A + DEC_THREADS - 1) / DEC_THREADS;
    if (decode_variant() == 1) {
        SEGS_CUDA_CHECK(cudaFuncSetAttribute(decode_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
        decode_forward_kernel<<<std::min(tiles, SM_COUNT * 2), DEC_THREADS, FWD_SMEM, stream>>>(A, visible_mask, anchor, anchor_feat, offset, scaling,
                                                                camera_center, p7, *params, xyz, color, opacity, out_scaling,
                                                                rot, neural_opacity, mask, st, hc.dev);
        SEGS_LAUNCH_CHECK();
    } else {
        // variant 2: list the visible anchors, then one persistent CTA per SM over tiles of 128 visible anchors
        if (visible_mask != nullptr) {
            decode_compact_kernel<<<(A + CMP_TILE - 1) / CMP_TILE, DEC_THREADS, 0, stream>>>(A, visible_mask, st);
            SEGS_LAUNCH_CHECK();
        }
        SEGS_CUDA_CHECK(cudaFuncSetAttribute(decode_forward_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2::SMEM));
        decode_forward_v2_kernel<<<std::min(tiles, SM_COUNT), tc2::THREADS, tc2::SMEM, stream>>>(
            A, visible_mask == nullptr ? 1 : 0, anchor, anchor_feat, offset, scaling, camera_center, p7, *params, xyz, color,
            opacity, out_scaling, rot, neural_opacity, mask, st, hc.dev);
        SEGS_LAUNCH_CHECK();
    }
    SEGS_CUDA_CHECK(cudaEventRecord(hc.ev, stream));
    SEGS_CUDA_CHECK(cudaEventSynchronize(hc.ev));
    counts[0] = (int)((volatile uint32_t*)hc.host)[0];
    counts[1] = (int)((volatile uint32_t*)hc.host)[1];
    return SEGS_OK;
}

extern "C" int segs_decode_backward_ex(
    int A, const unsigned char* visible_mask, const float* anchor, const float* anchor_feat, const float* offset,
    const float* scaling, const float* camera_center, const float* pose, const segs_decode_params* params,
    const char* state, int n_vis, int n_out, const float* g_xyz, const float* g_color, const float* g_opacity,
    const float* g_scaling, const float* g_rot, const float* g_neural_opacity, float* d_anchor, float* d_anchor_feat,
    float* d_offset, float* d_scaling, const segs_decode_grads* dp, segs_alloc_fn scratch_alloc, void* scratch_user,
    int flags, void* stream_)
{
    (void)visible_mask;
    const bool acc_mode = (flags & SEGS_DECODE_ACCUMULATE) != 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (A == 0) return SEGS_OK;
    int rc = check_params(params);
    if (rc) return rc;
    if (A < 0 || n_vis < 0 || n_vis > A || n_out < 0 || !anchor || !anchor_feat || !offset || !scaling || !camera_center ||
        !pose || !state || !d_anchor || !d_anchor_feat || !d_offset || !d_scaling || !dp || !scratch_alloc) {
        set_error("decode backward: invalid argument"); return SEGS_ERR_INVALID_ARG;
    }
    if (n_out > 0 && (!g_xyz || !g_color || !g_opacity || !g_scaling || !g_rot)) { set_error("decode backward: NULL row gradient"); return SEGS_ERR_INVALID_ARG; }
    const segs_decode_params& p = *params;
    const int in_o = 35 + (p.add_opacity_dist ? 1 : 0), in_s = 35 + (p.add_cov_dist ? 1 : 0);
    const int in_c = 35 + (p.add_color_dist ? 1 : 0), ld_c = in_c + p.appearance_dim;
    if (!dp->opacity_w1 || !dp->opacity_b1 || !dp->opacity_w2 || !dp->opacity_b2 || !dp->cov_w1 || !dp->cov_b1 || !dp->cov_w2 ||
        !dp->cov_b2 || !dp->color_w1 || !dp->color_b1 || !dp->color_w2 || !dp->color_b2 ||
        (p.appearance_dim > 0 && (!dp->app_w || !dp->app_b)) ||
        (p.use_feat_bank && (!dp->bank_w1 || !dp->bank_b1 || !dp->bank_w2 || !dp->bank_b2))) {
        set_error("decode backward: NULL weight-gradient output"); return SEGS_ERR_INVALID_ARG;
    }
    // invisible anchors and the weight accumulators start from zero: one launch for all 22 arrays
    {
        ZeroList z;
        int c = 0;
        auto zero = [&](float* ptr, size_t n) { if (ptr && n) { z.p[c] = ptr; z.n[c] = n; ++c; } };
        if (!acc_mode) {
            zero(d_anchor, size_t(A) * 3); zero(d_anchor_feat, size_t(A) * FEAT);
            zero(d_offset, size_t(A) * NOFF * 3); zero(d_scaling, size_t(A) * 6);
        }
        zero(dp->opacity_w1, size_t(FEAT) * in_o); zero(dp->opacity_b1, FEAT);
        zero(dp->opacity_w2, NOFF * FEAT);         zero(dp->opacity_b2, NOFF);
        zero(dp->cov_w1, size_t(FEAT) * in_s);     zero(dp->cov_b1, FEAT);
        zero(dp->cov_w2, 7 * NOFF * FEAT);         zero(dp->cov_b2, 7 * NOFF);
        zero(dp->color_w1, size_t(FEAT) * ld_c);   zero(dp->color_b1, FEAT);
        zero(dp->color_w2, 3 * NOFF * FEAT);       zero(dp->color_b2, 3 * NOFF);
        if (p.appearance_dim > 0) { zero(dp->app_w, size_t(p.appearance_dim) * 7); zero(dp->app_b, p.appearance_dim); }
        if (p.use_feat_bank) {
            zero(dp->bank_w1, FEAT * 4); zero(dp->bank_b1, FEAT);
            zero(dp->bank_w2, 3 * FEAT); zero(dp->bank_b2, 3);
        }
        z.count = c;
        zero_many_kernel<<<SM_COUNT * 4, 256, 0, stream>>>(z);
        SEGS_LAUNCH_CHECK();
    }
    if (n_vis == 0) return SEGS_OK;

    const size_t ntiles = (size_t(n_vis) + FT - 1) / FT;
    const size_t fact_bytes = ntiles * FACT_ROWS * FT * sizeof(float) + 256;
    char* scratch = scratch_alloc(scratch_user, fact_bytes);
    if (!scratch) { set_error("decode backward: scratch allocation of %zu bytes failed", fact_bytes); return SEGS_ERR_ALLOC; }
    float* fact = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 127) & ~uintptr_t(127));
    if (!p.use_feat_bank) {
        // the bank rows are never written in this mode; the job list below skips them
    }
    DecodeState st = DecodeState::carve(const_cast<char*>(state), A, nullptr);
    Pose7 p7;
    for (int i = 0; i < 7; ++i) p7.v[i] = pose[i];
    {
        // which of the two runs is decided on the device by the flag the forward left in the state; the other returns at once
        const int grid = std::min((n_vis + DEC_THREADS - 1) / DEC_THREADS, SM_COUNT * 4);
        decode_backward_kernel<true><<<grid, DEC_THREADS, 0, stream>>>(
            n_vis, anchor, anchor_feat, offset, scaling, camera_center, p7, p, st, g_xyz, g_color, g_opacity, g_scaling, g_rot,
            g_neural_opacity, d_anchor, d_anchor_feat, d_offset, d_scaling, fact, flags);
        SEGS_LAUNCH_CHECK();
        decode_backward_kernel<false><<<grid, DEC_THREADS, 0, stream>>>(
            n_vis, anchor, anchor_feat, offset, scaling, camera_center, p7, p, st, g_xyz, g_color, g_opacity, g_scaling, g_rot,
            g_neural_opacity, d_anchor, d_anchor_feat, d_offset, d_scaling, fact, flags);
        SEGS_LAUNCH_CHECK();
    }

    if (decode_wgrad_variant() == 2) {
        WGradOut o;
        o.w1[0] = dp->opacity_w1; o.w1[1] = dp->cov_w1; o.w1[2] = dp->color_w1;
        o.b1[0] = dp->opacity_b1; o.b1[1] = dp->cov_b1; o.b1[2] = dp->color_b1;
        o.ld1[0] = in_o; o.ld1[1] = in_s; o.ld1[2] = ld_c;
        o.in1[0] = in_o; o.in1[1] = in_s; o.in1[2] = in_c;         // the appearance columns follow from db1 (appgrad kernel)
        o.w2[0] = dp->opacity_w2; o.w2[1] = dp->cov_w2; o.w2[2] = dp->color_w2;
        o.b2[0] = dp->opacity_b2; o.b2[1] = dp->cov_b2; o.b2[2] = dp->color_b2;
        o.bank_w1 = p.use_feat_bank ? dp->bank_w1 : nullptr; o.bank_b1 = p.use_feat_bank ? dp->bank_b1 : nullptr;
        o.bank_w2 = p.use_feat_bank ? dp->bank_w2 : nullptr; o.bank_b2 = p.use_feat_bank ? dp->bank_b2 : nullptr;
        const int nstages = (n_vis + wg2::KS - 1) / wg2::KS;
        // every CTA pays for zeroing its operand tiles and for one flush: at least four stages each
        const int grid = std::max(1, std::min((nstages + 3) / 4, SM_COUNT));
        SEGS_CUDA_CHECK(cudaFuncSetAttribute(decode_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg2::SMEM));
        decode_wgrad_tc_kernel<<<grid, wg2::THREADS, wg2::SMEM, stream>>>(fact, n_vis, o, p.use_feat_bank ? FACT_ROWS : F_HB);
        SEGS_LAUNCH_CHECK();
        if (p.appearance_dim > 0) {
            decode_appgrad_kernel<<<1, 32, 0, stream>>>(p, *dp, p7);
            SEGS_LAUNCH_CHECK();
        }
        return SEGS_OK;
    }
    WJobs jobs;
    int n = 0;
    auto add = [&](int warp, int u0, int np, int v0, float* out, int ld, int nq, int tr = 0) {
        jobs.j[n++] = WJob{u0, np, v0, out, ld, nq, tr, warp};
    };
    // second layers: dW2[n][j] = sum d2[n] h[j]   (lanes = hidden unit j)
    add(0, F_D2S, 35, F_H + 1 * FEAT, dp->cov_w2, FEAT, FEAT);
    add(1, F_D2S + 35, 35, F_H + 1 * FEAT, dp->cov_w2 + 35 * FEAT, FEAT, FEAT);
    add(2, F_D2C, 3 * NOFF, F_H + 2 * FEAT, dp->color_w2, FEAT, FEAT);
    add(7, F_D2O, NOFF, F_H + 0 * FEAT, dp->opacity_w2, FEAT, FEAT);
    // first layers: dW1[j][i] = sum dpre[j] x[i]   (rows = hidden unit j, lanes = input column i < 32)
    add(3, F_DPRE + 0 * FEAT, FEAT, F_X, dp->opacity_w1, in_o, 32);
    add(4, F_DPRE + 1 * FEAT, FEAT, F_X, dp->cov_w1, in_s, 32);
    add(5, F_DPRE + 2 * FEAT, FEAT, F_X, dp->color_w1, ld_c, 32);
    // ... input columns 32.. : rows = input column, lanes = hidden unit, transposed store
    add(6, F_X + 32, in_o - 32, F_DPRE + 0 * FEAT, dp->opacity_w1 + 32, in_o, FEAT, 1);
    add(6, F_X + 32, in_s - 32, F_DPRE + 1 * FEAT, dp->cov_w1 + 32, in_s, FEAT, 1);
    add(6, F_X + 32, in_c - 32, F_DPRE + 2 * FEAT, dp->color_w1 + 32, ld_c, FEAT, 1);
    // biases
    add(6, F_D2O, NOFF, -1, dp->opacity_b2, 0, 0);
    add(6, F_D2S, 7 * NOFF, -1, dp->cov_b2, 0, 0);
    add(6, F_D2C, 3 * NOFF, -1, dp->color_b2, 0, 0);
    add(7, F_DPRE + 0 * FEAT, FEAT, -1, dp->opacity_b1, 0, 0);
    add(7, F_DPRE + 1 * FEAT, FEAT, -1, dp->cov_b1, 0, 0);
    add(7, F_DPRE + 2 * FEAT, FEAT, -1, dp->color_b1, 0, 0);
    if (p.use_feat_bank) {
        add(7, F_DLOG, 3, F_HB, dp->bank_w2, FEAT, FEAT);                 // dW2b[m][j] = sum dlog[m] hb[j]
        add(7, F_CAT, 4, F_DPREB, dp->bank_w1, 4, FEAT, 1);               // dW1b[j][i] = sum dpreb[j] cat[i]
        add(7, F_DLOG, 3, -1, dp->bank_b2, 0, 0);
        add(7, F_DPREB, FEAT, -1, dp->bank_b1, 0, 0);
    }
    jobs.n = n;
    const int nrows_used = p.use_feat_bank ? FACT_ROWS : F_HB;     // the bank rows are only written in bank mode
    const size_t smem = size_t(FACT_ROWS) * WG_STRIDE * sizeof(float);
    SEGS_CUDA_CHECK(cudaFuncSetAttribute(decode_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<size_t>(ntiles, SM_COUNT * 2);   // 2 x 111 KB of shared memory per SM
    decode_wgrad_kernel<<<grid, WG_THREADS, smem, stream>>>(fact, n_vis, jobs, nrows_used);
    SEGS_LAUNCH_CHECK();
    if (p.appearance_dim > 0) {
        decode_appgrad_kernel<<<1, 32, 0, stream>>>(p, *dp, p7);
        SEGS_LAUNCH_CHECK();
    }
    return SEGS_OK;
}

// =======================================================================================
// densification statistics of one view (GaussianModel::training_statis, gaussian_model.cpp:1459-1503)
// =======================================================================================
namespace segs {
namespace {
// thread = visible anchor.  opacity_accum[a] += sum_o max(neural_opacity, 0); anchor_demon[a] += 1; for every emitted
// offset whose Gaussian was rendered (radii > 0): offset_gradient_accum[a,o] += |dL_dmean2D.xy|, offset_denom[a,o] += 1.
__global__ void __launch_bounds__(256)
training_statis_kernel(int n_vis, DecodeState st, const float* __restrict__ neural_opacity, const int* __restrict__ radii,
                       const float* __restrict__ dL_dmean2D, float* __restrict__ opacity_accum,
                       float* __restrict__ anchor_demon, float* __restrict__ offset_gradient_accum,
                       float* __restrict__ offset_denom, const int atomic)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_vis) return;
    auto add = [&](float* dst, float v) { if (atomic) atomicAdd(dst, v); else *dst += v; };
    const size_t a = st.anchor_index[k];
    const uint32_t m = st.mask_bits[k];
    size_t r = st.row_start[k];
    float sum = 0.f;
#pragma unroll
    for (int o = 0; o < NOFF; ++o) {
        const float v = __ldg(neural_opacity + size_t(k) * NOFF + o);
        sum += v < 0.f ? 0.f : v;                       // torch::where(temp_opacity < 0, 0, temp_opacity)
    }
    add(opacity_accum + a, sum);
    add(anchor_demon + a, 1.f);
#pragma unroll 1
    for (int o = 0; o < NOFF; ++o) {
        if (!((m >> o) & 1u)) continue;
        if (__ldg(radii + r) > 0) {
            const float gx = __ldg(dL_dmean2D + 3 * r), gy = __ldg(dL_dmean2D + 3 * r + 1);
            add(offset_gradient_accum + a * NOFF + o, sqrtf(gx * gx + gy * gy));
            add(offset_denom + a * NOFF + o, 1.f);
        }
        ++r;
    }
}
}  // namespace
}  // namespace segs

extern "C" int segs_training_statis(int A, const char* decode_state, int n_vis, const float* neural_opacity, const int* radii,
                                    const float* dL_dmean2D, float* opacity_accum, float* anchor_demon,
                                    float* offset_gradient_accum, float* offset_denom, int atomic, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (A < 0 || n_vis < 0 || n_vis > A) { set_error("training_statis: invalid sizes"); return SEGS_ERR_INVALID_ARG; }
    if (n_vis == 0) return SEGS_OK;
    if (!decode_state || !neural_opacity || !radii || !dL_dmean2D || !opacity_accum || !anchor_demon || !offset_gradient_accum ||
        !offset_denom) { set_error("training_statis: NULL pointer"); return SEGS_ERR_INVALID_ARG; }
    DecodeState st = DecodeState::carve(const_cast<char*>(decode_state), A, nullptr);
    training_statis_kernel<<<(n_vis + 255) / 256, 256, 0, stream>>>(n_vis, st, neural_opacity, radii, dL_dmean2D, opacity_accum,
                                                                     anchor_demon, offset_gradient_accum, offset_denom, atomic);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

extern "C" int segs_decode_backward(
    int A, const unsigned char* visible_mask, const float* anchor, const float* anchor_feat, const float* offset,
    const float* scaling, const float* camera_center, const float* pose, const segs_decode_params* params,
    const char* state, int n_vis, int n_out, const float* g_xyz, const float* g_color, const float* g_opacity,
    const float* g_scaling, const float* g_rot, const float* g_neural_opacity, float* d_anchor, float* d_anchor_feat,
    float* d_offset, float* d_scaling, const segs_decode_grads* dp, segs_alloc_fn scratch_alloc, void* scratch_user,
    void* stream_)
{
    return segs_decode_backward_ex(A, visible_mask, anchor, anchor_feat, offset, scaling, camera_center, pose, params, state,
                                   n_vis, n_out, g_xyz, g_color, g_opacity, g_scaling, g_rot, g_neural_opacity, d_anchor,
                                   d_anchor_feat, d_offset, d_scaling, dp, scratch_alloc, scratch_user, 0, stream_);
}
