// Stable LSD radix sort of (u32 key, u32 value) pairs for sm_100a — single-pass-per-digit
// ("onesweep") formulation.  Used for the depth ordering of the Gaussians, the super-tile
// grouping of the binning candidates (binning.cu) and the Morton ordering of the kNN (knn.cu);
// together with binning.cu it replaces the reference's cub::DeviceRadixSort::SortPairs over
// 64-bit tile|depth keys (cuda_rasterizer/rasterizer_impl.cu:303-309) and simple-knn's
// cub sort (simple_knn.cu:204-209).
//
//   1. ONE histogram kernel reads the keys once and builds the 256-bin histogram of every digit
//      that will be sorted on.
//   2. One kernel per 8-bit digit.  A CTA owns a tile of TILE keys (tile index handed out by an
//      atomic ticket, so a CTA only ever waits for CTAs that are already running), ranks its keys
//      per warp with match.any (stable: lanes in order, warps in order, items in order), publishes
//      its per-digit counts and resolves its scatter base with a decoupled look-back over the
//      preceding tiles (flag + count packed in one 32-bit word, so no fences are needed), then
//      scatters.  Traffic per digit: one read + one write of the pairs — the algorithmic minimum —
//      and 1 launch instead of histogram + scan + scatter.
//
// All scratch is caller-provided (radix_sort_temp_words) and zeroed here with one memset.
#include "common.cuh"

namespace segs {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 12;                         // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // 3072 keys per CTA
constexpr int RS_MAX_PASSES = 4;
constexpr uint32_t FLAG_AGG = 1u << 30;              // tile's own count is available
constexpr uint32_t FLAG_PREFIX = 2u << 30;           // inclusive prefix over tiles [0, tile] is available
constexpr uint32_t FLAG_MASK = 3u << 30;
static_assert(RS_THREADS == RADIX_BINS, "one thread per radix bin");

inline int rs_tiles(size_t n) { return int((n + RS_TILE - 1) / RS_TILE); }

// temp layout (u32 words): [0,4) tile tickets | [4, 4+4*256) digit histograms | look-back words
constexpr size_t RS_HIST_OFF = 4;
constexpr size_t RS_LOOK_OFF = RS_HIST_OFF + size_t(RS_MAX_PASSES) * RADIX_BINS;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t* __restrict__ keys, size_t n, int begin_bit, int npasses, uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t s_hist[RS_MAX_PASSES][RADIX_BINS];
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RADIX_BINS; i += RS_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    for (size_t base = size_t(blockIdx.x) * RS_TILE; base < n; base += size_t(gridDim.x) * RS_TILE) {
        uint32_t k[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const size_t e = base + size_t(i) * RS_THREADS + threadIdx.x;
            k[i] = (e < n) ? __ldg(keys + e) : 0u;
        }
#pragma unroll
        for (int i = 0; i < RS_ITEMS; ++i) {
            const size_t e = base + size_t(i) * RS_THREADS + threadIdx.x;
            if (e < n) {
                const uint32_t v = k[i] >> begin_bit;
#pragma unroll
                for (int p = 0; p < RS_MAX_PASSES; ++p)
                    if (p < npasses) atomicAdd(&s_hist[p][(v >> (8 * p)) & (RADIX_BINS - 1)], 1u);
            }
        }
    }
    __syncthreads();
    for (int p = 0; p < npasses; ++p) {
        const uint32_t c = s_hist[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * RADIX_BINS + threadIdx.x], c);
    }
}

// CTA-wide exclusive scan of one value per thread (256 threads); contains two __syncthreads
__device__ __forceinline__ uint32_t scan256(uint32_t c, uint32_t* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();                     // s_warp may still be read from a previous scan
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w)
        if (w < warp) woff += s_warp[w];
    return woff + inc - c;
}

template <bool IOTA>
__global__ void __launch_bounds__(RS_THREADS)
rs_pass_kernel(const uint32_t* __restrict__ key_in, uint32_t* __restrict__ key_out,
               const uint32_t* __restrict__ val_in, uint32_t* __restrict__ val_out, size_t n, int shift,
               const uint32_t* __restrict__ hist /*[256] of this digit*/, uint32_t* __restrict__ ticket,
               volatile uint32_t* __restrict__ look /*[tiles][256]*/)
{
    __shared__ uint32_t s_wh[RS_WARPS][RADIX_BINS];   // per-warp digit counts -> per-warp offsets inside the tile
    __shared__ uint32_t s_key[RS_TILE], s_val[RS_TILE];   // the tile, reordered by digit
    __shared__ uint32_t s_tstart[RADIX_BINS];         // first slot of every digit inside the reordered tile
    __shared__ uint32_t s_gbase[RADIX_BINS];          // global position of that first slot
    __shared__ uint32_t s_warp[RS_WARPS];
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RADIX_BINS; i += RS_THREADS) (&s_wh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;

    // warp-striped: warp w owns [wbase, wbase + 32*ITEMS); item i of lane l is wbase + 32*i + l
    const size_t tbase = size_t(tile) * RS_TILE;
    const size_t wbase = tbase + size_t(warp) * (32 * RS_ITEMS);
    const uint32_t n_tile = (uint32_t)min(size_t(RS_TILE), n - tbase);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        key[i] = (e < n) ? __ldg(key_in + e) : 0xFFFFFFFFu;
        val[i] = IOTA ? (uint32_t)e : ((e < n) ? __ldg(val_in + e) : 0u);
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        const bool valid = e < n;
        const uint32_t digit = valid ? ((key[i] >> shift) & (RADIX_BINS - 1)) : RADIX_BINS;
        const uint32_t peers = __match_any_sync(FULL, digit);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = s_wh[warp][digit];
            s_wh[warp][digit] = old + __popc(peers);
        }
        old = __shfl_sync(FULL, old, leader);
        rank[i] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // thread t owns bin t
    const int bin = threadIdx.x;
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = s_wh[w][bin];
        s_wh[w][bin] = count;          // exclusive over warps (the digit's first slot is added below)
        count += c;
    }
    volatile uint32_t* mine = look + size_t(tile) * RADIX_BINS + bin;
    uint32_t excl = 0;
    if (tile == 0) {
        *mine = FLAG_PREFIX | count;
    } else {
        *mine = FLAG_AGG | count;
        // decoupled look-back: walk the preceding tiles until one has published its inclusive
        // prefix.  LOOK words are fetched per round so their L2 latencies overlap (all tiles of a
        // wave publish their aggregates at about the same time, so the walk is many tiles deep).
        constexpr int LOOK = 8;
        bool done = false;
        for (int t = (int)tile - 1; !done; t -= LOOK) {
            uint32_t w[LOOK];
#pragma unroll
            for (int i = 0; i < LOOK; ++i) w[i] = (t - i >= 0) ? look[size_t(t - i) * RADIX_BINS + bin] : uint32_t(FLAG_PREFIX);
#pragma unroll
            for (int i = 0; i < LOOK; ++i) {
                if (done) break;
                while ((w[i] & FLAG_MASK) == 0u) w[i] = look[size_t(t - i) * RADIX_BINS + bin];
                excl += w[i] & ~FLAG_MASK;
                done = (w[i] & FLAG_MASK) == FLAG_PREFIX;
            }
        }
        *mine = FLAG_PREFIX | (excl + count);
    }
    const uint32_t gbin = scan256(hist[bin], s_warp);          // first global slot of this digit
    const uint32_t tstart = scan256(count, s_warp);            // first slot of this digit inside the tile
    s_tstart[bin] = tstart;
    s_gbase[bin] = gbin + excl;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) s_wh[w][bin] += tstart;
    __syncthreads();

    // reorder the tile by digit in shared memory ...
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        if (e < n) {
            const uint32_t digit = (key[i] >> shift) & (RADIX_BINS - 1);
            const uint32_t slot = s_wh[warp][digit] + rank[i];
            s_key[slot] = key[i];
            s_val[slot] = val[i];
        }
    }
    __syncthreads();
    // ... so that consecutive threads store consecutive addresses inside every digit's run
    // (a direct scatter costs one 32-byte L2 sector write per 4-byte element)
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t slot = i * RS_THREADS + threadIdx.x;
        if (slot < n_tile) {
            const uint32_t k = s_key[slot];
            const uint32_t digit = (k >> shift) & (RADIX_BINS - 1);
            const uint32_t pos = s_gbase[digit] + (slot - s_tstart[digit]);
            key_out[pos] = k;
            val_out[pos] = s_val[slot];
        }
    }
}

}  // namespace

size_t radix_sort_temp_words(size_t n, int npasses)
{
    return RS_LOOK_OFF + size_t(npasses) * rs_tiles(n) * RADIX_BINS;
}

int radix_sort_pairs(uint32_t* key_a, uint32_t* key_b, uint32_t* val_a, uint32_t* val_b, size_t n,
                     int begin_bit, int npasses, bool iota_values, uint32_t* temp, cudaStream_t stream)
{
    if (n == 0 || npasses <= 0) return SEGS_OK;
    if (npasses > RS_MAX_PASSES) { set_error("radix_sort_pairs: at most %d digits", RS_MAX_PASSES); return SEGS_ERR_INVALID_ARG; }
    const int tiles = rs_tiles(n);
    SEGS_CUDA_CHECK(cudaMemsetAsync(temp, 0, radix_sort_temp_words(n, npasses) * sizeof(uint32_t), stream));
    uint32_t* ghist = temp + RS_HIST_OFF;
    rs_hist_kernel<<<min(tiles, SM_COUNT * 4), RS_THREADS, 0, stream>>>(key_a, n, begin_bit, npasses, ghist);
    SEGS_LAUNCH_CHECK();
    uint32_t *ki = key_a, *ko = key_b, *vi = val_a, *vo = val_b;
    for (int p = 0; p < npasses; ++p) {
        uint32_t* look = temp + RS_LOOK_OFF + size_t(p) * tiles * RADIX_BINS;
        if (p == 0 && iota_values)
            rs_pass_kernel<true><<<tiles, RS_THREADS, 0, stream>>>(ki, ko, vi, vo, n, begin_bit + 8 * p,
                                                                   ghist + p * RADIX_BINS, temp + p, look);
        else
            rs_pass_kernel<false><<<tiles, RS_THREADS, 0, stream>>>(ki, ko, vi, vo, n, begin_bit + 8 * p,
                                                                    ghist + p * RADIX_BINS, temp + p, look);
        SEGS_LAUNCH_CHECK();
        uint32_t* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    return SEGS_OK;   // result is in (key_a, val_a) for an even number of passes, (key_b, val_b) otherwise
}

}  // namespace segs
