// Stable LSD radix sort of (u32 key, u32 value) pairs for sm_100a — ONE persistent kernel for the histogram and all digit
// passes.  Used for the depth ordering of the Gaussians, the super-tile grouping of the binning candidates (binning.cu),
// the Morton ordering of the kNN (knn.cu) and the voxel ordering of the densification (densify.cu); together with
// binning.cu it replaces the reference's cub::DeviceRadixSort::SortPairs over 64-bit tile|depth keys
// (cuda_rasterizer/rasterizer_impl.cu:303-309) and simple-knn's cub sort (simple_knn.cu:204-209).
//
// At the sizes of this path (1e5 - 5e6 keys, 33 - 1600 tiles) a pass is one wave of CTAs, so what a sort costs is not
// bandwidth (8 bytes in + 8 bytes out per pair and pass: 2.5 us at 1M pairs) but launch gaps and the latency of the
// inter-tile prefix.  Measured on B200 with one kernel per pass and a classic decoupled look-back (round 1): 12 us per CTA,
// of which 4.3 us walking back over the tile counts, plus ~9 us of launch / ramp / drain per pass.  Hence:
//
//   * One launch.  CTAs draw tickets from a global counter: ticket -> (stage, tile).  Stage 0 tiles build the 256-bin
//     histograms of every digit; stage p+1 tiles run digit p.  A tile of stage s waits until all tiles of stage s-1 have
//     finished (a counter per stage).  Tickets are handed out in order, so a waiting CTA only ever waits for CTAs that
//     are already running: no co-residency assumption, no cooperative launch.
//   * No prefix chain.  Every tile publishes its digit counts EARLY (shared-memory atomics right after the load, before
//     the ranking); the last tile of every group of 32 also publishes the group's sum.  A tile's scatter base is then
//     digit base + the counts of the <= 31 preceding tiles of its group + the sums of the preceding groups: independent
//     loads issued together, not a walk that has to meet a published prefix.
//   * Peer masks for the stable in-warp ranking come from 8 ballots (one per digit bit) rather than match.any.
//   * Each digit's run of a tile is staged through shared memory so that it is stored coalesced.
//
// Data written earlier in the same launch is read with ld.global.cg / volatile loads (L1 is not coherent).
// All scratch is caller-provided (radix_sort_temp_words) and zeroed here with one memset.
#include <algorithm>
#include "common.cuh"

namespace segs {

namespace {

// dev aid (tools/cuda/sort_bench.cu -DSEGS_RS_PHASE_TIMING): cycles per phase of sort_tile, summed over tiles
#ifdef SEGS_RS_PHASE_TIMING
#define RS_T(i) do { __syncthreads(); if (threadIdx.x == 0) rs_t[i] = clock64(); } while (0)
#else
#define RS_T(i) do { } while (0)
#endif

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 12;                         // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // 3072 keys per tile
constexpr int RS_MAX_PASSES = 4;
constexpr int RS_GROUP = 32;                         // tiles per group (one published sum per group)
constexpr uint32_t FLAG_SET = 1u << 31;              // word published (counts stay below 2^31)
static_assert(RS_THREADS == RADIX_BINS, "one thread per radix bin");

inline int rs_tiles(size_t n) { return int((n + RS_TILE - 1) / RS_TILE); }
inline int rs_groups(int tiles) { return (tiles + RS_GROUP - 1) / RS_GROUP; }

// temp layout (u32 words): control [0,16): [0] ticket, [1 + s] tiles finished in stage s |
//                          digit histograms [4][256] | per pass: tile counts [tiles][256], group sums [groups][256]
constexpr size_t RS_CTRL_WORDS = 16;
constexpr size_t RS_HIST_OFF = RS_CTRL_WORDS;
constexpr size_t RS_LOOK_OFF = RS_HIST_OFF + size_t(RS_MAX_PASSES) * RADIX_BINS;
inline size_t rs_pass_words(int tiles) { return (size_t(tiles) + rs_groups(tiles)) * RADIX_BINS; }

struct RsShared {
    uint32_t wh[RS_WARPS][RADIX_BINS];    // per-warp digit counts -> per-warp offsets inside the tile
    uint32_t key[RS_TILE], val[RS_TILE];  // the tile, reordered by digit
    uint32_t tstart[RADIX_BINS];          // first slot of every digit inside the reordered tile
    uint32_t gbase[RADIX_BINS];           // global position of that first slot
    uint32_t cnt[RADIX_BINS];             // the tile's digit counts (published early)
    uint32_t warp_sums[RS_WARPS];
    uint32_t ticket;
};

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

// sum of `count` published words p[0], p[stride], ... : RS_BATCH loads in flight per round (their L2 latencies overlap);
// a word that is not published yet is polled
constexpr int RS_BATCH = 16;
__device__ __forceinline__ uint32_t sum_published(const volatile uint32_t* p, long long stride, uint32_t count) {
    uint32_t sum = 0;
    for (uint32_t j0 = 0; j0 < count; j0 += RS_BATCH) {
        uint32_t w[RS_BATCH];
#pragma unroll
        for (uint32_t j = 0; j < RS_BATCH; ++j) w[j] = (j0 + j < count) ? p[(long long)(j0 + j) * stride] : uint32_t(FLAG_SET);
#pragma unroll
        for (uint32_t j = 0; j < RS_BATCH; ++j) {
            if (j0 + j < count) {
                while (!(w[j] & FLAG_SET)) w[j] = p[(long long)(j0 + j) * stride];
                sum += w[j] & ~FLAG_SET;
            }
        }
    }
    return sum;
}

// stage 0: the 256-bin histogram of every digit over one tile of the INPUT keys
__device__ __forceinline__ void hist_tile(RsShared& sm, const uint32_t* __restrict__ keys, size_t n, uint32_t tile, int begin_bit,
                                          int npasses, uint32_t* __restrict__ ghist)
{
    uint32_t* h = sm.key;                // RS_MAX_PASSES * 256 words of scratch
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RADIX_BINS; i += RS_THREADS) h[i] = 0;
    __syncthreads();
    const size_t base = size_t(tile) * RS_TILE;
    uint32_t k[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = base + size_t(i) * RS_THREADS + threadIdx.x;
        k[i] = (e < n) ? __ldg(keys + e) : 0u;      // the input is not written by this launch
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = base + size_t(i) * RS_THREADS + threadIdx.x;
        if (e < n) {
            const uint32_t v = k[i] >> begin_bit;
#pragma unroll
            for (int p = 0; p < RS_MAX_PASSES; ++p)
                if (p < npasses) atomicAdd(&h[p * RADIX_BINS + ((v >> (8 * p)) & (RADIX_BINS - 1))], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < npasses; ++p) {
        const uint32_t c = h[p * RADIX_BINS + threadIdx.x];
        if (c) atomicAdd(&ghist[p * RADIX_BINS + threadIdx.x], c);
    }
}

// one tile of one digit pass
__device__ __forceinline__ void sort_tile(RsShared& sm, const uint32_t* key_in, uint32_t* key_out, const uint32_t* val_in,
                                          uint32_t* val_out, size_t n, int shift, bool iota, bool input_is_fresh,
                                          const uint32_t* hist /*[256] of this digit*/, volatile uint32_t* tile_counts /*[tiles][256]*/,
                                          volatile uint32_t* group_sums /*[groups][256]*/, uint32_t tile)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef SEGS_RS_PHASE_TIMING
    long long rs_t[10];
#endif
    RS_T(0);
    for (int i = threadIdx.x; i < RS_WARPS * RADIX_BINS; i += RS_THREADS) (&sm.wh[0][0])[i] = 0;
    sm.cnt[threadIdx.x] = 0;
    __syncthreads();

    // warp-striped: warp w owns [wbase, wbase + 32*ITEMS); item i of lane l is wbase + 32*i + l
    const size_t tbase = size_t(tile) * RS_TILE;
    const size_t wbase = tbase + size_t(warp) * (32 * RS_ITEMS);
    const uint32_t n_tile = (uint32_t)min(size_t(RS_TILE), n - tbase);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        // pass 0 reads the caller's input; later passes read what other CTAs of this launch wrote: L2, never L1
        key[i] = (e < n) ? (input_is_fresh ? __ldg(key_in + e) : __ldcg(key_in + e)) : 0xFFFFFFFFu;
        val[i] = iota ? (uint32_t)e : ((e < n) ? (input_is_fresh ? __ldg(val_in + e) : __ldcg(val_in + e)) : 0u);
    }
#ifdef SEGS_RS_PHASE_TIMING
    if (key[RS_ITEMS - 1] == 0x12345678u && val[0] == 77u) sm.ticket = 0;      // keep the loads in front of the timer
#endif
    RS_T(1);
    // the tile's digit counts, published as early as possible
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        if (e < n) atomicAdd(&sm.cnt[(key[i] >> shift) & (RADIX_BINS - 1)], 1u);
    }
    __syncthreads();
    const int bin = threadIdx.x;
    const uint32_t count = sm.cnt[bin];
    tile_counts[size_t(tile) * RADIX_BINS + bin] = FLAG_SET | count;
    const uint32_t in_group = tile % RS_GROUP, group = tile / RS_GROUP;
    uint32_t group_excl = 0;                  // counts of the preceding tiles of this tile's group
    const bool group_leader = in_group == RS_GROUP - 1;
    if (group_leader) {
        // last tile of a full group: publish the group's sum (its predecessors drew their tickets earlier and publish
        // their counts before they rank, so this wait is short)
        group_excl = sum_published(tile_counts + size_t(tile - 1) * RADIX_BINS + bin, -(long long)RADIX_BINS, RS_GROUP - 1);
        group_sums[size_t(group) * RADIX_BINS + bin] = FLAG_SET | (group_excl + count);
    }

    RS_T(2);
    // stable ranking inside each warp: peers = lanes holding the same digit, from one ballot per digit bit
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t peers[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        const uint32_t digit = (key[i] >> shift) & (RADIX_BINS - 1);
        uint32_t m = __ballot_sync(FULL, e < n);                // invalid items (tail of the last tile) are nobody's peer
#pragma unroll
        for (int b = 0; b < RADIX_BITS; ++b) {
            const bool bit = (digit >> b) & 1u;
            const uint32_t v = __ballot_sync(FULL, bit);
            m &= bit ? v : ~v;
        }
        peers[i] = m;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        const uint32_t digit = (key[i] >> shift) & (RADIX_BINS - 1);
        const int leader = __ffs(peers[i]) - 1;
        uint32_t old = 0;
        if (e < n && lane == leader) {
            old = sm.wh[warp][digit];
            sm.wh[warp][digit] = old + __popc(peers[i]);
        }
        old = __shfl_sync(FULL, old, leader < 0 ? 0 : leader);
        rank[i] = old + __popc(peers[i] & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    RS_T(3);

    // thread t owns bin t: exclusive offsets over the warps ...
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = sm.wh[w][bin];
        sm.wh[w][bin] = run;
        run += c;
    }
    // ... and over the preceding tiles: the counts of this group's earlier tiles + the sums of the earlier groups,
    // all loads independent (no prefix chain)
    if (!group_leader && in_group)
        group_excl = sum_published(tile_counts + size_t(tile - 1) * RADIX_BINS + bin, -(long long)RADIX_BINS, in_group);
    const uint32_t excl = group_excl + (group ? sum_published(group_sums + bin, RADIX_BINS, group) : 0u);
    RS_T(4);
    const uint32_t gbin = scan256(__ldcg(hist + bin), sm.warp_sums);      // first global slot of this digit
    const uint32_t tstart = scan256(count, sm.warp_sums);                 // first slot of this digit inside the tile
    sm.tstart[bin] = tstart;
    sm.gbase[bin] = gbin + excl;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) sm.wh[w][bin] += tstart;
    __syncthreads();
    RS_T(5);

    // reorder the tile by digit in shared memory ...
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        if (e < n) {
            const uint32_t digit = (key[i] >> shift) & (RADIX_BINS - 1);
            const uint32_t slot = sm.wh[warp][digit] + rank[i];
            sm.key[slot] = key[i];
            sm.val[slot] = val[i];
        }
    }
    __syncthreads();
    RS_T(6);
    // ... so that consecutive threads store consecutive addresses inside every digit's run
    // (a direct scatter costs one 32-byte L2 sector write per 4-byte element)
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t slot = i * RS_THREADS + threadIdx.x;
        if (slot < n_tile) {
            const uint32_t k = sm.key[slot];
            const uint32_t digit = (k >> shift) & (RADIX_BINS - 1);
            const uint32_t pos = sm.gbase[digit] + (slot - sm.tstart[digit]);
            key_out[pos] = k;
            val_out[pos] = sm.val[slot];
        }
    }
#ifdef SEGS_RS_PHASE_TIMING
    RS_T(7);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 7; ++i) atomicAdd(&g_rs_cycles[i], (unsigned long long)(rs_t[i + 1] - rs_t[i]));
        atomicAdd(&g_rs_cycles[7], 1ull);
    }
#endif
}

#ifndef RS_MIN_CTAS
#define RS_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_CTAS)
rs_sort_kernel(uint32_t* key_a, uint32_t* key_b, uint32_t* val_a, uint32_t* val_b, size_t n, int tiles, int begin_bit, int npasses,
               int iota_values, uint32_t* temp)
{
    __shared__ RsShared sm;
    volatile uint32_t* ctrl = temp;
    uint32_t* ghist = temp + RS_HIST_OFF;
    const uint32_t total = uint32_t(1 + npasses) * (uint32_t)tiles;
    const size_t pass_words = (size_t(tiles) + (tiles + RS_GROUP - 1) / RS_GROUP) * RADIX_BINS;
    for (;;) {
        __syncthreads();                                      // everyone is done with sm (incl. sm.ticket) of the previous tile
        if (threadIdx.x == 0) sm.ticket = atomicAdd(temp, 1u);
        __syncthreads();
        const uint32_t ticket = sm.ticket;
#ifdef SEGS_RS_PHASE_TIMING
        if (ticket == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); g_rs_phase[0] = t; }
#endif
        if (ticket >= total) return;
        const uint32_t stage = ticket / (uint32_t)tiles, tile = ticket % (uint32_t)tiles;
        if (stage == 0) {
            hist_tile(sm, key_a, n, tile, begin_bit, npasses, ghist);
        } else {
            // every tile of the previous stage has finished (its scatters and histogram updates are visible)
            if (threadIdx.x == 0) {
                while (ctrl[stage] < (uint32_t)tiles) { }
                __threadfence();
            }
            __syncthreads();
            const int p = int(stage) - 1;
            const bool even = (p & 1) == 0;
            uint32_t* look = temp + RS_LOOK_OFF + size_t(p) * pass_words;
            sort_tile(sm, even ? key_a : key_b, even ? key_b : key_a, even ? val_a : val_b, even ? val_b : val_a, n,
                      begin_bit + 8 * p, p == 0 && iota_values != 0, p == 0, ghist + p * RADIX_BINS, look,
                      look + size_t(tiles) * RADIX_BINS, tile);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();                                  // this tile's global writes before the stage counter
#ifdef SEGS_RS_PHASE_TIMING
            if (atomicAdd(temp + 1 + stage, 1u) == (uint32_t)tiles - 1) {
                unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                g_rs_phase[1 + stage] = t;
            }
#else
            atomicAdd(temp + 1 + stage, 1u);
#endif
        }
    }
}

}  // namespace

size_t radix_sort_temp_words(size_t n, int npasses)
{
    return RS_LOOK_OFF + size_t(npasses) * rs_pass_words(rs_tiles(n));
}

int radix_sort_pairs(uint32_t* key_a, uint32_t* key_b, uint32_t* val_a, uint32_t* val_b, size_t n,
                     int begin_bit, int npasses, bool iota_values, uint32_t* temp, cudaStream_t stream)
{
    if (n == 0 || npasses <= 0) return SEGS_OK;
    if (npasses > RS_MAX_PASSES) { set_error("radix_sort_pairs: at most %d digits", RS_MAX_PASSES); return SEGS_ERR_INVALID_ARG; }
    const int tiles = rs_tiles(n);
    SEGS_CUDA_CHECK(cudaMemsetAsync(temp, 0, radix_sort_temp_words(n, npasses) * sizeof(uint32_t), stream));
    // persistent grid: one CTA per tile while the tiles fit the machine (every CTA then runs exactly one tile per stage
    // and the hardware spreads them evenly over the SMs), else as many CTAs as can be resident
    const int grid = std::min(tiles, SM_COUNT * RS_MIN_CTAS);
    rs_sort_kernel<<<grid, RS_THREADS, 0, stream>>>(key_a, key_b, val_a, val_b, n, tiles, begin_bit, npasses, iota_values ? 1 : 0, temp);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;   // result is in (key_a, val_a) for an even number of passes, (key_b, val_b) otherwise
}

}  // namespace segs
