// Depth ordering and tile binning for sm_100a.
//
// Replaces cub::DeviceScan::InclusiveSum, duplicateWithKeys, the 64-bit-key
// cub::DeviceRadixSort::SortPairs and identifyTileRanges of the reference
// (cuda_rasterizer/rasterizer_impl.cu:70-138, 276-318).
//
// The reference materialises R (Gaussian x tile) 12-byte instances and sorts them by the
// 64-bit key  tile<<32 | depth_bits  with a stable LSD radix sort (6 CUB passes, ~150 B of
// traffic per instance).  The same total order (tile, depth_bits, Gaussian index) is produced
// here WITHOUT sorting the R instances, by hierarchical binning:
//
//   0. a stable 4-digit radix sort of the P Gaussians by depth_bits (radix_sort.cu; P is ~8x
//      smaller than R);
//   1. coarse level: the image is cut into super-tiles of 4x4 tiles.  The depth-ordered Gaussians
//      are expanded into (super-tile, Gaussian) candidates (R' ~ 0.25 R); each candidate carries
//      the 16-bit mask of the tiles of its super-tile that the Gaussian's rectangle covers
//      (key = super-tile | mask << 16, value = Gaussian id).  A stable radix sort on the
//      super-tile bits (one digit up to 256 super-tiles, two up to 65536) groups them: every
//      super-tile now owns a contiguous, depth-ordered candidate list;
//   2. one streaming pass over the grouped keys finds the super-tile boundaries and counts the
//      set mask bits per tile (packed counters + redux, ~16 atomics per 256 candidates); a scan
//      over the T tile counts gives `ranges`;
//   3. fine level: one warp per tile streams its super-tile's candidate list and
//      stream-compacts the candidates whose mask has the tile's bit with a ballot — compaction
//      preserves list order, so every tile's list is in (depth_bits, Gaussian index) order,
//      exactly the reference's sorted order, bit for bit.  point_list is the only R-sized
//      array ever written (4 B/instance).
#include "common.cuh"

namespace segs {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int SSHIFT = 2;                 // super-tile = 4x4 tiles -> 16-bit tile mask
constexpr int SSIDE = 1 << SSHIFT;
constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;

// ---------------------------------------------------------------------------------------
// coarse level: load-balanced expansion of depth-ordered Gaussians into candidates
// ---------------------------------------------------------------------------------------
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_ITEMS = 8;
constexpr int EMIT_TILE = EMIT_THREADS * EMIT_ITEMS;   // Gaussians per CTA

// number of super-tiles a tile rectangle overlaps (0 for an empty rectangle)
__device__ __forceinline__ uint32_t coarse_count(ushort4 rc) {
    if (rc.x >= rc.z || rc.y >= rc.w) return 0u;
    const uint32_t sx0 = rc.x >> SSHIFT, sx1 = (rc.z - 1) >> SSHIFT;
    const uint32_t sy0 = rc.y >> SSHIFT, sy1 = (rc.w - 1) >> SSHIFT;
    return (sx1 - sx0 + 1) * (sy1 - sy0 + 1);
}

// bits [lo, hi) of a 4-bit field
__device__ __forceinline__ uint32_t span4(uint32_t lo, uint32_t hi) { return ((1u << hi) - 1u) & ~((1u << lo) - 1u); }

// 16-bit mask (bit = ly*4 + lx) of the tiles of super-tile (sx, sy) inside the rectangle
__device__ __forceinline__ uint32_t tile_mask(ushort4 rc, uint32_t sx, uint32_t sy) {
    const uint32_t bx = sx << SSHIFT, by = sy << SSHIFT;
    const uint32_t lx0 = max((uint32_t)rc.x, bx) - bx, lx1 = min((uint32_t)rc.z, bx + SSIDE) - bx;
    const uint32_t ly0 = max((uint32_t)rc.y, by) - by, ly1 = min((uint32_t)rc.w, by + SSIDE) - by;
    const uint32_t rows = span4(ly0, ly1);
    const uint32_t rows4 = (rows & 1u) | ((rows & 2u) << 3) | ((rows & 4u) << 6) | ((rows & 8u) << 9);   // bit ly -> bit 4*ly
    return rows4 * span4(lx0, lx1);                              // times the row pattern (no overlap, no carries)
}

__global__ void __launch_bounds__(EMIT_THREADS)
coarse_emit_kernel(const uint32_t* __restrict__ order, const ushort4* __restrict__ rect, int P, int sgrid_x,
                   uint32_t* __restrict__ ticket, volatile uint32_t* __restrict__ look,
                   uint32_t* __restrict__ key_out, uint32_t* __restrict__ val_out)
{
    __shared__ uint32_t s_off[EMIT_TILE + 1];      // exclusive candidate offset of every item of the CTA
    __shared__ uint32_t s_id[EMIT_TILE];
    __shared__ ushort4 s_rect[EMIT_TILE];
    __shared__ uint32_t s_warp[EMIT_THREADS / 32];
    __shared__ uint32_t s_look[EMIT_THREADS / 32];
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;

    // blocked arrangement: thread t owns items [8t, 8t+8) of the CTA's depth-ordered run
    const int first = (int)tile * EMIT_TILE + threadIdx.x * EMIT_ITEMS;
    uint32_t id[EMIT_ITEMS], c[EMIT_ITEMS];
    ushort4 rc[EMIT_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < EMIT_ITEMS; ++i) id[i] = (first + i < P) ? __ldg(order + first + i) : 0u;
#pragma unroll
    for (int i = 0; i < EMIT_ITEMS; ++i) {
        rc[i] = (first + i < P) ? __ldg(rect + id[i]) : make_ushort4(0, 0, 0, 0);
        c[i] = coarse_count(rc[i]);
        sum += c[i];
    }
    // block exclusive scan of the per-thread sums
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < EMIT_THREADS / 32; ++w) {
        const uint32_t x = s_warp[w];
        if (w < warp) woff += x;
        total += x;
    }
    uint32_t run = woff + inc - sum;
#pragma unroll
    for (int i = 0; i < EMIT_ITEMS; ++i) {
        const int slot = threadIdx.x * EMIT_ITEMS + i;
        s_off[slot] = run;
        s_id[slot] = id[i];
        s_rect[slot] = rc[i];
        run += c[i];
    }
    // decoupled look-back over the preceding CTAs' totals (one word per CTA).  The CTAs of this single-wave kernel publish
    // their aggregates at about the same time, so a 32-wide walk meets a published prefix only after many dependent
    // global round trips (tile 488 of 489 at C2: up to 15); the 8 warps poll 8 windows of 32 tiles AT ONCE and every
    // thread combines them: the nearest window holding an inclusive prefix ends the walk.
    if (threadIdx.x == 0) {
        s_off[EMIT_TILE] = total;
        look[tile] = (tile == 0 ? FLAG_PREFIX : FLAG_AGG) | total;
    }
    uint32_t excl = 0;
    if (tile != 0) {
        for (int newest = (int)tile - 1;; newest -= EMIT_THREADS) {
            const int t = newest - 32 * warp - lane;
            uint32_t w = FLAG_PREFIX;                           // tiles before 0: an empty prefix
            if (t >= 0) {
                do { w = look[t]; } while ((w & FLAG_MASK) == 0u);
            }
            const unsigned pref = __ballot_sync(FULL, (w & FLAG_MASK) == FLAG_PREFIX);
            const int stop = __ffs(pref) - 1;                   // nearest published prefix of this window (-1: none)
            const uint32_t v = (stop < 0 || lane <= stop) ? (w & ~FLAG_MASK) : 0u;
            const uint32_t sum = __reduce_add_sync(FULL, v);
            if (lane == 0) s_look[warp] = sum | (stop >= 0 ? 0x80000000u : 0u);     // sums stay below 2^30
            __syncthreads();
            bool done = false;
#pragma unroll
            for (int q = 0; q < EMIT_THREADS / 32; ++q) {
                if (!done) {
                    const uint32_t x = s_look[q];
                    excl += x & 0x7FFFFFFFu;
                    done = (x & 0x80000000u) != 0u;
                }
            }
            if (done) break;
            __syncthreads();                                    // s_look is rewritten in the next round
        }
        if (threadIdx.x == 0) look[tile] = FLAG_PREFIX | (excl + total);
    }
    __syncthreads();                                            // s_off / s_id / s_rect are complete
    const uint32_t base = excl;

    // load-balanced expansion: output slot j belongs to the last item whose offset is <= j
    for (uint32_t j = threadIdx.x; j < total; j += EMIT_THREADS) {
        int lo = 0, hi = EMIT_TILE;            // invariant: s_off[lo] <= j < s_off[hi]
#pragma unroll
        for (int step = 0; step < 11; ++step) {       // log2(EMIT_TILE) = 11
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= j) lo = mid; else hi = mid;
        }
        const ushort4 r = s_rect[lo];
        const uint32_t k = j - s_off[lo];
        const uint32_t sx0 = r.x >> SSHIFT, sx1 = (r.z - 1) >> SSHIFT, sy0 = r.y >> SSHIFT;
        const uint32_t w = sx1 - sx0 + 1;
        const uint32_t sy = sy0 + k / w, sx = sx0 + k % w;      // row-major over the super-tile rectangle
        key_out[base + j] = (sy * (uint32_t)sgrid_x + sx) | (tile_mask(r, sx, sy) << 16);
        val_out[base + j] = s_id[lo];
    }
}

// ---------------------------------------------------------------------------------------
// super-tile boundaries + per-tile counts from the grouped candidate keys
// ---------------------------------------------------------------------------------------
constexpr int CNT_THREADS = 256;
constexpr int CNT_ITEMS = 8;                              // candidates per lane -> 4-bit counters suffice
constexpr int CNT_WARP_RUN = 32 * CNT_ITEMS;

// bit k of the low byte -> bit 4k
__device__ __forceinline__ uint32_t spread8x4(uint32_t x) {
    x &= 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}

__global__ void __launch_bounds__(CNT_THREADS)
count_bounds_kernel(const uint32_t* __restrict__ keys, uint32_t n, int sgrid_x, int grid_x, int grid_y,
                    uint32_t* __restrict__ st_begin, uint32_t* __restrict__ st_end,
                    uint32_t* __restrict__ tile_counts)
{
    const int lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * CNT_THREADS + threadIdx.x) >> 5;
    const uint32_t base = wid * CNT_WARP_RUN;
    if (base >= n) return;
    uint32_t key[CNT_ITEMS];
#pragma unroll
    for (int i = 0; i < CNT_ITEMS; ++i) {
        const uint32_t e = base + i * 32 + lane;
        key[i] = (e < n) ? __ldg(keys + e) : 0xFFFFFFFFu;
    }
    const uint32_t prev_run = (base > 0 && lane == 0) ? (__ldg(keys + base - 1) & 0xFFFFu) : 0xFFFFFFFFu;
    // boundaries: candidate e starts a super-tile if its id differs from candidate e-1's
    uint32_t carry = prev_run;   // id of the element before item i's lane 0
    uint32_t lo4 = 0, hi4 = 0;   // 16 x 4-bit per-tile counters (tiles 0-7, 8-15)
    uint32_t st_min = 0xFFFFFFFFu, st_max = 0u;
#pragma unroll
    for (int i = 0; i < CNT_ITEMS; ++i) {
        const uint32_t e = base + i * 32 + lane;
        const uint32_t st = key[i] & 0xFFFFu;
        uint32_t prev = __shfl_up_sync(FULL, st, 1);
        const uint32_t last = __shfl_sync(FULL, st, 31);
        if (lane == 0) prev = carry;
        carry = last;
        if (e < n) {
            if (e == 0 || st != prev) {
                st_begin[st] = e;
                if (e != 0) st_end[prev] = e;
            }
            if (e == n - 1) st_end[st] = n;
            st_min = min(st_min, st);
            st_max = max(st_max, st);
            lo4 += spread8x4(key[i] >> 16);
            hi4 += spread8x4(key[i] >> 24);
        }
    }
    // whole run inside one super-tile (the common case: runs are 256 long, lists thousands)?
    // the keys are sorted by super-tile, so it is enough to compare the smallest and largest id
    const uint32_t first_st = __reduce_min_sync(FULL, st_min);
    const uint32_t last_st = __reduce_max_sync(FULL, st_max);
    if (first_st == last_st) {
        // widen the nibble counters to 16 bits (warp totals reach 256) and reduce with redux.sync
        uint32_t tot[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t a = (lo4 >> (8 * q)) & 0xFFu, b = (hi4 >> (8 * q)) & 0xFFu;
            tot[q] = __reduce_add_sync(FULL, (a & 0xFu) | ((a >> 4) << 16));          // tiles 2q, 2q+1
            tot[4 + q] = __reduce_add_sync(FULL, (b & 0xFu) | ((b >> 4) << 16));      // tiles 8+2q, 9+2q
        }
        if (lane < 16) {
            uint32_t word = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((lane >> 1) == q) word = tot[q];
            const uint32_t cnt = (lane & 1) ? (word >> 16) : (word & 0xFFFFu);
            const int tx = (int)((first_st % (uint32_t)sgrid_x) << SSHIFT) + (lane & (SSIDE - 1));
            const int ty = (int)((first_st / (uint32_t)sgrid_x) << SSHIFT) + (lane >> SSHIFT);
            if (cnt && tx < grid_x && ty < grid_y) atomicAdd(&tile_counts[ty * grid_x + tx], cnt);
        }
    } else {
        // the run straddles a super-tile boundary: per-candidate updates
#pragma unroll
        for (int i = 0; i < CNT_ITEMS; ++i) {
            const uint32_t e = base + i * 32 + lane;
            if (e >= n) continue;
            const uint32_t st = key[i] & 0xFFFFu;
            const int bx = (int)((st % (uint32_t)sgrid_x) << SSHIFT), by = (int)((st / (uint32_t)sgrid_x) << SSHIFT);
            uint32_t m = key[i] >> 16;
            while (m) {
                const int li = __ffs(m) - 1;
                m &= m - 1;
                atomicAdd(&tile_counts[(by + (li >> SSHIFT)) * grid_x + bx + (li & (SSIDE - 1))], 1u);
            }
        }
    }
}

// single CTA: exclusive scan of the T tile counts -> ranges (rasterizer_impl.cu:116-138:
// [first, last+1), {0,0} for an empty tile); total -> *total_out
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int T, const uint32_t* __restrict__ tile_counts, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ total_out)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int start = 0; start < T; start += 1024) {
        const int t = start + threadIdx.x;
        const uint32_t total = (t < T) ? tile_counts[t] : 0u;
        uint32_t inc = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp[w];
        const uint32_t carry = s_carry;
        const uint32_t first = carry + woff + inc - total;
        if (t < T) ranges[t] = total ? make_uint2(first, first + total) : make_uint2(0u, 0u);
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

// ---------------------------------------------------------------------------------------
// fine level: warp = tile, ordered ballot compaction of the super-tile's candidate list
// ---------------------------------------------------------------------------------------
constexpr int FINE_SEGS = 8;        // the candidate list is cut into 8 contiguous segments ...
constexpr int FINE_THREADS = SSIDE * FINE_SEGS * 32;   // ... one warp per (tile of the row, segment)

// CTA = one tile row (4 tiles) of a super-tile.  Phase 1: every warp counts its tile's hits in its
// segment; phase 2: it compacts them behind the hits of the earlier segments.  32 warps share the
// list through L1, and no warp walks more than 1/8 of it.
__global__ void __launch_bounds__(FINE_THREADS)
fine_write_kernel(int sgrid_x, int grid_x, int grid_y, const uint32_t* __restrict__ st_begin,
                  const uint32_t* __restrict__ st_end, const uint32_t* __restrict__ keys,
                  const uint32_t* __restrict__ vals, const uint2* __restrict__ ranges,
                  uint32_t* __restrict__ point_list)
{
    __shared__ uint32_t s_cnt[SSIDE][FINE_SEGS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = warp & (SSIDE - 1), seg = warp >> SSHIFT;
    const int st = blockIdx.x >> SSHIFT, ly = blockIdx.x & (SSIDE - 1);
    const int tx = ((st % sgrid_x) << SSHIFT) + lx, ty = ((st / sgrid_x) << SSHIFT) + ly;
    const bool live = tx < grid_x && ty < grid_y;
    const uint32_t begin = st_begin[st], end = st_end[st];
    // segment length: a multiple of 32 so every warp reads whole aligned-to-begin chunks
    const uint32_t seg_len = (((end - begin + FINE_SEGS - 1) / FINE_SEGS) + 31u) & ~31u;
    const uint32_t s0 = min(end, begin + seg * seg_len), s1 = min(end, s0 + seg_len);
    const int bit = 16 + ly * SSIDE + lx;
    constexpr int UNROLL = 4;

    uint32_t cnt = 0;
    if (live) {
        for (uint32_t c = s0; c < s1; c += 32 * UNROLL) {
            uint32_t k[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t e = c + u * 32 + lane;
                k[u] = (e < s1) ? __ldg(keys + e) : 0u;
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) cnt += (k[u] >> bit) & 1u;
        }
        cnt = __reduce_add_sync(FULL, cnt);
    }
    if (lane == 0) s_cnt[lx][seg] = cnt;
    __syncthreads();
    if (!live || cnt == 0) return;
    uint32_t run = ranges[ty * grid_x + tx].x;
#pragma unroll
    for (int q = 0; q < FINE_SEGS; ++q)
        if (q < seg) run += s_cnt[lx][q];
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t c = s0; c < s1; c += 32 * UNROLL) {
        uint32_t k[UNROLL], v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t e = c + u * 32 + lane;
            k[u] = (e < s1) ? __ldg(keys + e) : 0u;
            v[u] = (e < s1) ? __ldg(vals + e) : 0u;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const bool hit = (k[u] >> bit) & 1u;
            const unsigned m = __ballot_sync(FULL, hit);
            if (hit) point_list[run + __popc(m & lt)] = v[u];
            run += __popc(m);
        }
    }
}

}  // namespace

int launch_depth_order(int P, GeomState& g, cudaStream_t stream)
{
    // 4 stable 8-bit digits over (depth_bits, index); values start as 0..P-1 (not read); the
    // result lands back in (key_a, val_a)
    return radix_sort_pairs(g.key_a, g.key_b, g.val_a, g.val_b, (size_t)P, 0, 4, true, g.sort_temp, stream);
}

BinningPlan plan_binning(int grid_x, int grid_y)
{
    BinningPlan p;
    p.sshift = SSHIFT;
    p.sgrid_x = (grid_x + SSIDE - 1) >> SSHIFT;
    p.sgrid_y = (grid_y + SSIDE - 1) >> SSHIFT;
    const int n = p.sgrid_x * p.sgrid_y;
    p.sort_passes = n <= 256 ? 1 : 2;
    return p;
}

int emit_blocks(int P) { return (P + EMIT_TILE - 1) / EMIT_TILE; }

int launch_binning(int P, int R, int Rc, const ViewParams& vp, GeomState& g, BinningState& b,
                   ImageState& img, cudaStream_t stream)
{
    const int T = vp.grid_x * vp.grid_y;
    if (R == 0) {
        SEGS_CUDA_CHECK(cudaMemsetAsync(img.ranges, 0, size_t(T) * sizeof(uint2), stream));
        return SEGS_OK;
    }
    const BinningPlan pl = plan_binning(vp.grid_x, vp.grid_y);
    const int nst = pl.sgrid_x * pl.sgrid_y;
    if (nst > 65536) { set_error("image too large for the two-level tile binning (%d x %d tiles)", vp.grid_x, vp.grid_y); return SEGS_ERR_INVALID_ARG; }

    // everything the kernels below accumulate into / poll: one memset
    SEGS_CUDA_CHECK(cudaMemsetAsync(b.zeroed, 0, b.zeroed_bytes, stream));

    // coarse: expand depth-ordered Gaussians into (super-tile | mask, id) candidates ...
    coarse_emit_kernel<<<emit_blocks(P), EMIT_THREADS, 0, stream>>>(g.val_a, g.rect, P, pl.sgrid_x, b.emit_ticket,
                                                                   b.emit_look, b.cand_key_a, b.cand_val_a);
    SEGS_LAUNCH_CHECK();
    // ... and group them by super-tile (stable; the mask bits ride along in the key)
    int rc = radix_sort_pairs(b.cand_key_a, b.cand_key_b, b.cand_val_a, b.cand_val_b, (size_t)Rc, 0, pl.sort_passes,
                              false, b.sort_temp, stream);
    if (rc) return rc;
    const uint32_t* keys = (pl.sort_passes & 1) ? b.cand_key_b : b.cand_key_a;
    const uint32_t* vals = (pl.sort_passes & 1) ? b.cand_val_b : b.cand_val_a;

    const int cnt_warps = (Rc + CNT_WARP_RUN - 1) / CNT_WARP_RUN;
    count_bounds_kernel<<<(cnt_warps * 32 + CNT_THREADS - 1) / CNT_THREADS, CNT_THREADS, 0, stream>>>(
        keys, (uint32_t)Rc, pl.sgrid_x, vp.grid_x, vp.grid_y, b.st_begin, b.st_end, b.tile_counts);
    SEGS_LAUNCH_CHECK();
    tile_scan_kernel<<<1, 1024, 0, stream>>>(T, b.tile_counts, img.ranges, g.counters + 3);
    SEGS_LAUNCH_CHECK();
    fine_write_kernel<<<nst * SSIDE, FINE_THREADS, 0, stream>>>(pl.sgrid_x, vp.grid_x, vp.grid_y, b.st_begin, b.st_end,
                                                                  keys, vals, img.ranges, b.point_list);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // namespace segs
