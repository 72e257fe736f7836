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
//   0. a stable 4-pass radix sort of the P Gaussians by depth_bits (P is ~8x smaller than R);
//   1. coarse level: the image is cut into super-tiles of SxS tiles (S a power of two chosen
//      so that there are <= 256 super-tiles).  The depth-ordered Gaussians are expanded into
//      (super-tile, Gaussian) candidates (R' ~ 0.25 R) and ONE stable radix pass on the
//      super-tile id groups them: every super-tile now has its depth-ordered candidate list;
//   2. fine level: one CTA per super-tile, one warp per tile.  The CTA streams its candidate
//      list through shared memory; each warp tests 32 candidates per step against its tile
//      and stream-compacts the hits with a ballot — compaction preserves the list order, so
//      every tile's list is in (depth_bits, Gaussian index) order, exactly the reference's
//      sorted order, bit for bit.  A count pass + a scan over the T tile counts gives
//      `ranges`; the write pass emits point_list (the only R-sized traffic: 4 B/instance).
//
// Radix pass (depth sort, coarse pass, kNN Morton sort) = histogram kernel (per-CTA digit
// counts, bin-major) + one CTA per bin turning them into scatter bases + a scatter kernel that
// ranks with warp __match_any_sync (stable, no atomics on the ranking path).
#include "common.cuh"

namespace segs {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
static_assert(SORT_THREADS == RADIX_BINS, "one thread per radix bin is assumed");

// ---------------------------------------------------------------------------------------
// radix pass
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ key_in, size_t n, int shift, int nblocks,
                  uint32_t* __restrict__ block_hist, uint32_t* __restrict__ global_hist)
{
    __shared__ uint32_t hist[RADIX_BINS];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = size_t(blockIdx.x) * SORT_TILE;
    uint32_t k[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const size_t e = base + size_t(i) * SORT_THREADS + threadIdx.x;
        k[i] = (e < n) ? __ldg(key_in + e) : 0u;
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const size_t e = base + size_t(i) * SORT_THREADS + threadIdx.x;
        if (e < n) atomicAdd(&hist[(k[i] >> shift) & (RADIX_BINS - 1)], 1u);
    }
    __syncthreads();
    const uint32_t c = hist[threadIdx.x];
    block_hist[size_t(threadIdx.x) * nblocks + blockIdx.x] = c;
    if (c) atomicAdd(&global_hist[threadIdx.x], c);
}

// One CTA per bin: exclusive scan of that bin's per-CTA counts, offset by the total of all
// lower bins.  In place.
__global__ void __launch_bounds__(256)
radix_scan_kernel(uint32_t* __restrict__ block_hist, const uint32_t* __restrict__ global_hist, int nblocks)
{
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_carry;
    const int bin = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    uint32_t v = (threadIdx.x < bin) ? global_hist[threadIdx.x] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += s_warp[w];
        s_carry = t;
    }
    __syncthreads();

    uint32_t* row = block_hist + size_t(bin) * nblocks;
    for (int start = 0; start < nblocks; start += 256) {
        const int i = start + threadIdx.x;
        const uint32_t x = (i < nblocks) ? row[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
        }
        __syncthreads();                   // s_warp / s_carry from the previous round consumed
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp[w];
        const uint32_t carry = s_carry;
        if (i < nblocks) row[i] = carry + woff + inc - x;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = carry + woff + inc;
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ key_in, uint32_t* __restrict__ key_out,
                     const uint32_t* __restrict__ val_in, uint32_t* __restrict__ val_out,
                     size_t n, int shift, int nblocks, const uint32_t* __restrict__ block_hist)
{
    constexpr int WARPS = SORT_THREADS / 32;
    __shared__ uint32_t warp_hist[WARPS][RADIX_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < WARPS * RADIX_BINS; i += SORT_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();

    // warp-striped: warp w owns [wbase, wbase + 32*ITEMS); item i of lane l is wbase + 32*i + l
    const size_t wbase = size_t(blockIdx.x) * SORT_TILE + size_t(warp) * (32 * SORT_ITEMS);
    uint32_t key[SORT_ITEMS], val[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        const bool valid = e < n;
        key[i] = valid ? __ldg(key_in + e) : 0u;
        val[i] = valid ? __ldg(val_in + e) : 0u;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        const bool valid = e < n;
        const uint32_t digit = valid ? ((key[i] >> shift) & (RADIX_BINS - 1)) : RADIX_BINS;
        const uint32_t peers = __match_any_sync(FULL, digit);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = warp_hist[warp][digit];
            warp_hist[warp][digit] = old + __popc(peers);
        }
        old = __shfl_sync(FULL, old, leader);
        rank[i] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread t owns bin t: turn per-warp counts into scatter bases
        const int bin = threadIdx.x;
        uint32_t running = block_hist[size_t(bin) * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = warp_hist[w][bin];
            warp_hist[w][bin] = running;
            running += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const size_t e = wbase + size_t(i) * 32 + lane;
        if (e < n) {
            const uint32_t digit = (key[i] >> shift) & (RADIX_BINS - 1);
            const uint32_t pos = warp_hist[warp][digit] + rank[i];
            key_out[pos] = key[i];
            val_out[pos] = val[i];
        }
    }
}

// ---------------------------------------------------------------------------------------
// exclusive scan of the coarse candidate counts in depth order
// ---------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// number of super-tiles a tile rectangle overlaps (0 for an empty rectangle)
__device__ __forceinline__ uint32_t coarse_count(ushort4 rc, int sshift) {
    if (rc.x >= rc.z || rc.y >= rc.w) return 0u;
    const uint32_t sx0 = rc.x >> sshift, sx1 = (rc.z - 1) >> sshift;
    const uint32_t sy0 = rc.y >> sshift, sy1 = (rc.w - 1) >> sshift;
    return (sx1 - sx0 + 1) * (sy1 - sy0 + 1);
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t x, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < warp) woff += c;
        tot += c;
    }
    *total = tot;
    __syncthreads();
    return woff + inc - x;
}

__global__ void __launch_bounds__(SCAN_THREADS)
coarse_reduce_kernel(const uint32_t* __restrict__ order, const ushort4* __restrict__ rect, int P, int sshift,
                     uint32_t* __restrict__ partials)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < P) sum += coarse_count(__ldg(rect + __ldg(order + base + i)), sshift);
    uint32_t total;
    block_exclusive_scan(sum, s_warp, &total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
coarse_partials_kernel(uint32_t* __restrict__ partials, int nb)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    uint32_t carry = 0;
    for (int start = 0; start < nb; start += SCAN_THREADS) {
        const int i = start + threadIdx.x;
        const uint32_t x = (i < nb) ? partials[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(x, s_warp, &total);
        if (i < nb) partials[i] = carry + ex;
        carry += total;
    }
}

// scan + emission fused: candidate j of depth-ordered Gaussian k goes to offsets[k] + j
// (row-major over its super-tile rectangle), key = super-tile id, value = Gaussian id
__global__ void __launch_bounds__(SCAN_THREADS)
coarse_emit_kernel(const uint32_t* __restrict__ order, const ushort4* __restrict__ rect, int P, int sshift,
                   int sgrid_x, const uint32_t* __restrict__ partials, uint32_t* __restrict__ key_out,
                   uint32_t* __restrict__ val_out)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t id[SCAN_ITEMS], c[SCAN_ITEMS];
    ushort4 rc[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        c[i] = 0;
        if (base + i < P) {
            id[i] = __ldg(order + base + i);
            rc[i] = __ldg(rect + id[i]);
            c[i] = coarse_count(rc[i], sshift);
        }
        sum += c[i];
    }
    uint32_t total;
    uint32_t run = partials[blockIdx.x] + block_exclusive_scan(sum, s_warp, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (c[i]) {
            const uint32_t sx0 = rc[i].x >> sshift, sx1 = (rc[i].z - 1) >> sshift;
            const uint32_t sy0 = rc[i].y >> sshift, sy1 = (rc[i].w - 1) >> sshift;
            for (uint32_t sy = sy0; sy <= sy1; ++sy)
                for (uint32_t sx = sx0; sx <= sx1; ++sx) {
                    key_out[run] = sy * (uint32_t)sgrid_x + sx;
                    val_out[run] = id[i];
                    ++run;
                }
        }
    }
}

// ---------------------------------------------------------------------------------------
// fine level: CTA = super-tile, warp = tile(s); ordered ballot compaction
// ---------------------------------------------------------------------------------------
constexpr int FINE_MAX_THREADS = 1024;

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// The candidate list is streamed in batches of blockDim.x candidates, double-buffered: the ids of
// batch b+2 are prefetched into registers and the 8-byte rectangles of batch b+1 are gathered
// with cp.async while the warps test batch b.
template <bool WRITE>
__global__ void __launch_bounds__(FINE_MAX_THREADS)
fine_bin_kernel(int sshift, int sgrid_x, int grid_x, int grid_y, const uint32_t* __restrict__ coarse_hist,
                const uint32_t* __restrict__ cand_id, const ushort4* __restrict__ rect,
                uint32_t* __restrict__ tile_counts, const uint2* __restrict__ ranges,
                uint32_t* __restrict__ point_list)
{
    extern __shared__ __align__(16) unsigned char s_fine[];
    const int NT = blockDim.x;
    ushort4* s_rect = reinterpret_cast<ushort4*>(s_fine);                  // [2][NT]
    uint32_t* s_id = reinterpret_cast<uint32_t*>(s_fine + size_t(2) * NT * sizeof(ushort4));   // [2][NT]
    __shared__ uint32_t s_red[32];
    __shared__ uint32_t s_start, s_len;
    const int st = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    // candidate range of this super-tile = prefix of the coarse histogram (<= 256 bins)
    {
        uint32_t v = 0;
        for (int i = tid; i < st; i += blockDim.x) v += coarse_hist[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < nwarps; ++w) t += s_red[w];
            s_start = t;
            s_len = coarse_hist[st];
        }
        __syncthreads();
    }
    const uint32_t clen = s_len;
    const uint32_t* list = cand_id + s_start;
    const int nbatches = (int)((clen + NT - 1) / NT);

    // tiles of this super-tile handled by this warp: local index li = warp, warp + nwarps, ...
    // (at most MAX_TPW per warp; kept in registers, all loops over them are fully unrolled)
    const int S = 1 << sshift;
    const int stx = (st % sgrid_x) << sshift, sty = (st / sgrid_x) << sshift;
    constexpr int MAX_TPW = 8;
    uint32_t run[MAX_TPW];
    unsigned ttx[MAX_TPW], tty[MAX_TPW];
    bool live[MAX_TPW];
#pragma unroll
    for (int q = 0; q < MAX_TPW; ++q) {
        const int li = warp + q * nwarps;
        const int tx = stx + (li & (S - 1)), ty = sty + (li >> sshift);
        live[q] = li < S * S && tx < grid_x && ty < grid_y;
        ttx[q] = (unsigned)tx; tty[q] = (unsigned)ty;
        run[q] = (WRITE && live[q]) ? ranges[ty * grid_x + tx].x : 0u;
    }
    const uint32_t lt = (1u << lane) - 1u;

    // pipeline prologue: batch 0 rectangles in flight, batch 1 ids in registers
    uint32_t id_next = ((uint32_t)tid < clen) ? __ldg(list + tid) : 0u;
    if ((uint32_t)tid < clen) {
        s_id[tid] = id_next;
        cp_async8(&s_rect[tid], rect + id_next);
    }
    cp_async_commit_group();
    id_next = ((uint32_t)(NT + tid) < clen) ? __ldg(list + NT + tid) : 0u;

    for (int b = 0; b < nbatches; ++b) {
        const int buf = b & 1;
        const uint32_t n_in = min((uint32_t)NT, clen - (uint32_t)b * NT);
        if (b + 1 < nbatches) {
            const uint32_t p = (uint32_t)(b + 1) * NT + tid;
            if (p < clen) {
                s_id[(buf ^ 1) * NT + tid] = id_next;
                cp_async8(&s_rect[(buf ^ 1) * NT + tid], rect + id_next);
            }
            cp_async_commit_group();
            const uint32_t p2 = (uint32_t)(b + 2) * NT + tid;
            id_next = (p2 < clen) ? __ldg(list + p2) : 0u;
            cp_async_wait_group<1>();
        } else {
            cp_async_wait_group<0>();
        }
        __syncthreads();
        const ushort4* b_rect = s_rect + buf * NT;
        const uint32_t* b_id = s_id + buf * NT;
        for (uint32_t c = 0; c < n_in; c += 32) {
            const uint32_t slot = c + lane;
            ushort4 rc = make_ushort4(1, 1, 0, 0);      // empty rectangle: never hits
            uint32_t id = 0;
            if (slot < n_in) {
                rc = b_rect[slot];
                if (WRITE) id = b_id[slot];
            }
#pragma unroll
            for (int q = 0; q < MAX_TPW; ++q) {
                if (live[q]) {                              // warp-uniform
                    const bool hit = (ttx[q] >= rc.x) && (ttx[q] < rc.z) && (tty[q] >= rc.y) && (tty[q] < rc.w);
                    const unsigned mask = __ballot_sync(FULL, hit);
                    if (WRITE && hit) point_list[run[q] + __popc(mask & lt)] = id;
                    run[q] += __popc(mask);
                }
            }
        }
        __syncthreads();   // everyone is done with this buffer before it is refilled
    }
    if (!WRITE && lane == 0) {
#pragma unroll
        for (int q = 0; q < MAX_TPW; ++q)
            if (live[q]) tile_counts[tty[q] * grid_x + ttx[q]] = run[q];
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

}  // namespace

int radix_pass(const uint32_t* key_in, uint32_t* key_out, const uint32_t* val_in,
               uint32_t* val_out, size_t n, int shift, uint32_t* block_hist,
               uint32_t* global_hist, cudaStream_t stream)
{
    if (n == 0) return SEGS_OK;
    const int nblocks = sort_blocks(n);
    SEGS_CUDA_CHECK(cudaMemsetAsync(global_hist, 0, RADIX_BINS * sizeof(uint32_t), stream));
    radix_hist_kernel<<<nblocks, SORT_THREADS, 0, stream>>>(key_in, n, shift, nblocks, block_hist, global_hist);
    SEGS_LAUNCH_CHECK();
    radix_scan_kernel<<<RADIX_BINS, 256, 0, stream>>>(block_hist, global_hist, nblocks);
    SEGS_LAUNCH_CHECK();
    radix_scatter_kernel<<<nblocks, SORT_THREADS, 0, stream>>>(key_in, key_out, val_in, val_out, n, shift,
                                                              nblocks, block_hist);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int launch_depth_order(int P, GeomState& g, cudaStream_t stream)
{
    // 4 stable 8-bit passes over (depth_bits, index): a -> b -> a -> b -> a
    int rc;
    if ((rc = radix_pass(g.key_a, g.key_b, g.val_a, g.val_b, P, 0, g.block_hist, g.global_hist, stream))) return rc;
    if ((rc = radix_pass(g.key_b, g.key_a, g.val_b, g.val_a, P, 8, g.block_hist, g.global_hist, stream))) return rc;
    if ((rc = radix_pass(g.key_a, g.key_b, g.val_a, g.val_b, P, 16, g.block_hist, g.global_hist, stream))) return rc;
    if ((rc = radix_pass(g.key_b, g.key_a, g.val_b, g.val_a, P, 24, g.block_hist, g.global_hist, stream))) return rc;
    return SEGS_OK;
}

// Super-tile side (in tiles, power of two) such that there are at most 256 super-tiles, i.e.
// one 8-bit radix pass groups the candidates.  Deterministic in the tile grid.
BinningPlan plan_binning(int grid_x, int grid_y)
{
    BinningPlan p;
    p.sshift = 1;
    while (((grid_x + (1 << p.sshift) - 1) >> p.sshift) * ((grid_y + (1 << p.sshift) - 1) >> p.sshift) > RADIX_BINS)
        ++p.sshift;
    p.sgrid_x = (grid_x + (1 << p.sshift) - 1) >> p.sshift;
    p.sgrid_y = (grid_y + (1 << p.sshift) - 1) >> p.sshift;
    return p;
}

int launch_binning(int P, int R, int Rc, const ViewParams& vp, GeomState& g, BinningState& b,
                   ImageState& img, cudaStream_t stream)
{
    const int T = vp.grid_x * vp.grid_y;
    if (R == 0) {
        SEGS_CUDA_CHECK(cudaMemsetAsync(img.ranges, 0, size_t(T) * sizeof(uint2), stream));
        return SEGS_OK;
    }
    const BinningPlan pl = plan_binning(vp.grid_x, vp.grid_y);
    if (pl.sshift > 4) { set_error("image too large for the two-level tile binning (%d x %d tiles)", vp.grid_x, vp.grid_y); return SEGS_ERR_INVALID_ARG; }

    // coarse: expand depth-ordered Gaussians into (super-tile, id) candidates ...
    const int nb = (P + SCAN_TILE - 1) / SCAN_TILE;
    coarse_reduce_kernel<<<nb, SCAN_THREADS, 0, stream>>>(g.val_a, g.rect, P, pl.sshift, b.partials);
    SEGS_LAUNCH_CHECK();
    coarse_partials_kernel<<<1, SCAN_THREADS, 0, stream>>>(b.partials, nb);
    SEGS_LAUNCH_CHECK();
    coarse_emit_kernel<<<nb, SCAN_THREADS, 0, stream>>>(g.val_a, g.rect, P, pl.sshift, pl.sgrid_x, b.partials,
                                                        b.cand_key_a, b.cand_val_a);
    SEGS_LAUNCH_CHECK();
    // ... and group them by super-tile with one stable 8-bit pass (global_hist = candidates per super-tile)
    int rc = radix_pass(b.cand_key_a, b.cand_key_b, b.cand_val_a, b.cand_val_b, (size_t)Rc, 0, b.block_hist,
                        b.global_hist, stream);
    if (rc) return rc;

    // fine: per tile count -> ranges -> ordered write
    const int nst = pl.sgrid_x * pl.sgrid_y;
    const int tiles_per_st = 1 << (2 * pl.sshift);
    // warps per CTA: each warp owns up to 8 tiles of the super-tile
    const int warps = tiles_per_st <= 4 ? 4 : (tiles_per_st <= 128 ? 16 : 32);
    const size_t fine_smem = size_t(2) * warps * 32 * (sizeof(ushort4) + sizeof(uint32_t));
    fine_bin_kernel<false><<<nst, warps * 32, fine_smem, stream>>>(pl.sshift, pl.sgrid_x, vp.grid_x, vp.grid_y, b.global_hist,
                                                          b.cand_val_b, g.rect, b.tile_counts, nullptr, nullptr);
    SEGS_LAUNCH_CHECK();
    tile_scan_kernel<<<1, 1024, 0, stream>>>(T, b.tile_counts, img.ranges, g.counters + 3);
    SEGS_LAUNCH_CHECK();
    fine_bin_kernel<true><<<nst, warps * 32, fine_smem, stream>>>(pl.sshift, pl.sgrid_x, vp.grid_x, vp.grid_y, b.global_hist,
                                                         b.cand_val_b, g.rect, nullptr, img.ranges, b.point_list);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // namespace segs
