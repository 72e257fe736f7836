// Depth ordering, instance emission, tile binning and tile ranges for sm_100a
// (SURVEY §8 rows F2-F6).
//
// Replaces cub::DeviceScan::InclusiveSum, duplicateWithKeys, the 64-bit-key
// cub::DeviceRadixSort::SortPairs and identifyTileRanges of the reference
// (cuda_rasterizer/rasterizer_impl.cu:70-138, 276-318).
//
// The reference sorts R (Gaussian x tile) instances by a 64-bit key  tile<<32 | depth_bits
// with a stable LSD radix sort over 32+msb(T) bits (6 CUB passes of 24 B/instance at C2).
// Here the same total order (tile, depth_bits, Gaussian index) is produced as
//   1. a stable radix sort of the P Gaussians by depth_bits (4 passes over 8 B/Gaussian,
//      P is ~8x smaller than R),
//   2. emission of the instances in that depth order (instances of one Gaussian are
//      row-major over its tile rectangle, as in duplicateWithKeys),
//   3. a stable radix sort of the instances by tile id only: ceil(log256(T)) passes of
//      8 B/instance (2 passes for every image up to 4096x4096).
// Stability of each step makes ties resolve by ascending Gaussian index, exactly like the
// reference's single stable sort, so point_list and ranges are bit-identical.
//
// Radix pass = histogram kernel (per-CTA digit counts, bin-major) + one CTA per bin that
// turns them into global scatter bases + a scatter kernel that ranks with warp
// __match_any_sync (stable, no atomics on the ranking path).
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
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const size_t e = base + size_t(i) * SORT_THREADS + threadIdx.x;
        if (e < n) atomicAdd(&hist[(__ldg(key_in + e) >> shift) & (RADIX_BINS - 1)], 1u);
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

    // base = sum of global_hist[0..bin)
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
        // inclusive warp scan
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
// exclusive scan of tiles_touched in depth order
// ---------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

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
scan_reduce_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles_touched,
                   int P, uint32_t* __restrict__ partials)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < P) sum += __ldg(tiles_touched + __ldg(order + base + i));
    uint32_t total;
    block_exclusive_scan(sum, s_warp, &total);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

// single CTA: exclusive scan of the per-CTA sums; total -> counters[0]
__global__ void __launch_bounds__(SCAN_THREADS)
scan_partials_kernel(uint32_t* __restrict__ partials, int nb, uint32_t* __restrict__ counters)
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
    if (threadIdx.x == 0) counters[0] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles_touched,
                  int P, const uint32_t* __restrict__ partials, uint32_t* __restrict__ offsets)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t c[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        c[i] = (base + i < P) ? __ldg(tiles_touched + __ldg(order + base + i)) : 0u;
        sum += c[i];
    }
    uint32_t total;
    uint32_t run = partials[blockIdx.x] + block_exclusive_scan(sum, s_warp, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < P) offsets[base + i] = run;
        run += c[i];
    }
}

// ---------------------------------------------------------------------------------------
// instance emission in depth order (duplicateWithKeys, rasterizer_impl.cu:70-111)
// ---------------------------------------------------------------------------------------
// One warp per 32 consecutive depth-ordered Gaussians; the warp walks the flat range of
// their instances so every store is a full coalesced 128-byte line regardless of how the
// instance counts are distributed (one Gaussian covers 1..900+ tiles).
constexpr int EMIT_THREADS = 256;

__global__ void __launch_bounds__(EMIT_THREADS)
emit_instances_kernel(int P, const uint32_t* __restrict__ order, const uint32_t* __restrict__ offsets,
                      const uint32_t* __restrict__ tiles_touched, const ushort4* __restrict__ rect,
                      int grid_x, uint32_t* __restrict__ tile_out, uint32_t* __restrict__ idx_out)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * EMIT_THREADS + threadIdx.x) >> 5;
    const int k = gwarp * 32 + lane;
    uint32_t id = 0, off = 0, cnt = 0;
    ushort4 rc = make_ushort4(0, 0, 0, 0);
    if (k < P) {
        id = __ldg(order + k);
        off = __ldg(offsets + k);
        cnt = __ldg(tiles_touched + id);
        rc = __ldg(rect + id);
    }
    const uint32_t begin = __shfl_sync(FULL, off, 0);
    const uint32_t end = __reduce_max_sync(FULL, (k < P) ? off + cnt : 0u);
    // lanes past P must never win the search below
    if (k >= P) off = 0xFFFFFFFFu;
    for (uint32_t sb = begin; sb < end; sb += 32) {          // warp-uniform trip count
        const uint32_t s = sb + lane;
        // largest j with off_j <= s (offsets are non-decreasing over lanes).  Gaussians with
        // cnt == 0 share their offset with the next one, so the LAST such lane owns s.
        int j = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const uint32_t o = __shfl_sync(FULL, off, (j + step) & 31);
            if (o <= s) j += step;                           // j + step <= 31 by construction
        }
        const uint32_t oj = __shfl_sync(FULL, off, j);
        const uint32_t idj = __shfl_sync(FULL, id, j);
        const uint32_t x0 = __shfl_sync(FULL, (uint32_t)rc.x, j);
        const uint32_t y0 = __shfl_sync(FULL, (uint32_t)rc.y, j);
        const uint32_t x1 = __shfl_sync(FULL, (uint32_t)rc.z, j);
        if (s < end) {
            const uint32_t w = x1 - x0;
            const uint32_t local = s - oj;
            const uint32_t ty = local / w, tx = local - ty * w;
            tile_out[s] = (y0 + ty) * (uint32_t)grid_x + (x0 + tx);
            idx_out[s] = idj;
        }
    }
}

// identifyTileRanges (rasterizer_impl.cu:116-138) on the sorted 32-bit tile ids
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int L, const uint32_t* __restrict__ tile_sorted, uint2* __restrict__ ranges)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = __ldg(tile_sorted + idx);
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = __ldg(tile_sorted + idx - 1);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
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
    // exclusive offsets of tiles_touched in depth order; total -> counters[0]
    const int nb = (P + SCAN_TILE - 1) / SCAN_TILE;
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, stream>>>(g.val_a, g.tiles_touched, P, g.scan_partials);
    SEGS_LAUNCH_CHECK();
    scan_partials_kernel<<<1, SCAN_THREADS, 0, stream>>>(g.scan_partials, nb, g.counters);
    SEGS_LAUNCH_CHECK();
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, stream>>>(g.val_a, g.tiles_touched, P, g.scan_partials, g.offsets);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int launch_binning(int P, int R, const ViewParams& vp, GeomState& g, BinningState& b,
                   ImageState& img, cudaStream_t stream)
{
    const int T = vp.grid_x * vp.grid_y;
    SEGS_CUDA_CHECK(cudaMemsetAsync(img.ranges, 0, size_t(T) * sizeof(uint2), stream));
    if (R == 0) return SEGS_OK;
    const int warps = (P + 31) / 32;
    const int blocks = (warps * 32 + EMIT_THREADS - 1) / EMIT_THREADS;
    emit_instances_kernel<<<blocks, EMIT_THREADS, 0, stream>>>(P, g.val_a, g.offsets, g.tiles_touched, g.rect,
                                                             vp.grid_x, b.tile_a, b.idx_a);
    SEGS_LAUNCH_CHECK();
    const int passes = num_tile_passes((uint32_t)T);
    uint32_t *ka = b.tile_a, *kb = b.tile_b, *va = b.idx_a, *vb = b.idx_b;
    for (int p = 0; p < passes; ++p) {
        int rc = radix_pass(ka, kb, va, vb, (size_t)R, 8 * p, b.block_hist, b.global_hist, stream);
        if (rc) return rc;
        uint32_t* t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    // sorted tile ids now in ka, point_list in va
    tile_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>(R, ka, img.ranges);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // namespace segs
