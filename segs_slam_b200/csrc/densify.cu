// Densification decisions of the anchor model for sm_100a.
//
// Replaces the ATen op sequences of GaussianModel::adjust_anchor / anchor_growing / prune_anchor
// (/root/reference/src/gaussian_model.cpp:1505-1762): ~120 small ATen launches per call, an O(U x A) chunked
// duplicate test (:1601-1616, 4096 anchors per chunk) and a torch_scatter scatter_max (:1635).
//
// Design (B200): the candidate offsets of one growing level are compacted, keyed by their integer voxel coordinate
// and ordered with three stable LSD rounds of the library's own radix sort (z, y, x with the sign bit flipped =
// at::unique_dim's lexicographic order); runs of equal keys are the unique voxels.  The duplicate test against the
// existing anchors is a binary search of every anchor's voxel in that sorted list (O(A log U) instead of O(U x A)),
// and scatter_max is a per-voxel maximum over its contiguous run — deterministic, no float atomics.  Pruning is one
// fused statistics pass + a scan, then row gathers.  Two host read-backs per growing level (candidate count, new
// anchor count), like the reference's boolean indexing; densification runs once per update_interval iterations.
//
// Arithmetic that decides integer outputs follows ATen's CUDA kernels operation by operation: `x / cur_size` with a
// host scalar is a multiplication by the FP32-rounded reciprocal (BinaryDivTrueKernel.cu), torch::round is
// round-half-even (rintf), exp is the accurate expf, `anchor + offset * scaling` is a separately rounded multiply and add.
#include "common.cuh"

namespace segs {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int DT = 256;

inline int blocks_for(size_t n, int t = DT) { return int((n + t - 1) / t); }

// ---- exclusive scan of u32 (n up to a few million): block sums -> one-block scan -> add ---------------------------
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = DT * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < DT / 32; ++w) {
        const uint32_t x = s_warp[w];
        if (w < warp) woff += x;
        tot += x;
    }
    if (total) *total = tot;
    return woff + inc - v;
}

// `n` is read from the device when n_dev != nullptr (sizes that never travel to the host)
__global__ void __launch_bounds__(DT)
scan_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ tile_sums, uint32_t n_host,
                 const uint32_t* __restrict__ n_dev)
{
    __shared__ uint32_t s_warp[DT / 32];
    const uint32_t n = n_dev ? min(*n_dev, n_host) : n_host;
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0u; sum += v[i]; }
    uint32_t tot;
    uint32_t run = block_exclusive_scan(sum, s_warp, &tot);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(DT)
scan_sums_kernel(uint32_t* __restrict__ tile_sums, int tiles, uint32_t* __restrict__ total_out)
{
    __shared__ uint32_t s_warp[DT / 32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int start = 0; start < tiles; start += DT) {
        const int t = start + threadIdx.x;
        const uint32_t v = t < tiles ? tile_sums[t] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan(v, s_warp, &tot);
        const uint32_t carry = s_carry;
        if (t < tiles) tile_sums[t] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

__global__ void __launch_bounds__(DT)
scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_sums, uint32_t n_host, const uint32_t* __restrict__ n_dev)
{
    const uint32_t n = n_dev ? min(*n_dev, n_host) : n_host;
    const uint32_t off = tile_sums[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) out[base + i] += off;
}

inline size_t scan_temp_words(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

int exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, const uint32_t* n_dev, uint32_t* temp, uint32_t* total_out,
                   cudaStream_t stream)
{
    const int tiles = int((size_t(n) + SCAN_TILE - 1) / SCAN_TILE);
    if (tiles == 0) { SEGS_CUDA_CHECK(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), stream)); return SEGS_OK; }
    scan_tile_kernel<<<tiles, DT, 0, stream>>>(in, out, temp, n, n_dev);
    SEGS_LAUNCH_CHECK();
    scan_sums_kernel<<<1, DT, 0, stream>>>(temp, tiles, total_out);
    SEGS_LAUNCH_CHECK();
    scan_add_kernel<<<tiles, DT, 0, stream>>>(out, temp, n, n_dev);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

// ---- anchor_growing, one level ------------------------------------------------------------------------------------
// candidate_mask of gaussian_model.cpp:1562-1568: grads >= cur_threshold && offset_mask && rand > 0.5^(i+1), with
// grads = |offset_gradient_accum / offset_denom| (NaN -> 0, :1710-1712) and offset_mask = offset_denom > threshold (:1713)
__global__ void __launch_bounds__(DT)
grow_mark_kernel(int init_slots, const float* __restrict__ accum, const float* __restrict__ denom, const float* __restrict__ rnd,
                 float denom_threshold, float cur_threshold, float rand_threshold, uint32_t* __restrict__ flag)
{
    const int s = blockIdx.x * DT + threadIdx.x;
    if (s >= init_slots) return;
    const float d = denom[s];
    float g = __fdiv_rn(accum[s], d);
    if (g != g) g = 0.f;
    g = fabsf(g);
    flag[s] = (g >= cur_threshold && d > denom_threshold && rnd[s] > rand_threshold) ? 1u : 0u;
}

__device__ __forceinline__ int voxel_of(float x, float inv_size) { return (int)rintf(__fmul_rn(x, inv_size)); }

// selected_xyz of :1582-1590 and its voxel coordinate, written at the candidate's compacted position
__global__ void __launch_bounds__(DT)
grow_keys_kernel(int init_slots, int n_offsets, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                 const float* __restrict__ anchor, const float* __restrict__ offset, const float* __restrict__ log_scaling,
                 float inv_size, uint32_t* __restrict__ cand_slot, int* __restrict__ kx, int* __restrict__ ky, int* __restrict__ kz)
{
    const int s = blockIdx.x * DT + threadIdx.x;
    if (s >= init_slots || !flag[s]) return;
    const int a = s / n_offsets;
    const uint32_t j = pos[s];
    int k[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float sc = expf(log_scaling[6 * a + c]);                                   // get_scaling()[:, :3]
        const float xyz = __fadd_rn(anchor[3 * a + c], __fmul_rn(offset[3 * size_t(s) + c], sc));
        k[c] = voxel_of(xyz, inv_size);
    }
    cand_slot[j] = (uint32_t)s;
    kx[j] = k[0]; ky[j] = k[1]; kz[j] = k[2];
}

// sort key of one LSD round: component `comp` of the candidate at the current permutation, sign bit flipped
__global__ void __launch_bounds__(DT)
grow_round_key_kernel(uint32_t M, const int* __restrict__ comp, const uint32_t* __restrict__ perm, uint32_t* __restrict__ key)
{
    const uint32_t i = blockIdx.x * DT + threadIdx.x;
    if (i >= M) return;
    const uint32_t j = perm ? perm[i] : i;
    key[i] = (uint32_t)comp[j] ^ 0x80000000u;
}

__global__ void __launch_bounds__(DT)
grow_heads_kernel(uint32_t M, const uint32_t* __restrict__ perm, const int* __restrict__ kx, const int* __restrict__ ky,
                  const int* __restrict__ kz, uint32_t* __restrict__ head)
{
    const uint32_t i = blockIdx.x * DT + threadIdx.x;
    if (i >= M) return;
    uint32_t h = 1u;
    if (i > 0) {
        const uint32_t a = perm[i], b = perm[i - 1];
        h = (kx[a] != kx[b] || ky[a] != ky[b] || kz[a] != kz[b]) ? 1u : 0u;
    }
    head[i] = h;
}

// ufirst[u] = first sorted position of unique voxel u (uid = exclusive scan of head + head - 1)
__global__ void __launch_bounds__(DT)
grow_first_kernel(uint32_t M, const uint32_t* __restrict__ head, const uint32_t* __restrict__ head_scan, uint32_t* __restrict__ ufirst)
{
    const uint32_t i = blockIdx.x * DT + threadIdx.x;
    if (i >= M || !head[i]) return;
    ufirst[head_scan[i]] = i;
}

// remove_duplicates of :1596-1621: a unique voxel is dropped when ANY existing anchor's voxel equals it.  Every anchor
// binary-searches the lexicographically sorted unique list.
__global__ void __launch_bounds__(DT)
grow_dup_kernel(int A_now, const float* __restrict__ anchor, float inv_size, const uint32_t* __restrict__ n_unique,
                const uint32_t* __restrict__ ufirst, const uint32_t* __restrict__ perm, const int* __restrict__ kx,
                const int* __restrict__ ky, const int* __restrict__ kz, uint32_t* __restrict__ keep)
{
    const int a = blockIdx.x * DT + threadIdx.x;
    if (a >= A_now) return;
    const int gx = voxel_of(anchor[3 * a], inv_size), gy = voxel_of(anchor[3 * a + 1], inv_size), gz = voxel_of(anchor[3 * a + 2], inv_size);
    int lo = 0, hi = (int)*n_unique - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t j = perm[ufirst[mid]];
        const int x = kx[j], y = ky[j], z = kz[j];
        const int cmp = (x != gx) ? (x < gx ? -1 : 1) : (y != gy) ? (y < gy ? -1 : 1) : (z != gz) ? (z < gz ? -1 : 1) : 0;
        if (cmp == 0) { keep[mid] = 0u; return; }
        if (cmp < 0) lo = mid + 1; else hi = mid - 1;
    }
}

__global__ void __launch_bounds__(DT)
fill_u32_kernel(uint32_t n, uint32_t v, uint32_t* __restrict__ p)
{
    const uint32_t i = blockIdx.x * DT + threadIdx.x;
    if (i < n) p[i] = v;
}

// candidate_anchor (:1623) and new_feat = scatter_max(...)[remove_duplicates] (:1632-1637): one warp per unique voxel,
// lane = feature channel (feat_dim <= 32), maximum over the voxel's contiguous run of candidates
__global__ void __launch_bounds__(DT)
grow_emit_kernel(uint32_t M, int n_offsets, int feat_dim, float cur_size, const uint32_t* __restrict__ n_unique,
                 const uint32_t* __restrict__ ufirst, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ new_index,
                 const uint32_t* __restrict__ perm, const uint32_t* __restrict__ cand_slot, const int* __restrict__ kx,
                 const int* __restrict__ ky, const int* __restrict__ kz, const float* __restrict__ anchor_feat,
                 float* __restrict__ new_anchor, float* __restrict__ new_feat)
{
    const uint32_t u = (blockIdx.x * DT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t U = *n_unique;
    if (u >= U || !keep[u]) return;
    const uint32_t begin = ufirst[u], end = (u + 1 < U) ? ufirst[u + 1] : M;
    const uint32_t dst = new_index[u];
    const uint32_t j0 = perm[begin];
    if (lane == 0) {
        new_anchor[3 * size_t(dst) + 0] = __fmul_rn((float)kx[j0], cur_size);
        new_anchor[3 * size_t(dst) + 1] = __fmul_rn((float)ky[j0], cur_size);
        new_anchor[3 * size_t(dst) + 2] = __fmul_rn((float)kz[j0], cur_size);
    }
    if (lane < feat_dim) {
        float m = -__int_as_float(0x7f800000);
        for (uint32_t i = begin; i < end; ++i) {
            const uint32_t a = cand_slot[perm[i]] / (uint32_t)n_offsets;
            m = fmaxf(m, anchor_feat[size_t(a) * feat_dim + lane]);
        }
        new_feat[size_t(dst) * feat_dim + lane] = m;
    }
}

// ---- adjust_anchor's statistics update and prune decision (:1716-1755) -------------------------------------------------
__global__ void __launch_bounds__(DT)
prune_offsets_kernel(int init_slots, float denom_threshold, float* __restrict__ offset_gradient_accum, float* __restrict__ offset_denom)
{
    const int s = blockIdx.x * DT + threadIdx.x;
    if (s >= init_slots) return;
    if (offset_denom[s] > denom_threshold) {            // offset_mask
        offset_denom[s] = 0.f;
        offset_gradient_accum[s] = 0.f;
    }
}

__global__ void __launch_bounds__(DT)
prune_anchors_kernel(int A, float anchor_threshold, float min_opacity, float* __restrict__ opacity_accum,
                     float* __restrict__ anchor_demon, uint32_t* __restrict__ keep)
{
    const int a = blockIdx.x * DT + threadIdx.x;
    if (a >= A) return;
    const float acc = opacity_accum[a], dem = anchor_demon[a];
    const bool anchors_mask = dem > anchor_threshold;
    const bool prune = (acc < __fmul_rn(min_opacity, dem)) && anchors_mask;
    if (anchors_mask) { opacity_accum[a] = 0.f; anchor_demon[a] = 0.f; }
    keep[a] = prune ? 0u : 1u;
}

__global__ void __launch_bounds__(DT)
compact_rows_kernel(int A, int row_floats, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ keep_index,
                    const float* __restrict__ src, float* __restrict__ dst, int clamp_from, float clamp_max)
{
    const size_t e = size_t(blockIdx.x) * DT + threadIdx.x;
    if (e >= size_t(A) * row_floats) return;
    const int a = int(e / row_floats), c = int(e % row_floats);
    if (!keep[a]) return;
    float v = src[e];
    if (clamp_from >= 0 && c >= clamp_from) v = fminf(v, clamp_max);          // prune_anchor's clamp (:1529-1532, :1542-1545)
    dst[size_t(keep_index[a]) * row_floats + c] = v;
}

}  // namespace
}  // namespace segs

using namespace segs;

extern "C" {

int segs_anchor_growing_level(
    int A_now, int init_slots, int n_offsets, int feat_dim,
    const float* anchor, const float* offset, const float* log_scaling, const float* anchor_feat,
    const float* offset_gradient_accum, const float* offset_denom, const float* rand_values,
    float denom_threshold, float cur_threshold, float rand_threshold, float cur_size,
    segs_alloc_fn scratch_alloc, void* scratch_user, segs_alloc_fn out_alloc, void* out_user,
    float** new_anchor, float** new_feat, int* n_candidates, int* n_new, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_new) *n_new = 0;
    if (n_candidates) *n_candidates = 0;
    if (new_anchor) *new_anchor = nullptr;
    if (new_feat) *new_feat = nullptr;
    if (A_now <= 0 || init_slots <= 0) return SEGS_OK;
    if (n_offsets <= 0 || feat_dim <= 0 || feat_dim > 32 || init_slots > A_now * n_offsets || !(cur_size > 0.f)) {
        set_error("anchor growing: invalid sizes"); return SEGS_ERR_INVALID_ARG;
    }
    if (!anchor || !offset || !log_scaling || !anchor_feat || !offset_gradient_accum || !offset_denom || !rand_values ||
        !scratch_alloc || !out_alloc || !new_anchor || !new_feat || !n_new) {
        set_error("anchor growing: NULL argument"); return SEGS_ERR_INVALID_ARG;
    }
    const float inv_size = (float)(1.0 / (double)cur_size);      // ATen: high_prec_t(1.0) / scalar, cast to float
    const size_t S = size_t(init_slots);
    int rc;

    // phase 1 scratch: flags + positions over the slots
    size_t w1 = 2 * S + scan_temp_words(S) + 8 + 64;
    uint32_t* s1 = reinterpret_cast<uint32_t*>(scratch_alloc(scratch_user, w1 * sizeof(uint32_t)));
    if (!s1) { set_error("anchor growing: scratch allocation failed"); return SEGS_ERR_ALLOC; }
    uint32_t* counters = s1;                 // [0] M, [1] U, [2] N_new
    uint32_t* flag = s1 + 8;
    uint32_t* pos = flag + S;
    uint32_t* scan_tmp = pos + S;
    grow_mark_kernel<<<blocks_for(S), DT, 0, stream>>>(init_slots, offset_gradient_accum, offset_denom, rand_values,
                                                       denom_threshold, cur_threshold, rand_threshold, flag);
    SEGS_LAUNCH_CHECK();
    if ((rc = exclusive_scan(flag, pos, (uint32_t)S, nullptr, scan_tmp, counters, stream))) return rc;
    uint32_t M = 0;
    SEGS_CUDA_CHECK(cudaMemcpyAsync(&M, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SEGS_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (n_candidates) *n_candidates = (int)M;
    if (M == 0) return SEGS_OK;

    // phase 2 scratch: a second block sized by the candidate count (both blocks must stay valid until this function
    // returns — the callback hands out a NEW block per call, see include/segs_raster.h)
    const size_t Mz = M;
    const size_t sort_words = radix_sort_temp_words(Mz, 4);
    const size_t w2 = Mz /*cand_slot*/ + 3 * Mz /*kx ky kz*/ + 4 * Mz /*key a/b, perm a/b*/ + 2 * Mz /*head, head_scan*/ +
                      3 * Mz /*ufirst, keep, new_index*/ + sort_words + scan_temp_words(Mz) + 64;
    uint32_t* s2 = reinterpret_cast<uint32_t*>(scratch_alloc(scratch_user, w2 * sizeof(uint32_t)));
    if (!s2) { set_error("anchor growing: scratch allocation failed"); return SEGS_ERR_ALLOC; }
    if (s2 == s1) { set_error("anchor growing: the scratch callback must return a new block per call"); return SEGS_ERR_INVALID_ARG; }
    uint32_t* p = s2;
    uint32_t* cand_slot = p; p += Mz;
    int* kx = reinterpret_cast<int*>(p); p += Mz;
    int* ky = reinterpret_cast<int*>(p); p += Mz;
    int* kz = reinterpret_cast<int*>(p); p += Mz;
    uint32_t* key_a = p; p += Mz;
    uint32_t* key_b = p; p += Mz;
    uint32_t* perm_a = p; p += Mz;
    uint32_t* perm_b = p; p += Mz;
    uint32_t* head = p; p += Mz;
    uint32_t* head_scan = p; p += Mz;
    uint32_t* ufirst = p; p += Mz;
    uint32_t* keep = p; p += Mz;
    uint32_t* new_index = p; p += Mz;
    uint32_t* sort_tmp = p; p += sort_words;
    uint32_t* scan_tmp2 = p;

    grow_keys_kernel<<<blocks_for(S), DT, 0, stream>>>(init_slots, n_offsets, flag, pos, anchor, offset, log_scaling, inv_size,
                                                       cand_slot, kx, ky, kz);
    SEGS_LAUNCH_CHECK();
    // lexicographic (x, y, z) order = three stable LSD rounds: z, then y, then x
    const int* comps[3] = {kz, ky, kx};
    for (int r = 0; r < 3; ++r) {
        grow_round_key_kernel<<<blocks_for(Mz), DT, 0, stream>>>(M, comps[r], r == 0 ? nullptr : perm_a, key_a);
        SEGS_LAUNCH_CHECK();
        if ((rc = radix_sort_pairs(key_a, key_b, perm_a, perm_b, Mz, 0, 4, r == 0, sort_tmp, stream))) return rc;   // result in the a-buffers
    }
    grow_heads_kernel<<<blocks_for(Mz), DT, 0, stream>>>(M, perm_a, kx, ky, kz, head);
    SEGS_LAUNCH_CHECK();
    if ((rc = exclusive_scan(head, head_scan, M, nullptr, scan_tmp2, counters + 1, stream))) return rc;            // U
    grow_first_kernel<<<blocks_for(Mz), DT, 0, stream>>>(M, head, head_scan, ufirst);
    SEGS_LAUNCH_CHECK();
    fill_u32_kernel<<<blocks_for(Mz), DT, 0, stream>>>(M, 1u, keep);
    SEGS_LAUNCH_CHECK();
    grow_dup_kernel<<<blocks_for(A_now), DT, 0, stream>>>(A_now, anchor, inv_size, counters + 1, ufirst, perm_a, kx, ky, kz, keep);
    SEGS_LAUNCH_CHECK();
    if ((rc = exclusive_scan(keep, new_index, M, counters + 1, scan_tmp2, counters + 2, stream))) return rc;       // over the first U entries
    uint32_t N = 0;
    SEGS_CUDA_CHECK(cudaMemcpyAsync(&N, counters + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SEGS_CUDA_CHECK(cudaStreamSynchronize(stream));
    *n_new = (int)N;
    if (N == 0) return SEGS_OK;
    float* out = reinterpret_cast<float*>(out_alloc(out_user, size_t(N) * (3 + feat_dim) * sizeof(float)));
    if (!out) { set_error("anchor growing: output allocation failed"); return SEGS_ERR_ALLOC; }
    *new_anchor = out;
    *new_feat = out + size_t(N) * 3;
    grow_emit_kernel<<<blocks_for(Mz * 32), DT, 0, stream>>>(M, n_offsets, feat_dim, cur_size, counters + 1, ufirst, keep, new_index,
                                                            perm_a, cand_slot, kx, ky, kz, anchor_feat, *new_anchor, *new_feat);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

int segs_prune_plan(
    int A, int init_slots, float* opacity_accum, float* anchor_demon, float* offset_gradient_accum, float* offset_denom,
    float denom_threshold, float anchor_threshold, float min_opacity,
    unsigned int* keep /* [A] out: 1 = the anchor survives */, unsigned int* keep_index /* [A] out: its new row */,
    unsigned int* scratch /* segs_prune_scratch_words(A) words */, int* n_keep, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_keep) *n_keep = 0;
    if (A <= 0) return SEGS_OK;
    if (!opacity_accum || !anchor_demon || !keep || !keep_index || !scratch || !n_keep ||
        (init_slots > 0 && (!offset_gradient_accum || !offset_denom))) {
        set_error("prune plan: NULL argument"); return SEGS_ERR_INVALID_ARG;
    }
    if (init_slots > 0) {
        prune_offsets_kernel<<<blocks_for(init_slots), DT, 0, stream>>>(init_slots, denom_threshold, offset_gradient_accum, offset_denom);
        SEGS_LAUNCH_CHECK();
    }
    prune_anchors_kernel<<<blocks_for(A), DT, 0, stream>>>(A, anchor_threshold, min_opacity, opacity_accum, anchor_demon, keep);
    SEGS_LAUNCH_CHECK();
    int rc;
    if ((rc = exclusive_scan(keep, keep_index, (uint32_t)A, nullptr, scratch + 1, scratch, stream))) return rc;
    uint32_t n = 0;
    SEGS_CUDA_CHECK(cudaMemcpyAsync(&n, scratch, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SEGS_CUDA_CHECK(cudaStreamSynchronize(stream));
    *n_keep = (int)n;
    return SEGS_OK;
}

size_t segs_prune_scratch_words(int A) { return 1 + scan_temp_words(size_t(A > 0 ? A : 0)); }

int segs_compact_rows(int A, int row_floats, const unsigned int* keep, const unsigned int* keep_index, const float* src,
                      float* dst, int clamp_from, float clamp_max, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (A <= 0 || row_floats <= 0) return SEGS_OK;
    if (!keep || !keep_index || !src || !dst) { set_error("compact rows: NULL argument"); return SEGS_ERR_INVALID_ARG; }
    compact_rows_kernel<<<blocks_for(size_t(A) * row_floats), DT, 0, stream>>>(A, row_floats, keep, keep_index, src, dst,
                                                                               clamp_from, clamp_max);
    SEGS_LAUNCH_CHECK();
    return SEGS_OK;
}

}  // extern "C"
