// Shared definitions for the sm_100a rasterizer kernels: opaque-buffer layout, error
// plumbing, small device helpers.  Private to the library (nothing here is ABI).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/segs_raster.h"

namespace segs {

constexpr int TILE_X = 16;              // reference config.h:16-17 (BLOCK_X/BLOCK_Y)
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr int NUM_CH = 3;               // reference config.h:15
constexpr int SM_COUNT = 148;           // B200

// ---- error plumbing -----------------------------------------------------------------
void set_error(const char* fmt, ...);
#define SEGS_CUDA_CHECK(expr)                                                            \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::segs::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,       \
                              cudaGetErrorString(_e));                                   \
            return SEGS_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)
void count_launch();
unsigned readback_event_flags();      // flags of the events the host read-backs wait on (segs_set_blocking_sync)
cudaEvent_t readback_event();         // the calling thread's read-back event on the CURRENT device (nullptr on failure)
#define SEGS_LAUNCH_CHECK()                  \
    do {                                     \
        ::segs::count_launch();              \
        SEGS_CUDA_CHECK(cudaGetLastError()); \
    } while (0)

// ---- opaque buffer layout -----------------------------------------------------------
// A bump allocator over the caller-owned byte buffers, 128-byte aligned sections
// (same idea as the reference's obtain()/required(), rasterizer_impl.h:22-72, but the
// contents are this library's own SoA layout).
struct Carver {
    char* p;
    explicit Carver(char* base) : p(base) {}
    template <typename T>
    T* take(size_t count) {
        uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127);
        T* r = reinterpret_cast<T*>(a);
        p = reinterpret_cast<char*>(r + count);
        return r;
    }
    size_t used(char* base) const { return size_t(p - base); }
};

constexpr int COUNTER_BASE = 8, COUNTER_SLOTS = 256, COUNTER_WORDS = COUNTER_BASE + 2 * COUNTER_SLOTS;

constexpr int RADIX_BITS = 8;
constexpr int RADIX_BINS = 1 << RADIX_BITS;

// stable LSD radix sort of (u32 key, u32 value) pairs (radix_sort.cu): `npasses` 8-bit digits
// starting at `begin_bit`; ping-pongs between (key_a,val_a) and (key_b,val_b), the result is in
// the a-buffers for an even number of passes and in the b-buffers otherwise.  `iota_values`:
// val_a is not read, values start as 0..n-1.  temp: radix_sort_temp_words(n, npasses) u32 words.
size_t radix_sort_temp_words(size_t n, int npasses);
int radix_sort_pairs(uint32_t* key_a, uint32_t* key_b, uint32_t* val_a, uint32_t* val_b, size_t n,
                     int begin_bit, int npasses, bool iota_values, uint32_t* temp, cudaStream_t stream);

// Per-Gaussian render record, 48 bytes (three float4):
//   rec[3i+0] = { x, y, hx, hy }            2D mean (pixels), conservative half extents
//   rec[3i+1] = { conic.x, conic.y, conic.z, opacity }
//   rec[3i+2] = { r, g, b, depth }
// Per-Gaussian gradient accumulator written by the blend backward, 48 bytes:
//   acc[3i+0] = { dL_dmean2D.x, dL_dmean2D.y, dL_dconic.x, dL_dconic.y }
//   acc[3i+1] = { dL_dconic.w(=zz), dL_dopacity, dL_dcolor.r, dL_dcolor.g }
//   acc[3i+2] = { dL_dcolor.b, 0, 0, 0 }
struct GeomState {
    float*    depths;         // [P]
    uint32_t* tiles_touched;  // [P]
    ushort4*  rect;           // [P] x0,y0,x1,y1 in tiles
    float4*   rec;            // [3P]
    float*    cov3D;          // [6][P] planes
    float4*   acc;            // [3P]
    uint8_t*  clamped;        // [3P] (SH path)
    uint32_t* key_a;          // [P] depth-sort ping
    uint32_t* key_b;          // [P] depth-sort pong
    uint32_t* val_a;          // [P]
    uint32_t* val_b;          // [P]  -> depth order after 4 passes lives in val_a
    uint32_t* sort_temp;      // [radix_sort_temp_words(P, 4)]
    uint32_t* counters;       // [COUNTER_WORDS]: 0 = num_rendered (sum of tiles_touched), 1 = error flag,
                              //      2 = number of coarse (super-tile, Gaussian) candidates,
                              //      3 = num_rendered as seen by the tile scan (cross-check),
                              //      4 = preprocess CTAs finished (the last one publishes 0-2 to the host),
                              //      COUNTER_BASE + 2k, + 2k + 1 = partial sums of 0 and 2 (slot k of COUNTER_SLOTS)
    static GeomState carve(char* base, size_t P, size_t* bytes);
};

// two-level tile binning (binning.cu): super-tiles of (1 << sshift)^2 tiles
struct BinningPlan { int sshift, sgrid_x, sgrid_y, sort_passes; };
BinningPlan plan_binning(int grid_x, int grid_y);
int emit_blocks(int P);       // CTAs of the candidate expansion (one look-back word each)

struct BinningState {
    uint32_t* point_list;     // [R]   Gaussian ids, tile-major, depth-minor
    uint32_t* cand_key_a;     // [Rc]  super-tile id | 16-bit tile mask << 16 of each candidate (depth order)
    uint32_t* cand_key_b;     // [Rc]
    uint32_t* cand_val_a;     // [Rc]  Gaussian id of each candidate
    uint32_t* cand_val_b;     // [Rc]  -> grouped by super-tile, depth order inside
    uint32_t* sort_temp;      // [radix_sort_temp_words(Rc, passes)]
    // ---- zeroed with one memset before the binning kernels run ----
    uint32_t* zeroed;
    size_t    zeroed_bytes;
    uint32_t* emit_ticket;    // [1]
    uint32_t* emit_look;      // [emit_blocks(P)]
    uint32_t* st_begin;       // [super-tiles] candidate range of every super-tile
    uint32_t* st_end;         // [super-tiles]
    uint32_t* tile_counts;    // [T]
    static BinningState carve(char* base, size_t R, size_t Rc, size_t P, int grid_x, int grid_y, size_t* bytes);
};

struct ImageState {
    float*    final_T;        // [N]
    uint32_t* n_contrib;      // [N]
    uint2*    ranges;         // [T]
    static ImageState carve(char* base, size_t N, size_t T, size_t* bytes);
};

// ---- kernel launchers (one per translation unit) ------------------------------------
struct ViewParams {
    int W, H;
    int grid_x, grid_y;
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;
    float scale_modifier;
    int sshift;               // log2 of the super-tile side in tiles (two-level binning)
};

int launch_preprocess(int P, int D, int M, const float* means3D, const float* scales,
                      const float* rotations, const float* opacities, const float* shs,
                      const float* cov3D_precomp, const float* colors_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                      const ViewParams& vp, bool prefiltered, int* radii, GeomState& g,
                      uint32_t* host_counters /* mapped pinned memory, device address */, cudaStream_t stream);

int launch_filter(int P, const float* means3D, const float* scales, const float* rotations,
                  const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                  const ViewParams& vp, bool prefiltered, int* radii, uint32_t* err_flag,
                  cudaStream_t stream);

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix,
                        unsigned char* present, cudaStream_t stream);

int launch_project(int P, int D, int M, const float* means3D, const float* scales,
                   const float* rotations, const float* opacities, const float* shs,
                   const float* cov3D_precomp, const float* colors_precomp,
                   const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                   const ViewParams& vp, bool prefiltered, float* out_rgb, float* points_image,
                   int* radii, cudaStream_t stream);

// depth ordering of Gaussians (stable radix sort of (depth_bits, index))
int launch_depth_order(int P, GeomState& g, cudaStream_t stream);
// chunked tile multisplit: point_list + ranges
int launch_binning(int P, int R, int Rc, const ViewParams& vp, GeomState& g, BinningState& b,
                   ImageState& img, cudaStream_t stream);

int debug_blend_stats(unsigned long long* out8, bool reset);   // zeros unless built with -DSEGS_BLEND_STATS

int launch_blend_forward(const ViewParams& vp, const GeomState& g, const BinningState& b,
                         ImageState& img, const float* background, float* out_color,
                         cudaStream_t stream);

int launch_blend_backward(const ViewParams& vp, const GeomState& g, const BinningState& b,
                          const ImageState& img, const float* background,
                          const float* dL_dpix, cudaStream_t stream);

int launch_preprocess_backward(int P, int D, int M, const float* means3D, const float* scales,
                               const float* rotations, const float* shs,
                               const float* cov3D_precomp, const float* viewmatrix,
                               const float* projmatrix, const float* campos,
                               const ViewParams& vp, const int* radii, GeomState& g,
                               float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                               float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                               float* dL_dsh, float* dL_dscale, float* dL_drot,
                               cudaStream_t stream);

}  // namespace segs
