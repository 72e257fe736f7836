// C-ABI entry points (include/segs_raster.h) and the opaque-buffer layout.
//
// Pipeline driver replacing CudaRasterizer::Rasterizer::{forward,backward,visible_filter,
// markVisible,project2_image} (cuda_rasterizer/rasterizer_impl.cu:141-153,198-336,339-393,
// 397-490,494-585).  Unlike the reference, every launch goes to the caller's stream, every
// CUDA call is checked, and nothing is allocated here: all device memory comes from the
// three allocation callbacks.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <string>
#include "common.cuh"

namespace segs {

// ---- error plumbing -----------------------------------------------------------------
static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};      // process-wide: lanes launch from their own host threads
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Host waits for the two small read-backs (num_rendered, decode counts) spin by default; with more waiting host
// threads than cores (8 ranks x lanes on one box) they should sleep instead: cudaEventBlockingSync.
static std::atomic<int> g_blocking_sync{0};
unsigned readback_event_flags() {
    return cudaEventDisableTiming | (g_blocking_sync.load(std::memory_order_relaxed) ? cudaEventBlockingSync : 0u);
}

// ---- layouts --------------------------------------------------------------------------
GeomState GeomState::carve(char* base, size_t P, size_t* bytes) {
    Carver c(base);
    GeomState g;
    g.depths = c.take<float>(P);
    g.tiles_touched = c.take<uint32_t>(P);
    g.rect = c.take<ushort4>(P);
    g.rec = c.take<float4>(3 * P);
    g.cov3D = c.take<float>(6 * P);
    g.acc = c.take<float4>(3 * P);
    g.clamped = c.take<uint8_t>(3 * P);
    g.key_a = c.take<uint32_t>(P);
    g.key_b = c.take<uint32_t>(P);
    g.val_a = c.take<uint32_t>(P);
    g.val_b = c.take<uint32_t>(P);
    g.sort_temp = c.take<uint32_t>(radix_sort_temp_words(P, 4));
    g.counters = c.take<uint32_t>(COUNTER_WORDS);
    if (bytes) *bytes = c.used(base) + 128;
    return g;
}

BinningState BinningState::carve(char* base, size_t R, size_t Rc, size_t P, int grid_x, int grid_y, size_t* bytes) {
    Carver c(base);
    BinningState b;
    const BinningPlan pl = plan_binning(grid_x, grid_y);
    const size_t nst = size_t(pl.sgrid_x) * pl.sgrid_y, T = size_t(grid_x) * grid_y;
    b.point_list = c.take<uint32_t>(R);      // first: the only section the backward reads
    b.cand_key_a = c.take<uint32_t>(Rc);
    b.cand_key_b = c.take<uint32_t>(Rc);
    b.cand_val_a = c.take<uint32_t>(Rc);
    b.cand_val_b = c.take<uint32_t>(Rc);
    b.sort_temp = c.take<uint32_t>(radix_sort_temp_words(Rc, pl.sort_passes));
    const size_t nz = 1 + size_t(emit_blocks((int)P)) + 2 * nst + T;
    b.zeroed = c.take<uint32_t>(nz);
    b.zeroed_bytes = nz * sizeof(uint32_t);
    b.emit_ticket = b.zeroed;
    b.emit_look = b.emit_ticket + 1;
    b.st_begin = b.emit_look + emit_blocks((int)P);
    b.st_end = b.st_begin + nst;
    b.tile_counts = b.st_end + nst;
    if (bytes) *bytes = c.used(base) + 128;
    return b;
}

ImageState ImageState::carve(char* base, size_t N, size_t T, size_t* bytes) {
    Carver c(base);
    ImageState s;
    s.final_T = c.take<float>(N);
    s.n_contrib = c.take<uint32_t>(N);
    s.ranges = c.take<uint2>(T);
    if (bytes) *bytes = c.used(base) + 128;
    return s;
}

static ViewParams make_view(int width, int height, float tan_fovx, float tan_fovy, float scale_modifier) {
    ViewParams vp;
    vp.W = width;
    vp.H = height;
    vp.grid_x = (width + TILE_X - 1) / TILE_X;
    vp.grid_y = (height + TILE_Y - 1) / TILE_Y;
    vp.tan_fovx = tan_fovx;
    vp.tan_fovy = tan_fovy;
    // rasterizer_impl.cu:221-222 (float arithmetic on the host)
    vp.focal_y = height / (2.0f * tan_fovy);
    vp.focal_x = width / (2.0f * tan_fovx);
    vp.scale_modifier = scale_modifier;
    vp.sshift = plan_binning(vp.grid_x, vp.grid_y).sshift;
    return vp;
}

// A few words of MAPPED pinned memory per host thread: the preprocess kernel stores
// num_rendered / the error flag / the candidate count there directly (no copy engine).
struct HostWords { uint32_t* host = nullptr; uint32_t* dev = nullptr; };
static HostWords pinned_words() {
    static thread_local HostWords w;
    if (!w.host) {
        void* h = nullptr;
        void* d = nullptr;
        if (cudaHostAlloc(&h, 8 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
            cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
            w.host = static_cast<uint32_t*>(h);
            w.dev = static_cast<uint32_t*>(d);
        }
    }
    return w;
}

// One event per (host thread, device) marking "the read-back words have landed in pinned memory".  Events belong to the
// device that was current when they were created, so the cache is keyed by cudaGetDevice(); an event is re-created when
// segs_set_blocking_sync() changed the flags it was created with.
cudaEvent_t readback_event() {
    struct Slot { cudaEvent_t ev = nullptr; unsigned flags = 0; };
    constexpr int MAX_DEVICES = 64;
    static thread_local Slot slots[MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
    Slot& s = slots[dev];
    const unsigned flags = readback_event_flags();
    if (s.ev && s.flags != flags) { cudaEventDestroy(s.ev); s.ev = nullptr; }
    if (!s.ev) {
        if (cudaEventCreateWithFlags(&s.ev, flags) != cudaSuccess) { s.ev = nullptr; return nullptr; }
        s.flags = flags;
    }
    return s.ev;
}

// ---- optional per-stage timing ----------------------------------------------------------
struct Profile {
    bool on = false;
    bool created = false;
    cudaEvent_t ev[2 * SEGS_PROFILE_STAGES];
    bool armed[SEGS_PROFILE_STAGES] = {false, false, false, false, false, false};
};
static thread_local Profile g_prof;

static void prof_begin(int stage, cudaStream_t s) {
    if (!g_prof.on) return;
    if (!g_prof.created) {
        for (auto& e : g_prof.ev) cudaEventCreate(&e);
        g_prof.created = true;
    }
    cudaEventRecord(g_prof.ev[2 * stage], s);
}
static void prof_end(int stage, cudaStream_t s) {
    if (!g_prof.on) return;
    cudaEventRecord(g_prof.ev[2 * stage + 1], s);
    g_prof.armed[stage] = true;
}

}  // namespace segs

using namespace segs;

extern "C" {

int segs_version(void) { return SEGS_ABI_VERSION; }

unsigned long long segs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int segs_set_blocking_sync(int on) { g_blocking_sync.store(on != 0); return SEGS_OK; }

int segs_profile_enable(int on) { g_prof.on = on != 0; return SEGS_OK; }

int segs_profile_read(float* ms) {
    if (!ms) { set_error("invalid argument"); return SEGS_ERR_INVALID_ARG; }
    for (int s = 0; s < SEGS_PROFILE_STAGES; ++s) {
        ms[s] = 0.f;
        if (!g_prof.created || !g_prof.armed[s]) continue;
        SEGS_CUDA_CHECK(cudaEventSynchronize(g_prof.ev[2 * s + 1]));
        SEGS_CUDA_CHECK(cudaEventElapsedTime(&ms[s], g_prof.ev[2 * s], g_prof.ev[2 * s + 1]));
        g_prof.armed[s] = false;
    }
    return SEGS_OK;
}
const char* segs_last_error(void) { return g_last_error; }

int segs_raster_forward(
    segs_alloc_fn geom_alloc, void* geom_user,
    segs_alloc_fn binning_alloc, void* binning_user,
    segs_alloc_fn image_alloc, void* image_user,
    int P, int D, int M,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, int* radii, int* num_rendered, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (num_rendered) *num_rendered = 0;
    if (!geom_alloc || !binning_alloc || !image_alloc) { set_error("allocation callbacks must not be NULL"); return SEGS_ERR_INVALID_ARG; }
    if (P < 0 || width <= 0 || height <= 0) { set_error("invalid sizes P=%d W=%d H=%d", P, width, height); return SEGS_ERR_INVALID_ARG; }
    if (!out_color || !background || !viewmatrix || !projmatrix) { set_error("NULL required pointer"); return SEGS_ERR_INVALID_ARG; }
    const size_t N = size_t(width) * height;
    if (P == 0) {
        // RasterizeGaussiansCUDA (rasterize_points.cu:81): nothing runs, the image stays zero
        SEGS_CUDA_CHECK(cudaMemsetAsync(out_color, 0, NUM_CH * N * sizeof(float), stream));
        return SEGS_OK;
    }
    if (!means3D || !opacities) { set_error("means3D/opacities must not be NULL"); return SEGS_ERR_INVALID_ARG; }
    if (!colors_precomp && !shs) {
        // rasterizer_impl.cu:241-244 analogue: no colour source at all
        set_error("Please provide exactly one of either SHs or precomputed colors!");
        return SEGS_ERR_INVALID_ARG;
    }
    if (!cov3D_precomp && (!scales || !rotations)) {
        set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return SEGS_ERR_INVALID_ARG;
    }
    if (!colors_precomp && (M <= 0 || (D + 1) * (D + 1) > M || D > 3 || D < 0)) {
        set_error("SH degree %d needs %d coefficients, got M=%d", D, (D + 1) * (D + 1), M);
        return SEGS_ERR_INVALID_ARG;
    }

    const ViewParams vp = make_view(width, height, tan_fovx, tan_fovy, scale_modifier);
    const size_t T = size_t(vp.grid_x) * vp.grid_y;
    if (vp.grid_x > 65535 || vp.grid_y > 65535) { set_error("image too large"); return SEGS_ERR_INVALID_ARG; }

    size_t geom_bytes = 0, img_bytes = 0, bin_bytes = 0;
    GeomState::carve(nullptr, P, &geom_bytes);
    char* geom_ptr = geom_alloc(geom_user, geom_bytes);
    if (!geom_ptr) { set_error("geometry buffer allocation of %zu bytes failed", geom_bytes); return SEGS_ERR_ALLOC; }
    GeomState g = GeomState::carve(geom_ptr, P, nullptr);

    ImageState::carve(nullptr, N, T, &img_bytes);
    char* img_ptr = image_alloc(image_user, img_bytes);
    if (!img_ptr) { set_error("image buffer allocation of %zu bytes failed", img_bytes); return SEGS_ERR_ALLOC; }
    ImageState img = ImageState::carve(img_ptr, N, T, nullptr);

    const HostWords hw = pinned_words();
    if (!hw.host) { set_error("cudaHostAlloc for the readback words failed"); return SEGS_ERR_CUDA; }
    volatile uint32_t* host_words = hw.host;

    SEGS_CUDA_CHECK(cudaMemsetAsync(g.counters, 0, COUNTER_WORDS * sizeof(uint32_t), stream));
    int rc;
    prof_begin(0, stream);
    if ((rc = launch_preprocess(P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp,
                                colors_precomp, viewmatrix, projmatrix, cam_pos, vp, prefiltered != 0,
                                radii, g, hw.dev, stream))) return rc;
    prof_end(0, stream);
    // num_rendered (accumulated by the preprocess and stored to mapped host memory by its last
    // CTA) sizes the binning buffer, so the host has to see it (rasterizer_impl.cu:279-285) — but
    // only the preprocess is waited for: the depth sort is already queued and runs while the
    // host allocates.
    cudaEvent_t ready = readback_event();
    if (!ready) { set_error("cudaEventCreate failed"); return SEGS_ERR_CUDA; }
    SEGS_CUDA_CHECK(cudaEventRecord(ready, stream));
    prof_begin(1, stream);
    if ((rc = launch_depth_order(P, g, stream))) return rc;
    prof_end(1, stream);
    SEGS_CUDA_CHECK(cudaEventSynchronize(ready));
    const uint32_t R = host_words[0], Rc = host_words[2];
    if (host_words[1] != 0) {
        set_error("Point is filtered although prefiltered is set. This shouldn't happen!");
        return SEGS_ERR_PREFILTERED;
    }
    if (R > 0x7FFFFFFFu) { set_error("num_rendered %u overflows int", R); return SEGS_ERR_INVALID_ARG; }

    BinningState::carve(nullptr, R, Rc, P, vp.grid_x, vp.grid_y, &bin_bytes);
    char* bin_ptr = binning_alloc(binning_user, bin_bytes);
    if (!bin_ptr) { set_error("binning buffer allocation of %zu bytes failed", bin_bytes); return SEGS_ERR_ALLOC; }
    BinningState b = BinningState::carve(bin_ptr, R, Rc, P, vp.grid_x, vp.grid_y, nullptr);

    prof_begin(2, stream);
    if ((rc = launch_binning(P, (int)R, (int)Rc, vp, g, b, img, stream))) return rc;
    prof_end(2, stream);
    prof_begin(3, stream);
    if ((rc = launch_blend_forward(vp, g, b, img, background, out_color, stream))) return rc;
    prof_end(3, stream);
    if (num_rendered) *num_rendered = (int)R;
    return SEGS_OK;
}

int segs_raster_backward(
    int P, int D, int M, int R,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* campos, float tan_fovx, float tan_fovy, const int* radii,
    char* geom_buffer, char* binning_buffer, char* image_buffer,
    const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
    float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot, void* stream_)
{
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    (void)colors_precomp;
    if (P == 0) return SEGS_OK;
    if (P < 0 || R < 0 || width <= 0 || height <= 0) { set_error("invalid sizes"); return SEGS_ERR_INVALID_ARG; }
    if (!geom_buffer || !image_buffer || (R > 0 && !binning_buffer)) { set_error("NULL state buffer"); return SEGS_ERR_INVALID_ARG; }
    if (!dL_dpix || !dL_dmean2D || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dcov3D || !dL_dscale || !dL_drot) {
        set_error("NULL gradient output"); return SEGS_ERR_INVALID_ARG;
    }
    if (shs && M > 0 && !dL_dsh) { set_error("dL_dsh must not be NULL on the SH path"); return SEGS_ERR_INVALID_ARG; }
    const ViewParams vp = make_view(width, height, tan_fovx, tan_fovy, scale_modifier);
    const size_t N = size_t(width) * height, T = size_t(vp.grid_x) * vp.grid_y;
    GeomState g = GeomState::carve(geom_buffer, P, nullptr);
    BinningState b = BinningState::carve(binning_buffer, R, 0, P, vp.grid_x, vp.grid_y, nullptr);
    ImageState img = ImageState::carve(image_buffer, N, T, nullptr);
    int rc;
    // the 48-byte gradient accumulators start from zero (one bulk memset instead of strided clears
    // in the per-Gaussian kernels; forward-only renders never pay for it)
    SEGS_CUDA_CHECK(cudaMemsetAsync(g.acc, 0, size_t(P) * 3 * sizeof(float4), stream));
    if (R > 0) {
        prof_begin(4, stream);
        if ((rc = launch_blend_backward(vp, g, b, img, background, dL_dpix, stream))) return rc;
        prof_end(4, stream);
    }
    prof_begin(5, stream);
    if ((rc = launch_preprocess_backward(P, D, M, means3D, scales, rotations, shs, cov3D_precomp, viewmatrix,
                                         projmatrix, campos, vp, radii, g, dL_dmean2D, dL_dconic, dL_dopacity,
                                         dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, stream)))
        return rc;
    prof_end(5, stream);
    return SEGS_OK;
}

int segs_visible_filter(
    int P, int M, int width, int height,
    const float* means3D, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix,
    float tan_fovx, float tan_fovy, int prefiltered, int* radii, void* stream_)
{
    (void)M; (void)prefiltered;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P == 0) return SEGS_OK;
    if (P < 0 || width <= 0 || height <= 0 || !means3D || !radii || !viewmatrix || !projmatrix) {
        set_error("invalid argument"); return SEGS_ERR_INVALID_ARG;
    }
    if (!cov3D_precomp && (!scales || !rotations)) {
        set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return SEGS_ERR_INVALID_ARG;
    }
    const ViewParams vp = make_view(width, height, tan_fovx, tan_fovy, scale_modifier);
    return launch_filter(P, means3D, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, vp,
                         false, radii, nullptr, stream);
}

int segs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                      unsigned char* present, void* stream_)
{
    (void)projmatrix;   // only the view-space depth test is live (auxiliary.h:155-156)
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P == 0) return SEGS_OK;
    if (P < 0 || !means3D || !viewmatrix || !present) { set_error("invalid argument"); return SEGS_ERR_INVALID_ARG; }
    return launch_mark_visible(P, means3D, viewmatrix, present, stream);
}

int segs_project(
    int P, int D, int M, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
    const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_rgb, float* points_image, int* radii, void* stream_)
{
    (void)prefiltered;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (P == 0) return SEGS_OK;
    if (P < 0 || width <= 0 || height <= 0 || !means3D || !radii || !out_rgb || !points_image || !viewmatrix || !projmatrix) {
        set_error("invalid argument"); return SEGS_ERR_INVALID_ARG;
    }
    if (!colors_precomp && !shs) { set_error("Please provide exactly one of either SHs or precomputed colors!"); return SEGS_ERR_INVALID_ARG; }
    if (!cov3D_precomp && (!scales || !rotations)) {
        set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return SEGS_ERR_INVALID_ARG;
    }
    const ViewParams vp = make_view(width, height, tan_fovx, tan_fovy, scale_modifier);
    return launch_project(P, D, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp,
                          viewmatrix, projmatrix, cam_pos, vp, false, out_rgb, points_image, radii, stream);
}

int segs_debug_blend_stats(unsigned long long* out8, int reset) { return out8 ? debug_blend_stats(out8, reset != 0) : SEGS_ERR_INVALID_ARG; }

int segs_buffer_section(const char* name, char* geom_buffer, char* binning_buffer, char* image_buffer,
                        int P, int R, int width, int height, void** ptr, size_t* bytes)
{
    if (!name || !ptr || !bytes) { set_error("invalid argument"); return SEGS_ERR_INVALID_ARG; }
    const ViewParams vp = make_view(width, height, 1.f, 1.f, 1.f);
    const size_t N = size_t(width) * height, T = size_t(vp.grid_x) * vp.grid_y;
    GeomState g = GeomState::carve(geom_buffer, P, nullptr);
    BinningState b = BinningState::carve(binning_buffer, R, 0, P, vp.grid_x, vp.grid_y, nullptr);
    ImageState img = ImageState::carve(image_buffer, N, T, nullptr);
    const std::string n(name);
    auto out = [&](void* p, size_t sz) { *ptr = p; *bytes = sz; return SEGS_OK; };
    if (n == "depths") return out(g.depths, sizeof(float) * P);
    if (n == "tiles_touched") return out(g.tiles_touched, sizeof(uint32_t) * P);
    if (n == "rect") return out(g.rect, sizeof(ushort4) * P);
    if (n == "rec") return out(g.rec, sizeof(float4) * 3 * P);
    if (n == "cov3D") return out(g.cov3D, sizeof(float) * 6 * P);
    if (n == "depth_order") return out(g.val_a, sizeof(uint32_t) * P);
    if (n == "point_list") return out(b.point_list, sizeof(uint32_t) * R);
    if (n == "ranges") return out(img.ranges, sizeof(uint2) * T);
    if (n == "final_T") return out(img.final_T, sizeof(float) * N);
    if (n == "n_contrib") return out(img.n_contrib, sizeof(uint32_t) * N);
    set_error("unknown section '%s'", name);
    return SEGS_ERR_INVALID_ARG;
}

}  // extern "C"
