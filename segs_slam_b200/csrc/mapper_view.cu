// One keyframe view of the batched mapping step, issued from C++ with no interpreter in the loop
// (BASELINE config 4): anchor prefilter -> fused decode -> rasterize -> L1+SSIM loss ->
// rasterizer backward -> decode backward that ACCUMULATES straight into the caller's gradient bucket.
//
// Restates the per-iteration body of GaussianMapper::trainForOneIteration
// (/root/reference/src/gaussian_mapper.cpp:870-950): prefilter_voxel (src/gaussian_renderer.cpp:131-199),
// GaussianRenderer::render (:40-127, generate_neural_gaussians :214-334), the loss (:908-925) and
// loss.backward() — as one host function over the kernels of this library.  The reference runs it as
// ~150 ATen launches with autograd bookkeeping; here it is 27 launches, two host read-backs (the
// decode's row count and num_rendered — both shape the next allocation, as in the reference) and no
// temporary that outlives the view: everything is carved from a reusable workspace arena.
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"

namespace segs {

namespace {

// Bump arena over a few large cudaMalloc'd chunks; reset per view, chunks are kept (after the first
// views of a run no allocation happens any more).
struct Chunk { char* base; size_t size; size_t used; };

}  // namespace

}  // namespace segs

struct segs_workspace {
    std::vector<segs::Chunk> chunks;
    size_t granule = size_t(256) << 20;
    size_t total = 0;
    bool failed = false;

    // frequency regularisation: cuFFT plans + working spectrum for the last (H, W, scales) used on this lane, and the
    // target magnitudes when the caller did not precompute them
    segs_freq_plan* freq = nullptr;
    int freq_H = 0, freq_W = 0, freq_ns = 0;
    float* freq_gt_mag = nullptr;

    // lane thread (segs_mapper_views): persistent, so the per-thread pinned read-back words and events of the
    // library are created once, not once per step
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, job_done = false, quit = false;
    int job_status = 0;
    std::string job_error;

    void start_worker() {
        if (worker.joinable()) return;
        worker = std::thread([this] {
            for (;;) {
                std::function<void()> j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [this] { return has_job || quit; });
                    if (quit) return;
                    j = std::move(job);
                    has_job = false;
                }
                j();
                {
                    std::lock_guard<std::mutex> lk(mu);
                    job_done = true;
                }
                cv.notify_all();
            }
        });
    }
    void submit(std::function<void()> j) {
        start_worker();
        {
            std::lock_guard<std::mutex> lk(mu);
            job = std::move(j);
            has_job = true;
            job_done = false;
        }
        cv.notify_all();
    }
    void wait_job() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return job_done; });
    }
    void stop_worker() {
        if (!worker.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv.notify_all();
        worker.join();
    }

    void reset() { for (auto& c : chunks) c.used = 0; failed = false; }
    char* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~size_t(255);
        if (bytes == 0) bytes = 256;
        for (auto& c : chunks)
            if (c.size - c.used >= bytes) { char* p = c.base + c.used; c.used += bytes; return p; }
        const size_t sz = bytes > granule ? bytes : granule;
        void* p = nullptr;
        if (cudaMalloc(&p, sz) != cudaSuccess) { cudaGetLastError(); failed = true; return nullptr; }
        chunks.push_back(segs::Chunk{static_cast<char*>(p), sz, bytes});
        total += sz;
        return static_cast<char*>(p);
    }
    template <typename T> T* take(size_t n) { return reinterpret_cast<T*>(alloc(n * sizeof(T))); }
};

namespace segs {
namespace {

char* ws_alloc_cb(void* user, size_t bytes) { return static_cast<segs_workspace*>(user)->alloc(bytes); }

__global__ void __launch_bounds__(256)
radii_to_mask_kernel(int n, const int* __restrict__ radii, unsigned char* __restrict__ mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = radii[i] > 0;       // visible_mask = radii_pure > 0 (gaussian_renderer.cpp:197)
}

__global__ void add_scalar_kernel(float* __restrict__ dst, const float* __restrict__ src) { atomicAdd(dst, *src); }

}  // namespace
}  // namespace segs

using namespace segs;

extern "C" {

int segs_workspace_create(segs_workspace** out)
{
    if (!out) { set_error("workspace: NULL output"); return SEGS_ERR_INVALID_ARG; }
    *out = new segs_workspace();
    return SEGS_OK;
}

int segs_workspace_destroy(segs_workspace* ws)
{
    if (!ws) return SEGS_OK;
    ws->stop_worker();
    if (ws->freq) segs_freq_plan_destroy(ws->freq);
    if (ws->freq_gt_mag) cudaFree(ws->freq_gt_mag);
    for (auto& c : ws->chunks) cudaFree(c.base);
    delete ws;
    return SEGS_OK;
}

size_t segs_workspace_bytes(const segs_workspace* ws) { return ws ? ws->total : 0; }

static int mapper_view_impl(segs_workspace* ws, const segs_mapper_view_args* a, segs_mapper_view_result* res,
                            bool concurrent, cudaStream_t stream)
{
    if (!ws || !a || !res) { set_error("mapper view: NULL argument"); return SEGS_ERR_INVALID_ARG; }
    res->n_visible = res->n_gaussians = res->num_rendered = 0;
    const int A = a->A, W = a->width, H = a->height;
    if (A <= 0 || W <= 0 || H <= 0) { set_error("mapper view: invalid sizes A=%d W=%d H=%d", A, W, H); return SEGS_ERR_INVALID_ARG; }
    if (!a->anchor || !a->anchor_feat || !a->offset || !a->scaling || !a->filter_scales || !a->filter_rotations ||
        !a->params || !a->viewmatrix || !a->projmatrix || !a->campos || !a->pose || !a->background || !a->gt_image ||
        !a->grad_anchor || !a->grad_anchor_feat || !a->grad_offset || !a->grad_scaling || !a->grad_params || !a->loss_accum) {
        set_error("mapper view: NULL required pointer"); return SEGS_ERR_INVALID_ARG;
    }
    ws->reset();
    int rc;
    auto oom = [&]() { set_error("mapper view: workspace allocation failed (%zu bytes held)", ws->total); return SEGS_ERR_ALLOC; };

    // ---- prefilter_voxel (gaussian_renderer.cpp:131-199): radii of the anchors themselves ----
    int* anchor_radii = ws->take<int>(A);
    unsigned char* visible = ws->take<unsigned char>(A);
    if (!anchor_radii || !visible) return oom();
    if ((rc = segs_visible_filter(A, 0, W, H, a->anchor, a->filter_scales, 1.0f, a->filter_rotations, nullptr, a->viewmatrix,
                                  a->projmatrix, a->tan_fovx, a->tan_fovy, 0, anchor_radii, stream))) return rc;
    radii_to_mask_kernel<<<(A + 255) / 256, 256, 0, stream>>>(A, anchor_radii, visible);
    SEGS_LAUNCH_CHECK();

    // ---- generate_neural_gaussians (:214-334) ----
    const size_t cap = size_t(A) * 10;
    float* xyz = ws->take<float>(cap * 3);
    float* color = ws->take<float>(cap * 3);
    float* opacity = ws->take<float>(cap);
    float* scaling = ws->take<float>(cap * 3);
    float* rot = ws->take<float>(cap * 4);
    float* neural_opacity = ws->take<float>(cap);
    unsigned char* offset_mask = ws->take<unsigned char>(cap);
    char* dstate = ws->alloc(segs_decode_state_bytes(A));
    if (!xyz || !color || !opacity || !scaling || !rot || !neural_opacity || !offset_mask || !dstate) return oom();
    int counts[2] = {0, 0};
    if ((rc = segs_decode_forward(A, visible, a->anchor, a->anchor_feat, a->offset, a->scaling, a->campos, a->pose, a->params,
                                  xyz, color, opacity, scaling, rot, neural_opacity, offset_mask, dstate, counts, stream)))
        return rc;
    const int n_vis = counts[0], P = counts[1];
    res->n_visible = n_vis;
    res->n_gaussians = P;

    // ---- rasterize (:40-127 -> RasterizeGaussiansCUDA) ----
    const size_t N = size_t(W) * H;
    float* image = a->image_out ? a->image_out : ws->take<float>(3 * N);
    int* radii = ws->take<int>(P > 0 ? P : 1);
    if (!image || !radii) return oom();
    // the three opaque buffers come from the arena through the same callback ABI the tensor-level API uses
    struct Slot { segs_workspace* ws; char* ptr; };
    Slot geom{ws, nullptr}, binning{ws, nullptr}, img{ws, nullptr};
    auto slot_cb = [](void* u, size_t bytes) -> char* {
        Slot* s = static_cast<Slot*>(u);
        s->ptr = s->ws->alloc(bytes);
        return s->ptr;
    };
    int R = 0;
    if ((rc = segs_raster_forward(slot_cb, &geom, slot_cb, &binning, slot_cb, &img, P, 0, 0, a->background, W, H, xyz, nullptr,
                                  color, opacity, scaling, 1.0f, rot, nullptr, a->viewmatrix, a->projmatrix, a->campos,
                                  a->tan_fovx, a->tan_fovy, 0, image, radii, &R, stream))) return rc;
    res->num_rendered = R;

    // ---- loss (gaussian_mapper.cpp:908-925) and its gradient w.r.t. the image ----
    char* lstate = ws->alloc(segs_loss_state_bytes(3, H, W));
    float* loss3 = ws->take<float>(4);
    float* dL_dimage = ws->take<float>(3 * N);
    if (!lstate || !loss3 || !dL_dimage) return oom();
    const float lam = a->lambda_dssim;
    if ((rc = segs_loss_l1_ssim_forward(3, H, W, image, a->gt_image, a->row_mask, 1.f - lam, -lam, lam, loss3, lstate, stream))) return rc;
    add_scalar_kernel<<<1, 1, 0, stream>>>(a->loss_accum, loss3 + 2);
    SEGS_LAUNCH_CHECK();
    if (a->loss_terms_out) SEGS_CUDA_CHECK(cudaMemcpyAsync(a->loss_terms_out, loss3, 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    // frequency regularisation (gaussian_mapper.cpp:930-945): value into the loss, gradient ADDED behind the SSIM backward
    auto freq_term = [&](float* grad_image) -> int {
        if (a->lambda_frequency_high == 0.f) return SEGS_OK;
        const int ns = a->use_multi_resolution ? (a->freq_scale_num < 1 ? 1 : (a->freq_scale_num > 4 ? 4 : a->freq_scale_num)) : 1;
        int frc;
        if (!ws->freq || ws->freq_H != H || ws->freq_W != W || ws->freq_ns != ns) {
            if (ws->freq) { segs_freq_plan_destroy(ws->freq); ws->freq = nullptr; }
            if (ws->freq_gt_mag) { cudaFree(ws->freq_gt_mag); ws->freq_gt_mag = nullptr; }
            float sc[4];
            for (int i = 0; i < ns; ++i) sc[i] = float(1.0 / pow(2.0, i));              // gaussian_mapper.cpp:514-517
            if ((frc = segs_freq_plan_create(3, H, W, ns, sc, &ws->freq))) return frc;
            ws->freq_H = H; ws->freq_W = W; ws->freq_ns = ns;
        }
        const float* gt_mag = a->gt_freq_mag;
        if (!gt_mag) {
            if (!ws->freq_gt_mag && cudaMalloc(&ws->freq_gt_mag, segs_freq_mag_floats(ws->freq) * sizeof(float)) != cudaSuccess) {
                cudaGetLastError(); return oom();
            }
            if ((frc = segs_freq_target(ws->freq, a->gt_image, a->row_mask, ws->freq_gt_mag, stream))) return frc;
            gt_mag = ws->freq_gt_mag;
        }
        return segs_freq_loss(ws->freq, image, a->row_mask, gt_mag, a->lambda_frequency_high, nullptr, a->loss_accum, grad_image, stream);
    };
    if (P == 0) {                        // nothing to back-propagate into; visible anchors still count in the statistics
        if ((rc = freq_term(nullptr))) return rc;
        if (n_vis > 0 && (a->stat_opacity_accum || a->stat_anchor_demon || a->stat_offset_gradient_accum || a->stat_offset_denom))
            return segs_training_statis(A, dstate, n_vis, neural_opacity, anchor_radii /* not read: no offset survived */, xyz,
                                        a->stat_opacity_accum, a->stat_anchor_demon, a->stat_offset_gradient_accum,
                                        a->stat_offset_denom, concurrent ? 1 : 0, stream);
        return SEGS_OK;
    }
    if ((rc = segs_loss_l1_ssim_backward(3, H, W, image, a->gt_image, a->row_mask, 1.f - lam, -lam, nullptr, lstate, dL_dimage, stream))) return rc;
    if ((rc = freq_term(dL_dimage))) return rc;

    // ---- rasterizer backward (RasterizeGaussiansBackwardCUDA) ----
    const size_t Pz = size_t(P);
    float* dL_dmean2D = a->dL_dmean2D_out ? a->dL_dmean2D_out : ws->take<float>(Pz * 3);
    float* dL_dconic = ws->take<float>(Pz * 4);
    float* dL_dopacity = ws->take<float>(Pz);
    float* dL_dcolor = ws->take<float>(Pz * 3);
    float* dL_dmean3D = ws->take<float>(Pz * 3);
    float* dL_dcov3D = ws->take<float>(Pz * 6);
    float* dL_dscale = ws->take<float>(Pz * 3);
    float* dL_drot = ws->take<float>(Pz * 4);
    if (!dL_dmean2D || !dL_dconic || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dcov3D || !dL_dscale || !dL_drot) return oom();
    if ((rc = segs_raster_backward(P, 0, 0, R, a->background, W, H, xyz, nullptr, color, scaling, 1.0f, rot, nullptr, a->viewmatrix,
                                   a->projmatrix, a->campos, a->tan_fovx, a->tan_fovy, radii, geom.ptr, binning.ptr, img.ptr,
                                   dL_dimage, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, nullptr,
                                   dL_dscale, dL_drot, stream))) return rc;
    if (a->radii_out) SEGS_CUDA_CHECK(cudaMemcpyAsync(a->radii_out, radii, Pz * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    // training_statis (gaussian_mapper.cpp:963): this view's contribution to the densification statistics
    if (a->stat_opacity_accum || a->stat_anchor_demon || a->stat_offset_gradient_accum || a->stat_offset_denom)
        if ((rc = segs_training_statis(A, dstate, n_vis, neural_opacity, radii, dL_dmean2D, a->stat_opacity_accum,
                                       a->stat_anchor_demon, a->stat_offset_gradient_accum, a->stat_offset_denom,
                                       concurrent ? 1 : 0, stream))) return rc;
    // 0.01 * scaling.prod(1).mean(): value into the loss, gradient on top of the rasterizer's dL_dscale
    if (a->scaling_reg_weight != 0.f)
        if ((rc = segs_scaling_reg(P, scaling, a->scaling_reg_weight, nullptr, dL_dscale, a->loss_accum, stream))) return rc;

    // ---- decode backward, accumulating into the bucket ----
    // MLP weight gradients of THIS view go to a small arena block first (the appearance-embedding gradient is
    // derived from this view's colour-bias gradient), then one launch adds all of them to the bucket.
    const segs_decode_params& p = *a->params;
    const int in_o = 35 + (p.add_opacity_dist ? 1 : 0), in_s = 35 + (p.add_cov_dist ? 1 : 0);
    const int in_c = 35 + (p.add_color_dist ? 1 : 0), ld_c = in_c + p.appearance_dim;
    const size_t wn[18] = {size_t(32) * in_o, 32, 10 * 32, 10, size_t(32) * in_s, 32, 70 * 32, 70, size_t(32) * ld_c, 32, 30 * 32, 30,
                           size_t(p.appearance_dim) * 7, size_t(p.appearance_dim), 32 * 4, 32, 3 * 32, 3};
    size_t wtotal = 0;
    for (int k = 0; k < 18; ++k) wtotal += (wn[k] + 3) & ~size_t(3);
    float* wtmp = ws->take<float>(wtotal);
    if (!wtmp) return oom();
    segs_decode_grads tmp;
    float** tmp_fields = reinterpret_cast<float**>(&tmp);
    float* const* dst_fields = reinterpret_cast<float* const*>(a->grad_params);
    {
        size_t off = 0;
        for (int k = 0; k < 18; ++k) { tmp_fields[k] = wtmp + off; off += (wn[k] + 3) & ~size_t(3); }
    }
    if ((rc = segs_decode_backward_ex(A, visible, a->anchor, a->anchor_feat, a->offset, a->scaling, a->campos, a->pose, a->params,
                                      dstate, n_vis, P, dL_dmean3D, dL_dcolor, dL_dopacity, dL_dscale, dL_drot, nullptr,
                                      a->grad_anchor, a->grad_anchor_feat, a->grad_offset, a->grad_scaling, &tmp, ws_alloc_cb, ws,
                                      SEGS_DECODE_ACCUMULATE | (a->scaling_is_log ? SEGS_DECODE_LOG_SCALING : 0) |
                                          (concurrent ? SEGS_DECODE_ATOMIC : 0), stream))) return rc;
    {
        float* dst[18]; const float* src[18]; unsigned long long cnt[18];
        int n = 0;
        for (int k = 0; k < 18; ++k) {
            const bool live = (k < 12) || (k < 14 && p.appearance_dim > 0) || (k >= 14 && p.use_feat_bank);
            if (!live || !wn[k]) continue;
            if (!dst_fields[k]) { set_error("mapper view: NULL weight-gradient destination %d", k); return SEGS_ERR_INVALID_ARG; }
            dst[n] = dst_fields[k]; src[n] = tmp_fields[k]; cnt[n] = wn[k]; ++n;
        }
        if ((rc = segs_accumulate(n, dst, src, cnt, concurrent ? 1 : 0, stream))) return rc;
    }
    return SEGS_OK;
}

int segs_mapper_view(segs_workspace* ws, const segs_mapper_view_args* a, segs_mapper_view_result* res, void* stream)
{
    return mapper_view_impl(ws, a, res, false, static_cast<cudaStream_t>(stream));
}

// run fn(view index, lane workspace, lane stream, concurrent) over n_views on n_lanes lanes (see segs_mapper_views)
extern "C++" {
template <class F>
static int run_lanes(int n_views, int n_lanes, segs_workspace* const* ws, void* const* streams, cudaStream_t main_stream, F fn)
{
    if (n_lanes > n_views) n_lanes = n_views;
    int device = 0;
    SEGS_CUDA_CHECK(cudaGetDevice(&device));
    const bool concurrent = n_lanes > 1;

    // RAII: the events are destroyed on every path out of this function (the SEGS_CUDA_CHECK early returns included)
    struct Events {
        cudaEvent_t start = nullptr;
        std::vector<cudaEvent_t> done;
        ~Events() {
            if (start) cudaEventDestroy(start);
            for (cudaEvent_t e : done) if (e) cudaEventDestroy(e);
        }
    } evs;
    evs.done.assign(n_lanes, nullptr);
    SEGS_CUDA_CHECK(cudaEventCreateWithFlags(&evs.start, cudaEventDisableTiming));
    SEGS_CUDA_CHECK(cudaEventRecord(evs.start, main_stream));
    for (int l = 0; l < n_lanes; ++l) SEGS_CUDA_CHECK(cudaEventCreateWithFlags(&evs.done[l], cudaEventDisableTiming));
    cudaEvent_t start = evs.start;
    std::vector<cudaEvent_t>& done = evs.done;

    auto run_lane = [&](int l) -> int {
        cudaStream_t st = static_cast<cudaStream_t>(streams[l]);
        cudaError_t e = cudaStreamWaitEvent(st, start, 0);
        if (e != cudaSuccess) { set_error("lanes: cudaStreamWaitEvent: %s", cudaGetErrorString(e)); return SEGS_ERR_CUDA; }
        int rc = SEGS_OK;
        for (int v = l; v < n_views && rc == SEGS_OK; v += n_lanes) rc = fn(v, ws[l], st, concurrent);
        cudaEventRecord(done[l], st);
        return rc;
    };
    for (int l = 1; l < n_lanes; ++l) {
        segs_workspace* w = ws[l];
        w->submit([w, l, device, &run_lane] {
            cudaSetDevice(device);
            w->job_status = run_lane(l);
            w->job_error = w->job_status ? segs_last_error() : "";
        });
    }
    int rc = run_lane(0);
    for (int l = 1; l < n_lanes; ++l) {
        ws[l]->wait_job();
        if (rc == SEGS_OK && ws[l]->job_status) {
            rc = ws[l]->job_status;
            set_error("lane %d: %s", l, ws[l]->job_error.c_str());
        }
    }
    for (int l = 0; l < n_lanes; ++l) cudaStreamWaitEvent(main_stream, done[l], 0);
    return rc;
}
}  // extern "C++"

static int check_lanes(int n_views, const void* args, const void* results, int n_lanes, segs_workspace* const* ws, void* const* streams)
{
    if (n_views < 0 || n_lanes < 1 || n_lanes > 8 || (n_views > 0 && (!args || !results)) || !ws || !streams) {
        set_error("views: invalid argument"); return SEGS_ERR_INVALID_ARG;
    }
    for (int l = 0; l < n_lanes; ++l)
        if (!ws[l]) { set_error("views: NULL workspace %d", l); return SEGS_ERR_INVALID_ARG; }
    return SEGS_OK;
}

int segs_mapper_views(int n_views, const segs_mapper_view_args* args, segs_mapper_view_result* results,
                      int n_lanes, segs_workspace* const* ws, void* const* streams, void* main_stream)
{
    int rc = check_lanes(n_views, args, results, n_lanes, ws, streams);
    if (rc || n_views == 0) return rc;
    return run_lanes(n_views, n_lanes, ws, streams, static_cast<cudaStream_t>(main_stream),
                     [&](int v, segs_workspace* w, cudaStream_t st, bool concurrent) {
                         return mapper_view_impl(w, args + v, results + v, concurrent, st);
                     });
}

// rasterize + back-propagate one view of explicit Gaussians, gradients accumulated (see segs_raster_views)
static int raster_view_impl(segs_workspace* ws, const segs_raster_view_args* a, segs_mapper_view_result* res,
                            bool concurrent, cudaStream_t stream)
{
    res->n_visible = res->n_gaussians = res->num_rendered = 0;
    const int P = a->P, W = a->width, H = a->height;
    if (P < 0 || W <= 0 || H <= 0) { set_error("raster view: invalid sizes P=%d W=%d H=%d", P, W, H); return SEGS_ERR_INVALID_ARG; }
    if (!a->background || !a->viewmatrix || !a->projmatrix || !a->campos || !a->image_out ||
        (P > 0 && (!a->means3D || !a->colors_precomp || !a->opacities || !a->scales || !a->rotations))) {
        set_error("raster view: NULL required pointer"); return SEGS_ERR_INVALID_ARG;
    }
    ws->reset();
    int rc;
    auto oom = [&]() { set_error("raster view: workspace allocation failed (%zu bytes held)", ws->total); return SEGS_ERR_ALLOC; };
    struct Slot { segs_workspace* ws; char* ptr; };
    Slot geom{ws, nullptr}, binning{ws, nullptr}, img{ws, nullptr};
    auto slot_cb = [](void* u, size_t bytes) -> char* {
        Slot* s = static_cast<Slot*>(u);
        s->ptr = s->ws->alloc(bytes);
        return s->ptr;
    };
    int* radii = a->radii_out ? a->radii_out : ws->take<int>(P > 0 ? P : 1);
    if (!radii) return oom();
    int R = 0;
    if ((rc = segs_raster_forward(slot_cb, &geom, slot_cb, &binning, slot_cb, &img, P, 0, 0, a->background, W, H, a->means3D,
                                  nullptr, a->colors_precomp, a->opacities, a->scales, 1.0f, a->rotations, nullptr,
                                  a->viewmatrix, a->projmatrix, a->campos, a->tan_fovx, a->tan_fovy, 0, a->image_out, radii,
                                  &R, stream))) return rc;
    res->n_gaussians = P;
    res->num_rendered = R;
    if (!a->dL_dout || P == 0) return SEGS_OK;            // forward only
    const size_t Pz = size_t(P);
    float* g2D = ws->take<float>(Pz * 3);
    float* gconic = ws->take<float>(Pz * 4);
    float* gop = ws->take<float>(Pz);
    float* gcol = ws->take<float>(Pz * 3);
    float* g3D = ws->take<float>(Pz * 3);
    float* gcov = ws->take<float>(Pz * 6);
    float* gsc = ws->take<float>(Pz * 3);
    float* grot = ws->take<float>(Pz * 4);
    if (!g2D || !gconic || !gop || !gcol || !g3D || !gcov || !gsc || !grot) return oom();
    if ((rc = segs_raster_backward(P, 0, 0, R, a->background, W, H, a->means3D, nullptr, a->colors_precomp, a->scales, 1.0f,
                                   a->rotations, nullptr, a->viewmatrix, a->projmatrix, a->campos, a->tan_fovx, a->tan_fovy,
                                   radii, geom.ptr, binning.ptr, img.ptr, a->dL_dout, g2D, gconic, gop, gcol, g3D, gcov, nullptr,
                                   gsc, grot, stream))) return rc;
    float* dst[6] = {a->grad_means3D, a->grad_means2D, a->grad_colors, a->grad_opacity, a->grad_scales, a->grad_rotations};
    const float* src[6] = {g3D, g2D, gcol, gop, gsc, grot};
    const unsigned long long w6[6] = {3, 3, 3, 1, 3, 4};
    float* d2[6]; const float* s2[6]; unsigned long long c2[6];
    int n = 0;
    for (int k = 0; k < 6; ++k)
        if (dst[k]) { d2[n] = dst[k]; s2[n] = src[k]; c2[n] = w6[k] * Pz; ++n; }
    if (n) return segs_accumulate(n, d2, s2, c2, concurrent ? 1 : 0, stream);
    return SEGS_OK;
}

int segs_raster_views(int n_views, const segs_raster_view_args* args, segs_mapper_view_result* results,
                      int n_lanes, segs_workspace* const* ws, void* const* streams, void* main_stream)
{
    int rc = check_lanes(n_views, args, results, n_lanes, ws, streams);
    if (rc || n_views == 0) return rc;
    return run_lanes(n_views, n_lanes, ws, streams, static_cast<cudaStream_t>(main_stream),
                     [&](int v, segs_workspace* w, cudaStream_t st, bool concurrent) {
                         return raster_view_impl(w, args + v, results + v, concurrent, st);
                     });
}

}  // extern "C"
