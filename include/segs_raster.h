/*
 * segs_raster.h — C ABI of the B200-native (sm_100a) Gaussian-splatting rasterizer,
 * anchor-init kNN and anchor decoder that drop in behind SEGS-SLAM's L6/L7 entry points.
 *
 * Every entry point takes plain device pointers, sizes, a cudaStream_t (as void*) and —
 * where the reference grows opaque byte tensors through std::function<char*(size_t)>
 * (/root/reference/cuda_rasterizer/rasterizer.h:32-34) — C allocation callbacks.
 * All functions return 0 on success and a non-zero status otherwise; the message of the
 * last failure on the calling thread is returned by segs_last_error().  No torch types,
 * no global mutable state: every call is self-contained in the buffers it is handed.
 *
 * Reference interface replaced by each function (file:line under /root/reference):
 *
 *   segs_raster_forward      CudaRasterizer::Rasterizer::forward        cuda_rasterizer/rasterizer.h:31-53,
 *                                                                       rasterizer_impl.cu:198-336
 *                            (called by RasterizeGaussiansCUDA,         src/rasterize_points.cu:36-114)
 *   segs_raster_backward     CudaRasterizer::Rasterizer::backward       cuda_rasterizer/rasterizer.h:73-101,
 *                                                                       rasterizer_impl.cu:397-490
 *                            (called by RasterizeGaussiansBackwardCUDA, src/rasterize_points.cu:116-193)
 *   segs_visible_filter      CudaRasterizer::Rasterizer::visible_filter cuda_rasterizer/rasterizer.h:55-71,
 *                                                                       rasterizer_impl.cu:339-393
 *                            (called by RasterizeGaussiansfilterCUDA,   src/rasterize_points.cu:216-276)
 *   segs_mark_visible        CudaRasterizer::Rasterizer::markVisible    cuda_rasterizer/rasterizer.h:24-29,
 *                                                                       rasterizer_impl.cu:141-153
 *                            (called by markVisible,                    src/rasterize_points.cu:195-214)
 *   segs_project             CudaRasterizer::Rasterizer::project2_image cuda_rasterizer/rasterizer.h:103-125,
 *                                                                       rasterizer_impl.cu:494-585
 *                            (called by RasterizeGaussiansprojectCUDA,  src/rasterize_points.cu:278-362)
 *   segs_knn_mean_dist2      SimpleKNN::knn                             third_party/simple-knn/simple_knn.h:15-19,
 *                                                                       simple_knn.cu:185-221
 *                            (called by distCUDA2,                      third_party/simple-knn/spatial.cu:16-25)
 *   segs_decode_forward /    GaussianRenderer::generate_neural_gaussians src/gaussian_renderer.cpp:214-334
 *   segs_decode_backward     with the module shapes of GaussianModel    src/gaussian_model.cpp:62-98
 *
 * Matrices are 16 floats indexed m[4*col+row] exactly as the reference kernels read them
 * (cuda_rasterizer/auxiliary.h:59-78).  An absent optional input is a null pointer
 * (the reference passes data_ptr() of a 0-element tensor, src/gaussian_rasterizer.cpp:183-193).
 */
#ifndef SEGS_RASTER_H_INCLUDED
#define SEGS_RASTER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGS_ABI_VERSION 1

/* Status codes. */
#define SEGS_OK              0
#define SEGS_ERR_INVALID_ARG 1   /* bad pointer / size / unsupported combination            */
#define SEGS_ERR_CUDA        2   /* a CUDA runtime call or kernel launch failed              */
#define SEGS_ERR_ALLOC       3   /* an allocation callback returned NULL                     */
#define SEGS_ERR_PREFILTERED 4   /* a point was culled although `prefiltered` was set
                                    (the reference __trap()s, auxiliary.h:158-161)           */

/* Grow-and-return-pointer callback: must return a device pointer to at least `bytes`
 * bytes (>= 128-byte aligned), valid until the matching backward call has finished.
 * Mirrors std::function<char*(size_t)> of rasterizer.h:32-34 / resizeFunctional of
 * src/rasterize_points.cu:28-34. */
typedef char* (*segs_alloc_fn)(void* user, size_t bytes);

int         segs_version(void);
const char* segs_last_error(void);

/* ---- forward ------------------------------------------------------------------------ */
/* Returns the number of rendered Gaussian/tile instances in *num_rendered (the int that
 * Rasterizer::forward returns).  out_color is [3,H,W] planar, radii is [P] (may be NULL,
 * as rasterizer.h:53). out_color and radii are fully written for P > 0. */
int segs_raster_forward(
    segs_alloc_fn geom_alloc,    void* geom_user,
    segs_alloc_fn binning_alloc, void* binning_user,
    segs_alloc_fn image_alloc,   void* image_user,
    int P, int D, int M,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* cam_pos,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    float* out_color,
    int* radii,
    int* num_rendered,
    void* stream);

/* ---- backward ----------------------------------------------------------------------- */
/* Gradient outputs are FULLY written (zeros for Gaussians that were not rendered), so the
 * caller may hand in uninitialised memory.  dL_dconic ([P,2,2]) is optional (NULL ok): the
 * reference only uses it internally (src/rasterize_points.cu:150,192).  dL_dmean2D is
 * [P,3] with z = 0 (backward.cu:412).  dL_dsh may be NULL when M == 0. */
int segs_raster_backward(
    int P, int D, int M, int R,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* campos,
    float tan_fovx, float tan_fovy,
    const int* radii,
    char* geom_buffer,
    char* binning_buffer,
    char* image_buffer,
    const float* dL_dpix,
    float* dL_dmean2D,
    float* dL_dconic,
    float* dL_dopacity,
    float* dL_dcolor,
    float* dL_dmean3D,
    float* dL_dcov3D,
    float* dL_dsh,
    float* dL_dscale,
    float* dL_drot,
    void* stream);

/* ---- anchor prefilter / frustum mark / debug projection ----------------------------- */
/* radii[P] is fully written (0 = not visible).  Needs no scratch memory (the reference
 * allocates — and never uses — full geometry and image buffers here). */
int segs_visible_filter(
    int P, int M,
    int width, int height,
    const float* means3D,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    int* radii,
    void* stream);

/* present[P] (1 byte each, C++ bool) = p_view.z > 0.2 (auxiliary.h:155-156). */
int segs_mark_visible(
    int P,
    const float* means3D,
    const float* viewmatrix,
    const float* projmatrix,
    unsigned char* present,
    void* stream);

/* Debug projection: points_image[P,2], radii[P], out_rgb[P,3] — all DEVICE pointers and
 * fully written (zeros for culled points; out_rgb is only written by the SH path, as in
 * the reference, and zero otherwise). */
int segs_project(
    int P, int D, int M,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* cam_pos,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    float* out_rgb,
    float* points_image,
    int* radii,
    void* stream);

/* ---- anchor-init kNN ---------------------------------------------------------------- */
/* mean_dists[P] = mean squared distance to the 3 nearest neighbours (self excluded).
 * scratch grows one opaque device buffer (all temporaries; no cudaMalloc inside). */
int segs_knn_mean_dist2(
    int P,
    const float* points,
    float* mean_dists,
    segs_alloc_fn scratch_alloc, void* scratch_user,
    void* stream);

/* ---- structure-enhanced anchor decode -------------------------------------------------- */
/* Replaces GaussianRenderer::generate_neural_gaussians (src/gaussian_renderer.cpp:214-334): the
 * visible-anchor gather, the optional feature-bank mix, the pose "appearance" Linear(7 -> app), the
 * opacity / covariance / colour MLPs (module shapes: src/gaussian_model.cpp:60-98; feat_dim = 32,
 * n_offsets = 10 as in every shipped config), the opacity > 0 mask and the assembly of the
 * surviving neural Gaussians — one fused kernel instead of ~40 ATen launches and the
 * [A*10, 22] temporary.
 *
 * Weights use torch::nn::Linear's layout: weight [out, in] row-major, bias [out]; all DEVICE
 * pointers.  Input column order of the first layers: [feat(32), ob_view(3), ob_dist(1, only with
 * the add_*_dist flag), appearance(appearance_dim, colour MLP only)].  app_* may be NULL when
 * appearance_dim == 0, bank_* when use_feat_bank == 0. */
typedef struct segs_decode_params {
    const float* opacity_w1; const float* opacity_b1; const float* opacity_w2; const float* opacity_b2; /* [32,35+d] [32] [10,32] [10] */
    const float* cov_w1;     const float* cov_b1;     const float* cov_w2;     const float* cov_b2;     /* [32,35+d] [32] [70,32] [70] */
    const float* color_w1;   const float* color_b1;   const float* color_w2;   const float* color_b2;   /* [32,35+d+app] [32] [30,32] [30] */
    const float* app_w;      const float* app_b;                                                        /* [app,7] [app] */
    const float* bank_w1;    const float* bank_b1;    const float* bank_w2;    const float* bank_b2;    /* [32,4] [32] [3,32] [3] */
    int appearance_dim;      /* 0..32 */
    int use_feat_bank;
    int add_opacity_dist, add_cov_dist, add_color_dist;
} segs_decode_params;

/* Gradient outputs of segs_decode_backward, same shapes as the weights; fully written. */
typedef struct segs_decode_grads {
    float* opacity_w1; float* opacity_b1; float* opacity_w2; float* opacity_b2;
    float* cov_w1;     float* cov_b1;     float* cov_w2;     float* cov_b2;
    float* color_w1;   float* color_b1;   float* color_w2;   float* color_b2;
    float* app_w;      float* app_b;
    float* bank_w1;    float* bank_b1;    float* bank_w2;    float* bank_b2;
} segs_decode_grads;

/* Bytes of caller-owned state that forward fills and backward reads (anchor ordinals, row starts,
 * masks, look-back words and — left by the default forward kernel for the backward — the layer-1
 * pre-activations [96] and second-layer outputs [110] of every visible anchor: 836 bytes per anchor). */
size_t segs_decode_state_bytes(int A);

/* visible_mask: [A] bytes (C++ bool) or NULL = all visible.  scaling = exp(_scaling) [A,6]
 * (GaussianModel::get_scaling).  pose = {t.x, t.y, t.z, q.w, q.x, q.y, q.z} of the keyframe (HOST
 * floats, gaussian_renderer.cpp:258-261).  Row outputs have capacity n_vis*10 <= A*10 rows and are
 * written compacted, in (anchor, offset) order: xyz [.,3], color [.,3], opacity [.,1],
 * out_scaling [.,3], rot [.,4].  neural_opacity [A*10] and mask [A*10] (bytes) are indexed by
 * (visible-anchor ordinal * 10 + offset).  counts (HOST) receives {visible anchors, emitted
 * Gaussians}; the call waits for the kernel, like the reference's boolean indexing does. */
int segs_decode_forward(
    int A, const unsigned char* visible_mask,
    const float* anchor, const float* anchor_feat, const float* offset, const float* scaling,
    const float* camera_center, const float* pose,
    const segs_decode_params* params,
    float* xyz, float* color, float* opacity, float* out_scaling, float* rot,
    float* neural_opacity, unsigned char* mask,
    char* state, int* counts,
    void* stream);

/* Which forward kernel segs_decode_forward launches (both are sm_100a tcgen05 kernels, same results to ~1e-6):
 *   2 (default)  visible anchors are listed first, tile = 128 visible anchors, 512 threads per CTA, BOTH layers of the
 *                three MLPs on the tensor cores (51 tcgen05.mma per tile, 3xTF32), rows assembled by thread = row;
 *   1            round 1's kernel: tiles in anchor-index order, thread = anchor, first layers on the tensor cores,
 *                second layers as FP32 FFMA chains.
 * The environment variable SEGS_DECODE_VARIANT (1 | 2) sets the initial value. */
int segs_decode_set_variant(int variant);
int segs_decode_get_variant(void);

/* g_* are the gradients w.r.t. the compacted row outputs (n_out rows); g_neural_opacity
 * ([n_vis*10], may be NULL) the gradient w.r.t. the un-masked opacity output.  d_anchor [A,3],
 * d_anchor_feat [A,32], d_offset [A,10,3], d_scaling [A,6] (w.r.t. the `scaling` input) and every
 * tensor of *dparams are fully written (zeros for invisible anchors).  scratch grows one opaque
 * device buffer for the per-anchor factors of the weight gradients. */
int segs_decode_backward(
    int A, const unsigned char* visible_mask,
    const float* anchor, const float* anchor_feat, const float* offset, const float* scaling,
    const float* camera_center, const float* pose,
    const segs_decode_params* params,
    const char* state, int n_vis, int n_out,
    const float* g_xyz, const float* g_color, const float* g_opacity, const float* g_scaling, const float* g_rot,
    const float* g_neural_opacity,
    float* d_anchor, float* d_anchor_feat, float* d_offset, float* d_scaling,
    const segs_decode_grads* dparams,
    segs_alloc_fn scratch_alloc, void* scratch_user,
    void* stream);

/* segs_decode_backward with options (flags, OR-ed):
 *   SEGS_DECODE_ACCUMULATE   d_anchor / d_anchor_feat / d_offset / d_scaling are ADDED to instead of being zeroed
 *                            and written (gradient accumulation over the keyframe batch; every anchor row is owned
 *                            by one thread, no atomics).  *dparams is still zeroed and written.
 *   SEGS_DECODE_LOG_SCALING  d_scaling is the gradient w.r.t. _scaling = log(scaling) (the trainable tensor,
 *                            gaussian_model.cpp:186-189), i.e. multiplied by `scaling`.
 *   SEGS_DECODE_ATOMIC       with ACCUMULATE: add with RED.ADD.F32 (several views accumulate into the same arrays
 *                            from concurrent streams; the summation order then varies from run to run). */
#define SEGS_DECODE_ACCUMULATE  1
#define SEGS_DECODE_LOG_SCALING 2
#define SEGS_DECODE_ATOMIC      4
int segs_decode_backward_ex(
    int A, const unsigned char* visible_mask,
    const float* anchor, const float* anchor_feat, const float* offset, const float* scaling,
    const float* camera_center, const float* pose,
    const segs_decode_params* params,
    const char* state, int n_vis, int n_out,
    const float* g_xyz, const float* g_color, const float* g_opacity, const float* g_scaling, const float* g_rot,
    const float* g_neural_opacity,
    float* d_anchor, float* d_anchor_feat, float* d_offset, float* d_scaling,
    const segs_decode_grads* dparams,
    segs_alloc_fn scratch_alloc, void* scratch_user,
    int flags, void* stream);

/* Densification statistics of one view (GaussianModel::training_statis, src/gaussian_model.cpp:1459-1503), from the
 * state segs_decode_forward left: for every visible anchor a, opacity_accum[a] += sum over its 10 offsets of
 * max(neural_opacity, 0) and anchor_demon[a] += 1; for every emitted offset o whose Gaussian was rendered (radii > 0),
 * offset_gradient_accum[a*10+o] += |dL_dmean2D.xy| and offset_denom[a*10+o] += 1.  radii [n_out], dL_dmean2D [n_out,3]
 * are the rasterizer's outputs for the decoded Gaussians.  atomic != 0: RED.ADD (concurrent views). */
int segs_training_statis(int A, const char* decode_state, int n_vis, const float* neural_opacity, const int* radii,
                         const float* dL_dmean2D, float* opacity_accum, float* anchor_demon,
                         float* offset_gradient_accum, float* offset_denom, int atomic, void* stream);

/* ---- frequency-domain regularisation ---------------------------------------------------
 *   segs_freq_*   loss_utils::high_frequency_loss / multi_scale_loss      include/loss_utils.h:125-169, 210-237,
 *                 as called at                                             src/gaussian_mapper.cpp:930-945
 * value = weight * sum_s scales[s] * mean over [C,h_s,w_s] of | |fft2(D_s image)| - |fft2(D_s gt)| |, D_s = ATen's bilinear
 * interpolate(scale_factor = s, align_corners = false, recompute_scale_factor = true) ([h_s, w_s] = floor([H, W] * s));
 * high_frequency_loss is the single scale 1.  The reference's frequency mask is empty for C = 3 (it is indexed on the
 * channel dimension), so the whole spectrum counts; low_freq_loss is identically zero for the same reason (see freq.cu).
 * A plan owns the cuFFT plans of its scales and a working spectrum on the device it was created on; it serves one stream
 * at a time. */
typedef struct segs_freq_plan segs_freq_plan;
int    segs_freq_plan_create(int C, int H, int W, int n_scales /* <= 4 */, const float* scales /* HOST, each <= 1 */, segs_freq_plan** out);
int    segs_freq_plan_destroy(segs_freq_plan* plan);
size_t segs_freq_mag_floats(const segs_freq_plan* plan);          /* floats of a target-magnitude buffer */
/* gt_mag (DEVICE, segs_freq_mag_floats floats) = |fft2(D_s (gt * row_mask))| of every scale: per keyframe, reusable */
int    segs_freq_target(segs_freq_plan* plan, const float* gt, const float* row_mask /* [C,H] or NULL */, float* gt_mag, void* stream);
/* loss_out (DEVICE scalar, may be NULL) += value;  dL_dimage [C,H,W] (may be NULL) += dL_dloss (DEVICE scalar, NULL = 1) *
 * d value / d image (through the row mask).  Deterministic. */
int    segs_freq_loss(segs_freq_plan* plan, const float* image, const float* row_mask, const float* gt_mag, float weight,
                      const float* dL_dloss, float* loss_out, float* dL_dimage, void* stream);

/* ---- keyframe image ingest ----------------------------------------------------------------
 *   segs_ingest_image      camera.undistortImage (cv::remap, INTER_LINEAR)      include/camera.h:106-115
 *                          + tensor_utils::cvMat2TorchTensor_Float32            include/tensor_utils.h:40-69
 *   segs_resize_bilinear   the Gaussian-pyramid levels (cv::cuda::resize)       src/gaussian_mapper.cpp:621-632, 1289-1298
 * src_hwc: DEVICE [src_H, src_W, C] interleaved FP32 (the uploaded cv::Mat); map_x / map_y: DEVICE [H, W] FP32 from
 * cv::initUndistortRectifyMap(..., CV_32F, ...) (both NULL = no undistortion, src size = output size);
 * dst_chw: DEVICE [C, H, W].  OpenCV's arithmetic: maps quantised to 1/32 pixel, FP32 bilinear table, constant-0 border. */
int segs_ingest_image(int H, int W, int C, int src_H, int src_W, const float* src_hwc, const float* map_x, const float* map_y,
                      float* dst_chw, void* stream);
/* cv::resize(INTER_LINEAR) of a planar FP32 image [C,H,W] -> [C,h,w] */
int segs_resize_bilinear(int C, int H, int W, const float* src_chw, int h, int w, float* dst_chw, void* stream);

/* ---- densification decisions ---------------------------------------------------------
 *   segs_anchor_growing_level   one level i of GaussianModel::anchor_growing     src/gaussian_model.cpp:1556-1703
 *   segs_prune_plan             the statistics update + prune decision of        src/gaussian_model.cpp:1716-1755
 *                               GaussianModel::adjust_anchor
 *   segs_compact_rows           the row gathers of GaussianModel::prune_anchor   src/gaussian_model.cpp:1505-1555
 * The host (segs_slam_b200/densify.py) owns the tensors and their resizing, like the reference's torch::cat /
 * index sequences; every decision and every new value is computed here.
 *
 * segs_anchor_growing_level: candidates are the first `init_slots` (= anchors before growing x n_offsets) offset slots
 * with |offset_gradient_accum / offset_denom| >= cur_threshold, offset_denom > denom_threshold and
 * rand_values > rand_threshold (the reference draws rand_values with torch::rand_like).  Their positions
 * anchor + offset * exp(log_scaling[:, :3]) are snapped to voxels of side cur_size; voxels already occupied by one of the
 * A_now current anchors are dropped; the survivors come out in lexicographic voxel order (at::unique_dim):
 *   *new_anchor [n_new,3] = voxel * cur_size,  *new_feat [n_new,feat_dim] = per-voxel maximum of the candidates' anchor
 *   features (torch_scatter's scatter_max), both inside ONE block obtained from out_alloc.
 * scratch_alloc is called at most twice and must return a NEW block each time; all blocks stay owned by the caller and
 * must remain valid until the call returns.  Synchronises `stream` twice (candidate count, n_new). */
int segs_anchor_growing_level(
    int A_now, int init_slots, int n_offsets, int feat_dim,
    const float* anchor, const float* offset, const float* log_scaling, const float* anchor_feat,
    const float* offset_gradient_accum, const float* offset_denom, const float* rand_values,
    float denom_threshold, float cur_threshold, float rand_threshold, float cur_size,
    segs_alloc_fn scratch_alloc, void* scratch_user, segs_alloc_fn out_alloc, void* out_user,
    float** new_anchor, float** new_feat, int* n_candidates, int* n_new, void* stream);

/* In place over A anchors (A = after growing; the first init_slots offset slots carry statistics):
 *   offset slots with offset_denom > denom_threshold: offset_denom = offset_gradient_accum = 0;
 *   keep[a] = !(opacity_accum[a] < min_opacity * anchor_demon[a] && anchor_demon[a] > anchor_threshold);
 *   anchors with anchor_demon > anchor_threshold: opacity_accum = anchor_demon = 0;
 *   keep_index = exclusive scan of keep, *n_keep = its total (one stream synchronisation). */
size_t segs_prune_scratch_words(int A);
int segs_prune_plan(
    int A, int init_slots, float* opacity_accum, float* anchor_demon, float* offset_gradient_accum, float* offset_denom,
    float denom_threshold, float anchor_threshold, float min_opacity,
    unsigned int* keep, unsigned int* keep_index, unsigned int* scratch, int* n_keep, void* stream);

/* dst[keep_index[a]][:] = src[a][:] for every a with keep[a] (rows of row_floats floats); clamp_from >= 0: columns
 * >= clamp_from are clamped to <= clamp_max on the way (prune_anchor clamps _scaling[:, 3:] to 0.05). */
int segs_compact_rows(int A, int row_floats, const unsigned int* keep, const unsigned int* keep_index, const float* src,
                      float* dst, int clamp_from, float clamp_max, void* stream);

/* ---- one keyframe view of the batched mapping step (BASELINE config 4) -------------
 *   segs_mapper_view      the body of GaussianMapper::trainForOneIteration  src/gaussian_mapper.cpp:870-950:
 *                         prefilter_voxel (src/gaussian_renderer.cpp:131-199) -> generate_neural_gaussians
 *                         (:214-334) -> render (:40-127) -> loss (gaussian_mapper.cpp:908-925) -> backward,
 *                         with the parameter gradients ACCUMULATED into the caller's bucket.
 * A workspace is a reusable device arena (grown with cudaMalloc on demand, reset per view); one workspace
 * serves one stream at a time. */
typedef struct segs_workspace segs_workspace;
int    segs_workspace_create(segs_workspace** out);
int    segs_workspace_destroy(segs_workspace* ws);
size_t segs_workspace_bytes(const segs_workspace* ws);

typedef struct segs_mapper_view_args {
    /* replicated model (DEVICE) */
    int A;
    const float* anchor;            /* [A,3]    */
    const float* anchor_feat;       /* [A,32]   */
    const float* offset;            /* [A,10,3] */
    const float* scaling;           /* [A,6] = exp(_scaling) (GaussianModel::get_scaling)                   */
    int scaling_is_log;             /* != 0: grad_scaling is w.r.t. _scaling (chain through the exp)         */
    const float* filter_scales;     /* [A,3] = get_scaling()[:, :3] contiguous (gaussian_renderer.cpp:172)  */
    const float* filter_rotations;  /* [A,4] = get_rotation() (normalised)                                   */
    const segs_decode_params* params;
    /* keyframe */
    int width, height;
    float tan_fovx, tan_fovy;
    const float* viewmatrix;        /* DEVICE, 16 floats, m[4*col+row] */
    const float* projmatrix;        /* DEVICE */
    const float* campos;            /* DEVICE, 3 floats */
    const float* pose;              /* HOST, {t.xyz, q.wxyz} */
    const float* background;        /* DEVICE, 3 floats */
    const float* gt_image;          /* DEVICE [3,H,W] */
    const float* row_mask;          /* DEVICE [3,H] or NULL (mask_rgb, gaussian_mapper.cpp:911-915) */
    float lambda_dssim;             /* loss = (1-l) L1 + l (1-SSIM) + scaling_reg_weight * mean(prod(scaling)) */
    float scaling_reg_weight;       /* 0.01 in the reference (gaussian_mapper.cpp:921) */
    /* accumulated outputs (DEVICE; += ) */
    float* grad_anchor;             /* [A,3]    */
    float* grad_anchor_feat;        /* [A,32]   */
    float* grad_offset;             /* [A,10,3] */
    float* grad_scaling;            /* [A,6]    */
    const segs_decode_grads* grad_params;   /* every live MLP tensor */
    float* loss_accum;              /* scalar */
    /* optional per-view outputs (DEVICE, may be NULL) */
    float* image_out;               /* [3,H,W] rendered image */
    float* loss_terms_out;          /* {Ll1, ssim, photometric loss} */
    float* dL_dmean2D_out;          /* [A*10,3] capacity: screen-space gradient of the emitted Gaussians (densification) */
    int*   radii_out;               /* [A*10] capacity */
    /* optional densification statistics (segs_training_statis), all four or none; += */
    float* stat_opacity_accum;          /* [A]    */
    float* stat_anchor_demon;           /* [A]    */
    float* stat_offset_gradient_accum;  /* [A*10] */
    float* stat_offset_denom;           /* [A*10] */
    /* optional frequency regularisation (src/gaussian_mapper.cpp:930-945): loss += lambda_frequency_high *
     * (use_multi_resolution ? multi_scale_loss(scales 1, 1/2, ..., 1/2^(freq_scale_num-1)) : high_frequency_loss);
     * 0 = off.  gt_freq_mag: segs_freq_target of this keyframe's gt_image for the same scales, or NULL (computed on the fly). */
    float lambda_frequency_high;
    int   use_multi_resolution;
    int   freq_scale_num;               /* 1..4 */
    const float* gt_freq_mag;
} segs_mapper_view_args;

typedef struct segs_mapper_view_result {
    int n_visible;                  /* anchors that passed the prefilter */
    int n_gaussians;                /* neural Gaussians emitted by the decode */
    int num_rendered;               /* tile instances */
} segs_mapper_view_result;

int segs_mapper_view(segs_workspace* ws, const segs_mapper_view_args* args, segs_mapper_view_result* result, void* stream);

/* A batch of views on `n_lanes` concurrent lanes (lane l = workspace ws[l] + stream streams[l] + one persistent host
 * thread owned by the workspace; lane 0 runs on the calling thread): lane l issues views l, l + n_lanes, ...  The
 * latency-bound stages of one view (sorts, decode, the two host read-backs) overlap the issue-bound blend kernels of
 * another.  With n_lanes > 1 the accumulated outputs are updated with RED.ADD.F32.  Every lane stream first waits for
 * the work already queued on `main_stream`, and `main_stream` waits for every lane before the call returns (the host
 * does not wait for the GPU beyond the read-backs).  args: HOST array [n_views]; results: HOST array [n_views]. */
int segs_mapper_views(int n_views, const segs_mapper_view_args* args, segs_mapper_view_result* results,
                      int n_lanes, segs_workspace* const* ws, void* const* streams, void* main_stream);

/* One view of a keyframe batch over EXPLICIT Gaussians (precomputed colours, scale + quaternion — the
 * GaussianRasterizer call of src/gaussian_renderer.cpp:86-127 followed by its backward, src/gaussian_rasterizer.cpp:
 * 88-154): forward into image_out, and, when dL_dout is not NULL, backward with the parameter gradients ADDED to the
 * grad_* accumulators (NULL = not wanted).  All pointers DEVICE. */
typedef struct segs_raster_view_args {
    int P;
    const float* means3D;           /* [P,3] */
    const float* colors_precomp;    /* [P,3] */
    const float* opacities;         /* [P,1] */
    const float* scales;            /* [P,3] */
    const float* rotations;         /* [P,4] */
    const float* background;        /* [3]   */
    int width, height;
    float tan_fovx, tan_fovy;
    const float* viewmatrix;        /* 16 floats, m[4*col+row] */
    const float* projmatrix;
    const float* campos;            /* [3] */
    const float* dL_dout;           /* [3,H,W] or NULL (forward only) */
    float* image_out;               /* [3,H,W] */
    int*   radii_out;               /* [P] or NULL */
    float* grad_means3D;            /* [P,3] += */
    float* grad_means2D;            /* [P,3] += (screen-space gradient; drives densification, gaussian_model.cpp:1488-1499) */
    float* grad_colors;             /* [P,3] += */
    float* grad_opacity;            /* [P,1] += */
    float* grad_scales;             /* [P,3] += */
    float* grad_rotations;          /* [P,4] += */
} segs_raster_view_args;

/* A batch of such views on n_lanes concurrent lanes; lanes, streams and ordering exactly as segs_mapper_views.
 * results[v].num_rendered is filled (n_visible is unused). */
int segs_raster_views(int n_views, const segs_raster_view_args* args, segs_mapper_view_result* results,
                      int n_lanes, segs_workspace* const* ws, void* const* streams, void* main_stream);

/* ---- mapper loss and optimizer ---------------------------------------
 *   segs_loss_l1_ssim_*   loss_utils::l1_loss / ssim / _ssim           include/loss_utils.h:29-32, 50-127,
 *                         as combined at                                src/gaussian_mapper.cpp:917-925
 *   segs_scaling_reg      0.01 * scaling.prod(1).mean()                 src/gaussian_mapper.cpp:922-925
 *   segs_adam_step        torch::optim::Adam::step over the groups of   src/gaussian_model.cpp:620-872,
 *                         GaussianModel::trainingSetup                  stepped at src/gaussian_mapper.cpp:1003-1006
 *   segs_accumulate       gradient accumulation over the keyframe batch (new construct)     */

/* Bytes of caller-owned state the loss forward fills and the backward reads (three derivative maps of
 * the SSIM index, the per-CTA partial sums). */
size_t segs_loss_state_bytes(int C, int H, int W);

/* loss = w_l1 * mean|x - y| + w_ssim * mean(SSIM_11x11,sigma1.5(x, y)) + bias over image, gt [C,H,W]
 * (both multiplied by row_mask [C,H] when it is not NULL — the `mask_rgb` of gaussian_mapper.cpp:911-915).
 * The mapper's loss is (w_l1, w_ssim, bias) = (1 - lambda_dssim, -lambda_dssim, lambda_dssim).
 * loss_out (DEVICE, 3 floats) = { mean|x-y|, mean SSIM, loss }.  Deterministic (fixed-order sums). */
int segs_loss_l1_ssim_forward(
    int C, int H, int W, const float* image, const float* gt, const float* row_mask,
    float w_l1, float w_ssim, float bias, float* loss_out, char* state, void* stream);

/* dL_dimage [C,H,W] = dL_dloss * d loss / d image.  dL_dloss: DEVICE scalar, NULL = 1. */
int segs_loss_l1_ssim_backward(
    int C, int H, int W, const float* image, const float* gt, const float* row_mask,
    float w_l1, float w_ssim, const float* dL_dloss, char* state, float* dL_dimage, void* stream);

/* reg = weight * mean_i(s_i0 * s_i1 * s_i2) over scaling [n,3].  reg_out (DEVICE scalar, may be NULL) is
 * ADDED to; dL_dscaling [n,3] (may be NULL) is ADDED to with dL_dloss (DEVICE scalar, NULL = 1) * d reg. */
int segs_scaling_reg(int n, const float* scaling, float weight, const float* dL_dloss, float* dL_dscaling,
                     float* reg_out, void* stream);

/* One tensor (= one parameter group of the reference's optimizer) of a fused Adam step.  `offset`/`count`
 * locate its gradient and moments in the flat arrays; consecutive tensors must tile them contiguously. */
typedef struct segs_adam_tensor {
    float* param;                  /* DEVICE, count floats, updated in place */
    unsigned long long offset;     /* first element in grad_flat / exp_avg_flat / exp_avg_sq_flat */
    unsigned long long count;
    float lr, beta1, beta2, eps, weight_decay;
    long long step;                /* 1-based step number used for the bias corrections */
} segs_adam_tensor;

/* tensors: HOST array.  g = grad_flat * grad_scale; Adam (amsgrad off) as LibTorch 2.0.1;
 * zero_grad != 0 clears grad_flat behind the read (ready for the next batch). */
int segs_adam_step(int n_tensors, const segs_adam_tensor* tensors, float* grad_flat, float* exp_avg_flat,
                   float* exp_avg_sq_flat, float grad_scale, int zero_grad, void* stream);

/* dst[k][i] += src[k][i], i < counts[k], for all k in one launch (HOST arrays of DEVICE pointers);
 * atomic != 0: with RED.ADD.F32 (concurrent accumulators). */
int segs_accumulate(int n_arrays, float* const* dst, const float* const* src, const unsigned long long* counts,
                    int atomic, void* stream);

/* ---- per-stage device timing (bench.py roofline) -------------------------------------- */
/* When enabled (per host thread), segs_raster_forward / segs_raster_backward bracket their
 * stages with CUDA events on the caller's stream.  segs_profile_read synchronises those
 * events and returns the milliseconds of the most recent forward and backward call:
 *   ms[0] preprocess   ms[1] depth order + offsets   ms[2] instance emission + tile sort + ranges
 *   ms[3] blend forward   ms[4] blend backward   ms[5] preprocess backward
 * (a stage that did not run since the last read reports 0). */
#define SEGS_PROFILE_STAGES 6
/* Number of kernels this library has launched in this process so far (all host threads). */
unsigned long long segs_launch_count(void);
int segs_profile_enable(int on);
/* Host threads waiting for the library's small read-backs (num_rendered, decode counts) spin by default; on != 0
 * makes every host thread that has not waited yet sleep instead (cudaEventBlockingSync) — for boxes that run more
 * waiting threads (ranks x lanes) than they have cores.  Call before the first forward. */
int segs_set_blocking_sync(int on);
int segs_profile_read(float* ms /* [SEGS_PROFILE_STAGES] */);

/* Work counters of the blend kernels (development builds with -DSEGS_BLEND_STATS; zeros otherwise):
 * forward {records staged, (Gaussian, sub-tile) pairs evaluated, pairs reaching a pixel, blended (Gaussian, pixel) pairs},
 * backward likewise. */
int segs_debug_blend_stats(unsigned long long* out8, int reset);

/* ---- inspection of the opaque buffers (parity tests, debugging) --------------------- */
/* Returns in *ptr / *bytes the device address and size of a named section of the
 * buffers produced by segs_raster_forward for the given (P, R, width, height).
 * Sections: "depths" f32[P], "tiles_touched" u32[P], "rect" u16[P,4] (x0,y0,x1,y1),
 * "rec" f32[P,12] (x,y,hx,hy | conic.x,conic.y,conic.z,opacity | r,g,b,depth),
 * "cov3D" f32[6,P] (planes), "depth_order" u32[P], "point_list" u32[R], "ranges" u32[T,2],
 * "final_T" f32[N], "n_contrib" u32[N]. */
int segs_buffer_section(
    const char* name,
    char* geom_buffer, char* binning_buffer, char* image_buffer,
    int P, int R, int width, int height,
    void** ptr, size_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* SEGS_RASTER_H_INCLUDED */
