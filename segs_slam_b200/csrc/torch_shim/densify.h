// C++/LibTorch host layer of the densification decisions — the C++ twin of segs_slam_b200/densify.py, i.e. what a
// SEGS-SLAM maintainer puts in place of the bodies of GaussianModel::anchor_growing / adjust_anchor / prune_anchor
// (/root/reference/src/gaussian_model.cpp:1505-1762).  Decisions and new values come from the kernels behind
// include/segs_raster.h (segs_anchor_growing_level, segs_prune_plan, segs_compact_rows); torch owns and resizes the
// tensors (cat / new allocations), fills the constant rows of new anchors and draws the random numbers
// (torch::rand, the reference's torch::rand_like sequence).  Every output is bit-identical to the reference's.
#pragma once

#include <torch/torch.h>

#include <array>
#include <vector>

namespace densify {

// the reference's member names; moments: exp_avg / exp_avg_sq of the six optimizer groups in the reference's group order
// (anchor, offset, anchor_feat, opacity, scaling, rotation), undefined = that tensor has no Adam state
struct AnchorState {
    torch::Tensor anchor, offset, anchor_feat, opacity, scaling /* log */, rotation;
    torch::Tensor opacity_accum, anchor_demon, offset_gradient_accum, offset_denom;
    std::array<torch::Tensor, 6> exp_avg, exp_avg_sq;
    std::vector<std::array<int, 2>> growing_report;   // per level {candidates, new anchors}
    std::array<int, 2> prune_report{{0, 0}};          // {anchors before pruning, after}
};

struct ModelParams {                                  // defaults of GaussianModelParams (include/gaussian_parameters.h:35-41)
    int n_offsets = 10, update_depth = 3, update_init_factor = 16, update_hierachy_factor = 4;
    float voxel_size = 0.001f;
};

// GaussianModel::adjust_anchor (gaussian_model.cpp:1705-1762); `gen`: generator of the random draw (default: the device's
// default generator, like the reference)
void adjust_anchor(AnchorState& st, const ModelParams& mp, int check_interval = 100, float success_threshold = 0.8f,
                   float grad_threshold = 0.0002f, float min_opacity = 0.005f,
                   c10::optional<at::Generator> gen = c10::nullopt);

}  // namespace densify
