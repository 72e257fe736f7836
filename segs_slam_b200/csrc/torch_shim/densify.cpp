// See densify.h.
#include "densify.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <cmath>
#include <string>

#include "../../../include/segs_raster.h"

namespace densify {

namespace {

void raise_if(int status) {
    if (status == SEGS_OK) return;
    TORCH_CHECK(false, "segs_raster: ", std::string(segs_last_error()), " (status ", status, ")");
}

float* fptr(const torch::Tensor& t) { return t.numel() == 0 ? nullptr : t.data_ptr<float>(); }

// allocation callbacks: a NEW block per call, kept alive by the list
struct Blocks { std::vector<torch::Tensor> list; torch::TensorOptions opt; };
char* block_cb(void* user, size_t bytes) {
    auto* b = static_cast<Blocks*>(user);
    b->list.push_back(torch::empty({static_cast<int64_t>(bytes > 0 ? bytes : 1)}, b->opt.dtype(torch::kUInt8)));
    return static_cast<char*>(b->list.back().data_ptr());
}

std::array<torch::Tensor*, 6> params_of(AnchorState& st) {
    return {&st.anchor, &st.offset, &st.anchor_feat, &st.opacity, &st.scaling, &st.rotation};
}

}  // namespace

void adjust_anchor(AnchorState& st, const ModelParams& mp, int check_interval, float success_threshold, float grad_threshold,
                   float min_opacity, c10::optional<at::Generator> gen)
{
    torch::NoGradGuard ng;
    TORCH_CHECK(st.anchor.is_cuda() && st.anchor.scalar_type() == torch::kFloat32, "densify: FP32 CUDA tensors expected (no CPU path)");
    const c10::cuda::CUDAGuard guard(st.anchor.device());
    void* stream = static_cast<void*>(at::cuda::getCurrentCUDAStream().stream());
    const auto opt = st.anchor.options();
    const int k = mp.n_offsets;
    const int64_t A0 = st.anchor.size(0);
    const int feat_dim = static_cast<int>(st.anchor_feat.size(1));
    const int init_slots = static_cast<int>(A0 * k);
    // host scalars in the reference's types: int * float -> float (:1742), ... * 0.5 (double) compared in FP32 (:1713)
    const float anchor_threshold = static_cast<float>(check_interval) * success_threshold;
    const float denom_threshold = static_cast<float>(static_cast<double>(anchor_threshold) * 0.5);
    st.offset_gradient_accum = st.offset_gradient_accum.contiguous();
    st.offset_denom = st.offset_denom.contiguous();
    st.growing_report.clear();

    // ---- anchor_growing (:1556-1703) ----
    for (int i = 0; i < mp.update_depth; ++i) {
        // the reference draws before it decides to skip the level (:1566 vs :1570-1575)
        torch::Tensor rnd = torch::rand({init_slots}, gen, opt);
        const int64_t A_now = st.anchor.size(0);
        if (A_now * k - init_slots == 0 && i > 0) { st.growing_report.push_back({0, 0}); continue; }
        const float cur_threshold = static_cast<float>(grad_threshold * std::pow(std::floor(mp.update_hierachy_factor / 2), i));   // :1561
        const float rand_threshold = static_cast<float>(std::pow(0.5, i + 1));                                                      // :1566
        const int size_factor = static_cast<int>(std::floor(mp.update_init_factor / std::pow(mp.update_hierachy_factor, i)));     // :1585
        const float cur_size = mp.voxel_size * size_factor;                                                                         // :1586
        torch::Tensor anchor = st.anchor.contiguous(), offset = st.offset.contiguous(), scaling = st.scaling.contiguous(),
                      feat = st.anchor_feat.contiguous();
        Blocks scratch{{}, opt}, out{{}, opt};
        float *new_anchor_p = nullptr, *new_feat_p = nullptr;
        int n_cand = 0, n_new = 0;
        raise_if(segs_anchor_growing_level(static_cast<int>(A_now), init_slots, k, feat_dim, fptr(anchor), fptr(offset), fptr(scaling),
                                           fptr(feat), fptr(st.offset_gradient_accum), fptr(st.offset_denom), fptr(rnd), denom_threshold,
                                           cur_threshold, rand_threshold, cur_size, block_cb, &scratch, block_cb, &out, &new_anchor_p,
                                           &new_feat_p, &n_cand, &n_new, stream));
        st.growing_report.push_back({n_cand, n_new});
        if (n_new == 0) continue;
        const int64_t n = n_new;
        torch::Tensor candidate_anchor = torch::from_blob(new_anchor_p, {n, 3}, opt);
        torch::Tensor new_feat = torch::from_blob(new_feat_p, {n, feat_dim}, opt);
        torch::Tensor new_scaling = torch::log(torch::full({n, 6}, cur_size, opt));                     // :1625-1626
        torch::Tensor new_rotation = torch::zeros({n, 4}, opt);
        new_rotation.index_put_({torch::indexing::Slice(), 0}, 1.0);                                    // :1627-1628
        torch::Tensor tenth = 0.1f * torch::ones({n, 1}, opt);
        torch::Tensor new_opacities = torch::log(tenth / (1 - tenth));                                  // :1630-1631
        torch::Tensor new_offsets = torch::zeros({n, k, 3}, opt);                                       // :1639
        st.anchor_demon = torch::cat({st.anchor_demon, torch::zeros({n, 1}, opt)}, 0);
        st.opacity_accum = torch::cat({st.opacity_accum, torch::zeros({n, 1}, opt)}, 0);
        std::array<torch::Tensor, 6> ext = {candidate_anchor, new_offsets, new_feat, new_opacities, new_scaling, new_rotation};
        auto ps = params_of(st);
        for (int g = 0; g < 6; ++g) {
            if (st.exp_avg[g].defined()) st.exp_avg[g] = torch::cat({st.exp_avg[g], torch::zeros_like(ext[g])}, 0);
            if (st.exp_avg_sq[g].defined()) st.exp_avg_sq[g] = torch::cat({st.exp_avg_sq[g], torch::zeros_like(ext[g])}, 0);
            *ps[g] = torch::cat({*ps[g], ext[g]}, 0);       // copies out of the kernel's output block
        }
    }

    // ---- statistics update + prune decision (:1716-1755), prune_anchor (:1505-1555) ----
    const int64_t A = st.anchor.size(0);
    if (A > A0) {
        torch::Tensor pad = torch::zeros({(A - A0) * k, 1}, opt);
        st.offset_denom = torch::cat({st.offset_denom, pad}, 0);
        st.offset_gradient_accum = torch::cat({st.offset_gradient_accum, pad}, 0);
    }
    st.opacity_accum = st.opacity_accum.contiguous();
    st.anchor_demon = st.anchor_demon.contiguous();
    torch::Tensor keep = torch::empty({A}, opt.dtype(torch::kInt32)), keep_index = torch::empty({A}, opt.dtype(torch::kInt32));
    torch::Tensor scratch = torch::empty({static_cast<int64_t>(segs_prune_scratch_words(static_cast<int>(A)))}, opt.dtype(torch::kInt32));
    int n_keep = 0;
    raise_if(segs_prune_plan(static_cast<int>(A), init_slots, fptr(st.opacity_accum), fptr(st.anchor_demon), fptr(st.offset_gradient_accum),
                             fptr(st.offset_denom), denom_threshold, anchor_threshold, min_opacity,
                             reinterpret_cast<unsigned int*>(keep.data_ptr<int>()), reinterpret_cast<unsigned int*>(keep_index.data_ptr<int>()),
                             reinterpret_cast<unsigned int*>(scratch.data_ptr<int>()), &n_keep, stream));
    auto compact = [&](const torch::Tensor& t, int row_floats, int clamp_from, float clamp_max) {
        torch::Tensor src = t.contiguous();
        torch::Tensor dst = torch::empty({static_cast<int64_t>(n_keep), row_floats}, opt);
        if (n_keep > 0)
            raise_if(segs_compact_rows(static_cast<int>(A), row_floats, reinterpret_cast<unsigned int*>(keep.data_ptr<int>()),
                                       reinterpret_cast<unsigned int*>(keep_index.data_ptr<int>()), fptr(src), dst.data_ptr<float>(),
                                       clamp_from, clamp_max, stream));
        return dst;
    };
    st.offset_denom = compact(st.offset_denom, k, -1, 0.f).view({-1, 1});
    st.offset_gradient_accum = compact(st.offset_gradient_accum, k, -1, 0.f).view({-1, 1});
    st.opacity_accum = compact(st.opacity_accum, 1, -1, 0.f);
    st.anchor_demon = compact(st.anchor_demon, 1, -1, 0.f);
    auto ps = params_of(st);
    for (int g = 0; g < 6; ++g) {
        std::vector<int64_t> shape = ps[g]->sizes().vec();
        int64_t rf = 1;
        for (size_t d = 1; d < shape.size(); ++d) rf *= shape[d];
        shape[0] = n_keep;
        if (st.exp_avg[g].defined()) st.exp_avg[g] = compact(st.exp_avg[g], static_cast<int>(rf), -1, 0.f).view(shape);
        if (st.exp_avg_sq[g].defined()) st.exp_avg_sq[g] = compact(st.exp_avg_sq[g], static_cast<int>(rf), -1, 0.f).view(shape);
        // prune_anchor clamps _scaling[:, 3:] to <= 0.05 on the way (:1529-1532)
        *ps[g] = (g == 4 ? compact(*ps[g], static_cast<int>(rf), 3, 0.05f) : compact(*ps[g], static_cast<int>(rf), -1, 0.f)).view(shape);
    }
    st.prune_report = {static_cast<int>(A), n_keep};
}

}  // namespace densify
