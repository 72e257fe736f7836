// See fused_mapper.h.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "fused_mapper.h"

#include <string>

#include "../../../include/segs_raster.h"
#include "loss_utils.h"

namespace {

void raise_if(int status) {
    if (status == SEGS_OK) return;
    TORCH_CHECK(false, "segs_raster: ", std::string(segs_last_error()), " (status ", status, ")");
}

const float* fptr(const torch::Tensor& t) { return (!t.defined() || t.numel() == 0) ? nullptr : t.data_ptr<float>(); }

void check_f32_cuda(const torch::Tensor& t, const char* what) {
    TORCH_CHECK(t.defined() && t.is_cuda() && t.scalar_type() == torch::kFloat32 && t.is_contiguous(), "FusedMapper: ", what,
                " must be a contiguous FP32 CUDA tensor (no CPU path)");
}

}  // namespace

FusedMapper::FusedMapper(std::vector<torch::Tensor> model, std::vector<torch::Tensor> weights, std::array<int, 5> cfg,
                         int image_height, int image_width, float tanfovx, float tanfovy, torch::Tensor bg,
                         double lambda_dssim, double scaling_reg_weight, std::vector<double> lrs, double eps, int lanes)
    : model_(std::move(model)), weights_(std::move(weights)), cfg_(cfg), H_(image_height), W_(image_width), tanx_(tanfovx),
      tany_(tanfovy), bg_(bg.to(torch::kFloat32).contiguous()), lambda_(lambda_dssim), reg_w_(scaling_reg_weight), eps_(eps),
      lrs_(std::move(lrs)), lanes_(lanes < 1 ? 1 : lanes)
{
    TORCH_CHECK(model_.size() == 5, "FusedMapper: model = {_anchor, _offset, _anchor_feat, _scaling, _rotation}");
    TORCH_CHECK(weights_.size() == 18, "FusedMapper: 18 weight tensors in segs_decode_params order (undefined = absent)");
    for (int k = 0; k < 5; ++k) check_f32_cuda(model_[k], "a model tensor");
    const c10::cuda::CUDAGuard guard(model_[0].device());
    params_ = {model_[0], model_[1], model_[2], model_[3]};
    for (auto& w : weights_)
        if (w.defined()) { check_f32_cuda(w, "an MLP weight"); params_.push_back(w); }
    TORCH_CHECK(lrs_.size() == params_.size(), "FusedMapper: ", params_.size(), " learning rates expected, got ", lrs_.size());
    int64_t total = 0;
    for (auto& p : params_) total += p.numel();
    auto opt = model_[0].options();
    grad_flat_ = torch::zeros({total}, opt);
    exp_avg_ = torch::zeros({total}, opt);
    exp_avg_sq_ = torch::zeros({total}, opt);
    loss_accum_ = torch::zeros({}, opt);
    int64_t off = 0;
    for (auto& p : params_) { grad_views_.push_back(grad_flat_.narrow(0, off, p.numel())); off += p.numel(); }
    for (int l = 0; l < lanes_; ++l) {
        segs_workspace* w = nullptr;
        raise_if(segs_workspace_create(&w));
        ws_.push_back(w);
        if (lanes_ > 1) streams_.push_back(c10::cuda::getStreamFromPool(false, model_[0].device().index()));
    }
}

FusedMapper::~FusedMapper() {
    for (auto* w : ws_) segs_workspace_destroy(w);
}

int64_t FusedMapper::workspace_bytes() const {
    int64_t n = 0;
    for (auto* w : ws_) n += static_cast<int64_t>(segs_workspace_bytes(w));
    return n;
}

torch::Tensor FusedMapper::render_views(const std::vector<KeyframeView>& views) {
    const c10::cuda::CUDAGuard guard(model_[0].device());
    torch::NoGradGuard no_grad;
    loss_accum_.zero_();
    if (views.empty()) return loss_accum_.clone();
    // per-step derived tensors (gaussian_model.cpp:186-189, 213; gaussian_renderer.cpp:168-197)
    torch::Tensor scaling = torch::exp(model_[3]);
    torch::Tensor fscales = scaling.slice(1, 0, 3).contiguous();
    torch::Tensor frot = torch::nn::functional::normalize(model_[4]).contiguous();

    segs_decode_params dp{};
    const float** wp = reinterpret_cast<const float**>(&dp);
    for (int k = 0; k < 18; ++k) wp[k] = fptr(weights_[k]);
    dp.appearance_dim = cfg_[0]; dp.use_feat_bank = cfg_[1]; dp.add_opacity_dist = cfg_[2]; dp.add_cov_dist = cfg_[3];
    dp.add_color_dist = cfg_[4];
    segs_decode_grads dg{};
    float** gp = reinterpret_cast<float**>(&dg);
    {
        size_t next = 4;
        for (int k = 0; k < 18; ++k) gp[k] = weights_[k].defined() ? grad_views_[next++].data_ptr<float>() : nullptr;
    }

    std::vector<segs_mapper_view_args> args(views.size());
    std::vector<segs_mapper_view_result> res(views.size());
    std::vector<torch::Tensor> keep;                      // contiguous copies stay alive until the call returns
    auto dense = [&](const torch::Tensor& t) { keep.push_back(t.to(torch::kFloat32).contiguous()); return keep.back().data_ptr<float>(); };
    for (size_t v = 0; v < views.size(); ++v) {
        segs_mapper_view_args a{};
        a.A = static_cast<int>(model_[0].size(0));
        a.anchor = model_[0].data_ptr<float>(); a.offset = model_[1].data_ptr<float>(); a.anchor_feat = model_[2].data_ptr<float>();
        a.scaling = scaling.data_ptr<float>(); a.scaling_is_log = 1;
        a.filter_scales = fscales.data_ptr<float>(); a.filter_rotations = frot.data_ptr<float>();
        a.params = &dp;
        a.width = W_; a.height = H_; a.tan_fovx = tanx_; a.tan_fovy = tany_;
        a.viewmatrix = dense(views[v].world_view_transform); a.projmatrix = dense(views[v].full_proj_transform);
        a.campos = dense(views[v].camera_center);
        a.pose = views[v].pose.data();
        a.background = bg_.data_ptr<float>();
        a.gt_image = dense(views[v].gt_image);
        a.row_mask = views[v].row_mask.defined() && views[v].row_mask.numel() ? dense(views[v].row_mask) : nullptr;
        a.lambda_dssim = static_cast<float>(lambda_); a.scaling_reg_weight = static_cast<float>(reg_w_);
        a.grad_anchor = grad_views_[0].data_ptr<float>(); a.grad_offset = grad_views_[1].data_ptr<float>();
        a.grad_anchor_feat = grad_views_[2].data_ptr<float>(); a.grad_scaling = grad_views_[3].data_ptr<float>();
        a.grad_params = &dg;
        a.loss_accum = loss_accum_.data_ptr<float>();
        a.lambda_frequency_high = static_cast<float>(freq_lambda_);
        a.use_multi_resolution = freq_multi_ ? 1 : 0;
        a.freq_scale_num = freq_scales_;
        a.gt_freq_mag = nullptr;                            // computed per view in the lane's workspace
        args[v] = a;
    }
    const int lanes = static_cast<int>(std::min<size_t>(lanes_, views.size()));
    void* main_stream = static_cast<void*>(at::cuda::getCurrentCUDAStream().stream());
    std::vector<void*> streams(lanes);
    for (int l = 0; l < lanes; ++l) streams[l] = lanes == 1 ? main_stream : static_cast<void*>(streams_[l].stream());
    raise_if(segs_mapper_views(static_cast<int>(views.size()), args.data(), res.data(), lanes, ws_.data(), streams.data(), main_stream));
    return loss_accum_.clone();
}

void FusedMapper::adam_step(double grad_scale) {
    ++step_;
    loss_utils::adam_step(params_, lrs_, grad_flat_, exp_avg_, exp_avg_sq_, step_, 0.9, 0.999, eps_, 0.0, grad_scale, true);
}
