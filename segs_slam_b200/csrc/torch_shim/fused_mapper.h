// C++/LibTorch host class of the keyframe-batched mapping step — the C++ twin of segs_slam_b200/mapper.py:FusedMapper,
// i.e. what a SEGS-SLAM maintainer drops into GaussianMapper in place of the body of trainForOneIteration
// (/root/reference/src/gaussian_mapper.cpp:823-1032) when optimising a batch of keyframes per step:
//   render_views()  prefilter -> decode -> render -> L1+SSIM(+scaling regulariser) -> backward for every view, issued
//                   through segs_mapper_views on `lanes` concurrent lanes; gradients land in ONE flat FP32 bucket
//   grad_flat()     the bucket: what the caller all-reduces over its ranks (ncclAllReduce / c10d, sum)
//   adam_step()     one fused Adam launch over every trainable tensor, 1/batch scale and bucket clear included
// torch is used for device memory and streams only.
#pragma once

#include <c10/cuda/CUDAStream.h>
#include <torch/torch.h>

#include <array>
#include <vector>

struct segs_workspace;

struct KeyframeView {
    torch::Tensor world_view_transform;   // [4,4] as GaussianKeyframe::world_view_transform_ (contiguous)
    torch::Tensor full_proj_transform;    // [4,4]
    torch::Tensor camera_center;          // [3]
    std::array<float, 7> pose;            // {t.xyz, R_quaternion wxyz}
    torch::Tensor gt_image;               // [3,H,W]
    torch::Tensor row_mask;               // [3,H] (mask_rgb) or undefined
};

class FusedMapper {
public:
    // model = {_anchor [A,3], _offset [A,10,3], _anchor_feat [A,32], _scaling [A,6] (log), _rotation [A,4]};
    // weights = the 18 tensors of segs_decode_params order (undefined = absent); cfg = {appearance_dim, use_feat_bank,
    // add_opacity_dist, add_cov_dist, add_color_dist}; lrs: one per trainable tensor in bucket order
    // (_anchor, _offset, _anchor_feat, _scaling, then the defined weights)
    FusedMapper(std::vector<torch::Tensor> model, std::vector<torch::Tensor> weights, std::array<int, 5> cfg,
                int image_height, int image_width, float tanfovx, float tanfovy, torch::Tensor bg, double lambda_dssim,
                double scaling_reg_weight, std::vector<double> lrs, double eps, int lanes);
    ~FusedMapper();
    FusedMapper(const FusedMapper&) = delete;
    FusedMapper& operator=(const FusedMapper&) = delete;

    // accumulates the views' gradients into the bucket; -> loss summed over the views (device scalar)
    torch::Tensor render_views(const std::vector<KeyframeView>& views);
    torch::Tensor grad_flat() { return grad_flat_; }
    std::vector<torch::Tensor> params() { return params_; }
    void adam_step(double grad_scale);
    // frequency regularisation of the loss (src/gaussian_mapper.cpp:930-945): 0 = off; Replica yamls: (0.01, true, 3)
    void set_frequency(double lambda_frequency_high, bool use_multi_resolution, int scale_num) {
        freq_lambda_ = lambda_frequency_high; freq_multi_ = use_multi_resolution; freq_scales_ = scale_num;
    }
    int64_t workspace_bytes() const;

private:
    std::vector<torch::Tensor> model_, weights_, params_;
    std::array<int, 5> cfg_;
    int H_, W_;
    float tanx_, tany_;
    torch::Tensor bg_, grad_flat_, exp_avg_, exp_avg_sq_, loss_accum_;
    std::vector<torch::Tensor> grad_views_;      // slices of grad_flat_, one per entry of params_
    double lambda_, reg_w_, eps_;
    std::vector<double> lrs_;
    int lanes_;
    double freq_lambda_ = 0.0;
    bool freq_multi_ = true;
    int freq_scales_ = 3;
    int64_t step_ = 0;
    std::vector<segs_workspace*> ws_;
    std::vector<c10::cuda::CUDAStream> streams_;
};
