// LibTorch drop-in for the `loss_utils` namespace of the reference (/root/reference/include/loss_utils.h:26-127):
// same function names, argument lists and return values — l1_loss, psnr, ssim — computed by the fused sm_100a
// kernels behind include/segs_raster.h (segs_loss_l1_ssim_forward / _backward) instead of 5 grouped conv2d and ~25
// elementwise ATen kernels each way.  `l1_ssim` is the fused combination the mapper's loss line wants
// (src/gaussian_mapper.cpp:917-921) and `adam_step` the fused optimizer step (src/gaussian_mapper.cpp:1003-1006).
// A SEGS-SLAM build puts this directory before include/ on the include path; gaussian_mapper.cpp compiles unchanged.
#pragma once

#include <torch/torch.h>

#include <vector>

namespace loss_utils {

torch::Tensor l1_loss(torch::Tensor& network_output, torch::Tensor& gt);

torch::Tensor psnr(torch::Tensor& img1, torch::Tensor& img2);

// window_size / size_average: only (11, true) — every call site of the reference — is built; anything else throws
torch::Tensor ssim(torch::Tensor& img1, torch::Tensor& img2, torch::DeviceType device_type = torch::kCUDA,
                   int window_size = 11, bool size_average = true);

// (1 - lambda) * l1_loss(image * mask, gt * mask) + lambda * (1 - ssim(image * mask, gt * mask)) in one kernel each
// way; row_mask = mask_rgb.squeeze(-1) [C,H] of gaussian_mapper.cpp:911-912, or an empty tensor
torch::Tensor l1_ssim(const torch::Tensor& image, const torch::Tensor& gt, double lambda_dssim,
                      const torch::Tensor& row_mask = torch::Tensor());

// One fused Adam step over `params` whose gradients / moments live in flat FP32 tensors laid out in the order of
// `params` (one learning rate per tensor = one parameter group of gaussian_model.cpp:620-872); grad_flat is scaled by
// grad_scale on the fly and cleared when zero_grad is set.
void adam_step(const std::vector<torch::Tensor>& params, const std::vector<double>& lrs, torch::Tensor& grad_flat,
               torch::Tensor& exp_avg_flat, torch::Tensor& exp_avg_sq_flat, int64_t step, double beta1 = 0.9,
               double beta2 = 0.999, double eps = 1e-15, double weight_decay = 0.0, double grad_scale = 1.0,
               bool zero_grad = true);

}  // namespace loss_utils
