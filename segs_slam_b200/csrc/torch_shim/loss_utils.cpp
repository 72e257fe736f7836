// See loss_utils.h.  torch is used for device memory, the current stream and autograd bookkeeping only.
#include "loss_utils.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <string>

#include "../../../include/segs_raster.h"

namespace {

void raise_if(int status) {
    if (status == SEGS_OK) return;
    TORCH_CHECK(false, "segs_raster: ", std::string(segs_last_error()), " (status ", status, ")");
}

void* current_stream() { return static_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

const float* fptr(const torch::Tensor& t) { return (!t.defined() || t.numel() == 0) ? nullptr : t.data_ptr<float>(); }

void check_images(const torch::Tensor& a, const torch::Tensor& b) {
    TORCH_CHECK(a.is_cuda() && b.is_cuda(), "segs_raster has no CPU path: tensors must live on a CUDA device");
    TORCH_CHECK(a.dim() == 3 && a.sizes() == b.sizes(), "loss: image and gt must both be [C,H,W]");
    TORCH_CHECK(a.scalar_type() == torch::kFloat32 && b.scalar_type() == torch::kFloat32, "loss: FP32 images expected");
}

// (l1, ssim, w_l1 * l1 + w_ssim * ssim + bias); gradients flow to `image` only
struct L1SSIMFunction : public torch::autograd::Function<L1SSIMFunction> {
    static torch::autograd::tensor_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor image, torch::Tensor gt,
                                                torch::Tensor row_mask, double w_l1, double w_ssim, double bias) {
        check_images(image, gt);
        const c10::cuda::CUDAGuard guard(image.device());
        image = image.contiguous();
        gt = gt.contiguous();
        if (row_mask.defined() && row_mask.numel()) row_mask = row_mask.to(torch::kFloat32).contiguous();
        else row_mask = torch::empty({0}, image.options());
        const int C = image.size(0), H = image.size(1), W = image.size(2);
        TORCH_CHECK(row_mask.numel() == 0 || row_mask.numel() == int64_t(C) * H, "loss: row_mask must be [C,H]");
        auto state = torch::empty({static_cast<int64_t>(segs_loss_state_bytes(C, H, W))}, image.options().dtype(torch::kByte));
        auto out = torch::empty({3}, image.options());
        raise_if(segs_loss_l1_ssim_forward(C, H, W, image.data_ptr<float>(), gt.data_ptr<float>(), fptr(row_mask),
                                           static_cast<float>(w_l1), static_cast<float>(w_ssim), static_cast<float>(bias),
                                           out.data_ptr<float>(), reinterpret_cast<char*>(state.data_ptr()), current_stream()));
        ctx->save_for_backward({image, gt, row_mask, state});
        ctx->saved_data["w_l1"] = w_l1;
        ctx->saved_data["w_ssim"] = w_ssim;
        return {out[0], out[1], out[2]};
    }

    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::tensor_list g) {
        auto sv = ctx->get_saved_variables();
        const torch::Tensor &image = sv[0], &gt = sv[1], &row_mask = sv[2], &state = sv[3];
        const c10::cuda::CUDAGuard guard(image.device());
        const int C = image.size(0), H = image.size(1), W = image.size(2);
        const double w[3][2] = {{1.0, 0.0}, {0.0, 1.0}, {ctx->saved_data["w_l1"].toDouble(), ctx->saved_data["w_ssim"].toDouble()}};
        torch::Tensor grad;
        for (int k = 0; k < 3; ++k) {
            if (!g[k].defined()) continue;
            auto up = g[k].to(torch::kFloat32).contiguous();           // the upstream gradient stays on the device
            auto d = torch::empty_like(image);
            raise_if(segs_loss_l1_ssim_backward(C, H, W, image.data_ptr<float>(), gt.data_ptr<float>(), fptr(row_mask),
                                                static_cast<float>(w[k][0]), static_cast<float>(w[k][1]), up.data_ptr<float>(),
                                                reinterpret_cast<char*>(state.data_ptr()), d.data_ptr<float>(), current_stream()));
            grad = grad.defined() ? grad + d : d;
        }
        return {grad, torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor()};
    }
};

}  // namespace

namespace loss_utils {

torch::Tensor l1_loss(torch::Tensor& network_output, torch::Tensor& gt) {
    return L1SSIMFunction::apply(network_output, gt, torch::empty({0}, network_output.options()), 1.0, 0.0, 0.0)[0];
}

torch::Tensor psnr(torch::Tensor& img1, torch::Tensor& img2) {
    check_images(img1, img2);
    auto mse = torch::pow(img1 - img2, 2).mean();                       // loss_utils.h:39-43 (evaluation only)
    return 10.0f * torch::log10(1.0f / mse);
}

torch::Tensor ssim(torch::Tensor& img1, torch::Tensor& img2, torch::DeviceType device_type, int window_size, bool size_average) {
    TORCH_CHECK(device_type == torch::kCUDA, "segs_raster has no CPU path");
    TORCH_CHECK(window_size == 11 && size_average, "ssim: only window_size = 11, size_average = true is built");
    return L1SSIMFunction::apply(img1, img2, torch::empty({0}, img1.options()), 0.0, 1.0, 0.0)[1];
}

torch::Tensor l1_ssim(const torch::Tensor& image, const torch::Tensor& gt, double lambda_dssim, const torch::Tensor& row_mask) {
    // (an undefined tensor cannot travel through Function::apply)
    const torch::Tensor mask = row_mask.defined() ? row_mask : torch::empty({0}, image.options());
    return L1SSIMFunction::apply(image, gt, mask, 1.0 - lambda_dssim, -lambda_dssim, lambda_dssim)[2];
}

void adam_step(const std::vector<torch::Tensor>& params, const std::vector<double>& lrs, torch::Tensor& grad_flat,
               torch::Tensor& exp_avg_flat, torch::Tensor& exp_avg_sq_flat, int64_t step, double beta1, double beta2,
               double eps, double weight_decay, double grad_scale, bool zero_grad) {
    TORCH_CHECK(params.size() == lrs.size(), "adam_step: one learning rate per tensor");
    TORCH_CHECK(grad_flat.is_cuda() && grad_flat.is_contiguous() && grad_flat.scalar_type() == torch::kFloat32,
                "adam_step: flat FP32 CUDA buffers expected (no CPU path)");
    const c10::cuda::CUDAGuard guard(grad_flat.device());
    torch::NoGradGuard no_grad;
    std::vector<segs_adam_tensor> t(params.size());
    unsigned long long off = 0;
    for (size_t k = 0; k < params.size(); ++k) {
        TORCH_CHECK(params[k].is_cuda() && params[k].is_contiguous() && params[k].scalar_type() == torch::kFloat32,
                    "adam_step: parameters must be contiguous FP32 CUDA tensors");
        t[k] = segs_adam_tensor{params[k].data_ptr<float>(), off, static_cast<unsigned long long>(params[k].numel()),
                                static_cast<float>(lrs[k]), static_cast<float>(beta1), static_cast<float>(beta2),
                                static_cast<float>(eps), static_cast<float>(weight_decay), static_cast<long long>(step)};
        off += static_cast<unsigned long long>(params[k].numel());
    }
    TORCH_CHECK(static_cast<int64_t>(off) == grad_flat.numel() && grad_flat.numel() == exp_avg_flat.numel() &&
                    grad_flat.numel() == exp_avg_sq_flat.numel(), "adam_step: flat buffers do not match the parameters");
    raise_if(segs_adam_step(static_cast<int>(t.size()), t.data(), grad_flat.data_ptr<float>(), exp_avg_flat.data_ptr<float>(),
                            exp_avg_sq_flat.data_ptr<float>(), static_cast<float>(grad_scale), zero_grad ? 1 : 0,
                            current_stream()));
}

}  // namespace loss_utils
