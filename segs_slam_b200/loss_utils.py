"""Photometric loss of the mapper, mirroring the reference's `loss_utils` namespace
(/root/reference/include/loss_utils.h:29-127: l1_loss, psnr, ssim) and the way the mapper combines it
(/root/reference/src/gaussian_mapper.cpp:908-925) on top of the fused CUDA kernels of
segs_slam_b200/csrc/loss.cu (C ABI: segs_loss_l1_ssim_forward / _backward, segs_scaling_reg).

torch is used for device memory, the current stream and autograd bookkeeping only; there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from .rasterize_points import _ptr, _stream


def _check(image: torch.Tensor, gt: torch.Tensor):
    if not image.is_cuda or not gt.is_cuda:
        raise RuntimeError("segs_slam_b200 has no CPU path: tensors must live on a CUDA device")
    if image.dim() != 3 or image.shape != gt.shape:
        raise RuntimeError(f"loss: image {tuple(image.shape)} and gt {tuple(gt.shape)} must both be [C,H,W]")
    if image.dtype != torch.float32 or gt.dtype != torch.float32:
        raise RuntimeError("loss: FP32 images expected")


class _L1SSIMFunction(torch.autograd.Function):
    """(l1, ssim, loss) = f(image); loss = w_l1*l1 + w_ssim*ssim + bias.  Gradients flow to `image` only."""

    @staticmethod
    def forward(ctx, image, gt, row_mask, w_l1, w_ssim, bias):
        _check(image, gt)
        lib = _lib.load()
        C_, H, W = image.shape
        img_c, gt_c = image.contiguous(), gt.contiguous()
        mask_c = row_mask.contiguous() if row_mask is not None else None
        if mask_c is not None and tuple(mask_c.shape) != (C_, H):
            raise RuntimeError(f"loss: row_mask must be [C,H] = {(C_, H)}, got {tuple(mask_c.shape)}")
        state = torch.empty((int(lib.segs_loss_state_bytes(C_, H, W)),), dtype=torch.uint8, device=image.device)
        out = torch.empty((3,), dtype=torch.float32, device=image.device)
        with torch.cuda.device(image.device):
            _lib.check(lib.segs_loss_l1_ssim_forward(C_, H, W, _ptr(img_c), _ptr(gt_c), _ptr(mask_c), w_l1, w_ssim,
                                                     bias, _ptr(out), state.data_ptr(), _stream()))
        ctx.save_for_backward(img_c, gt_c, mask_c if mask_c is not None else torch.empty(0, device=image.device),
                              state)
        ctx.w = (float(w_l1), float(w_ssim))
        ctx.set_materialize_grads(False)          # unused outputs arrive as None instead of zero tensors: one launch per used one
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_l1, g_ssim, g_loss):
        lib = _lib.load()
        img_c, gt_c, mask_c, state = ctx.saved_tensors
        C_, H, W = img_c.shape
        w_l1, w_ssim = ctx.w
        grad = None
        # each output is the same kernel with different weights; the usual case is one call (g_loss only)
        for g, (a, b) in ((g_loss, (w_l1, w_ssim)), (g_l1, (1.0, 0.0)), (g_ssim, (0.0, 1.0))):
            if g is None:
                continue
            d = torch.empty_like(img_c)
            gc = g.to(torch.float32).contiguous()
            with torch.cuda.device(img_c.device):
                _lib.check(lib.segs_loss_l1_ssim_backward(C_, H, W, _ptr(img_c), _ptr(gt_c), _ptr(mask_c), a, b,
                                                          _ptr(gc), state.data_ptr(), _ptr(d), _stream()))
            grad = d if grad is None else grad + d
        if grad is None:
            grad = torch.zeros_like(img_c)
        return grad, None, None, None, None, None


def l1_ssim_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float, row_mask: torch.Tensor | None = None):
    """(1 - lambda) * l1_loss + lambda * (1 - ssim)  (gaussian_mapper.cpp:917-921), one fused kernel each way.
    row_mask [C,H]: the `mask_rgb` of :911-915 (see `mask_rgb`).  -> (loss, Ll1, ssim) scalars on the device."""
    l1, ss, loss = _L1SSIMFunction.apply(image, gt, row_mask, 1.0 - lambda_dssim, -lambda_dssim, lambda_dssim)
    return loss, l1, ss


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """loss_utils::l1_loss (loss_utils.h:29-32)."""
    return _L1SSIMFunction.apply(network_output, gt, None, 1.0, 0.0, 0.0)[0]


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """loss_utils::ssim (loss_utils.h:112-127) with the only configuration the reference uses."""
    if window_size != 11 or not size_average:
        raise RuntimeError("ssim: only window_size = 11, size_average = true (the reference's call sites) is built")
    return _L1SSIMFunction.apply(img1, img2, None, 0.0, 1.0, 0.0)[1]


def psnr(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """loss_utils::psnr (loss_utils.h:39-43); evaluation only, plain device arithmetic."""
    _check(img1, img2)
    mse = torch.pow(img1 - img2, 2).mean()
    return 10.0 * torch.log10(1.0 / mse)


def mask_rgb(gt_image: torch.Tensor) -> torch.Tensor:
    """(gt != 0).any(-1) as float, [C,H] (gaussian_mapper.cpp:911-912); constant per keyframe."""
    return (gt_image != 0.0).any(-1).to(torch.float32)


class _ScalingReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scaling, weight):
        if not scaling.is_cuda:
            raise RuntimeError("segs_slam_b200 has no CPU path: tensors must live on a CUDA device")
        lib = _lib.load()
        s = scaling.contiguous()
        out = torch.zeros((), dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            _lib.check(lib.segs_scaling_reg(s.size(0), _ptr(s), weight, None, None, _ptr(out), _stream()))
        ctx.save_for_backward(s)
        ctx.weight = float(weight)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (s,) = ctx.saved_tensors
        d = torch.zeros_like(s)
        gc = g.to(torch.float32).contiguous()
        with torch.cuda.device(s.device):
            _lib.check(lib.segs_scaling_reg(s.size(0), _ptr(s), ctx.weight, _ptr(gc), _ptr(d), None, _stream()))
        return d, None


def scaling_reg(scaling: torch.Tensor, weight: float = 0.01) -> torch.Tensor:
    """weight * scaling.prod(1).mean() (gaussian_mapper.cpp:919-921)."""
    return _ScalingReg.apply(scaling, weight)


# ---- frequency-domain terms (loss_utils.h:125-237) ---------------------------------------------------------------
# On the C ABI (csrc/freq.cu: segs_freq_plan_*, segs_freq_target, segs_freq_loss): fused resampling / magnitude-loss /
# adjoint kernels around cuFFT plans.  The reference's indexing quirk is part of the contract: its [C,H,W] frequency
# masks are sliced on dims 0 and 1 (channel, row) with [H/2 - r, H/2 + r), which is EMPTY for C = 3 unless the image is
# only a few pixels high — the "high-pass" mask stays all ones (a full-spectrum magnitude loss) and the "low-pass" mask
# all zeros (low_freq_loss is identically zero, with zero gradient).  Sizes where the slice would not be empty are
# refused rather than silently computed differently.
_freq_plans = {}


def _mask_is_empty(C: int, H: int, W: int, cutoff_ratio: float) -> bool:
    r = int(cutoff_ratio * min(H, W) / 2)
    lo = H // 2 - r
    return r <= 0 or lo >= C                      # Slice(crow - r, crow + r) on the channel dimension (size C)


def _freq_plan(device, C: int, H: int, W: int, scales):
    import ctypes as C_
    key = (device.index if device.index is not None else torch.cuda.current_device(), C, H, W, tuple(float(x) for x in scales))
    plan = _freq_plans.get(key)
    if plan is None:
        lib = _lib.load()
        arr = (C_.c_float * len(scales))(*[float(x) for x in scales])
        plan = C_.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.segs_freq_plan_create(C, H, W, len(scales), arr, C_.byref(plan)))
        _freq_plans[key] = plan
    return plan


def freq_target(gt: torch.Tensor, scales=(1.0,), row_mask: torch.Tensor | None = None) -> torch.Tensor:
    """|fft2(D_s (gt * row_mask))| of every scale, flattened: the per-keyframe half of the frequency terms (pass it to
    FusedMapper / segs_mapper_view as `gt_freq_mag` to avoid recomputing it every view)."""
    _check(gt, gt)
    lib = _lib.load()
    C_, H, W = gt.shape
    plan = _freq_plan(gt.device, C_, H, W, scales)
    mag = torch.empty((int(lib.segs_freq_mag_floats(plan)),), dtype=torch.float32, device=gt.device)
    g = gt.contiguous()
    m = row_mask.contiguous() if row_mask is not None else None
    with torch.cuda.device(gt.device):
        _lib.check(lib.segs_freq_target(plan, _ptr(g), _ptr(m), _ptr(mag), _stream()))
    return mag


class _FreqLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, scales, weight):
        _check(image, gt)
        lib = _lib.load()
        C_, H, W = image.shape
        plan = _freq_plan(image.device, C_, H, W, scales)
        img_c = image.contiguous()
        mag = freq_target(gt, scales)
        out = torch.zeros((), dtype=torch.float32, device=image.device)
        with torch.cuda.device(image.device):
            _lib.check(lib.segs_freq_loss(plan, _ptr(img_c), None, _ptr(mag), float(weight), None, _ptr(out), None, _stream()))
        ctx.save_for_backward(img_c, mag)
        ctx.plan, ctx.weight = plan, float(weight)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        img_c, mag = ctx.saved_tensors
        d = torch.zeros_like(img_c)
        gc = g.to(torch.float32).contiguous()
        with torch.cuda.device(img_c.device):
            _lib.check(lib.segs_freq_loss(ctx.plan, _ptr(img_c), None, _ptr(mag), ctx.weight, _ptr(gc), None, _ptr(d), _stream()))
        return d, None, None, None


def high_frequency_loss(img1: torch.Tensor, img2: torch.Tensor, cutoff_ratio: float = 0.4) -> torch.Tensor:
    """loss_utils::high_frequency_loss (loss_utils.h:150-169): mean | |fft2(img1)| - |fft2(img2)| | (differentiable in img1)."""
    C_, H, W = img1.shape
    if not _mask_is_empty(C_, H, W, cutoff_ratio):
        raise NotImplementedError("high_frequency_loss: for this size the reference's mask slice over the channel dimension is not "
                                  f"empty (C={C_}, H={H}, r={int(cutoff_ratio * min(H, W) / 2)}); only the empty-mask case is built")
    return _FreqLoss.apply(img1, img2, (1.0,), 1.0)


def low_freq_loss(img1: torch.Tensor, img2: torch.Tensor, cutoff_ratio: float = 0.2) -> torch.Tensor:
    """loss_utils::low_freq_loss (loss_utils.h:190-207): with the reference's (empty) mask slice the low-pass mask is all
    zeros, so both of its terms are sums of |0 - 0|: identically zero, zero gradient."""
    C_, H, W = img1.shape
    if not _mask_is_empty(C_, H, W, cutoff_ratio):
        raise NotImplementedError("low_freq_loss: non-empty mask slice (tiny image); only the empty-mask case is built")
    _check(img1, img2)
    return (img1 * 0.0).sum()


def multi_scale_loss(gen_img: torch.Tensor, target_img: torch.Tensor, scales) -> torch.Tensor:
    """loss_utils::multi_scale_loss (loss_utils.h:210-237): sum_s s * high_frequency_loss(D_s gen, D_s target), D_s = bilinear
    interpolate(scale_factor = s, align_corners = False, recompute_scale_factor = True); one fused call for all scales."""
    scales = tuple(float(x) for x in scales)
    C_, H, W = gen_img.shape
    for sc in scales:
        if not _mask_is_empty(C_, int(H * sc), int(W * sc), 0.4):
            raise NotImplementedError("multi_scale_loss: non-empty mask slice at one of the scales (tiny image)")
    return _FreqLoss.apply(gen_img, target_img, scales, 1.0)
