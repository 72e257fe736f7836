"""TEST INFRASTRUCTURE — CPU restatement of the mapper's photometric loss and Adam step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the
product (segs_slam_b200/) never does.

Restates, in PyTorch FP32 on the CPU (same ATen ops in the same order as the reference):
  * loss_utils::l1_loss            /root/reference/include/loss_utils.h:29-32
  * loss_utils::gaussian           :50-65   (11 taps, sigma 1.5, normalised)
  * loss_utils::create_window      :67-76   (outer product, expanded to [C,1,11,11])
  * loss_utils::_ssim / ssim       :78-127  (5 grouped conv2d with zero padding 5, C1 = 0.01^2, C2 = 0.03^2)
  * loss_utils::psnr               :39-43
  * loss_utils::high_frequency_loss / low_freq_loss / multi_scale_loss   :129-237 (torch.fft, op for op)
  * the call site                  /root/reference/src/gaussian_mapper.cpp:908-925
      mask_rgb = (gt != 0).any(-1) ; loss = (1-l)*L1 + l*(1-ssim) + 0.01*scaling.prod(1).mean()
  * torch::optim::Adam::step       LibTorch torch/csrc/api/src/optim/adam.cpp (amsgrad off), the optimizer the
                                   reference builds at src/gaussian_model.cpp:620-872 (eps = 1e-15, :634)

Pinned: tests/golden/loss_*.npz, freq_*.npz and adam_*.npz were produced by the reference's own loss_utils.h and by
torch::optim::Adam (oracle/_ref/libloss_ref.so via tests/golden/make_loss_golden.py);
tests/test_loss_cpu.py holds this file to them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def gaussian(window_size: int, sigma: float) -> torch.Tensor:
    v = [math.exp(-float((x - window_size // 2) ** 2) / (2.0 * sigma * sigma)) for x in range(window_size)]
    g = torch.tensor(np.asarray(v, dtype=np.float32))
    return g / g.sum()


def create_window(window_size: int, channel: int) -> torch.Tensor:
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def l1_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return (a - b).abs().mean()


def psnr(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    mse = ((a - b) ** 2).mean()
    return 10.0 * torch.log10(1.0 / mse)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    channel = img1.size(-3)
    window = create_window(window_size, channel).to(img1)
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, window, padding=pad, groups=channel)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = conv(img1 * img1) - mu1_sq
    sigma2_sq = conv(img2 * img2) - mu2_sq
    sigma12 = conv(img1 * img2) - mu1_mu2
    C1, C2 = 0.01 * 0.01, 0.03 * 0.03
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def row_mask(gt: torch.Tensor) -> torch.Tensor:
    """mask_rgb of gaussian_mapper.cpp:911-912, squeezed to [C,H]."""
    return (gt != 0.0).any(-1).to(torch.float32)


def mapper_loss(image: np.ndarray, gt: np.ndarray, lambda_dssim: float, apply_mask: bool = False,
                scaling: np.ndarray | None = None):
    """-> dict(l1, ssim, loss, dL_dimage[, dL_dscaling]) as float32 numpy."""
    x = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).clone().requires_grad_(True)
    y = torch.from_numpy(np.ascontiguousarray(gt, dtype=np.float32)).clone()
    rendered, masked = x, x
    if apply_mask:
        m = row_mask(y).unsqueeze(-1)
        masked, rendered, y = masked * m, rendered * m, y * m
    Ll1 = l1_loss(rendered, y)
    ss = ssim(masked, y)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ss)
    sc = None
    if scaling is not None:
        sc = torch.from_numpy(np.ascontiguousarray(scaling, dtype=np.float32)).clone().requires_grad_(True)
        loss = loss + 0.01 * sc.prod(1).mean()
    loss.backward()
    out = dict(l1=np.float32(Ll1.item()), ssim=np.float32(ss.item()), loss=np.float32(loss.item()),
               dL_dimage=x.grad.numpy().copy())
    if sc is not None:
        out["dL_dscaling"] = sc.grad.numpy().copy()
    return out


def adam(param: np.ndarray, grads: np.ndarray, lr: float, beta1: float = 0.9, beta2: float = 0.999,
         eps: float = 1e-8, weight_decay: float = 0.0):
    """`len(grads)` Adam steps in FP32 numpy, adam.cpp's operation order.  -> (param, exp_avg, exp_avg_sq)."""
    f = np.float32
    p = param.astype(f).copy()
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for step, g in enumerate(grads.astype(f), start=1):
        if weight_decay != 0.0:
            g = g + f(weight_decay) * p
        m = m * f(beta1) + f(1.0 - beta1) * g
        v = v * f(beta2) + f(1.0 - beta2) * g * g
        bc1 = 1.0 - beta1 ** step
        bc2 = 1.0 - beta2 ** step
        denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
        p = p - f(lr / bc1) * (m / denom)
    return p, m, v


# ---- frequency-domain terms (loss_utils.h:129-237), op for op -----------------------------------------------------
# NOTE (reference behaviour, reproduced on purpose): the images are [C,H,W] but the masks are indexed with
# index_put_({Slice(crow-r, crow+r), Slice(ccol-r, ccol+r)}, v), i.e. on dims 0 and 1 — the CHANNEL and ROW dims.  With
# C = 3 the first slice is empty, so the high-pass mask stays all ones (the "high frequency" loss is the mean
# magnitude difference over the FULL spectrum) and the low-pass mask stays all zeros (low_freq_loss has zero gradient;
# its value is pi/(HWC) times the number of spectrum bins whose signed zeros differ in angle).  fftshift() without
# dims also rolls the channel dim; neither loss depends on the bin order.
def _filtered_fft(img: torch.Tensor, cutoff_ratio: float, high: bool) -> torch.Tensor:
    f = torch.fft.fftshift(torch.fft.fft2(img))
    H, W = img.shape[1], img.shape[2]
    crow, ccol = H // 2, W // 2
    mask = torch.ones_like(f) if high else torch.zeros_like(f)
    r = int(cutoff_ratio * min(H, W) / 2)
    mask[crow - r:crow + r, ccol - r:ccol + r] = 0 if high else 1
    return f * mask


def high_frequency_loss(img1: torch.Tensor, img2: torch.Tensor, cutoff_ratio: float = 0.4) -> torch.Tensor:
    a, b = _filtered_fft(img1, cutoff_ratio, True), _filtered_fft(img2, cutoff_ratio, True)
    return torch.mean(torch.abs(torch.abs(a) - torch.abs(b)))


def low_freq_loss(img1: torch.Tensor, img2: torch.Tensor, cutoff_ratio: float = 0.2) -> torch.Tensor:
    norm = float(img1.shape[0] * img1.shape[1] * img1.shape[2])
    a, b = _filtered_fft(img1, cutoff_ratio, False), _filtered_fft(img2, cutoff_ratio, False)
    la = torch.sum(torch.abs(torch.abs(a) - torch.abs(b))) / norm
    lp = torch.sum(torch.abs(torch.angle(a) - torch.angle(b))) / norm
    return la + lp


def multi_scale_loss(gen: torch.Tensor, target: torch.Tensor, scales) -> torch.Tensor:
    loss = torch.zeros((), device=gen.device)
    for s in scales:
        g = F.interpolate(gen.unsqueeze(0), scale_factor=(float(s), float(s)), mode="bilinear", align_corners=False,
                          recompute_scale_factor=True)
        t = F.interpolate(target.unsqueeze(0), scale_factor=(float(s), float(s)), mode="bilinear", align_corners=False,
                          recompute_scale_factor=True)
        loss = loss + s * high_frequency_loss(g.squeeze(0), t.squeeze(0))
    return loss
