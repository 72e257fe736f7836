"""GPU parity of the fused loss (csrc/loss.cu) and fused Adam (csrc/optim.cu), through the C ABI.

Bars (floating point): loss scalars 1e-5 relative; dL_dimage 1e-4 relative to the gradient's scale (the
reference sums 121 taps in cuDNN/ATen order, the kernel 11 + 11 separably); Adam updates 1e-5 relative."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import loss_oracle  # noqa: E402
from segs_slam_b200 import loss_utils, mapper, optim  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
LOSS = sorted(glob.glob(os.path.join(GOLD, "loss_*.npz")))
ADAM = sorted(glob.glob(os.path.join(GOLD, "adam_*.npz")))


def _grad_close(mine, ref, rel=1e-4):
    scale = float(np.abs(ref).max()) + 1e-30
    np.testing.assert_allclose(mine, ref, rtol=rel, atol=rel * scale)


def _run(image, gt, lam, apply_mask, scaling, dev):
    x = torch.from_numpy(image).to(dev).requires_grad_(True)
    y = torch.from_numpy(gt).to(dev)
    m = loss_utils.mask_rgb(y) if apply_mask else None
    loss, l1, ss = loss_utils.l1_ssim_loss(x, y, lam, m)
    sc = None
    if scaling is not None:
        sc = torch.from_numpy(scaling).to(dev).requires_grad_(True)
        loss = loss + loss_utils.scaling_reg(sc, 0.01)
    loss.backward()
    return dict(l1=l1.item(), ssim=ss.item(), loss=loss.item(), dL_dimage=x.grad.cpu().numpy(),
                dL_dscaling=None if sc is None else sc.grad.cpu().numpy())


@pytest.mark.parametrize("path", LOSS, ids=[os.path.basename(p) for p in LOSS])
def test_loss_matches_reference_golden(path, device):
    g = np.load(path)
    o = _run(g["image"], g["gt"], float(g["lambda_dssim"]), bool(g["apply_mask"]),
             g["scaling"] if "scaling" in g.files else None, device)
    for k in ("l1", "ssim", "loss"):
        np.testing.assert_allclose(o[k], float(g[k]), rtol=1e-5)
    _grad_close(o["dL_dimage"], g["dL_dimage"])
    if "scaling" in g.files:
        np.testing.assert_allclose(o["dL_dscaling"], g["dL_dscaling"], rtol=1e-5)
    ps = loss_utils.psnr(torch.from_numpy(g["image"]).to(device), torch.from_numpy(g["gt"]).to(device))
    np.testing.assert_allclose(float(ps), float(g["psnr"]), rtol=1e-5)


def test_loss_full_size_against_oracle(device):
    """1200x680 (the size of BASELINE configs 2-4): CUDA vs the CPU oracle on the same seeded images."""
    rng = np.random.default_rng(5)
    H, W = 680, 1200
    gt = rng.uniform(0, 1, (3, H, W)).astype(np.float32)
    gt[:, 100:103, :] = 0.0
    img = np.clip(gt + rng.normal(0, 0.2, gt.shape), 0, 1).astype(np.float32)
    o = _run(img, gt, 0.2, True, None, device)
    r = loss_oracle.mapper_loss(img, gt, 0.2, True)
    for k in ("l1", "ssim", "loss"):
        np.testing.assert_allclose(o[k], float(r[k]), rtol=1e-5)
    _grad_close(o["dL_dimage"], r["dL_dimage"])
    assert np.all(o["dL_dimage"][:, 100:103, :] == 0.0)     # masked rows receive no gradient


def test_loss_properties(device):
    """Size-independent properties: ssim(x, x) = 1 with zero gradient; l1 is symmetric; determinism."""
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.rand(3, 97, 131, generator=g).to(device)
    y = torch.rand(3, 97, 131, generator=g).to(device)
    xs = x.clone().requires_grad_(True)
    s = loss_utils.ssim(xs, x)
    s.backward()
    assert abs(s.item() - 1.0) < 1e-6
    assert float(xs.grad.abs().max()) < 1e-7
    assert loss_utils.l1_loss(x, y).item() == loss_utils.l1_loss(y, x).item()
    a = loss_utils.l1_ssim_loss(x, y, 0.2)[0].item()
    assert all(loss_utils.l1_ssim_loss(x, y, 0.2)[0].item() == a for _ in range(3))
    # individual pieces agree with the fused combination
    l1, ss = loss_utils.l1_loss(x, y).item(), loss_utils.ssim(x, y).item()
    np.testing.assert_allclose(a, 0.8 * l1 + 0.2 * (1.0 - ss), rtol=1e-6)


def test_loss_upstream_gradient_scales(device):
    g = torch.Generator(device="cpu").manual_seed(4)
    x = torch.rand(3, 40, 50, generator=g).to(device).requires_grad_(True)
    y = torch.rand(3, 40, 50, generator=g).to(device)
    (loss_utils.l1_ssim_loss(x, y, 0.2)[0] * 3.0).backward()
    g3 = x.grad.clone()
    x.grad = None
    loss_utils.l1_ssim_loss(x, y, 0.2)[0].backward()
    torch.testing.assert_close(g3, 3.0 * x.grad, rtol=1e-6, atol=0)


def test_loss_rejects_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU path"):
        loss_utils.l1_loss(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8))


def _bucket_for(chunks, dev):
    params = [torch.from_numpy(c.copy()).to(dev) for c in chunks]
    return params, mapper.GradBucket(params)


@pytest.mark.parametrize("path", ADAM, ids=[os.path.basename(p) for p in ADAM])
def test_adam_matches_torch_optim_adam_golden(path, device):
    g = np.load(path)
    n = g["param0"].size
    cuts = [0, n // 3, n // 3 + 5, n]                          # three tensors, one of them tiny
    params, bucket = _bucket_for([g["param0"][a:b] for a, b in zip(cuts[:-1], cuts[1:])], device)
    opt = optim.FusedAdam(bucket, float(g["lr"]), (float(g["beta1"]), float(g["beta2"])), float(g["eps"]),
                          float(g["weight_decay"]))
    for grad in g["grads"]:
        bucket.flat.copy_(torch.from_numpy(grad * 4.0).to(device))
        opt.step(grad_scale=0.25, zero_grad=True)              # power-of-two scale: exact
        assert float(bucket.flat.abs().max()) == 0.0
    mine = torch.cat([p.flatten() for p in params]).cpu().numpy()
    d_ref, d_mine = g["param"] - g["param0"], mine - g["param0"]
    # the deltas are differences of FP32 parameters: they carry the parameters' own rounding (1-2 ulp of |p|)
    ulp = float(np.spacing(np.float32(np.abs(g["param0"]).max())))
    np.testing.assert_allclose(d_mine, d_ref, rtol=2e-5, atol=2.0 * ulp)
    if float(g["weight_decay"]) == 0.0:
        assert np.array_equal(mine[:7], g["param0"][:7])


def test_adam_per_tensor_lr_against_torch(device):
    """71 floats per anchor + MLP-sized tensors, per-tensor learning rates, 3 steps, vs torch.optim.Adam groups."""
    gen = torch.Generator(device="cpu").manual_seed(9)
    shapes = [(5000, 3), (5000, 10, 3), (5000, 32), (5000, 6), (32, 35), (32,), (10, 32), (10,)]
    lrs = [0.0, 0.01, 0.0075, 0.007, 0.002, 0.002, 0.004, 0.004]
    ref_params = [torch.randn(s, generator=gen).to(device).requires_grad_(True) for s in shapes]
    my_params = [p.detach().clone() for p in ref_params]
    ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref_params, lrs)], eps=1e-15)
    bucket = mapper.GradBucket(my_params)
    opt = optim.FusedAdam(bucket, lrs, eps=1e-15)
    for _ in range(3):
        grads = [torch.randn(s, generator=gen).to(device) for s in shapes]
        for p, gr, v in zip(ref_params, grads, bucket.views):
            p.grad = gr.clone()
            v.copy_(gr)
        ref.step()
        opt.step()
    for a, b, p0 in zip(my_params, ref_params, shapes):
        torch.testing.assert_close(a, b.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(my_params[0], ref_params[0].detach())    # lr = 0: bit-identical (no update)


FREQ = sorted(glob.glob(os.path.join(GOLD, "freq_*.npz")))


@pytest.mark.parametrize("path", FREQ, ids=[os.path.basename(p) for p in FREQ])
def test_frequency_losses_match_reference_golden(path, device):
    """loss_utils::high_frequency_loss / multi_scale_loss / low_freq_loss on the fused kernels of csrc/freq.cu (cuFFT plans
    inside).  FFT sums of ~2000 FP32 terms: value 1e-4, gradient 1e-3 of its scale.  low_freq_loss has zero gradient by
    construction (the reference's low-pass mask is empty); the reference's VALUE only counts sign-of-zero flips of
    angle(+-0), so the product returns 0 and only the zero gradient is held."""
    g = np.load(path)
    x = torch.from_numpy(g["image"]).to(device).requires_grad_(True)
    y = torch.from_numpy(g["gt"]).to(device)
    hi = loss_utils.high_frequency_loss(x, y)
    hi.backward()
    np.testing.assert_allclose(hi.item(), float(g["high"]), rtol=1e-4)
    ref = g["d_high"]
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref, rtol=1e-3, atol=1e-3 * float(np.abs(ref).max()))
    x.grad = None
    lo = loss_utils.low_freq_loss(x, y)
    lo.backward()
    assert np.isfinite(lo.item()) and float(x.grad.abs().max()) == 0.0
    x.grad = None
    ms = loss_utils.multi_scale_loss(x, y, [1.0, 0.5])
    r = loss_oracle.multi_scale_loss(torch.from_numpy(g["image"]), torch.from_numpy(g["gt"]), [1.0, 0.5])
    np.testing.assert_allclose(ms.item(), float(r), rtol=1e-4)


@pytest.mark.parametrize("shape,scales", [((3, 40, 50), [1.0, 0.5, 0.25]), ((3, 33, 47), [1.0, 0.5]), ((3, 680, 1200), [1.0, 0.5, 0.25]),
                                          ((3, 120, 208), [1.0]), ((1, 64, 96), [1.0, 0.5, 0.25])])
def test_frequency_losses_match_the_compiled_reference(shape, scales, device):
    """multi_scale_loss / high_frequency_loss against the reference's OWN loss_utils.h running on the GPU (compiled
    unmodified into oracle/_ref/_model_ref.so), value and gradient, up to the Replica image size."""
    import model_ref
    if not model_ref.available():
        pytest.skip("oracle/_ref/_model_ref.so not built")
    mr = model_ref.load()
    gen = torch.Generator(device="cpu").manual_seed(shape[1])
    x0 = torch.rand(shape, generator=gen).to(device)
    y = (x0.cpu() + 0.1 * torch.randn(shape, generator=gen)).clamp(0, 1).to(device)
    for fn_mine, fn_ref in ((lambda a: loss_utils.multi_scale_loss(a, y, scales), lambda a: mr.multi_scale_loss(a, y, scales)),
                            (lambda a: loss_utils.high_frequency_loss(a, y), lambda a: mr.high_frequency_loss(a, y))):
        x = x0.clone().requires_grad_(True)
        v = fn_mine(x)
        v.backward()
        v_ref, g_ref = fn_ref(x0)
        np.testing.assert_allclose(v.item(), float(v_ref), rtol=1e-4)
        scale = float(g_ref.abs().max())
        assert scale > 0
        err = float((x.grad - g_ref).abs().max()) / scale
        rel_l2 = float((x.grad - g_ref).norm() / g_ref.norm())
        assert err < 1e-3 and rel_l2 < 1e-4, (err, rel_l2)
    # determinism of the fused path
    x = x0.clone().requires_grad_(True)
    v2 = loss_utils.multi_scale_loss(x, y, scales)
    v2.backward()
    x3 = x0.clone().requires_grad_(True)
    v3 = loss_utils.multi_scale_loss(x3, y, scales)
    v3.backward()
    assert v2.item() == v3.item() and torch.equal(x.grad, x3.grad)


@pytest.mark.parametrize("shape", [(3, 1, 1), (3, 7, 300), (1, 17, 31), (3, 16, 32), (3, 33, 65), (2, 200, 45), (3, 5, 5)])
def test_loss_ragged_sizes_against_oracle(device, shape):
    """Tile edges of the 32x16 CTA tiles, images smaller than the 11x11 window, single pixels, non-RGB channel counts."""
    C_, H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    gt = rng.uniform(0, 1, shape).astype(np.float32)
    img = np.clip(gt + rng.normal(0, 0.15, shape), 0, 1).astype(np.float32)
    o = _run(img, gt, 0.2, False, None, device)
    r = loss_oracle.mapper_loss(img, gt, 0.2, False)
    for k in ("l1", "ssim", "loss"):
        np.testing.assert_allclose(o[k], float(r[k]), rtol=2e-5)
    _grad_close(o["dL_dimage"], r["dL_dimage"])
