"""Generate tests/golden/loss_*.npz and adam_*.npz by running the REFERENCE's own loss code
(/root/reference/include/loss_utils.h, compiled into oracle/_ref/libloss_ref.so by `make -C oracle lossref`)
and torch::optim::Adam (the optimizer class the reference instantiates) on the CPU, in the build container.

    python tests/golden/make_loss_golden.py

The fixtures pin oracle/loss_oracle.py (-m "not gpu") and check segs_slam_b200/csrc/{loss,optim}.cu (-m gpu).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
import torch  # noqa: F401,E402  (libtorch must be resident before the veneer is loaded)

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libloss_ref.so"))
fp = C.POINTER(C.c_float)
lib.ref_mapper_loss.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp, C.c_float, C.c_int, C.c_int, fp, fp, fp, fp]
lib.ref_mapper_loss.restype = C.c_int
lib.ref_psnr.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp]
lib.ref_psnr.restype = C.c_float
lib.ref_adam.argtypes = [C.c_int, fp, fp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
lib.ref_adam.restype = C.c_int


lib.ref_frequency_losses.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp]
lib.ref_frequency_losses.restype = C.c_int


def p(a):
    return a.ctypes.data_as(fp)


def images(C_, H, W, seed, zero_rows=False):
    """A smooth-ish 'rendered' image and a ground truth that differs from it at every scale."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H, dtype=np.float32), np.linspace(0, 1, W, dtype=np.float32), indexing="ij")
    gt = np.stack([0.5 + 0.4 * np.sin(7 * xx + c) * np.cos(5 * yy - c) for c in range(C_)]).astype(np.float32)
    gt += rng.normal(0, 0.05, gt.shape).astype(np.float32)
    gt = np.clip(gt, 0.0, 1.0).astype(np.float32)
    img = np.clip(gt + rng.normal(0, 0.1, gt.shape).astype(np.float32) + 0.05, 0.0, 1.0).astype(np.float32)
    img[:, : H // 4, : W // 3] = gt[:, : H // 4, : W // 3]            # exact agreement somewhere (sign(0) = 0)
    if zero_rows:
        gt[:, 3:6, :] = 0.0                                            # rows the mask_rgb of the mapper removes
        gt[1, 10, :] = 0.0
    return np.ascontiguousarray(img), np.ascontiguousarray(gt)


def loss_case(name, C_, H, W, seed, lam, apply_mask, n_scaling):
    img, gt = images(C_, H, W, seed, zero_rows=apply_mask)
    rng = np.random.default_rng(seed + 1)
    sc = rng.uniform(0.005, 0.03, (max(n_scaling, 1), 3)).astype(np.float32)
    out3 = np.zeros(3, np.float32)
    dimg = np.zeros_like(img)
    dsc = np.zeros_like(sc)
    rc = lib.ref_mapper_loss(C_, H, W, p(img), p(gt), lam, int(apply_mask), n_scaling, p(sc), p(out3), p(dimg), p(dsc))
    assert rc == 0
    d = dict(image=img, gt=gt, lambda_dssim=np.float32(lam), apply_mask=np.int32(apply_mask), l1=out3[0], ssim=out3[1],
             loss=out3[2], dL_dimage=dimg, psnr=np.float32(lib.ref_psnr(C_, H, W, p(img), p(gt))))
    if n_scaling:
        d.update(scaling=sc, dL_dscaling=dsc)
    np.savez_compressed(os.path.join(HERE, f"loss_{name}.npz"), **d)
    print(name, out3)


def freq_case(name, C_, H, W, seed):
    img, gt = images(C_, H, W, seed)
    out3 = np.zeros(3, np.float32)
    dh, dm = np.zeros_like(img), np.zeros_like(img)
    assert lib.ref_frequency_losses(C_, H, W, p(img), p(gt), p(out3), p(dh), p(dm)) == 0
    np.savez_compressed(os.path.join(HERE, f"freq_{name}.npz"), image=img, gt=gt, high=out3[0], low=out3[1], multi=out3[2],
                        d_high=dh, d_multi=dm)
    print(name, out3)


def adam_case(name, n, steps, seed, lr, b1, b2, eps, wd):
    rng = np.random.default_rng(seed)
    p0 = rng.normal(0, 1, n).astype(np.float32)
    g = (rng.normal(0, 1, (steps, n)) * rng.uniform(1e-4, 1.0, (1, n))).astype(np.float32)
    g[:, :7] = 0.0                                                     # parameters that never see a gradient
    p1 = p0.copy()
    assert lib.ref_adam(n, p(p1), p(g), steps, lr, b1, b2, eps, wd) == 0
    np.savez_compressed(os.path.join(HERE, f"adam_{name}.npz"), param0=p0, grads=g, param=p1, lr=lr, beta1=b1, beta2=b2,
                        eps=eps, weight_decay=wd)
    print(name, float(np.abs(p1 - p0).max()))


if __name__ == "__main__":
    loss_case("rgb_53x75", 3, 53, 75, 11, 0.2, False, 0)              # ragged tile edges, the mapper's lambda
    loss_case("masked_48x64", 3, 48, 64, 12, 0.2, True, 500)          # mask_rgb rows + scaling regulariser
    loss_case("ssim_only_20x9", 3, 20, 9, 13, 1.0, False, 0)          # narrower than the window
    loss_case("l1_only_33x40", 1, 33, 40, 14, 0.0, False, 0)
    freq_case("rgb_40x50", 3, 40, 50, 15)                             # even sizes (bilinear 0.5x -> 20x25)
    freq_case("rgb_33x47", 3, 33, 47, 16)                             # odd sizes
    adam_case("default", 1000, 5, 21, 1e-3, 0.9, 0.999, 1e-8, 0.0)
    adam_case("segs", 777, 4, 22, 0.0075, 0.9, 0.999, 1e-15, 0.0)     # eps of gaussian_model.cpp:634
    adam_case("decay", 300, 3, 23, 2e-3, 0.8, 0.99, 1e-8, 0.01)
