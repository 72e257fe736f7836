"""Dev tool (GPU): where does the fused mapping view differ from the reference chain at full size?  Stage by stage, with
cross-feeding: image, dL/dimage (our fused loss vs the reference's autograd through loss_utils.h), and the rasterizer
backward of OUR kernels vs the reference kernels on IDENTICAL inputs (the decode outputs and dL/dimage of the product)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import common  # noqa: E402
import model_ref  # noqa: E402
import refimpl  # noqa: E402
from segs_slam_b200 import anchor_model, generate_neural_gaussians, loss_utils  # noqa: E402
from segs_slam_b200 import rasterize_points as rp  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300)), float((a - b).abs().max() / (b.abs().max() + 1e-300))


def main():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    A = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    W, H, fx = 1200, 680, 600.0
    mr = model_ref.load()
    fovx, fovy = 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fx))
    tanx, tany = float(mr.tan_half_fov(fovx)), float(mr.tan_half_fov(fovy))
    model = anchor_model.synth_anchor_model(A, W, H, fx, fx, 1003, device=dev)
    cam = anchor_model.circle_keyframes(8, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)[0]
    g = torch.Generator(device="cpu").manual_seed(1)
    gt = (torch.rand(3, H, W, generator=g) * 0.5).to(dev)
    bg = torch.zeros(3, device=dev)
    e = torch.empty(0, dtype=torch.float32, device=dev)
    ref = model_ref.from_model(model, reference_ctor=True)
    out = ref.view_gradients(cam.world_view_transform_, cam.full_proj_transform_, cam.camera_center_, list(cam.t_),
                             list(cam.R_quaternion_), fovx, fovy, H, W, bg, gt, 0.2)
    n_par = 4 + len(ref.mlp_parameters())
    image_r, vs_grad_r, dimg_r = out[1 + n_par], out[2 + n_par], out[-1]
    # product, tensor level
    with torch.no_grad():
        radii = rp.RasterizeGaussiansfilterCUDA(model.get_anchor(), model.get_scaling()[:, :3].contiguous(), model.get_rotation(), 1.0, e,
                                                cam.world_view_transform_, cam.full_proj_transform_, tanx, tany, H, W, False)
    xyz, color, opacity, scaling, rots, _n, _m = [t.detach() for t in generate_neural_gaussians(cam, model, radii > 0)]
    a = (bg, xyz, color, opacity, scaling, rots, 1.0, e, cam.world_view_transform_, cam.full_proj_transform_, tanx, tany, H, W, e, 0,
         cam.camera_center_, False)
    R, image_p, radii_p, gb, bb, ib = rp.RasterizeGaussiansCUDA(*a)
    print("P", xyz.size(0), "R", R)
    print("image: relL2, max", rel(image_p, image_r))
    img = image_p.clone().requires_grad_(True)
    loss = loss_utils.l1_ssim_loss(img, gt, 0.2)[0]
    loss.backward()
    print("dL/dimage (fused loss on our image vs reference autograd on its image): relL2, max", rel(img.grad, dimg_r))
    import loss_oracle
    img2 = image_p.clone().requires_grad_(True)
    l2 = 0.8 * loss_oracle.l1_loss(img2, gt) + 0.2 * (1.0 - loss_oracle.ssim(img2, gt))
    l2.backward()
    print("dL/dimage (fused loss vs ATen restatement, same image): relL2, max", rel(img.grad, img2.grad))
    print("dL/dimage (ATen restatement on our image vs reference): relL2, max", rel(img2.grad, dimg_r))
    names = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations")
    for label, dimg in (("reference dL/dimage", dimg_r.contiguous()), ("our dL/dimage", img.grad.contiguous())):
        mine = rp.RasterizeGaussiansBackwardCUDA(bg, xyz, radii_p, color, scaling, rots, 1.0, e, cam.world_view_transform_,
                                                 cam.full_proj_transform_, tanx, tany, dimg, e, 0, cam.camera_center_, gb, R, bb, ib)
        Rr, color_r, radii_r, g2, b2, i2 = refimpl.forward(bg, xyz, color, opacity, scaling, rots, 1.0, e, cam.world_view_transform_,
                                                           cam.full_proj_transform_, tanx, tany, H, W, e, 0, cam.camera_center_)
        d = refimpl.backward(bg, xyz, radii_r, color, scaling, rots, 1.0, e, cam.world_view_transform_, cam.full_proj_transform_, tanx,
                             tany, dimg, e, 0, cam.camera_center_, g2, Rr, b2, i2)
        d2 = refimpl.backward(bg, xyz, radii_r, color, scaling, rots, 1.0, e, cam.world_view_transform_, cam.full_proj_transform_, tanx,
                              tany, dimg, e, 0, cam.camera_center_, g2, Rr, b2, i2)
        torch.cuda.synchronize()
        print(f"rasterizer backward, same inputs, {label}:")
        for n, m_ in zip(names, mine):
            if n == "dL_dsh":
                continue
            print(f"   {n:14s} ours vs ref relL2 {rel(m_, d[n])[0]:.2e} max {rel(m_, d[n])[1]:.2e} | ref vs ref relL2 {rel(d2[n], d[n])[0]:.2e} | "
                  f"{common.grad_close(m_.reshape(d[n].shape), d[n])[1]}")


if __name__ == "__main__":
    main()
