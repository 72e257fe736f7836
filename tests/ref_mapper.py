"""TEST / BASELINE INFRASTRUCTURE — the reference's per-iteration mapping work, restated on the reference's own
CUDA rasterizer (oracle/_ref/libsegs_ref.so through tests/refimpl.py) and the ATen op sequences of its LibTorch
host code:

  GaussianMapper::trainForOneIteration   /root/reference/src/gaussian_mapper.cpp:823-1032
    prefilter_voxel                      src/gaussian_renderer.cpp:131-199      -> refimpl.visible_filter
    generate_neural_gaussians            src/gaussian_renderer.cpp:214-334      -> oracle/decode_oracle.py (same ATen ops)
    GaussianRasterizerFunction           src/gaussian_rasterizer.cpp:28-154     -> _RefRasterize below (reference kernels)
    loss                                 src/gaussian_mapper.cpp:908-925        -> oracle/loss_oracle.py (ATen conv2d SSIM)
    loss.backward(); optimizer->step()   :950, :1003-1006                       -> torch autograd, torch.optim.Adam

One optimizer step per keyframe, as the reference does.  Used by `bench.py --impl reference` for the mapping
figure and by nothing in the product (segs_slam_b200/ never imports it)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import decode_oracle  # noqa: E402
import loss_oracle  # noqa: E402
import refimpl  # noqa: E402


class _RefRasterize(torch.autograd.Function):
    """GaussianRasterizerFunction (gaussian_rasterizer.cpp:28-154) over the unmodified reference kernels."""

    @staticmethod
    def forward(ctx, means3D, means2D, colors, opacity, scales, rotations, bg, view, proj, campos, tanx, tany, H, W):
        e = torch.empty(0, dtype=torch.float32, device=means3D.device)
        a = [t.contiguous() for t in (means3D, colors, opacity, scales, rotations)]
        R, color, radii, geom, binning, img = refimpl.forward(bg, a[0], a[1], a[2], a[3], a[4], 1.0, e, view, proj, tanx,
                                                              tany, H, W, e, 0, campos, False)
        ctx.save_for_backward(*a, radii, geom, binning, img, bg, view, proj, campos)
        ctx.meta = (R, tanx, tany)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, g_color, _g_radii):
        means3D, colors, opacity, scales, rotations, radii, geom, binning, img, bg, view, proj, campos = ctx.saved_tensors
        R, tanx, tany = ctx.meta
        e = torch.empty(0, dtype=torch.float32, device=means3D.device)
        d = refimpl.backward(bg, means3D, radii, colors, scales, rotations, 1.0, e, view, proj, tanx, tany,
                             g_color.contiguous(), e, 0, campos, geom, R, binning, img)
        return (d["dL_dmeans3D"], d["dL_dmeans2D"], d["dL_dcolors"], d["dL_dopacity"], d["dL_dscales"], d["dL_drotations"],
                None, None, None, None, None, None, None, None)


def iteration(pc, cam, gt_image, H, W, tanx, tany, bg, optimizer, lambda_dssim=0.2):
    """One reference training iteration on one keyframe.  `pc` is a decode_oracle.DecodeModel-like module with
    `_rotation` optional.  -> loss (device scalar)."""
    dev = bg.device
    e = torch.empty(0, dtype=torch.float32, device=dev)
    with torch.no_grad():
        scal = pc.get_scaling()[:, :3].contiguous()
        if hasattr(pc, "_rotation"):
            rot = torch.nn.functional.normalize(pc._rotation)
        else:
            rot = torch.tensor([1.0, 0.0, 0.0, 0.0], device=dev).repeat(pc._anchor.size(0), 1)
        radii = refimpl.visible_filter(pc._anchor.detach().contiguous(), scal, rot.contiguous(), 1.0, e,
                                       cam.world_view_transform_, cam.full_proj_transform_, tanx, tany, H, W)
        visible = radii > 0
    xyz, color, opacity, scaling, rots, _nop, _mask = decode_oracle.generate_neural_gaussians(
        pc, cam.camera_center_, cam.t_, cam.R_quaternion_, visible)
    means2D = torch.zeros_like(xyz, requires_grad=True)                      # screenspace_points, gaussian_renderer.cpp:55-65
    image, _radii = _RefRasterize.apply(xyz, means2D, color, opacity, scaling, rots, bg, cam.world_view_transform_,
                                        cam.full_proj_transform_, cam.camera_center_, tanx, tany, H, W)
    mask_rgb = (gt_image != 0.0).any(-1).to(torch.float32).unsqueeze(-1)     # gaussian_mapper.cpp:911-915
    masked, rendered, gt = image * mask_rgb, image * mask_rgb, gt_image * mask_rgb
    Ll1 = loss_oracle.l1_loss(rendered, gt)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - loss_oracle.ssim(masked, gt)) + 0.01 * scaling.prod(1).mean()
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss.detach()
