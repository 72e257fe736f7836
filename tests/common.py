"""Shared helpers for the parity tests: run the product and the reference on one scene."""
from __future__ import annotations

import numpy as np
import torch

import refimpl
from segs_slam_b200 import rasterize_points as rp


def empty(dev):
    return torch.empty(0, dtype=torch.float32, device=dev)


def scene_args(t, scene, dev, use_sh=False, use_cov=None):
    """Argument dict shared by both implementations."""
    sh = t["sh"] if use_sh else empty(dev)
    colors = empty(dev) if use_sh else t["colors"]
    degree = int(scene.extras["sh_degree"][0]) if use_sh else 0
    if use_cov is not None:
        scales, rots, cov = empty(dev), empty(dev), use_cov
    else:
        scales, rots, cov = t["scales"], t["rotations"], empty(dev)
    return dict(bg=t["bg"], means3D=t["means3D"], colors=colors, opacity=t["opacities"], scales=scales,
                rotations=rots, scale_modifier=scene.scale_modifier, cov3D_precomp=cov,
                viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"], tan_fovx=scene.tanfovx,
                tan_fovy=scene.tanfovy, H=scene.H, W=scene.W, sh=sh, degree=degree, campos=t["campos"])


def run_mine(a, dL_dout=None):
    out = {}
    (R, color, radii, geom, binning, img) = rp.RasterizeGaussiansCUDA(
        a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], a["scale_modifier"],
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"],
        a["sh"], a["degree"], a["campos"], False)
    out.update(R=R, color=color, radii=radii, geom=geom, binning=binning, img=img)
    if dL_dout is not None:
        g = rp.RasterizeGaussiansBackwardCUDA(
            a["bg"], a["means3D"], radii, a["colors"], a["scales"], a["rotations"], a["scale_modifier"],
            a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], dL_dout,
            a["sh"], a["degree"], a["campos"], geom, R, binning, img)
        names = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
                 "dL_dscales", "dL_drotations")
        out["grads"] = dict(zip(names, g))
    return out


def run_ref(a, dL_dout=None):
    out = {}
    (R, color, radii, geom, binning, img) = refimpl.forward(
        a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], a["scale_modifier"],
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"],
        a["sh"], a["degree"], a["campos"], False)
    refimpl.load().ref_sync()
    out.update(R=R, color=color, radii=radii, geom=geom, binning=binning, img=img)
    if dL_dout is not None:
        out["grads"] = refimpl.backward(
            a["bg"], a["means3D"], radii, a["colors"], a["scales"], a["rotations"], a["scale_modifier"],
            a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], dL_dout,
            a["sh"], a["degree"], a["campos"], geom, R, binning, img)
        refimpl.load().ref_sync()
    return out


def mine_sections(m, P, W, H):
    """Internal state of the product in the reference's vocabulary."""
    R = m["R"]
    sec = lambda n, dt, shape=None: rp.buffer_section(n, m["geom"], m["binning"], m["img"], P, R, W, H, dt, shape)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    rec = sec("rec", torch.float32, (P, 12))
    s = dict(depths=sec("depths", torch.float32), tiles_touched=sec("tiles_touched", torch.int32),
             rect=sec("rect", torch.int16, (P, 4)), means2D=rec[:, 0:2], extent=rec[:, 2:4],
             conic_opacity=rec[:, 4:8], rgb=rec[:, 8:11],
             cov3D=sec("cov3D", torch.float32, (6, P)).t(),
             point_list=sec("point_list", torch.int32) if R else torch.empty(0, dtype=torch.int32, device=rec.device),
             ranges=sec("ranges", torch.int32, (T, 2)), final_T=sec("final_T", torch.float32),
             n_contrib=sec("n_contrib", torch.int32))
    # tile id of every list position (the product never materialises tile|depth keys; they are
    # reconstructed for the parity check from ranges + depth bits)
    counts = (s["ranges"][:, 1] - s["ranges"][:, 0]).long()
    s["tile_ids"] = torch.repeat_interleave(torch.arange(T, device=rec.device, dtype=torch.int32), counts)
    return s


def bits(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous().view(torch.int32)


def grad_close(mine: torch.Tensor, ref: torch.Tensor, ref2: torch.Tensor | None = None, rel=1e-4):
    """Gradient parity at `rel` relative (north_star: 1e-4).

    The reference accumulates with order-nondeterministic float atomics (9 per blended pair,
    backward.cu:523-554) and then pushes the sums through ill-conditioned per-Gaussian algebra
    (dL_dconic -> dL_dcov3D has catastrophic cancellation), so two runs OF THE REFERENCE differ
    from each other by more than 1e-4 on ~1e-5 of the elements (measured on B200, see
    tools/grad_diag.py / DESIGN.md).  The check is therefore:
      * every element within rel * (|ref| + mean|ref|), except at most max(2, 1e-4 * numel)
        outliers (the reference-vs-itself outlier rate, with head-room), and
      * the relative L2 error of the whole tensor <= `rel` (the reference's own run-to-run
        relative L2 spread reaches 6e-5 on dL_dcov3D / dL_drotations at C1).
    Returns (ok, description)."""
    mine, ref = mine.double().flatten(), ref.double().flatten()
    if ref.numel() == 0:
        return True, "empty"
    scale = ref.abs().mean() + 1e-30
    tol = rel * (ref.abs() + scale)
    ratio = (mine - ref).abs() / tol
    n_bad = int((ratio > 1.0).sum().item())
    allowed = max(2, int(1e-4 * ref.numel()))
    rel_l2 = ((mine - ref).norm() / (ref.norm() + 1e-30)).item()
    ok = n_bad <= allowed and rel_l2 <= rel
    return ok, f"outliers {n_bad}/{ref.numel()} (allowed {allowed}), worst {ratio.max().item():.1f}x, relL2 {rel_l2:.2e}"


def to_np(t):
    return t.detach().cpu().numpy()
