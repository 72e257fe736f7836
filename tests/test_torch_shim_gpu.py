"""GPU parity through the C++/LibTorch host layer (segs_slam_b200/csrc/torch_shim): the SAME entry
points a SEGS-SLAM build links instead of src/rasterize_points.cu — RasterizeGaussiansCUDA,
RasterizeGaussiansBackwardCUDA, RasterizeGaussiansfilterCUDA, RasterizeGaussiansprojectCUDA,
markVisible, distCUDA2 (/root/reference/include/rasterize_points.h:18-102, spatial.h:14) — and the
C++ autograd function (src/gaussian_rasterizer.cpp:28-154), compared with the unmodified
reference CUDA rasterizer (oracle/_ref) on identical inputs."""
import pytest
import torch

import common
import refimpl
from segs_slam_b200 import synth

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refimpl.available(), reason="oracle/_ref/libsegs_ref.so not built")


@pytest.fixture(scope="module")
def shim():
    from segs_slam_b200 import _segs_torch   # raises if the extension has not been built: no fallback
    return _segs_torch


def _fwd_args(a):
    return (a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], a["scale_modifier"],
            a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"],
            a["sh"], a["degree"], a["campos"], False)


@needs_ref
@pytest.mark.parametrize("name", ["tiny", "small", "C1"])
def test_cpp_entry_points_match_reference(device, shim, name):
    scene = synth.config(name)
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device)
    R, color, radii, geom, binning, img = shim.RasterizeGaussiansCUDA(*_fwd_args(a))
    grads = shim.RasterizeGaussiansBackwardCUDA(
        a["bg"], a["means3D"], radii, a["colors"], a["scales"], a["rotations"], a["scale_modifier"],
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], t["dL_dout"],
        a["sh"], a["degree"], a["campos"], geom, R, binning, img)
    r = common.run_ref(a, t["dL_dout"])
    r2 = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    assert R == r["R"]
    assert torch.equal(radii, r["radii"])
    assert torch.equal(common.bits(color), common.bits(r["color"]))
    assert geom.dtype == torch.uint8 and binning.dtype == torch.uint8 and img.dtype == torch.uint8
    names = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
             "dL_drotations")   # tuple order of src/rasterize_points.cu:192
    P = scene.P
    shapes = dict(dL_dmeans2D=(P, 3), dL_dcolors=(P, 3), dL_dopacity=(P, 1), dL_dmeans3D=(P, 3), dL_dcov3D=(P, 6),
                  dL_dsh=(P, 0, 3), dL_dscales=(P, 3), dL_drotations=(P, 4))
    for k, g in zip(names, grads):
        assert tuple(g.shape) == shapes[k], k
        ok, why = common.grad_close(g, r["grads"][k], r2["grads"][k], 1e-4)
        assert ok, (k, why)


@needs_ref
def test_cpp_autograd_function(device, shim):
    """GaussianRasterizer::forward -> GaussianRasterizerFunction (C++ autograd): gradients land on
    the right leaves in the order of src/gaussian_rasterizer.cpp:143-153."""
    scene = synth.config("small", bg=(0.3, 0.1, 0.6))
    t = scene.to_torch(device)
    leaf = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "colors", "opacities", "scales", "rotations")}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
    e = common.empty(device)
    color, radii = shim.rasterizer_forward(scene.H, scene.W, scene.tanfovx, scene.tanfovy, t["bg"], 1.0,
                                           t["viewmatrix"], t["projmatrix"], 0, t["campos"], False,
                                           leaf["means3D"], means2D, leaf["opacities"], e, leaf["colors"],
                                           leaf["scales"], leaf["rotations"], e)
    (color * t["dL_dout"]).sum().backward()
    a = common.scene_args(t, scene, device)
    r = common.run_ref(a, t["dL_dout"])
    r2 = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    assert torch.equal(radii, r["radii"])
    assert torch.equal(common.bits(color.detach()), common.bits(r["color"]))
    for leaf_name, gname in [("means3D", "dL_dmeans3D"), ("colors", "dL_dcolors"), ("opacities", "dL_dopacity"),
                             ("scales", "dL_dscales"), ("rotations", "dL_drotations")]:
        ok, why = common.grad_close(leaf[leaf_name].grad, r["grads"][gname], r2["grads"][gname], 1e-4)
        assert ok, (leaf_name, why)
    ok, why = common.grad_close(means2D.grad, r["grads"]["dL_dmeans2D"], r2["grads"]["dL_dmeans2D"], 1e-4)
    assert ok, why
    with pytest.raises(RuntimeError, match="excatly one of either SHs or precomputed colors"):
        shim.rasterizer_forward(scene.H, scene.W, scene.tanfovx, scene.tanfovy, t["bg"], 1.0, t["viewmatrix"],
                                t["projmatrix"], 0, t["campos"], False, leaf["means3D"], means2D, leaf["opacities"],
                                e, e, leaf["scales"], leaf["rotations"], e)


@needs_ref
def test_cpp_aux_entry_points(device, shim):
    scene = synth.config("small")
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device)
    e = common.empty(device)
    radii = shim.RasterizeGaussiansfilterCUDA(a["means3D"], a["scales"], a["rotations"], 1.0, e, a["viewmatrix"],
                                              a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"], False,
                                              False)
    ref = refimpl.visible_filter(a["means3D"], a["scales"], a["rotations"], 1.0, e, a["viewmatrix"], a["projmatrix"],
                                 a["tan_fovx"], a["tan_fovy"], a["H"], a["W"])
    assert torch.equal(radii, ref)
    present = shim.markVisible(a["means3D"], a["viewmatrix"], a["projmatrix"])
    assert present.dtype == torch.bool
    assert torch.equal(present, refimpl.mark_visible(a["means3D"], a["viewmatrix"], a["projmatrix"]))
    pts, rad, col = shim.RasterizeGaussiansprojectCUDA(*_fwd_args(a))
    assert pts.shape == (scene.P, 2) and col.shape == (scene.P, 3)
    assert torch.equal(rad, ref)
    d = shim.distCUDA2(a["means3D"][:5000].contiguous())
    assert torch.equal(common.bits(d), common.bits(refimpl.knn(a["means3D"][:5000].contiguous())))


def test_cpp_empty_and_errors(device, shim):
    e = common.empty(device)
    bg = torch.zeros(3, device=device)
    eye = torch.eye(4, device=device)
    R, color, radii, g, b, i = shim.RasterizeGaussiansCUDA(bg, torch.zeros((0, 3), device=device), e, e, e, e, 1.0, e,
                                                          eye, eye, 1.0, 1.0, 32, 48, e, 0, bg, False)
    assert R == 0 and color.abs().max().item() == 0.0 and g.numel() == 0
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        shim.RasterizeGaussiansCUDA(bg, torch.zeros((5, 4), device=device), e, e, e, e, 1.0, e, eye, eye, 1.0, 1.0,
                                    16, 16, e, 0, bg, False)


# ---- loss_utils drop-in (torch_shim/loss_utils.{h,cpp}) against the goldens of the reference's own header ----------
import glob  # noqa: E402
import os  # noqa: E402

import numpy as np  # noqa: E402

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_LOSS = sorted(glob.glob(os.path.join(_GOLD, "loss_*.npz")))
_ADAM = sorted(glob.glob(os.path.join(_GOLD, "adam_*.npz")))


@pytest.mark.parametrize("path", _LOSS, ids=[os.path.basename(p) for p in _LOSS])
def test_cpp_loss_utils_match_reference_golden(device, shim, path):
    g = np.load(path)
    x = torch.from_numpy(g["image"]).to(device).requires_grad_(True)
    y = torch.from_numpy(g["gt"]).to(device)
    lam = float(g["lambda_dssim"])
    mask = (y != 0).any(-1).float() if bool(g["apply_mask"]) else torch.empty(0, device=device)
    loss = shim.l1_ssim(x, y, lam, mask)
    if "scaling" in g.files:
        sc = torch.from_numpy(g["scaling"]).to(device)
        loss = loss + 0.01 * sc.prod(1).mean()
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-5)
    ref = g["dL_dimage"]
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-4 * float(np.abs(ref).max()))
    if not bool(g["apply_mask"]):
        np.testing.assert_allclose(shim.l1_loss(x.detach(), y).item(), float(g["l1"]), rtol=1e-5)
        np.testing.assert_allclose(shim.ssim(x.detach(), y).item(), float(g["ssim"]), rtol=1e-5)
        np.testing.assert_allclose(shim.psnr(x.detach(), y).item(), float(g["psnr"]), rtol=1e-5)


@pytest.mark.parametrize("path", _ADAM, ids=[os.path.basename(p) for p in _ADAM])
def test_cpp_adam_step_matches_torch_optim_adam_golden(device, shim, path):
    g = np.load(path)
    n = g["param0"].size
    cuts = [0, n // 2, n]
    params = [torch.from_numpy(g["param0"][a:b].copy()).to(device) for a, b in zip(cuts[:-1], cuts[1:])]
    grad, m, v = (torch.zeros(n, device=device) for _ in range(3))
    for step, gr in enumerate(g["grads"], start=1):
        grad.copy_(torch.from_numpy(gr).to(device))
        shim.adam_step(params, [float(g["lr"])] * 2, grad, m, v, step, float(g["beta1"]), float(g["beta2"]), float(g["eps"]),
                       float(g["weight_decay"]), 1.0, True)
        assert float(grad.abs().max()) == 0.0
    mine = torch.cat(params).cpu().numpy()
    ulp = float(np.spacing(np.float32(np.abs(g["param0"]).max())))
    np.testing.assert_allclose(mine - g["param0"], g["param"] - g["param0"], rtol=2e-5, atol=2.0 * ulp)


def test_cpp_loss_rejects_cpu_tensors(shim):
    with pytest.raises(RuntimeError, match="no CPU path"):
        shim.l1_loss(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8))


def test_cpp_fused_mapper_matches_the_python_host_class(device, shim):
    """torch_shim/fused_mapper.{h,cpp} (the C++/LibTorch host class of the keyframe-batched step) against
    segs_slam_b200.mapper.FusedMapper on the same model and views: accumulated gradient bucket, loss, one Adam step."""
    import copy
    from segs_slam_b200 import anchor_model, mapper
    from segs_slam_b200.gaussian_renderer import _weights
    W, H, fx = 208, 120, 150.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    model = anchor_model.synth_anchor_model(3000, W, H, fx, fx, 1003, device=device)
    model_c = copy.deepcopy(model)
    cams = anchor_model.circle_keyframes(8, 1.5, (0.0, 0.0, 3.25), tanx, tany, device)[:3]
    g = torch.Generator(device="cpu").manual_seed(1)
    targets = [(torch.rand(3, H, W, generator=g) * 0.5).to(device) for _ in cams]
    bg = torch.tensor([0.1, 0.2, 0.3], device=device)

    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, scaling_reg_weight=0.01, lrs=2e-3, lanes=1)
    loss_py = fm.step(cams, targets, optimize=False) * len(cams)
    bucket_py = fm.bucket.flat.clone()
    fm.optimizer.step(grad_scale=1.0 / len(cams), zero_grad=True)

    w = _weights(model_c)
    n_par = 4 + sum(1 for t in w if t is not None)
    cfg = [model_c.appearance_dim, int(model_c.use_feat_bank), 0, 0, 0]
    cm = shim.FusedMapper([model_c._anchor.data, model_c._offset.data, model_c._anchor_feat.data, model_c._scaling.data,
                           model_c._rotation.data], [None if t is None else t.data for t in w], cfg, H, W, tanx, tany, bg, 0.2,
                          0.01, [2e-3] * n_par, 1e-15, 2)
    e = torch.empty(0, device=device)
    views = [(c.world_view_transform_, c.full_proj_transform_, c.camera_center_, list(c.t_) + list(c.R_quaternion_), t, e)
             for c, t in zip(cams, targets)]
    loss_c = cm.render_views(views)
    torch.testing.assert_close(loss_c, loss_py, rtol=1e-5, atol=0)
    bucket_c = cm.grad_flat().clone()
    scale = float(bucket_py.abs().max())
    assert scale > 0 and float((bucket_c - bucket_py).abs().max()) < 1e-4 * scale
    cm.adam_step(1.0 / len(cams))
    assert float(cm.grad_flat().abs().max()) == 0.0
    for a, b in zip(cm.params(), fm.params):
        assert float((a - b).abs().max()) < 2e-5, float((a - b).abs().max())
    assert cm.workspace_bytes() > 0
    # the Replica configuration of the loss (frequency regularisation) through the C++ class
    fq = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, scaling_reg_weight=0.01, lrs=2e-3, lanes=1,
                            lambda_frequency_high=0.01, use_multi_resolution=True, freq_scale_num=3)
    loss_fq = fq.step(cams, targets, optimize=False) * len(cams)
    cm.set_frequency(0.01, True, 3)
    # NOTE: cm's parameters took one Adam step above, like fm's (both hold the same values to 2e-5)
    loss_cq = cm.render_views(views)
    torch.testing.assert_close(loss_cq, loss_fq, rtol=1e-4, atol=0)
    assert float(loss_cq) > float(loss_c)
    scale = float(fq.bucket.flat.abs().max())
    assert float((cm.grad_flat() - fq.bucket.flat).abs().max()) < 2e-3 * scale
