"""GPU: the full per-view pipeline of the batched mapping step on this package's kernels
(anchor prefilter -> fused decode -> rasterize -> L1), single rank."""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import decode_oracle as do  # noqa: E402

import test_decode_gpu as td  # noqa: E402
from segs_slam_b200 import mapper, synth  # noqa: E402

pytestmark = pytest.mark.gpu


class Keyframe:
    """The GaussianKeyframe members the renderer reads (gaussian_keyframe.cpp:151-184)."""

    def __init__(self, scene, dev, t=(0.0, 0.0, 0.0)):
        R = np.eye(3, dtype=np.float32)
        s = synth.with_camera(scene, R, np.asarray(t, dtype=np.float32))
        self.world_view_transform_ = torch.from_numpy(s.viewmatrix).to(dev)
        self.full_proj_transform_ = torch.from_numpy(s.projmatrix).to(dev)
        self.camera_center_ = torch.from_numpy(s.campos).to(dev)
        self.t_ = tuple(float(x) for x in t)
        self.R_quaternion_ = (1.0, 0.0, 0.0, 0.0)


def test_mapping_step_reduces_loss(device):
    torch.manual_seed(0)
    W, H, fx = 200, 120, 150.0
    scene = synth.synth(16, W, H, fx, fx, 3)                    # only the camera model is used
    model = td._adapt(do.synth_model(3000, W, H, fx, fx, 17, do.DecodeConfig(), device=device))
    cams = [Keyframe(scene, device, t=(0.02 * v, 0.01 * v, 0.0)) for v in range(4)]
    bg = torch.zeros(3, device=device)
    g = torch.Generator(device="cpu").manual_seed(1)
    targets = [torch.rand(3, H, W, generator=g).to(device) * 0.5 for _ in cams]
    render_loss = mapper.make_render_loss(model, cams, targets, H, W, scene.tanfovx, scene.tanfovy, bg)
    params = [p for p in model.parameters()]
    opt = torch.optim.Adam(params, lr=2e-3)
    losses = []
    bucket = None
    for _ in range(6):
        loss, bucket = mapper.mapping_step(params, render_loss, len(cams), opt, bucket)
        losses.append(float(loss))
    assert all(math.isfinite(x) for x in losses)
    assert bucket.flat.abs().sum().item() > 0
    assert losses[-1] < losses[0], losses


def _small_setup(device, A=4000, n_views=3):
    from segs_slam_b200 import anchor_model
    W, H, fx = 208, 120, 150.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    model = anchor_model.synth_anchor_model(A, W, H, fx, fx, 1003, device=device)
    cams = anchor_model.circle_keyframes(8, 1.5, (0.0, 0.0, 3.25), tanx, tany, device)[:n_views]
    g = torch.Generator(device="cpu").manual_seed(1)
    targets = [(torch.rand(3, H, W, generator=g) * 0.5).to(device) for _ in cams]
    if len(targets) > 1:
        targets[1][:, 7:9, :] = 0.0                                  # rows the mapper's mask_rgb removes
    return model, cams, targets, (W, H, tanx, tany)


@pytest.mark.parametrize("lanes", [1, 2])
def test_fused_view_matches_autograd_composition(device, lanes):
    """segs_mapper_view (C++-issued view, gradients accumulated in place) == the autograd composition of the
    tensor-level API (prefilter, decode, rasterize, fused loss, scaling regulariser) on the same views."""
    from segs_slam_b200 import loss_utils
    model, cams, targets, (W, H, tanx, tany) = _small_setup(device)
    bg = torch.tensor([0.1, 0.0, 0.2], device=device)
    masks = [loss_utils.mask_rgb(t) for t in targets]
    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, scaling_reg_weight=0.01, lanes=lanes)
    loss_f = fm.step(cams, targets, masks, optimize=False)
    grads_f = [v.clone() / len(cams) for v in fm.bucket.views]
    res = fm.last_result
    assert res.n_visible > 0 and res.n_gaussians > 0 and res.num_rendered > 0

    render_loss = mapper.make_render_loss(model, cams, targets, H, W, tanx, tany, bg, loss="l1_ssim", lambda_dssim=0.2,
                                          scaling_reg_weight=0.01, row_masks=masks)
    loss_a, bucket = mapper.mapping_step(fm.params, render_loss, len(cams), optimizer=None)
    np.testing.assert_allclose(float(loss_f), float(loss_a), rtol=1e-5)
    names = ["_anchor", "_offset", "_anchor_feat", "_scaling"] + [f"w{i}" for i in range(len(grads_f) - 4)]
    for name, gf, ga in zip(names, grads_f, bucket.views):
        scale = float(ga.abs().max()) + 1e-30
        err = float((gf - ga).abs().max()) / scale
        assert err < 1e-4, (name, err)
        assert float(ga.abs().max()) > 0, name


def test_fused_mapper_optimises_like_torch_adam(device):
    """Two steps of FusedMapper (fused Adam, per-tensor learning rates) against the autograd path + torch.optim.Adam."""
    import copy
    from segs_slam_b200 import loss_utils
    model, cams, targets, (W, H, tanx, tany) = _small_setup(device, A=3000, n_views=2)
    model_b = copy.deepcopy(model)
    bg = torch.zeros(3, device=device)
    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lrs=2e-3, eps=1e-15)
    params_b = [model_b._anchor, model_b._offset, model_b._anchor_feat, model_b._scaling] + \
        [w for w in __import__("segs_slam_b200.gaussian_renderer", fromlist=["_weights"])._weights(model_b) if w is not None]
    opt_b = torch.optim.Adam(params_b, lr=2e-3, eps=1e-15)
    render_loss = mapper.make_render_loss(model_b, cams, targets, H, W, tanx, tany, bg, loss="l1_ssim")
    bucket = None
    losses = []
    for _ in range(2):
        lf = fm.step(cams, targets)
        lb, bucket = mapper.mapping_step(params_b, render_loss, len(cams), opt_b, bucket)
        np.testing.assert_allclose(float(lf), float(lb), rtol=2e-4)
        losses.append(float(lf))
    for pa, pb in zip(fm.params, params_b):
        # Adam's first steps move every touched element by ~lr regardless of the gradient's size, so tiny
        # gradient differences (atomic order) can flip nothing but can shift m/sqrt(v) slightly
        assert float((pa - pb).abs().max()) < 2e-4 * 2e-3 * 50 + 1e-6, float((pa - pb).abs().max())
    assert fm.workspace_bytes() > 0



@pytest.mark.parametrize("lanes", [1, 2])
def test_raster_views_match_the_tensor_level_api(device, lanes):
    """segs_raster_views (batched views on lanes, accumulated gradients) == per-view RasterizeGaussiansCUDA /
    RasterizeGaussiansBackwardCUDA summed over the views: images and num_rendered bit-exact, gradients 1e-4."""
    import common
    from segs_slam_b200 import rasterize_points as rp
    scene = synth.config("small")
    t = scene.to_torch(device)
    P, W, H = scene.P, scene.W, scene.H
    views = [synth.with_camera(scene, np.eye(3, dtype=np.float32), np.array([0.03 * v, -0.02 * v, 0.0], np.float32))
             for v in range(3)]
    cams = [{k: torch.from_numpy(np.ascontiguousarray(getattr(s, k))).to(device) for k in ("viewmatrix", "projmatrix", "campos")}
            for s in views]
    g = torch.Generator(device="cpu").manual_seed(2)
    dLs = [torch.randn(3, H, W, generator=g).to(device) for _ in cams]
    e = common.empty(device)
    ref_imgs, ref_R, ref_sum = [], [], None
    for cam, dL in zip(cams, dLs):
        R, color, radii, gb_, bb_, ib_ = rp.RasterizeGaussiansCUDA(t["bg"], t["means3D"], t["colors"], t["opacities"], t["scales"],
                                                                   t["rotations"], 1.0, e, cam["viewmatrix"], cam["projmatrix"],
                                                                   scene.tanfovx, scene.tanfovy, H, W, e, 0, cam["campos"], False)
        gr = rp.RasterizeGaussiansBackwardCUDA(t["bg"], t["means3D"], radii, t["colors"], t["scales"], t["rotations"], 1.0, e,
                                               cam["viewmatrix"], cam["projmatrix"], scene.tanfovx, scene.tanfovy, dL, e, 0,
                                               cam["campos"], gb_, R, bb_, ib_)
        six = [gr[3], gr[0], gr[1], gr[2], gr[6], gr[7]]
        ref_sum = [x.clone() for x in six] if ref_sum is None else [a + b for a, b in zip(ref_sum, six)]
        ref_imgs.append(color)
        ref_R.append(R)
    rb = mapper.RasterBatch(device, lanes=lanes)
    imgs = [torch.empty(3, H, W, device=device) for _ in cams]
    acc = [torch.zeros(P, w, device=device) for w in (3, 3, 3, 1, 3, 4)]
    Rs = rb.run(t["means3D"], t["colors"], t["opacities"], t["scales"], t["rotations"], t["bg"], cams, H, W, scene.tanfovx,
                scene.tanfovy, imgs, dLs, acc)
    torch.cuda.synchronize()
    assert Rs == ref_R
    for a, b in zip(imgs, ref_imgs):
        assert torch.equal(a, b)
    for a, b in zip(acc, ref_sum):
        ok, why = common.grad_close(a.reshape(b.shape), b, None, rel=1e-4)
        assert ok, why
    # forward only: no dL_dout, no accumulators
    imgs2 = [torch.empty(3, H, W, device=device) for _ in cams]
    rb.run(t["means3D"], t["colors"], t["opacities"], t["scales"], t["rotations"], t["bg"], cams, H, W, scene.tanfovx,
           scene.tanfovy, imgs2)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(imgs2, ref_imgs))


def test_mapper_view_edge_cases(device):
    """No visible anchor (camera looking away): the view contributes the loss of the background image and no
    gradient; an empty Gaussian set through segs_raster_views leaves a zero image like RasterizeGaussiansCUDA."""
    from segs_slam_b200 import anchor_model, loss_utils
    model, cams, targets, (W, H, tanx, tany) = _small_setup(device, A=2000, n_views=1)
    R = np.diag([-1.0, 1.0, -1.0])                                   # look along -z: every anchor is behind the camera
    away = anchor_model.Keyframe(R, np.zeros(3), tanx, tany, device)
    bg = torch.tensor([0.2, 0.4, 0.6], device=device)
    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, lanes=1)
    loss = fm.step([away], targets, optimize=False)
    res = fm.last_result
    assert res.n_visible == 0 and res.n_gaussians == 0 and res.num_rendered == 0
    assert float(fm.bucket.flat.abs().max()) == 0.0
    expected = loss_utils.l1_ssim_loss(torch.zeros(3, H, W, device=device), targets[0], 0.2)[0]
    # RasterizeGaussiansCUDA with P == 0 leaves the image ZERO, not background (rasterize_points.cu:81)
    np.testing.assert_allclose(float(loss), float(expected), rtol=1e-6)

    rb = mapper.RasterBatch(device, lanes=2)
    e3, e1, e4 = torch.empty(0, 3, device=device), torch.empty(0, 1, device=device), torch.empty(0, 4, device=device)
    cam = {"viewmatrix": away.world_view_transform_, "projmatrix": away.full_proj_transform_, "campos": away.camera_center_}
    imgs = [torch.full((3, H, W), 7.0, device=device) for _ in range(2)]
    Rs = rb.run(e3, e3, e1, e3, e4, bg, [cam, cam], H, W, tanx, tany, imgs)
    torch.cuda.synchronize()
    assert Rs == [0, 0] and all(float(i.abs().max()) == 0.0 for i in imgs)


@pytest.mark.parametrize("lanes", [1, 2])
def test_fused_mapper_training_statistics(device, lanes):
    """The densification statistics the fused view accumulates (segs_training_statis) against the restatement of
    GaussianModel::training_statis (oracle/statis_oracle.py) fed by the autograd composition of the same views."""
    import statis_oracle
    from segs_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer, generate_neural_gaussians, loss_utils
    from segs_slam_b200.rasterize_points import RasterizeGaussiansfilterCUDA
    model, cams, targets, (W, H, tanx, tany) = _small_setup(device, A=4000, n_views=3)
    bg = torch.zeros(3, device=device)
    A = model._anchor.size(0)
    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lanes=lanes, statistics=True)
    fm.step(cams, targets, optimize=False)
    fm.step(cams, targets, optimize=False)                          # running accumulators: two steps = twice the views

    z = lambda n: torch.zeros(n, 1, device=device)
    ref = dict(opacity_accum=z(A), anchor_demon=z(A), offset_gradient_accum=z(10 * A), offset_denom=z(10 * A))
    e = torch.empty(0, device=device)
    for cam, tgt in zip(cams, targets):
        with torch.no_grad():
            radii_a = RasterizeGaussiansfilterCUDA(model.get_anchor(), model.get_scaling()[:, :3].contiguous(),
                                                   torch.nn.functional.normalize(model._rotation), 1.0, e,
                                                   cam.world_view_transform_, cam.full_proj_transform_, tanx, tany, H, W, False)
            visible = radii_a > 0
        xyz, color, opacity, scaling, rots, nop, mask = generate_neural_gaussians(cam, model, visible)
        settings = GaussianRasterizationSettings(H, W, tanx, tany, bg, 1.0, cam.world_view_transform_, cam.full_proj_transform_,
                                                 0, cam.camera_center_, False)
        means2D = torch.zeros_like(xyz, requires_grad=True)
        image, radii = GaussianRasterizer(settings)(xyz, means2D, opacity, False, True, True, True, False, e, color, scaling,
                                                    rots, e)
        loss = loss_utils.l1_ssim_loss(image, tgt, 0.2)[0] + loss_utils.scaling_reg(scaling, 0.01)
        loss.backward()
        for _ in range(2):
            statis_oracle.training_statis(ref, means2D.grad, nop.detach(), radii > 0, mask, visible)
    assert torch.equal(fm.anchor_demon, ref["anchor_demon"])
    assert torch.equal(fm.offset_denom, ref["offset_denom"])
    torch.testing.assert_close(fm.opacity_accum, ref["opacity_accum"], rtol=1e-5, atol=1e-6)
    scale = float(ref["offset_gradient_accum"].max())
    assert scale > 0
    torch.testing.assert_close(fm.offset_gradient_accum, ref["offset_gradient_accum"], rtol=1e-4, atol=1e-4 * scale)
