"""GPU: the product against the reference's OWN host layers — GaussianModel, GaussianRenderer (prefilter_voxel,
generate_neural_gaussians, render), GaussianRasterizer, rasterize_points.cu, the CUDA kernels and loss_utils.h —
compiled unmodified from /root/reference/src into oracle/_ref/_model_ref.so (oracle/Makefile `modelref`,
tests/model_ref.py).  Covers the compositions the per-stage parity tests leave open:

  * BASELINE configs 3 and 4: the fused mapping view (segs_mapper_view through FusedMapper) == the reference chain
    prefilter -> decode -> rasterize -> (1-l) L1 + l (1-SSIM) + 0.01 scaling regulariser -> backward
    (src/gaussian_mapper.cpp:870-950), loss 1e-5, every gradient tensor within 1e-4 of its scale, 1 and 2 lanes and one
    full-size C4 view;
  * the CUDA decode against the compiled generate_neural_gaussians, forward + backward, up to the C3 size;
  * project2_image (row V3) and the other tensor-level entry points through the reference's rasterize_points.cu."""
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import model_ref  # noqa: E402

from segs_slam_b200 import anchor_model, generate_neural_gaussians, mapper, synth  # noqa: E402
from segs_slam_b200 import rasterize_points as rp  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not model_ref.available(), reason="oracle/_ref/_model_ref.so not built")]


def _fov_pair(tan_wanted):
    """(fov, tan) such that tan == the float the reference derives from fov (std::tan(FoVx_ * 0.5f))."""
    fov = 2.0 * math.atan(tan_wanted)
    return fov, float(model_ref.load().tan_half_fov(fov))


def _setup(device, A, W, H, fx, n_views):
    fovx, tanx = _fov_pair(W / (2 * fx))
    fovy, tany = _fov_pair(H / (2 * fx))
    model = anchor_model.synth_anchor_model(A, W, H, fx, fx, 1003, device=device)
    cams = anchor_model.circle_keyframes(8, 1.5, (0.0, 0.0, 3.25), tanx, tany, device)[:n_views]
    g = torch.Generator(device="cpu").manual_seed(1)
    targets = [(torch.rand(3, H, W, generator=g) * 0.5).to(device) for _ in cams]
    if len(targets) > 1:
        targets[1][:, 7:9, :] = 0.0                 # rows the mapper's mask_rgb removes (gaussian_mapper.cpp:911-915)
    return model, cams, targets, (fovx, fovy, tanx, tany)


def _reference_views(model, cams, targets, fovx, fovy, H, W, bg, lam, one_ulp=False, frequency=None):
    """Loss sum and gradient sums of the reference chain over the views.  one_ulp: every anchor feature moved to the next
    representable float — an equally valid rounding of the inputs, used to measure how much the REFERENCE's own gradients
    move under a 1-ulp change (the chain is discontinuous: alpha < 1/255 skips, T < 1e-4 stops, sign() in the L1 term)."""
    if one_ulp:
        import copy
        model = copy.deepcopy(model)
        with torch.no_grad():
            model._anchor_feat.copy_(torch.nextafter(model._anchor_feat, torch.full_like(model._anchor_feat, float("inf"))))
    ref = model_ref.from_model(model, reference_ctor=True)
    if frequency is not None:
        ref.set_frequency(*frequency)
    total, loss = None, 0.0
    for cam, tgt in zip(cams, targets):
        out = ref.view_gradients(cam.world_view_transform_, cam.full_proj_transform_, cam.camera_center_, list(cam.t_),
                                 list(cam.R_quaternion_), fovx, fovy, H, W, bg, tgt, lam)
        n_par = 4 + len(ref.mlp_parameters())
        grads = out[1:1 + n_par]
        total = [g.clone() for g in grads] if total is None else [a + b for a, b in zip(total, grads)]
        loss += float(out[0])
    return loss, total


def _check_bucket(fm, n_views, loss_f, loss_r, grads_r, rel=1e-4, grads_r2=None):
    """Loss 1e-5; every gradient tensor: max |mine - ref| <= rel * max |ref| and relative L2 error <= rel.  With
    `grads_r2` (the reference evaluated again with its inputs moved by ONE ULP, see _reference_views) the bars are the
    larger of `rel` and 3x what that 1-ulp change does to the reference's own gradients: at full size a handful of
    alpha >= 1/255 decisions flip under any 1-ulp change of the decode (cuBLAS version, summation order), each moving
    a pixel by up to 4e-3 and through sign() / SSIM its dL/dimage by 10 % (measured: tools/diag_chain.py)."""
    np.testing.assert_allclose(float(loss_f), loss_r / n_views, rtol=1e-5)
    names = ["_anchor", "_offset", "_anchor_feat", "_scaling"] + [f"w{i}" for i in range(len(fm.bucket.views) - 4)]
    assert len(grads_r) == len(fm.bucket.views)
    report, bad = {}, {}
    for k, (name, gf, gr) in enumerate(zip(names, fm.bucket.views, grads_r)):
        gr = gr.view_as(gf)
        scale = float(gr.abs().max()) + 1e-30
        assert scale > 1e-30, name
        err = float((gf - gr).abs().max()) / scale
        rel_l2 = float((gf - gr).norm() / (gr.norm() + 1e-30))
        bar, bar_l2 = rel, rel
        if grads_r2 is not None:
            g2 = grads_r2[k].view_as(gf)
            noise = float((g2 - gr).abs().max()) / scale
            noise_l2 = float((g2 - gr).norm() / (gr.norm() + 1e-30))
            bar, bar_l2 = max(rel, 3.0 * noise), max(rel, 3.0 * noise_l2)
            report[name] = (err, rel_l2, noise, noise_l2)
        else:
            report[name] = (err, rel_l2)
        if err >= bar or rel_l2 > bar_l2:
            bad[name] = report[name]
    print("max-abs error / max|ref|, relative L2" + (", the same two for the reference under a 1-ulp input change" if grads_r2 is not None else "") + ":",
          {k: tuple(f"{x:.1e}" for x in v) for k, v in report.items()})
    assert not bad, bad


@pytest.mark.parametrize("lanes", [1, 2])
def test_fused_mapper_matches_reference_chain(device, lanes):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    W, H, fx = 208, 120, 150.0
    model, cams, targets, (fovx, fovy, tanx, tany) = _setup(device, 4000, W, H, fx, 4)
    bg = torch.tensor([0.1, 0.0, 0.2], device=device)
    from segs_slam_b200 import loss_utils
    masks = [loss_utils.mask_rgb(t) for t in targets]
    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, scaling_reg_weight=0.01, lanes=lanes)
    loss_f = fm.step(cams, targets, masks, optimize=False)
    loss_r, grads_r = _reference_views(model, cams, targets, fovx, fovy, H, W, bg, 0.2)
    _check_bucket(fm, len(cams), loss_f, loss_r, grads_r)


def test_fused_mapper_with_frequency_regularisation_matches_reference_chain(device):
    """The Replica configuration of the loss (cfg/gaussian_mapper/RGB-D/Replica/office0.yaml:140-146: lambda_frequency_high
    0.01, multi-resolution, 3 scales; src/gaussian_mapper.cpp:930-945): the fused view with the frequency term against the
    reference chain with the reference's own multi_scale_loss."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    W, H, fx = 208, 120, 150.0
    model, cams, targets, (fovx, fovy, tanx, tany) = _setup(device, 4000, W, H, fx, 3)
    bg = torch.tensor([0.1, 0.0, 0.2], device=device)
    from segs_slam_b200 import loss_utils
    masks = [loss_utils.mask_rgb(t) for t in targets]
    for multi in (True, False):
        fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, scaling_reg_weight=0.01, lanes=2,
                                lambda_frequency_high=0.01, use_multi_resolution=multi, freq_scale_num=3)
        loss_f = fm.step(cams, targets, masks, optimize=False)
        loss_r, grads_r = _reference_views(model, cams, targets, fovx, fovy, H, W, bg, 0.2, frequency=(0.01, multi, 3))
        fm0 = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, scaling_reg_weight=0.01, lanes=1)
        loss_0 = fm0.step(cams, targets, masks, optimize=False)
        assert float(loss_f) > float(loss_0) * 1.01           # the term is there (it is not a rounding-level contribution)
        _check_bucket(fm, len(cams), loss_f, loss_r, grads_r)


@pytest.mark.slow
def test_fused_mapper_matches_reference_chain_C4_view(device):
    """One full-size view of BASELINE config 4 (C3 anchor model, 1200x680)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    W, H, fx = 1200, 680, 600.0
    model, cams, targets, (fovx, fovy, tanx, tany) = _setup(device, 200_000, W, H, fx, 1)
    bg = torch.zeros(3, device=device)
    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lambda_dssim=0.2, scaling_reg_weight=0.01, lanes=1)
    loss_f = fm.step(cams, targets, None, optimize=False)
    loss_r, grads_r = _reference_views(model, cams, targets, fovx, fovy, H, W, bg, 0.2)
    _loss_r2, grads_r2 = _reference_views(model, cams, targets, fovx, fovy, H, W, bg, 0.2, one_ulp=True)
    _check_bucket(fm, 1, loss_f, loss_r, grads_r, grads_r2=grads_r2)


def test_reference_constructor_equals_veneer_constructor(device):
    """GaussianModel(const GaussianModelParams&) (gaussian_model.cpp:33-180) builds the same modules as the veneer's
    device-neutral construction that wrote the CPU goldens."""
    mr = model_ref.load()
    a, b = mr.RefModel(reference_ctor=True), mr.RefModel(reference_ctor=False)
    sa, sb = [tuple(p.shape) for p in a.mlp_parameters()], [tuple(p.shape) for p in b.mlp_parameters()]
    assert sa == sb and len(sa) == 18


def _decode_pair(device, A, seed, vis_frac, cfg_kw=None):
    import decode_oracle as do
    import test_decode_gpu as td
    cfg = do.DecodeConfig(**(cfg_kw or {}))
    model = td._adapt(do.synth_model(A, 1200, 680, 600.0, 600.0, seed, cfg, device=device))
    cam = td.Cam(device)
    g = torch.Generator(device="cpu").manual_seed(seed)
    vm = (torch.rand(A, generator=g) < vis_frac).to(device) if vis_frac is not None else torch.ones(A, dtype=torch.bool, device=device)
    ref_m = model_ref.from_model(model, reference_ctor=True)
    ref = ref_m.generate_neural_gaussians(torch.eye(4, device=device), torch.eye(4, device=device), cam.camera_center_, list(cam.t_),
                                          list(cam.R_quaternion_), vm)
    mine = generate_neural_gaussians(cam, model, vm)
    return model, ref_m, ref, mine, g


def _decode_backward_check(device, model, ref_m, ref, mine, g):
    import make_model_golden as mg
    import test_decode_gpu as td
    both, idx_r, idx_m = td._compare_forward(ref, mine)
    n_slots = ref[6].numel()
    G = torch.randn(n_slots, 14, generator=g).to(device)
    Gn = (torch.randn(n_slots, generator=g) * 0.1).to(device)
    slot = torch.nonzero(both).view(-1)

    def loss_of(out, idx):
        cols = torch.cat([out[0], out[1], out[2], out[3], out[4]], dim=1)
        return (cols[idx] * G[slot]).sum() + (out[5].view(-1) * Gn).sum()

    loss_of(ref, idx_r).backward()
    st = ref_m.state()
    g_ref = [t.grad for t in st[:4]] + [p.grad for p in ref_m.mlp_parameters()]
    params = [model._anchor, model._offset, model._anchor_feat, model._scaling] + model_ref.mlp_tensors(model)
    params = [[q for q in model.parameters() if q.data_ptr() == p.data_ptr()][0] for p in params]
    g_mine = torch.autograd.grad(loss_of(mine, idx_m), params, allow_unused=True)
    assert len(g_mine) == len(g_ref)
    for i, (a, b) in enumerate(zip(g_mine, g_ref)):
        assert a is not None and b is not None, i
        tol = 1e-4 * (b.abs() + b.abs().mean() + 1e-30)
        bad = ((a - b.view_as(a)).abs() > tol.view_as(a))
        assert bad.float().mean().item() <= 1e-4, (i, int(bad.sum()), float(((a - b.view_as(a)).abs() / tol.view_as(a)).max()))


def test_decode_matches_compiled_reference(device):
    torch.backends.cuda.matmul.allow_tf32 = False
    model, ref_m, ref, mine, g = _decode_pair(device, 4000, 3, 0.7)
    _decode_backward_check(device, model, ref_m, ref, mine, g)


@pytest.mark.slow
def test_decode_matches_compiled_reference_C3(device):
    """BASELINE config 3 size: 200k anchors x 10 offsets, forward AND backward."""
    torch.backends.cuda.matmul.allow_tf32 = False
    model, ref_m, ref, mine, g = _decode_pair(device, 200_000, 1003, None)
    _decode_backward_check(device, model, ref_m, ref, mine, g)
    frac = mine[6].float().mean().item()
    assert 0.2 < frac < 0.8, frac


def test_project2_image_matches_reference(device):
    """Row V3: RasterizeGaussiansprojectCUDA (rasterize_points.cu:278-362 -> rasterizer_impl.cu:494-585) of the reference
    itself against the product: pixel means and radii bit-exact, SH colours 1e-5."""
    mr = model_ref.load()
    scene = synth.sh_variant(synth.config("small"), 3)
    t = scene.to_torch(device)
    e = torch.empty(0, dtype=torch.float32, device=device)
    for use_sh in (True, False):
        sh = t["sh"] if use_sh else e
        colors = e if use_sh else t["colors"]
        deg = 3 if use_sh else 0
        args = (t["bg"], t["means3D"], colors, t["opacities"], t["scales"], t["rotations"], 1.0, e, t["viewmatrix"],
                t["projmatrix"], scene.tanfovx, scene.tanfovy, scene.H, scene.W, sh, deg, t["campos"], False)
        pts_r, radii_r, rgb_r = mr.RasterizeGaussiansprojectCUDA(*args)
        pts_m, radii_m, rgb_m = rp.RasterizeGaussiansprojectCUDA(*args)
        torch.cuda.synchronize()
        assert torch.equal(radii_m, radii_r)
        vis = radii_r > 0
        assert int(vis.sum()) > 100
        assert torch.equal(pts_m[vis], pts_r[vis])
        if use_sh:
            torch.testing.assert_close(rgb_m[vis], rgb_r[vis], rtol=1e-5, atol=1e-6)


def test_tensor_level_entry_points_of_the_reference(device):
    """distCUDA2, markVisible and RasterizeGaussiansfilterCUDA through the reference's own LibTorch functions."""
    mr = model_ref.load()
    scene = synth.config("small")
    t = scene.to_torch(device)
    e = torch.empty(0, dtype=torch.float32, device=device)
    pts = t["means3D"][:5000].contiguous()
    assert torch.equal(rp.distCUDA2(pts), mr.distCUDA2(pts))
    assert torch.equal(rp.markVisible(t["means3D"], t["viewmatrix"], t["projmatrix"]),
                       mr.markVisible(t["means3D"], t["viewmatrix"], t["projmatrix"]))
    a = (t["means3D"], t["scales"], t["rotations"], 1.0, e, t["viewmatrix"], t["projmatrix"], scene.tanfovx, scene.tanfovy,
         scene.H, scene.W, False)
    assert torch.equal(rp.RasterizeGaussiansfilterCUDA(*a, False), mr.RasterizeGaussiansfilterCUDA(*a, False))
