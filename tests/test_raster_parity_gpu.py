"""GPU parity: product (C-ABI, sm_100a kernels) vs the UNMODIFIED reference CUDA rasterizer
(oracle/_ref) on identical inputs.

Contract (BASELINE.md §4 / north_star): radii, tiles_touched, tile keys, sorted order
(point_list), tile ranges, num_rendered and n_contrib bit-exact; rendered RGB and final_T
within 1e-5 relative (asserted bit-exact here, which is stronger); gradients within 1e-4
relative.
"""
import math

import numpy as np
import pytest
import torch

import common
import refimpl
from segs_slam_b200 import synth

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not refimpl.available(), reason="oracle/_ref/libsegs_ref.so not built")

RGB_RTOL = 1e-5     # north_star: rendered RGB / final_T within 1e-5 relative
GRAD_RTOL = 1e-4    # north_star: parameter gradients within 1e-4 relative


def _rot_cam(angle_y=0.35, angle_x=-0.2, t=(0.3, -0.2, 0.4)):
    cy, sy = math.cos(angle_y), math.sin(angle_y)
    cx, sx = math.cos(angle_x), math.sin(angle_x)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float32)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float32)
    return (Ry @ Rx).astype(np.float32), np.asarray(t, dtype=np.float32)


def _scenes():
    out = {}
    out["tiny"] = synth.config("tiny")
    out["small_bg"] = synth.config("small", bg=(0.2, 0.5, 0.9))
    # odd image size (partial tiles), rotated camera, big Gaussians -> long per-tile lists
    s = synth.synth(30_000, 203, 117, 160.0, 150.0, 21, bg=(1.0, 1.0, 1.0))
    s.scales *= 4.0
    R, t = _rot_cam()
    out["odd_rot"] = synth.with_camera(s, R, t)
    out["C1"] = synth.config("C1")
    return out


SCENES = None


def scenes():
    global SCENES
    if SCENES is None:
        SCENES = _scenes()
    return SCENES


def _compare_forward(scene, a, m, r, name):
    P, W, H = scene.P, scene.W, scene.H
    N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
    ms = common.mine_sections(m, P, W, H)
    rg = refimpl.parse_geom(r["geom"], P)
    rb = refimpl.parse_binning(r["binning"], r["R"])
    ri = refimpl.parse_image(r["img"], N, T)

    assert m["R"] == r["R"], f"{name}: num_rendered {m['R']} != {r['R']}"
    assert torch.equal(m["radii"], r["radii"]), f"{name}: radii differ"
    assert torch.equal(ms["tiles_touched"], rg["tiles_touched"]), f"{name}: tiles_touched differ"
    vis = r["radii"] > 0
    assert torch.equal(common.bits(ms["depths"])[vis], common.bits(rg["depths"])[vis]), f"{name}: depths"
    assert torch.equal(common.bits(ms["means2D"])[vis], common.bits(rg["means2D"])[vis]), f"{name}: means2D"
    assert torch.equal(common.bits(ms["conic_opacity"])[vis], common.bits(rg["conic_opacity"])[vis]), \
        f"{name}: conic_opacity"
    if a["cov3D_precomp"].numel() == 0:
        assert torch.equal(common.bits(ms["cov3D"])[vis], common.bits(rg["cov3D"])[vis]), f"{name}: cov3D"
    if a["colors"].numel() == 0:   # SH path: colours are floats -> tolerance
        torch.testing.assert_close(ms["rgb"][vis], rg["rgb"][vis], rtol=RGB_RTOL, atol=1e-6)

    # sorted order, tile keys, ranges
    assert torch.equal(ms["point_list"], rb["point_list"]), f"{name}: point_list (sorted order) differs"
    if m["R"]:
        my_keys = (ms["tile_ids"].long() << 32) | (common.bits(ms["depths"])[ms["point_list"].long()].long() & 0xFFFFFFFF)
        assert torch.equal(my_keys, rb["point_list_keys"]), f"{name}: tile|depth keys differ"
    assert torch.equal(ms["ranges"], ri["ranges"]), f"{name}: tile ranges differ"

    # blend
    assert torch.equal(ms["n_contrib"], ri["n_contrib"]), f"{name}: n_contrib differs"
    torch.testing.assert_close(ms["final_T"], ri["final_T"], rtol=RGB_RTOL, atol=0)
    torch.testing.assert_close(m["color"], r["color"], rtol=RGB_RTOL, atol=1e-7)
    # stronger than the contract: the blend arithmetic is reproduced operation by operation
    assert torch.equal(common.bits(ms["final_T"]), common.bits(ri["final_T"])), f"{name}: final_T not bit-exact"
    assert torch.equal(common.bits(m["color"]), common.bits(r["color"])), f"{name}: colour not bit-exact"


def _compare_backward(m, r, r2, name):
    bad = []
    for k, g in m["grads"].items():
        ok, ratio = common.grad_close(g, r["grads"][k], r2["grads"][k] if r2 else None, GRAD_RTOL)
        if not ok:
            bad.append((k, ratio))
    assert not bad, f"{name}: gradients outside {GRAD_RTOL} relative: {bad}"
    # gradients of non-rendered Gaussians are exactly zero
    inv = ~(r["radii"] > 0)
    for k in ("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dcolors", "dL_dopacity", "dL_dmeans2D"):
        if inv.any():
            assert m["grads"][k][inv].abs().max().item() == 0.0


@needs_ref
@pytest.mark.parametrize("name", ["tiny", "small_bg", "odd_rot", "C1"])
def test_forward_backward_parity(device, name):
    scene = scenes()[name]
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device)
    m = common.run_mine(a, t["dL_dout"])
    r = common.run_ref(a, t["dL_dout"])
    r2 = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    _compare_forward(scene, a, m, r, name)
    _compare_backward(m, r, r2, name)


@needs_ref
@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_sh_path_parity(device, degree):
    scene = synth.sh_variant(synth.config("small"), degree)
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device, use_sh=True)
    m = common.run_mine(a, t["dL_dout"])
    r = common.run_ref(a, t["dL_dout"])
    r2 = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    name = f"sh{degree}"
    P, W, H = scene.P, scene.W, scene.H
    assert m["R"] == r["R"]
    assert torch.equal(m["radii"], r["radii"])
    ms = common.mine_sections(m, P, W, H)
    rb = refimpl.parse_binning(r["binning"], r["R"])
    ri = refimpl.parse_image(r["img"], W * H, ((W + 15) // 16) * ((H + 15) // 16))
    assert torch.equal(ms["point_list"], rb["point_list"])
    assert torch.equal(ms["n_contrib"], ri["n_contrib"])
    torch.testing.assert_close(m["color"], r["color"], rtol=RGB_RTOL, atol=1e-6)
    _compare_backward(m, r, r2, name)


@needs_ref
def test_cov3d_precomp_parity(device):
    scene = synth.config("small")
    t = scene.to_torch(device)
    # take the reference's own cov3D as the precomputed input
    a0 = common.scene_args(t, scene, device)
    r0 = common.run_ref(a0)
    cov = refimpl.parse_geom(r0["geom"], scene.P)["cov3D"].clone()
    # the reference leaves cov3D of culled points uninitialised: give them a small isotropic one
    cov[~(r0["radii"] > 0)] = torch.tensor([0.01, 0.0, 0.0, 0.01, 0.0, 0.01], device=device)
    a = common.scene_args(t, scene, device, use_cov=cov)
    m = common.run_mine(a, t["dL_dout"])
    r = common.run_ref(a, t["dL_dout"])
    r2 = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    _compare_forward(scene, a, m, r, "cov3d")
    _compare_backward(m, r, r2, "cov3d")


def test_empty_and_culled(device):
    """P == 0 leaves a zero image (not even background, src/rasterize_points.cu:81); a scene that
    is entirely behind the near plane renders pure background with zero instances."""
    from segs_slam_b200 import rasterize_points as rp
    e = common.empty(device)
    bg = torch.tensor([0.1, 0.2, 0.3], device=device)
    eye = torch.eye(4, device=device)
    R, color, radii, g, b, i = rp.RasterizeGaussiansCUDA(bg, torch.zeros((0, 3), device=device), e, e, e, e, 1.0,
                                                        e, eye, eye, 1.0, 1.0, 32, 48, e, 0, torch.zeros(3, device=device), False)
    assert R == 0 and color.shape == (3, 32, 48) and color.abs().max().item() == 0.0 and radii.numel() == 0

    scene = synth.config("tiny", bg=(0.1, 0.2, 0.3))
    scene.means3D[:, 2] = -1.0
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device)
    m = common.run_mine(a, t["dL_dout"])
    assert m["R"] == 0 and m["radii"].abs().max().item() == 0
    expect = t["bg"].view(3, 1, 1).expand(3, scene.H, scene.W)
    assert torch.equal(m["color"], expect)
    for g_ in m["grads"].values():
        if g_.numel():
            assert g_.abs().max().item() == 0.0
    if refimpl.available():
        r = common.run_ref(a)
        assert torch.equal(m["color"], r["color"])


def test_means3d_shape_error(device):
    from segs_slam_b200 import rasterize_points as rp
    e = common.empty(device)
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        rp.RasterizeGaussiansCUDA(torch.zeros(3, device=device), torch.zeros((5, 4), device=device), e, e, e, e, 1.0,
                                  e, torch.eye(4, device=device), torch.eye(4, device=device), 1.0, 1.0, 16, 16, e, 0,
                                  torch.zeros(3, device=device), False)


def test_second_backward_is_identical(device):
    """The gradient accumulator is re-zeroed by the backward, so calling backward twice on the
    same forward state (retain_graph) gives the same result up to atomic ordering."""
    scene = synth.config("small")
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device)
    from segs_slam_b200 import rasterize_points as rp
    m = common.run_mine(a, t["dL_dout"])
    g2 = rp.RasterizeGaussiansBackwardCUDA(
        a["bg"], a["means3D"], m["radii"], a["colors"], a["scales"], a["rotations"], a["scale_modifier"],
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], t["dL_dout"],
        a["sh"], a["degree"], a["campos"], m["geom"], m["R"], m["binning"], m["img"])
    for (k, g1), gb in zip(m["grads"].items(), g2):
        ok, ratio = common.grad_close(gb, g1, None, GRAD_RTOL)
        assert ok, (k, ratio)


def test_autograd_wrapper(device):
    """GaussianRasterizer.forward / autograd backward order (src/gaussian_rasterizer.cpp:143-153)."""
    from segs_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer
    scene = synth.config("tiny")
    t = scene.to_torch(device)
    leaf = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "colors", "opacities", "scales", "rotations")}
    means2D = torch.zeros_like(leaf["means3D"], requires_grad=True)
    settings = GaussianRasterizationSettings(scene.H, scene.W, scene.tanfovx, scene.tanfovy, t["bg"], 1.0,
                                             t["viewmatrix"], t["projmatrix"], 0, t["campos"], False)
    rast = GaussianRasterizer(settings)
    e = common.empty(device)
    color, radii = rast(leaf["means3D"], means2D, leaf["opacities"], False, True, True, True, False, e,
                        leaf["colors"], leaf["scales"], leaf["rotations"], e)
    (color * t["dL_dout"]).sum().backward()
    a = common.scene_args(t, scene, device)
    m = common.run_mine(a, t["dL_dout"])
    pairs = [("means3D", "dL_dmeans3D"), ("colors", "dL_dcolors"), ("opacities", "dL_dopacity"),
             ("scales", "dL_dscales"), ("rotations", "dL_drotations")]
    for leaf_name, gname in pairs:
        ok, ratio = common.grad_close(leaf[leaf_name].grad, m["grads"][gname], None, GRAD_RTOL)
        assert ok, (leaf_name, ratio)
    ok, ratio = common.grad_close(means2D.grad, m["grads"]["dL_dmeans2D"], None, GRAD_RTOL)
    assert ok
    with pytest.raises(RuntimeError, match="excatly one of either SHs or precomputed colors"):
        rast(leaf["means3D"], means2D, leaf["opacities"], False, False, True, True, False, e, e,
             leaf["scales"], leaf["rotations"], e)
    with pytest.raises(RuntimeError, match="exactly one of either scale/rotation pair"):
        rast(leaf["means3D"], means2D, leaf["opacities"], False, True, True, True, True, e, leaf["colors"],
             leaf["scales"], leaf["rotations"], e)


@needs_ref
@pytest.mark.slow
def test_full_size_C2(device):
    """BASELINE config 2 (1M Gaussians @ 1200x680): bit-exact sort/tile outputs at full size."""
    scene = synth.config("C2")
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device)
    m = common.run_mine(a, t["dL_dout"])
    r = common.run_ref(a, t["dL_dout"])
    r2 = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    _compare_forward(scene, a, m, r, "C2")
    _compare_backward(m, r, r2, "C2")


@needs_ref
@pytest.mark.slow
def test_large_scene_C5(device):
    """BASELINE config 5 (5M Gaussians @ 1920x1080, ~81M instances, lists of ~10k per tile; 510
    super-tiles => the two-digit super-tile sort): bit-exact sort/tile outputs at stress size."""
    scene = synth.config("C5")
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device)
    m = common.run_mine(a, t["dL_dout"])
    r = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    _compare_forward(scene, a, m, r, "C5")
    r2 = common.run_ref(a, t["dL_dout"])
    torch.cuda.synchronize()
    _compare_backward(m, r, r2, "C5")
