"""GPU parity of the auxiliary entry points: anchor prefilter (visible_filter), markVisible,
debug projection and the anchor-init kNN (distCUDA2) against the reference (oracle/_ref)."""
import numpy as np
import pytest
import torch

import common
import refimpl
from segs_slam_b200 import rasterize_points as rp
from segs_slam_b200 import synth

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refimpl.available(), reason="oracle/_ref/libsegs_ref.so not built")


@needs_ref
@pytest.mark.parametrize("name", ["tiny", "small", "C1"])
def test_visible_filter_and_mark_visible(device, name):
    scene = synth.config(name)
    scene.means3D[::7, 2] *= -1.0          # some points behind the camera
    t = scene.to_torch(device)
    e = common.empty(device)
    mine = rp.RasterizeGaussiansfilterCUDA(t["means3D"], t["scales"], t["rotations"], 1.0, e, t["viewmatrix"],
                                           t["projmatrix"], scene.tanfovx, scene.tanfovy, scene.H, scene.W, False)
    ref = refimpl.visible_filter(t["means3D"], t["scales"], t["rotations"], 1.0, e, t["viewmatrix"],
                                 t["projmatrix"], scene.tanfovx, scene.tanfovy, scene.H, scene.W)
    assert torch.equal(mine, ref)
    # the prefilter agrees with the radii of a full forward
    a = common.scene_args(t, scene, device)
    m = common.run_mine(a)
    assert torch.equal(mine, m["radii"])
    pm = rp.markVisible(t["means3D"], t["viewmatrix"], t["projmatrix"])
    pr = refimpl.mark_visible(t["means3D"], t["viewmatrix"], t["projmatrix"])
    assert torch.equal(pm, pr)


def test_project_matches_forward_state(device):
    scene = synth.sh_variant(synth.config("small"), 2)
    t = scene.to_torch(device)
    a = common.scene_args(t, scene, device, use_sh=True)
    m = common.run_mine(a)
    ms = common.mine_sections(m, scene.P, scene.W, scene.H)
    pts, radii, rgb = rp.RasterizeGaussiansprojectCUDA(
        a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], 1.0, a["cov3D_precomp"],
        a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"], a["sh"], a["degree"],
        a["campos"], False)
    vis = m["radii"] > 0
    assert torch.equal(radii, m["radii"])
    assert torch.equal(pts[vis], ms["means2D"][vis])
    assert torch.equal(rgb[vis], ms["rgb"][vis])
    assert pts[~vis].abs().max().item() == 0.0


def _point_sets():
    rng = np.random.default_rng(3)
    out = {
        "uniform_50k": rng.uniform(-2, 2, (50_000, 3)),
        "clustered_30k": np.concatenate([rng.normal(c, 0.05, (10_000, 3)) for c in ((0, 0, 0), (1, 1, 1), (-2, 0.5, 3))]),
        "voxel_grid": np.unique(np.round(rng.uniform(0, 1, (40_000, 3)) / 0.05), axis=0) * 0.05,
        "line_1025": np.stack([np.linspace(0, 1, 1025), np.zeros(1025), np.zeros(1025)], 1),
        "few_5": rng.uniform(0, 1, (5, 3)),
        "positive_only": rng.uniform(3, 4, (3000, 3)),     # exercises the {0,0,0}-initialised bbox
    }
    return {k: v.astype(np.float32) for k, v in out.items()}


@needs_ref
@pytest.mark.parametrize("name", list(_point_sets()))
def test_knn_parity(device, name):
    pts = torch.from_numpy(_point_sets()[name]).to(device)
    mine = rp.distCUDA2(pts)
    ref = refimpl.knn(pts)
    torch.cuda.synchronize()
    # exact 3-NN with the same per-distance arithmetic -> bit-exact
    assert torch.equal(common.bits(mine), common.bits(ref)), \
        f"max rel diff {((mine - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()}"


def test_knn_bruteforce_small(device):
    rng = np.random.default_rng(9)
    p = rng.uniform(-1, 1, (3000, 3)).astype(np.float32)
    pts = torch.from_numpy(p).to(device)
    mine = rp.distCUDA2(pts).cpu().numpy()
    d2 = ((p[:, None, :].astype(np.float64) - p[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    want = np.sort(d2, axis=1)[:, :3].mean(1)
    np.testing.assert_allclose(mine, want, rtol=1e-5)
