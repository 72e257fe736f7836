"""GPU: keyframe image ingest (segs_ingest_image / segs_resize_bilinear through segs_slam_b200.keyframe_ingest) against
goldens written by the real OpenCV (tests/golden/ingest_*.npz) and, at the Replica / TUM image sizes, against the numpy
oracle that is itself pinned to OpenCV (tests/test_ingest_cpu.py): undistorted planar image and undistortion mask
bit-exact, pyramid levels within the tolerance OpenCV's own IPP and generic paths differ by."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ingest_oracle as io  # noqa: E402
from segs_slam_b200 import keyframe_ingest as ki  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ingest_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_ingest_matches_opencv_golden(path, device):
    g = np.load(path)
    H, W, _ = g["image"].shape
    sizes = [tuple(int(v) for v in g[f"size_{i}"]) for i in range(2)]
    ing = ki.KeyframeIngest(device, H, W, g["K"], g["dist"], pyramid_sizes=sizes)
    out = ing.ingest(g["image"])
    assert out.shape == (3, H, W) and np.array_equal(out.cpu().numpy(), g["undistorted"].transpose(2, 0, 1))
    assert np.array_equal(ing.undistort_mask().cpu().numpy(), g["mask"].transpose(2, 0, 1))
    for lvl, i in zip(ing.pyramid(out), range(2)):
        np.testing.assert_allclose(lvl.cpu().numpy(), g[f"resized_{i}"].transpose(2, 0, 1), rtol=0, atol=5e-6)
    # a device-resident source works too, and no distortion = a pure [H,W,C] -> [C,H,W] conversion (tensor_utils.h:40-69)
    assert torch.equal(ing.ingest(torch.from_numpy(g["image"]).to(device)), out)
    plain = ki.KeyframeIngest(device, H, W)
    assert np.array_equal(plain.ingest(g["image"]).cpu().numpy(), g["image"].transpose(2, 0, 1))


@pytest.mark.parametrize("H,W,fx,dist", [(680, 1200, 600.0, (0.05, -0.12, 0.001, -0.0007)), (480, 640, 517.3, (0.2624, -0.9531, -0.0054, 0.0026))])
def test_ingest_full_size_matches_oracle(device, H, W, fx, dist):
    rng = np.random.default_rng(H)
    img = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    K = np.array([[fx, 0, W / 2 - 0.5], [0, fx, H / 2 - 0.5], [0, 0, 1]])
    sizes = [(H // 2, W // 2), (int(H * 0.8), int(W * 0.8))]
    ing = ki.KeyframeIngest(device, H, W, K, dist, pyramid_sizes=sizes)
    out = ing.ingest(img)
    mx, my = ki.init_undistort_rectify_map(K, dist, K, W, H)
    ref = io.remap_bilinear(img, mx, my).transpose(2, 0, 1)
    assert np.array_equal(out.cpu().numpy(), ref)
    for lvl, (h, w) in zip(ing.pyramid(out), sizes):
        assert np.array_equal(lvl.cpu().numpy(), io.resize_bilinear(ref, h, w))       # same arithmetic as the oracle: exact
    with pytest.raises(RuntimeError):
        ing.ingest(img.astype(np.float64))
