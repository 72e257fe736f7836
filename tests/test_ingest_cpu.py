"""CPU: the keyframe-ingest oracle (oracle/ingest_oracle.py) and the host-side map computation
(segs_slam_b200/keyframe_ingest.py:init_undistort_rectify_map) against the REAL OpenCV: tests/golden/ingest_*.npz were
written with cv2 4.13 (tests/golden/make_ingest_golden.py: cv2.initUndistortRectifyMap, cv2.remap, cv2.resize — the calls of
include/camera.h:70-115 and src/gaussian_mapper.cpp:621-632); where cv2 is importable the comparison is repeated live."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ingest_oracle as io  # noqa: E402
from segs_slam_b200.keyframe_ingest import init_undistort_rectify_map  # noqa: E402

GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ingest_*.npz")))
# cv2.resize of the pip wheel runs Intel IPP's linear kernel, whose coefficients differ from OpenCV's generic C++ path
# (restated by the oracle) by a few 1e-6 for non-integer ratios; exact 2x ratios agree to half an ulp.
RESIZE_ATOL = 5e-6


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_ingest_oracle_matches_opencv_golden(path):
    g = np.load(path)
    H, W, _ = g["image"].shape
    mx, my = init_undistort_rectify_map(g["K"], g["dist"], g["K"], W, H)
    assert np.array_equal(mx, g["map_x"]) and np.array_equal(my, g["map_y"])            # bit-exact maps
    und = io.remap_bilinear(g["image"], g["map_x"], g["map_y"])
    assert np.array_equal(und, g["undistorted"])                                        # bit-exact remap
    assert np.array_equal(io.remap_bilinear(np.ones_like(g["image"]), g["map_x"], g["map_y"]), g["mask"])
    for i in range(2):
        h, w = (int(v) for v in g[f"size_{i}"])
        r = io.resize_bilinear(g["undistorted"].transpose(2, 0, 1), h, w)
        np.testing.assert_allclose(r, g[f"resized_{i}"].transpose(2, 0, 1), rtol=0, atol=RESIZE_ATOL)
    assert len(GOLD) >= 2


def test_ingest_oracle_matches_opencv_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    H, W = 75, 110
    img = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    K = np.array([[88.0, 0, 54.2], [0, 87.5, 37.9], [0, 0, 1]])
    dist = np.array([0.11, -0.31, 0.002, -0.001])
    m1, m2 = cv2.initUndistortRectifyMap(K, dist, np.eye(3), K, (W, H), cv2.CV_32F)
    mx, my = init_undistort_rectify_map(K, dist, K, W, H)
    assert np.array_equal(mx, m1) and np.array_equal(my, m2)
    assert np.array_equal(io.remap_bilinear(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    for h, w in ((37, 55), (50, 73), (75, 110)):
        ref = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR).transpose(2, 0, 1)
        np.testing.assert_allclose(io.resize_bilinear(img.transpose(2, 0, 1), h, w), ref, rtol=0, atol=RESIZE_ATOL)
