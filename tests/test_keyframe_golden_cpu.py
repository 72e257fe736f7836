"""CPU suite: the per-view constants (row T0: world_view_transform_, projection_matrix_, full_proj_transform_,
camera_center_) against the reference's own dump of a real sequence (check_colmap.md -> tests/golden/
keyframe_transforms.json).  The dump has 4 decimals, so inputs carry 5e-5 of rounding; the products are compared
at 2e-3 absolute (|proj| <= 2.3, |t| <= 5)."""
import json
import math
import os

import numpy as np

from segs_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KF = json.load(open(os.path.join(ROOT, "tests", "golden", "keyframe_transforms.json")))


def test_fixture_present():
    assert len(KF) >= 8


def test_projection_matrix_matches_reference_dump():
    for k in KF:
        P = synth.projection_matrix(0.01, 100.0, k["FoVx"], k["FoVy"])
        ref = np.asarray(k["projection_matrix_"], dtype=np.float64).reshape(4, 4)     # dumped transposed
        np.testing.assert_allclose(P.T, ref, atol=1.5e-4)


def test_transform_tensors_match_reference_dump():
    for k in KF:
        wvt_ref = np.asarray(k["world_view_transform_"], dtype=np.float64).reshape(4, 4)
        R, t = wvt_ref[:3, :3].T.astype(np.float32), wvt_ref[3, :3].astype(np.float32)
        wvt, full, campos = synth.camera_matrices(R, t, math.tan(k["FoVx"] / 2), math.tan(k["FoVy"] / 2))
        np.testing.assert_allclose(wvt, wvt_ref, atol=1e-6)
        np.testing.assert_allclose(full, np.asarray(k["full_proj_transform_"]).reshape(4, 4), atol=2e-3)
        # camera_center_ = inverse(world_view_transform_)[3, :3]; R in the dump is orthonormal only to 4 decimals
        np.testing.assert_allclose(campos, np.asarray(k["camera_center_"]), atol=2e-3)


def test_cpp_transform_tensors_match_reference_dump():
    """torch_shim/keyframe_transforms.h (C++ twin of GaussianKeyframe::computeTransformTensors) on the same dump."""
    import torch  # noqa: F401
    from segs_slam_b200 import _segs_torch as shim
    for k in KF:
        wvt_ref = np.asarray(k["world_view_transform_"], dtype=np.float64).reshape(4, 4)
        R, t = wvt_ref[:3, :3].T, wvt_ref[3, :3]
        wvt, proj, full, center = shim.computeTransformTensors([float(x) for x in R.reshape(-1)], [float(x) for x in t],
                                                               k["FoVx"], k["FoVy"])
        np.testing.assert_allclose(wvt.numpy(), wvt_ref, atol=1e-4)        # two inversions of a 4-decimal matrix
        np.testing.assert_allclose(proj.numpy(), np.asarray(k["projection_matrix_"]).reshape(4, 4), atol=1.5e-4)
        np.testing.assert_allclose(full.numpy(), np.asarray(k["full_proj_transform_"]).reshape(4, 4), atol=2e-3)
        np.testing.assert_allclose(center.numpy(), np.asarray(k["camera_center_"]), atol=2e-3)
    r = shim.quaternionToRotation(0.5, 0.5, 0.5, 0.5)                      # 120 degrees about (1,1,1): x -> y -> z
    np.testing.assert_allclose(np.asarray(r).reshape(3, 3), [[0, 0, 1], [1, 0, 0], [0, 1, 0]], atol=1e-7)
