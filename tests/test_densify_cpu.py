"""CPU: the densification oracle (oracle/densify_oracle.py) against outputs of the reference ITSELF:
tests/golden/densify_*.npz hold the state of the reference's GaussianModel before and after its own adjust_anchor
(src/gaussian_model.cpp:1505-1762, compiled unmodified, run on a B200 by tests/golden/make_densify_golden.py) and the
random numbers it drew.  Decisions (which anchors appear, in which order, which are pruned) must match exactly; values
that went through exp / log on the GPU (the new anchors' log-scaling and opacity) within 1 ulp of the CPU's libm."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import densify_cases as dc  # noqa: E402
import densify_oracle  # noqa: E402


@pytest.mark.parametrize("case", ["a257"])
def test_densify_oracle_matches_reference_golden(case):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"densify_{case}.npz"))
    st = {k[len("before/"):]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("before/")}
    after = {k[len("after/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("after/")}
    rands = [torch.from_numpy(g[f"rand/{i}"]) for i in range(dc.MODEL["update_depth"])]
    out = densify_oracle.adjust_anchor(st, rands, 100, 0.8, 0.0002, 0.005, **dc.MODEL)
    assert out["_anchor"].shape == after["_anchor"].shape          # same number of anchors grown and pruned
    for k, ref in after.items():
        a = out[k]
        assert tuple(a.shape) == tuple(ref.shape), k
        if k in ("_scaling", "_opacity"):                           # logf on the GPU vs the CPU's: 1 ulp
            np.testing.assert_allclose(a.numpy(), ref.numpy(), rtol=3e-7, atol=0, err_msg=k)
        else:
            assert torch.equal(a, ref), (k, int((a != ref).sum()))


def test_level_constants_follow_the_reference_arithmetic():
    """cur_threshold = threshold * floor(4 / 2)^i, rand threshold 0.5^(i+1), cur_size = voxel_size * floor(16 / 4^i)."""
    from segs_slam_b200 import densify
    got = [densify.level_constants(i, 0.0002, 16, 4, 0.001) for i in range(3)]
    f = np.float32
    assert [g[0] for g in got] == [float(f(0.0002)), float(f(float(f(0.0002)) * 2)), float(f(float(f(0.0002)) * 4))]
    assert [g[1] for g in got] == [0.5, 0.25, 0.125]
    assert [g[2] for g in got] == [float(f(0.001) * f(16)), float(f(0.001) * f(4)), float(f(0.001) * f(1))]
