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
