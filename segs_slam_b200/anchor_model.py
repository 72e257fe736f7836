"""Minimal container for the trainable state the anchor decode reads, under the member names of the
reference's GaussianModel (/root/reference/include/gaussian_model.h, src/gaussian_model.cpp:60-98,
186-230): `_anchor`, `_offset`, `_anchor_feat`, `_scaling`, `_rotation`, the five MLPs, and the
configuration flags.  It exists so that the batched mapper (mapper.py), bench.py and the examples can
run the full per-view pipeline (prefilter -> decode -> rasterize) without the rest of SEGS-SLAM;
densification, checkpoints and the optimizer schedule are out of scope.

`synth_anchor_model` builds BASELINE.md §3 config C3; `circle_keyframes` the C4 poses.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import synth

FEAT_DIM = 32
N_OFFSETS = 10


class AnchorModel(nn.Module):
    def __init__(self, A: int, appearance_dim: int = 32, use_feat_bank: bool = True, add_opacity_dist: bool = False,
                 add_cov_dist: bool = False, add_color_dist: bool = False):
        super().__init__()
        self.feat_dim, self.n_offsets = FEAT_DIM, N_OFFSETS
        # densification parameters, defaults of GaussianModelParams (include/gaussian_parameters.h:35-41)
        self.voxel_size, self.update_depth, self.update_init_factor, self.update_hierachy_factor = 0.001, 3, 16, 4
        self.appearance_dim, self.use_feat_bank = appearance_dim, use_feat_bank
        self.add_opacity_dist, self.add_cov_dist, self.add_color_dist = add_opacity_dist, add_cov_dist, add_color_dist
        od, cd, kd = int(add_opacity_dist), int(add_cov_dist), int(add_color_dist)
        self.mlp_opacity = nn.Sequential(nn.Linear(FEAT_DIM + 3 + od, FEAT_DIM), nn.ReLU(True),
                                         nn.Linear(FEAT_DIM, N_OFFSETS), nn.Tanh())
        self.mlp_cov = nn.Sequential(nn.Linear(FEAT_DIM + 3 + cd, FEAT_DIM), nn.ReLU(True),
                                     nn.Linear(FEAT_DIM, 7 * N_OFFSETS))
        self.mlp_color = nn.Sequential(nn.Linear(FEAT_DIM + 3 + kd + appearance_dim, FEAT_DIM), nn.ReLU(True),
                                       nn.Linear(FEAT_DIM, 3 * N_OFFSETS), nn.Sigmoid())
        self.mlp_apperance = nn.Sequential(nn.Linear(7, appearance_dim)) if appearance_dim > 0 else None
        self.mlp_feature_bank = nn.Sequential(nn.Linear(4, FEAT_DIM), nn.ReLU(True), nn.Linear(FEAT_DIM, 3),
                                              nn.Softmax(dim=1)) if use_feat_bank else None
        self._anchor = nn.Parameter(torch.zeros(A, 3))
        self._offset = nn.Parameter(torch.zeros(A, N_OFFSETS, 3))
        self._anchor_feat = nn.Parameter(torch.zeros(A, FEAT_DIM))
        self._scaling = nn.Parameter(torch.zeros(A, 6))
        self._rotation = nn.Parameter(torch.zeros(A, 4), requires_grad=False)     # gaussian_model.cpp:372
        self._opacity = nn.Parameter(torch.zeros(A, 1), requires_grad=False)      # :373 (carried by the checkpoint only)

    def get_anchor(self):
        return self._anchor

    def get_scaling(self):          # exp activation, gaussian_model.cpp:186-189
        return torch.exp(self._scaling)

    def get_rotation(self):         # normalize activation, gaussian_model.cpp:213
        return torch.nn.functional.normalize(self._rotation)


def synth_anchor_model(A: int, W: int, H: int, fx: float, fy: float, seed: int, device="cuda", **cfg) -> AnchorModel:
    """C3 of BASELINE.md §3: anchors placed like the C2 points, `_anchor_feat ~ N(0, 0.1)`,
    `_offset ~ U(-1, 1)`, `_scaling = ln U(0.005, 0.03)`, `_rotation = (1,0,0,0)`, MLPs = torch's default
    Linear init under torch.manual_seed(0)."""
    torch.manual_seed(0)
    m = AnchorModel(A, **cfg)
    rng = np.random.default_rng(seed)
    f32 = np.float32
    tanx, tany = W / (2.0 * fx), H / (2.0 * fy)
    z = rng.uniform(0.5, 6.0, A).astype(f32)
    xn = rng.uniform(-1.1, 1.1, A).astype(f32)
    yn = rng.uniform(-1.1, 1.1, A).astype(f32)
    anchors = np.stack([xn * f32(tanx) * z, yn * f32(tany) * z, z], axis=1).astype(f32)
    with torch.no_grad():
        m._anchor.copy_(torch.from_numpy(anchors))
        m._anchor_feat.copy_(torch.from_numpy(rng.normal(0.0, 0.1, (A, FEAT_DIM)).astype(f32)))
        m._offset.copy_(torch.from_numpy(rng.uniform(-1.0, 1.0, (A, N_OFFSETS, 3)).astype(f32)))
        m._scaling.copy_(torch.from_numpy(np.log(rng.uniform(0.005, 0.03, (A, 6))).astype(f32)))
        m._rotation[:, 0] = 1.0
    return m.to(device)


class Keyframe:
    """The GaussianKeyframe members the renderer reads (gaussian_keyframe.cpp:151-184): world->camera
    pose (R, t), the transposed view / full projection matrices, the camera centre."""

    def __init__(self, R: np.ndarray, t: np.ndarray, tanfovx: float, tanfovy: float, device):
        wvt, full, campos = synth.camera_matrices(R.astype(np.float32), t.astype(np.float32), tanfovx, tanfovy)
        self.world_view_transform_ = torch.from_numpy(wvt).to(device)
        self.full_proj_transform_ = torch.from_numpy(full).to(device)
        self.camera_center_ = torch.from_numpy(campos).to(device)
        self.t_ = tuple(float(x) for x in t)
        self.R_quaternion_ = _quat_wxyz(R)


def _quat_wxyz(R: np.ndarray):
    tr = float(R[0, 0] + R[1, 1] + R[2, 2])
    w = math.sqrt(max(0.0, 1.0 + tr)) / 2.0
    if w < 1e-6:
        return (0.0, 1.0, 0.0, 0.0)
    return (w, float(R[2, 1] - R[1, 2]) / (4 * w), float(R[0, 2] - R[2, 0]) / (4 * w), float(R[1, 0] - R[0, 1]) / (4 * w))


def circle_keyframes(n: int, radius: float, centroid, tanfovx: float, tanfovy: float, device, seed: int = 1004):
    """C4: n poses on a circle of `radius` metres around the scene centroid, each looking at it."""
    rng = np.random.default_rng(seed)
    c = np.asarray(centroid, dtype=np.float64)
    out = []
    for i in range(n):
        ang = 2.0 * math.pi * i / n + rng.uniform(-0.02, 0.02)
        eye = c + np.array([radius * math.sin(ang), 0.15 * math.sin(3 * ang), -radius * math.cos(ang)])
        fwd = c - eye
        fwd /= np.linalg.norm(fwd)
        right = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
        right /= np.linalg.norm(right)
        up = np.cross(fwd, right)
        R = np.stack([right, up, fwd], axis=0)              # world -> camera rotation (x right, y down-ish, z forward)
        t = -R @ eye
        out.append(Keyframe(R, t, tanfovx, tanfovy, device))
    return out
