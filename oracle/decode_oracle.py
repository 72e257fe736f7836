"""TEST INFRASTRUCTURE — oracle for the structure-enhanced anchor decode.

Restates GaussianRenderer::generate_neural_gaussians
(/root/reference/src/gaussian_renderer.cpp:214-334) and the module shapes built in
GaussianModel::GaussianModel (/root/reference/src/gaussian_model.cpp:60-98) with the same ATen
operations in the same order (index / cat / repeat / Linear / ReLU / Tanh / Sigmoid / Softmax /
normalize), FP32, TF32 off.

PINNED to the reference itself: /root/reference/src/gaussian_renderer.cpp and gaussian_model.cpp compile UNMODIFIED in
this image once declaration-only stand-ins replace the absent Eigen / Sophus / OpenCV / PCL / torch_scatter headers
(oracle/stub_include/, `make -C oracle modelref` -> oracle/_ref/_model_ref.so).  tests/golden/decode_*.npz hold the
outputs and gradients of the reference's own generate_neural_gaussians (tests/golden/make_model_golden.py);
tests/test_decode_cpu.py holds this file to them (bit-identical on the generating machine), and on the GPU box the CUDA
decode is compared with the compiled reference live.  Only tests/, __graft_entry__.smoke() and bench.py's reference
arm may import this file; the product never does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

FEAT_DIM = 32      # Model.feat_dim   (every cfg/gaussian_mapper/**/*.yaml)
N_OFFSETS = 10     # Model.n_offsets


@dataclass
class DecodeConfig:
    appearance_dim: int = 32
    use_feat_bank: bool = True
    add_opacity_dist: bool = False
    add_cov_dist: bool = False
    add_color_dist: bool = False


class DecodeModel(nn.Module):
    """The trainable state the decode reads: anchors + the five MLPs (gaussian_model.cpp:60-98)."""

    def __init__(self, cfg: DecodeConfig, A: int):
        super().__init__()
        self.cfg = cfg
        od, cd, kd = int(cfg.add_opacity_dist), int(cfg.add_cov_dist), int(cfg.add_color_dist)
        self.mlp_opacity = nn.Sequential(nn.Linear(FEAT_DIM + 3 + od, FEAT_DIM), nn.ReLU(True),
                                         nn.Linear(FEAT_DIM, N_OFFSETS), nn.Tanh())
        self.mlp_cov = nn.Sequential(nn.Linear(FEAT_DIM + 3 + cd, FEAT_DIM), nn.ReLU(True),
                                     nn.Linear(FEAT_DIM, 7 * N_OFFSETS))
        self.mlp_color = nn.Sequential(nn.Linear(FEAT_DIM + 3 + kd + cfg.appearance_dim, FEAT_DIM), nn.ReLU(True),
                                       nn.Linear(FEAT_DIM, 3 * N_OFFSETS), nn.Sigmoid())
        self.mlp_apperance = nn.Sequential(nn.Linear(7, cfg.appearance_dim)) if cfg.appearance_dim > 0 else None
        self.mlp_feature_bank = nn.Sequential(nn.Linear(3 + 1, FEAT_DIM), nn.ReLU(True), nn.Linear(FEAT_DIM, 3),
                                              nn.Softmax(dim=1)) if cfg.use_feat_bank else None
        self._anchor = nn.Parameter(torch.zeros(A, 3))
        self._offset = nn.Parameter(torch.zeros(A, N_OFFSETS, 3))
        self._anchor_feat = nn.Parameter(torch.zeros(A, FEAT_DIM))
        self._scaling = nn.Parameter(torch.zeros(A, 6))

    def get_scaling(self):                      # gaussian_model.cpp:186-189 (exp activation)
        return torch.exp(self._scaling)


def generate_neural_gaussians(pc: DecodeModel, camera_center, pose_t, pose_q_wxyz, visible_mask=None):
    """-> (xyz, color, opacity, scaling, rot, neural_opacity, mask), gaussian_renderer.cpp:214-334."""
    cfg = pc.cfg
    dev = pc._anchor.device
    if visible_mask is None:
        visible_mask = torch.ones(pc._anchor.size(0), dtype=torch.bool, device=dev)
    feat = pc._anchor_feat[visible_mask]
    anchor = pc._anchor[visible_mask]
    grid_offsets = pc._offset[visible_mask]
    grid_scaling = pc.get_scaling()[visible_mask]
    ob_view = anchor - camera_center
    ob_dist = ob_view.norm(dim=1, keepdim=True)
    ob_view = ob_view / ob_dist

    if cfg.use_feat_bank:                       # :236-249
        cat_view = torch.cat([ob_view, ob_dist], dim=1)
        bank_weight = pc.mlp_feature_bank(cat_view).unsqueeze(1)
        feat = feat.unsqueeze(-1)
        feat = (feat[:, ::4, :1].repeat(1, 4, 1) * bank_weight[:, :, :1]
                + feat[:, ::2, :1].repeat(1, 2, 1) * bank_weight[:, :, 1:2]
                + feat[:, ::1, :1] * bank_weight[:, :, 2:])
        feat = feat.squeeze(-1)

    cat_local_view = torch.cat([feat, ob_view, ob_dist], dim=1)
    cat_local_view_wodist = torch.cat([feat, ob_view], dim=1)
    appearance_feat = None
    if cfg.appearance_dim > 0:                  # :256-270
        pose = [float(pose_t[0]), float(pose_t[1]), float(pose_t[2]), float(pose_q_wxyz[0]), float(pose_q_wxyz[1]),
                float(pose_q_wxyz[2]), float(pose_q_wxyz[3])]
        ob_pose = torch.tensor(pose, dtype=torch.float32).view(7, 1).to(dev).transpose(0, 1)
        appearance_feat = pc.mlp_apperance(ob_pose.expand(cat_local_view.size(0), -1))

    neural_opacity = pc.mlp_opacity(cat_local_view if cfg.add_opacity_dist else cat_local_view_wodist)
    neural_opacity = neural_opacity.reshape(-1, 1)
    mask = (neural_opacity > 0.0).view(-1)
    opacity = neural_opacity[mask]

    base = cat_local_view if cfg.add_color_dist else cat_local_view_wodist
    color = pc.mlp_color(torch.cat([base, appearance_feat], dim=1) if cfg.appearance_dim > 0 else base)
    color = color.reshape(anchor.size(0) * N_OFFSETS, 3)
    scale_rot = pc.mlp_cov(cat_local_view if cfg.add_cov_dist else cat_local_view_wodist)
    scale_rot = scale_rot.reshape(anchor.size(0) * N_OFFSETS, 7)

    offsets = grid_offsets.view(-1, 3)
    concatenated = torch.cat([grid_scaling, anchor], dim=-1)
    n, c = concatenated.shape
    concatenated_repeated = concatenated.repeat(1, N_OFFSETS).view(n * N_OFFSETS, c)
    concatenated_all = torch.cat([concatenated_repeated, color, scale_rot, offsets], dim=-1)
    masked = concatenated_all[mask]
    scaling_repeat, repeat_anchor, color, scale_rot, offsets = masked.split([6, 3, 3, 7, 3], dim=-1)
    scaling = scaling_repeat[:, 3:] * torch.sigmoid(scale_rot[:, :3])
    rot = F.normalize(scale_rot[:, 3:7])
    offsets = offsets * scaling_repeat[:, :3]
    xyz = repeat_anchor + offsets
    return xyz, color, opacity, scaling, rot, neural_opacity, mask


def synth_model(A: int, W: int, H: int, fx: float, fy: float, seed: int, cfg: DecodeConfig | None = None,
                device="cpu") -> DecodeModel:
    """anchors placed like the C2 points, _anchor_feat ~ N(0, 0.1),
    _offset ~ U(-1, 1), _scaling = ln U(0.005, 0.03), MLPs = torch default Linear init under
    torch.manual_seed(0)."""
    import numpy as np
    cfg = cfg or DecodeConfig()
    torch.manual_seed(0)
    m = DecodeModel(cfg, A)
    rng = np.random.default_rng(seed)
    f32 = np.float32
    tanx, tany = W / (2.0 * fx), H / (2.0 * fy)
    z = rng.uniform(0.5, 6.0, A).astype(f32)
    xn = rng.uniform(-1.1, 1.1, A).astype(f32)
    yn = rng.uniform(-1.1, 1.1, A).astype(f32)
    anchors = np.stack([xn * f32(tanx) * z, yn * f32(tany) * z, z], axis=1).astype(f32)
    feat = rng.normal(0.0, 0.1, (A, FEAT_DIM)).astype(f32)
    off = rng.uniform(-1.0, 1.0, (A, N_OFFSETS, 3)).astype(f32)
    sc = np.log(rng.uniform(0.005, 0.03, (A, 6))).astype(f32)
    with torch.no_grad():
        m._anchor.copy_(torch.from_numpy(anchors))
        m._anchor_feat.copy_(torch.from_numpy(feat))
        m._offset.copy_(torch.from_numpy(off))
        m._scaling.copy_(torch.from_numpy(sc))
    return m.to(device)
