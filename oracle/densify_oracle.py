"""TEST INFRASTRUCTURE — PyTorch restatement of the reference's densification decisions, op for op:

    GaussianModel::adjust_anchor   /root/reference/src/gaussian_model.cpp:1705-1762
    GaussianModel::anchor_growing  :1556-1703   (use_chunk = true; the chunked duplicate test is an O(U x A) equality)
    GaussianModel::prune_anchor    :1505-1555   (incl. the clamp of _scaling[:, 3:] to <= 0.05 it applies on the way)

PINNED to the reference itself: on the GPU box tests/test_densify_gpu.py runs the reference's own functions (compiled
unmodified into oracle/_ref/_model_ref.so) on the same state with the same torch RNG seed and compares every tensor
exactly; tests/golden/densify_*.npz hold outputs of that run for the CPU suite (tests/golden/make_densify_golden.py).

State = dict of tensors under the reference's member names: _anchor [A,3], _offset [A,k,3], _anchor_feat [A,32],
_opacity [A,1], _scaling [A,6], _rotation [A,4], opacity_accum [A,1], anchor_demon [A,1], offset_gradient_accum [A*k,1],
offset_denom [A*k,1], and optionally the Adam moments exp_avg / exp_avg_sq of the six anchor tensors under
"m_<name>" / "v_<name>".  Only tests/ imports this file; the product never does.

Two CUDA-vs-CPU facts of ATen that matter for bit-exactness and are restated explicitly (the reference runs on CUDA):
  * `tensor / python_float` on CUDA multiplies by the FP32 reciprocal of the scalar (BinaryDivTrueKernel.cu: "if the
    second operand is a CPU scalar, compute a * reciprocal(b)"), on the CPU it divides.  `div_by_scalar` below follows
    the CUDA behaviour on any device.
  * torch::rand_like draws from the device generator; the caller passes the random tensors in (`rands`), so the same
    numbers can be fed to the reference (torch.manual_seed) and to the product.
"""
from __future__ import annotations

import math

import numpy as np
import torch

ANCHOR_TENSORS = ("_anchor", "_offset", "_anchor_feat", "_opacity", "_scaling", "_rotation")   # optimizer groups 0-5


def div_by_scalar(t: torch.Tensor, s: float) -> torch.Tensor:
    inv = np.float32(1.0 / float(np.float32(s)))        # high_prec_t(1.0) / scalar, then cast to opmath_t (float)
    return t * float(inv)


def _cat_moments(st, name, ext):
    for mv in ("m_", "v_"):
        if mv + name in st:
            st[mv + name] = torch.cat([st[mv + name], torch.zeros_like(ext)], 0)


def anchor_growing(st, grads, threshold, offset_mask, rands, *, n_offsets=10, update_depth=3, update_init_factor=16,
                   update_hierachy_factor=4, voxel_size=0.001, feat_dim=32):
    """gaussian_model.cpp:1556-1703.  rands[i]: the torch::rand_like draw of level i ([init_length] floats)."""
    dev = grads.device
    init_length = st["_anchor"].size(0) * n_offsets
    for i in range(update_depth):
        cur_threshold = np.float32(float(np.float32(threshold)) * math.pow(math.floor(update_hierachy_factor // 2), i))
        candidate_mask = grads >= float(cur_threshold)
        candidate_mask = torch.logical_and(candidate_mask, offset_mask)
        rand_mask = rands[i] > math.pow(0.5, i + 1)
        candidate_mask = torch.logical_and(candidate_mask, rand_mask)
        length_inc = st["_anchor"].size(0) * n_offsets - init_length
        if length_inc == 0:
            if i > 0:
                continue
        else:
            candidate_mask = torch.cat([candidate_mask, torch.zeros(length_inc, dtype=torch.bool, device=dev)], 0)
        scaling = torch.exp(st["_scaling"])
        all_xyz = st["_anchor"].unsqueeze(1) + st["_offset"] * scaling[:, :3].unsqueeze(1)
        size_factor = math.floor(update_init_factor / math.pow(update_hierachy_factor, i))
        cur_size = float(np.float32(voxel_size) * np.float32(size_factor))
        grid_coords = torch.round(div_by_scalar(st["_anchor"], cur_size)).to(torch.int32)
        selected_xyz = all_xyz.view(-1, 3)[candidate_mask]
        selected_grid_coords = torch.round(div_by_scalar(selected_xyz, cur_size)).to(torch.int32)
        uniq, inverse = torch.unique(selected_grid_coords, dim=0, sorted=True, return_inverse=True)
        # the chunked test of :1601-1616 is `any over anchors of (row == grid_coords[a]).all()`
        if uniq.size(0) > 0:
            dup = torch.zeros(uniq.size(0), dtype=torch.bool, device=dev)
            for j in range(0, grid_coords.size(0), 4096):
                dup |= (uniq.unsqueeze(1) == grid_coords[j:j + 4096]).all(-1).any(-1).view(-1)
        else:
            dup = torch.zeros(0, dtype=torch.bool, device=dev)
        keep = ~dup
        candidate_anchor = uniq[keep] * cur_size
        n = candidate_anchor.size(0)
        if n > 0:
            new_scaling = torch.log(torch.ones_like(candidate_anchor).repeat(1, 2).float() * cur_size)
            new_rotation = torch.zeros(n, 4, device=dev)
            new_rotation[:, 0] = 1.0
            new_opacities = torch.log((0.1 * torch.ones(n, 1, device=dev)) / (1 - 0.1 * torch.ones(n, 1, device=dev)))
            new_feat = st["_anchor_feat"].unsqueeze(1).repeat(1, n_offsets, 1).view(-1, feat_dim)[candidate_mask]
            out = torch.zeros(uniq.size(0), feat_dim, device=dev)
            out = out.scatter_reduce(0, inverse.unsqueeze(1).expand(-1, feat_dim), new_feat, "amax", include_self=False)
            new_feat = out[keep]
            new_offsets = torch.zeros_like(candidate_anchor).unsqueeze(1).repeat(1, n_offsets, 1).float()
            st["anchor_demon"] = torch.cat([st["anchor_demon"], torch.zeros(n, 1, device=dev)], 0)
            st["opacity_accum"] = torch.cat([st["opacity_accum"], torch.zeros(n, 1, device=dev)], 0)
            for name, ext in zip(ANCHOR_TENSORS, (candidate_anchor, new_offsets, new_feat, new_opacities, new_scaling, new_rotation)):
                _cat_moments(st, name, ext)
                st[name] = torch.cat([st[name], ext], 0)
    return st


def prune_anchor(st, mask):
    """gaussian_model.cpp:1505-1555."""
    valid = ~mask
    for name in ANCHOR_TENSORS:
        for mv in ("m_", "v_"):
            if mv + name in st:
                st[mv + name] = st[mv + name][valid].clone()
        p = st[name][valid]
        if name == "_scaling":
            p = p.clone()
            p[:, 3:] = torch.clamp(p[:, 3:], -float(np.finfo(np.float32).max), 0.05)
        st[name] = p
    return st


def adjust_anchor(st, rands, check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005, **model):
    """gaussian_model.cpp:1705-1762.  Mutates and returns `st`."""
    n_offsets = model.get("n_offsets", 10)
    dev = st["_anchor"].device
    grads = st["offset_gradient_accum"] / st["offset_denom"]
    grads[grads.isnan()] = 0.0
    grads_norm = torch.linalg.vector_norm(grads, dim=-1)
    offset_mask = (st["offset_denom"] > float(np.float32(check_interval) * np.float32(success_threshold)) * 0.5).squeeze(1)
    anchor_growing(st, grads_norm, grad_threshold, offset_mask, rands, **model)
    A = st["_anchor"].size(0)
    st["offset_denom"][offset_mask] = 0
    st["offset_denom"] = torch.cat([st["offset_denom"], torch.zeros(A * n_offsets - st["offset_denom"].size(0), 1, device=dev)], 0)
    st["offset_gradient_accum"][offset_mask] = 0
    st["offset_gradient_accum"] = torch.cat(
        [st["offset_gradient_accum"], torch.zeros(A * n_offsets - st["offset_gradient_accum"].size(0), 1, device=dev)], 0)
    prune_mask = (st["opacity_accum"] < min_opacity * st["anchor_demon"]).squeeze(1)
    anchors_mask = (st["anchor_demon"] > float(np.float32(check_interval) * np.float32(success_threshold))).squeeze(1)
    prune_mask = torch.logical_and(prune_mask, anchors_mask)
    st["offset_denom"] = st["offset_denom"].view(-1, n_offsets)[~prune_mask].view(-1, 1)
    st["offset_gradient_accum"] = st["offset_gradient_accum"].view(-1, n_offsets)[~prune_mask].view(-1, 1)
    if int(anchors_mask.sum()) > 0:
        st["opacity_accum"][anchors_mask] = 0.0
        st["anchor_demon"][anchors_mask] = 0.0
    st["opacity_accum"] = st["opacity_accum"][~prune_mask]
    st["anchor_demon"] = st["anchor_demon"][~prune_mask]
    if prune_mask.size(0) > 0:
        prune_anchor(st, prune_mask)
    return st
