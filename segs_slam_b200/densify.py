"""Densification of the anchor model: GaussianModel::adjust_anchor / anchor_growing / prune_anchor of the reference
(/root/reference/src/gaussian_model.cpp:1505-1762) on this library's kernels (csrc/densify.cu through the C ABI:
segs_anchor_growing_level, segs_prune_plan, segs_compact_rows).

Every decision (candidate selection, voxel snapping, de-duplication against the existing anchors, per-voxel feature
maximum, prune mask) and every gathered value is computed by the kernels; torch is used for what the reference uses
its tensor library for as well — owning and resizing the tensors (cat / new allocations), the constant fills of the new
anchors' scaling / opacity / rotation, and the random draw (`torch.rand`, the same generator call sequence as the
reference's torch::rand_like, so a shared seed gives every replica — and the reference — the same numbers).

`st` is a dict of CUDA tensors under the reference's member names: _anchor [A,3], _offset [A,k,3], _anchor_feat [A,F],
_opacity [A,1], _scaling [A,6] (log), _rotation [A,4], opacity_accum [A,1], anchor_demon [A,1],
offset_gradient_accum [A*k,1], offset_denom [A*k,1], plus optional Adam moments "m_<name>" / "v_<name>"."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .rasterize_points import _stream

ANCHOR_TENSORS = ("_anchor", "_offset", "_anchor_feat", "_opacity", "_scaling", "_rotation")   # optimizer groups 0-5
_F = np.float32


class _Blocks:
    """Allocation callback that hands out a NEW device block per call and keeps them alive (segs_anchor_growing_level)."""

    def __init__(self, device, dtype=torch.uint8):
        self.device, self.blocks = device, []
        self.cb = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, _user, nbytes):
        t = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=self.device)
        self.blocks.append(t)
        return t.data_ptr()

    def done(self):
        self.cb = None
        return self.blocks


def level_constants(i: int, grad_threshold: float, update_init_factor: int, update_hierachy_factor: int, voxel_size: float):
    """(cur_threshold, rand_threshold, cur_size) of growing level i in the reference's FP32/double host arithmetic
    (gaussian_model.cpp:1561, :1566, :1585-1586)."""
    cur_threshold = float(_F(float(_F(grad_threshold)) * math.pow(math.floor(update_hierachy_factor // 2), i)))
    rand_threshold = float(_F(math.pow(0.5, i + 1)))
    size_factor = math.floor(update_init_factor / math.pow(update_hierachy_factor, i))
    cur_size = float(_F(voxel_size) * _F(size_factor))
    return cur_threshold, rand_threshold, cur_size


def _check(st):
    for k in ANCHOR_TENSORS + ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom"):
        t = st[k]
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError(f"densify: '{k}' must be an FP32 CUDA tensor (there is no CPU path)")


def anchor_growing(st, grad_threshold: float, denom_threshold: float, rands=None, generator=None, *, n_offsets=10,
                   update_depth=3, update_init_factor=16, update_hierachy_factor=4, voxel_size=0.001):
    """GaussianModel::anchor_growing (:1556-1703) for all levels; mutates and returns `st`.  rands: optional list of
    `update_depth` tensors [A*k]; default: torch.rand per level, drawn (like the reference) even for a skipped level."""
    lib = _lib.load()
    dev = st["_anchor"].device
    A0 = st["_anchor"].size(0)
    feat_dim = st["_anchor_feat"].size(1)
    init_slots = A0 * n_offsets
    accum = st["offset_gradient_accum"].contiguous()
    denom = st["offset_denom"].contiguous()
    report = []
    for i in range(update_depth):
        if rands is not None:
            rnd = rands[i].to(device=dev, dtype=torch.float32).contiguous()
        else:
            rnd = torch.rand(init_slots, device=dev, dtype=torch.float32, generator=generator)
        A_now = st["_anchor"].size(0)
        if A_now * n_offsets - init_slots == 0 and i > 0:          # :1570-1575: nothing was added so far -> level skipped
            report.append((0, 0))
            continue
        cur_threshold, rand_threshold, cur_size = level_constants(i, grad_threshold, update_init_factor, update_hierachy_factor,
                                                                  voxel_size)
        anchor, offset = st["_anchor"].contiguous(), st["_offset"].contiguous()
        scaling, feat = st["_scaling"].contiguous(), st["_anchor_feat"].contiguous()
        scratch, out = _Blocks(dev), _Blocks(dev)
        p_anchor, p_feat = C.c_void_p(), C.c_void_p()
        n_cand, n_new = C.c_int(0), C.c_int(0)
        with torch.cuda.device(dev):
            _lib.check(lib.segs_anchor_growing_level(
                A_now, init_slots, n_offsets, feat_dim, anchor.data_ptr(), offset.data_ptr(), scaling.data_ptr(), feat.data_ptr(),
                accum.data_ptr(), denom.data_ptr(), rnd.data_ptr(), float(denom_threshold), cur_threshold, rand_threshold,
                cur_size, scratch.cb, None, out.cb, None, C.byref(p_anchor), C.byref(p_feat), C.byref(n_cand), C.byref(n_new),
                _stream()))
        scratch.done()
        blocks = out.done()
        n = int(n_new.value)
        report.append((int(n_cand.value), n))
        if n == 0:
            continue
        new = blocks[0].view(torch.float32)
        candidate_anchor = new[:3 * n].view(n, 3)
        new_feat = new[3 * n:3 * n + feat_dim * n].view(n, feat_dim)
        f32 = dict(dtype=torch.float32, device=dev)
        new_scaling = torch.log(torch.full((n, 6), cur_size, **f32))                        # :1625-1626
        new_rotation = torch.zeros((n, 4), **f32)
        new_rotation[:, 0] = 1.0                                                            # :1627-1628
        tenth = 0.1 * torch.ones((n, 1), **f32)
        new_opacities = torch.log(tenth / (1 - tenth))                                      # :1630-1631 inverse_sigmoid(0.1)
        new_offsets = torch.zeros((n, n_offsets, 3), **f32)                                 # :1639
        st["anchor_demon"] = torch.cat([st["anchor_demon"], torch.zeros((n, 1), **f32)], 0)
        st["opacity_accum"] = torch.cat([st["opacity_accum"], torch.zeros((n, 1), **f32)], 0)
        for name, ext in zip(ANCHOR_TENSORS, (candidate_anchor, new_offsets, new_feat, new_opacities, new_scaling, new_rotation)):
            for mv in ("m_", "v_"):
                if mv + name in st:
                    st[mv + name] = torch.cat([st[mv + name], torch.zeros_like(ext)], 0)
            st[name] = torch.cat([st[name], ext], 0)
    st["_growing_report"] = report
    return st


def adjust_anchor(st, check_interval: int = 100, success_threshold: float = 0.8, grad_threshold: float = 0.0002,
                  min_opacity: float = 0.005, rands=None, generator=None, **model):
    """GaussianModel::adjust_anchor (:1705-1762).  Mutates and returns `st` (new tensors for everything that changed
    size).  `model`: n_offsets, update_depth, update_init_factor, update_hierachy_factor, voxel_size."""
    _check(st)
    lib = _lib.load()
    n_offsets = int(model.get("n_offsets", 10))
    dev = st["_anchor"].device
    A0 = st["_anchor"].size(0)
    init_slots = A0 * n_offsets
    anchor_threshold = float(_F(check_interval) * _F(success_threshold))                   # int * float -> float (:1742)
    denom_threshold = float(_F(anchor_threshold * 0.5))                                     # ... * 0.5 (double), compared in FP32 (:1713)
    st["offset_gradient_accum"] = st["offset_gradient_accum"].contiguous()
    st["offset_denom"] = st["offset_denom"].contiguous()
    anchor_growing(st, grad_threshold, denom_threshold, rands, generator, **model)
    A = st["_anchor"].size(0)
    f32 = dict(dtype=torch.float32, device=dev)
    if A > A0:                                                                               # :1717-1727 padding
        pad = torch.zeros(((A - A0) * n_offsets, 1), **f32)
        st["offset_denom"] = torch.cat([st["offset_denom"], pad], 0)
        st["offset_gradient_accum"] = torch.cat([st["offset_gradient_accum"], pad], 0)
    st["opacity_accum"] = st["opacity_accum"].contiguous()
    st["anchor_demon"] = st["anchor_demon"].contiguous()
    keep = torch.empty(A, dtype=torch.int32, device=dev)
    keep_index = torch.empty(A, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.segs_prune_scratch_words(A)), dtype=torch.int32, device=dev)
    n_keep = C.c_int(0)
    with torch.cuda.device(dev):
        _lib.check(lib.segs_prune_plan(A, init_slots, st["opacity_accum"].data_ptr(), st["anchor_demon"].data_ptr(),
                                       st["offset_gradient_accum"].data_ptr(), st["offset_denom"].data_ptr(), denom_threshold,
                                       anchor_threshold, float(min_opacity), keep.data_ptr(), keep_index.data_ptr(),
                                       scratch.data_ptr(), C.byref(n_keep), _stream()))
        n = int(n_keep.value)

        def compact(t, row_floats, clamp_from=-1, clamp_max=0.0):
            src = t.contiguous()
            dst = torch.empty((n, row_floats), **f32)
            if n == 0:
                return dst
            _lib.check(lib.segs_compact_rows(A, row_floats, keep.data_ptr(), keep_index.data_ptr(), src.data_ptr(), dst.data_ptr(),
                                             clamp_from, clamp_max, _stream()))
            return dst

        if True:                                 # prune_anchor runs whenever the mask is non-empty (:1757-1760): the clamp too
            st["offset_denom"] = compact(st["offset_denom"], n_offsets).view(-1, 1)
            st["offset_gradient_accum"] = compact(st["offset_gradient_accum"], n_offsets).view(-1, 1)
            st["opacity_accum"] = compact(st["opacity_accum"], 1)
            st["anchor_demon"] = compact(st["anchor_demon"], 1)
            for name in ANCHOR_TENSORS:
                shape = st[name].shape[1:]
                rf = int(np.prod(shape))
                for mv in ("m_", "v_"):
                    if mv + name in st:
                        st[mv + name] = compact(st[mv + name], rf).view(n, *shape)
                if name == "_scaling":
                    st[name] = compact(st[name], rf, 3, 0.05).view(n, *shape)               # clamp(_scaling[:, 3:], max = 0.05)
                else:
                    st[name] = compact(st[name], rf).view(n, *shape)
    st["_prune_report"] = (A, n)
    return st
