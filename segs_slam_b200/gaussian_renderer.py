"""Anchor decode, mirroring GaussianRenderer::generate_neural_gaussians of the reference
(/root/reference/src/gaussian_renderer.cpp:214-334) on top of the C-ABI's fused kernels
(segs_decode_forward / segs_decode_backward, segs_slam_b200/csrc/decode.cu).

`pc` is any object that carries the reference GaussianModel's members under the same names
(`_anchor_feat`, `_offset`, `get_anchor()`, `get_scaling()`, `mlp_opacity`, `mlp_cov`, `mlp_color`,
`mlp_apperance`, `mlp_feature_bank`, `use_feat_bank`, `appearance_dim`, `add_opacity_dist`,
`add_cov_dist`, `add_color_dist`); `viewpoint_camera` carries `camera_center_`, `t_`,
`R_quaternion_` (w, x, y, z) like GaussianKeyframe.  torch is used for device memory, the current
stream and autograd bookkeeping only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .rasterize_points import _Grower, _ptr, _stream

N_OFFSETS = 10
FEAT_DIM = 32


def _linear_layers(seq):
    return [m for m in seq if isinstance(m, torch.nn.Linear)]


def _weights(pc):
    """18 tensors in the order of segs_decode_params (absent ones are None)."""
    o, c, k = _linear_layers(pc.mlp_opacity), _linear_layers(pc.mlp_cov), _linear_layers(pc.mlp_color)
    w = [o[0].weight, o[0].bias, o[1].weight, o[1].bias, c[0].weight, c[0].bias, c[1].weight, c[1].bias,
         k[0].weight, k[0].bias, k[1].weight, k[1].bias]
    if getattr(pc, "appearance_dim", 0) > 0:
        a = _linear_layers(pc.mlp_apperance)
        w += [a[0].weight, a[0].bias]
    else:
        w += [None, None]
    if getattr(pc, "use_feat_bank", False):
        b = _linear_layers(pc.mlp_feature_bank)
        w += [b[0].weight, b[0].bias, b[1].weight, b[1].bias]
    else:
        w += [None, None, None, None]
    return w


class _DecodeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, visible_mask, camera_center, pose, anchor, anchor_feat, offset, scaling, *weights):
        lib = _lib.load()
        dev = anchor.device
        A = anchor.size(0)
        if anchor_feat.size(1) != FEAT_DIM or offset.size(1) != N_OFFSETS:
            raise RuntimeError("decode: feat_dim must be 32 and n_offsets 10 (every shipped SEGS-SLAM config)")
        keep = [t.contiguous() if t is not None else None for t in (anchor, anchor_feat, offset, scaling, *weights)]
        anchor_c, feat_c, offset_c, scaling_c = keep[:4]
        w_c = keep[4:]
        vm = visible_mask.contiguous() if visible_mask is not None else None
        cam = camera_center.to(device=dev, dtype=torch.float32).contiguous()
        params = _lib.DecodeParams(*[_ptr(t) for t in w_c], *cfg)
        cap = A * N_OFFSETS
        f32 = dict(dtype=torch.float32, device=dev)
        xyz, color = torch.empty((cap, 3), **f32), torch.empty((cap, 3), **f32)
        opacity, out_scaling = torch.empty((cap, 1), **f32), torch.empty((cap, 3), **f32)
        rot = torch.empty((cap, 4), **f32)
        neural_opacity = torch.empty((cap, 1), **f32)
        mask = torch.empty((cap,), dtype=torch.bool, device=dev)
        state = torch.empty((int(lib.segs_decode_state_bytes(A)),), dtype=torch.uint8, device=dev)
        counts = (C.c_int * 2)()
        pose_c = (C.c_float * 7)(*[float(v) for v in pose])
        with torch.cuda.device(dev):
            _lib.check(lib.segs_decode_forward(
                A, _ptr(vm), _ptr(anchor_c), _ptr(feat_c), _ptr(offset_c), _ptr(scaling_c), _ptr(cam), pose_c,
                C.byref(params), _ptr(xyz), _ptr(color), _ptr(opacity), _ptr(out_scaling), _ptr(rot),
                _ptr(neural_opacity), mask.data_ptr(), state.data_ptr(), counts, _stream()))
        n_vis, n_out = int(counts[0]), int(counts[1])
        ctx.cfg, ctx.pose, ctx.n_vis, ctx.n_out = cfg, tuple(float(v) for v in pose), n_vis, n_out
        ctx.has = [t is not None for t in weights]
        ctx.save_for_backward(vm if vm is not None else torch.empty(0, device=dev), cam, anchor_c, feat_c, offset_c,
                              scaling_c, state, *[t for t in w_c if t is not None])
        outs = (xyz[:n_out], color[:n_out], opacity[:n_out], out_scaling[:n_out], rot[:n_out],
                neural_opacity[:n_vis * N_OFFSETS], mask[:n_vis * N_OFFSETS])
        ctx.mark_non_differentiable(outs[6])
        return outs

    @staticmethod
    def backward(ctx, g_xyz, g_color, g_opacity, g_scaling, g_rot, g_nop, _g_mask):
        lib = _lib.load()
        saved = ctx.saved_tensors
        vm, cam, anchor, feat, offset, scaling, state = saved[:7]
        w_it = iter(saved[7:])
        weights = [next(w_it) if h else None for h in ctx.has]
        dev = anchor.device
        A = anchor.size(0)
        f32 = dict(dtype=torch.float32, device=dev)
        zero_rows = lambda g, w: (g.contiguous() if g is not None else torch.zeros((ctx.n_out, w), **f32))
        g_xyz, g_color, g_opacity = zero_rows(g_xyz, 3), zero_rows(g_color, 3), zero_rows(g_opacity, 1)
        g_scaling, g_rot = zero_rows(g_scaling, 3), zero_rows(g_rot, 4)
        g_nop = g_nop.contiguous() if g_nop is not None else None
        d_anchor, d_feat = torch.empty((A, 3), **f32), torch.empty((A, FEAT_DIM), **f32)
        d_offset, d_scaling = torch.empty((A, N_OFFSETS, 3), **f32), torch.empty((A, 6), **f32)
        d_w = [torch.empty_like(t) if t is not None else None for t in weights]
        params = _lib.DecodeParams(*[_ptr(t) for t in weights], *ctx.cfg)
        grads = _lib.DecodeGrads(*[_ptr(t) for t in d_w])
        scratch = _Grower(dev)
        pose_c = (C.c_float * 7)(*ctx.pose)
        with torch.cuda.device(dev):
            _lib.check(lib.segs_decode_backward(
                A, _ptr(vm), _ptr(anchor), _ptr(feat), _ptr(offset), _ptr(scaling), _ptr(cam), pose_c, C.byref(params),
                state.data_ptr(), ctx.n_vis, ctx.n_out, _ptr(g_xyz), _ptr(g_color), _ptr(g_opacity), _ptr(g_scaling),
                _ptr(g_rot), _ptr(g_nop), _ptr(d_anchor), _ptr(d_feat), _ptr(d_offset), _ptr(d_scaling),
                C.byref(grads), scratch.cb, None, _stream()))
        scratch.done()
        return (None, None, None, None, d_anchor, d_feat, d_offset, d_scaling, *d_w)


def generate_neural_gaussians(viewpoint_camera, pc, visible_mask=None, is_training=False):
    """-> (xyz, color, opacity, scaling, rot, neural_opacity, mask), as
    GaussianRenderer::generate_neural_gaussians (src/gaussian_renderer.cpp:214-334)."""
    anchor = pc.get_anchor() if callable(getattr(pc, "get_anchor", None)) else pc._anchor
    if not anchor.is_cuda:
        raise RuntimeError("segs_slam_b200 has no CPU path: tensors must live on a CUDA device")
    cfg = (int(getattr(pc, "appearance_dim", 0)), int(bool(getattr(pc, "use_feat_bank", False))),
           int(bool(getattr(pc, "add_opacity_dist", False))), int(bool(getattr(pc, "add_cov_dist", False))),
           int(bool(getattr(pc, "add_color_dist", False))))
    t, q = viewpoint_camera.t_, viewpoint_camera.R_quaternion_
    pose = [float(t[0]), float(t[1]), float(t[2]), float(q[0]), float(q[1]), float(q[2]), float(q[3])]
    vm = None
    if visible_mask is not None:
        vm = visible_mask if visible_mask.dtype == torch.bool else (visible_mask != 0)
    return _DecodeFunction.apply(cfg, vm, viewpoint_camera.camera_center_, pose, anchor, pc._anchor_feat, pc._offset,
                                 pc.get_scaling(), *_weights(pc))
