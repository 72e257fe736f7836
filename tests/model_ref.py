"""TEST / BASELINE INFRASTRUCTURE: the reference's own GaussianModel / GaussianRenderer / GaussianRasterizer /
loss_utils, compiled UNMODIFIED from /root/reference/src into oracle/_ref/_model_ref.so (`make -C oracle modelref`,
veneer oracle/model_ref_wrap.cpp), driven from Python.  Used by tests/, tests/golden/make_model_golden.py, smoke() and
bench.py's reference arm; never by the product."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_mod = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "_model_ref.so"))


def load():
    global _mod
    if _mod is None:
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import _model_ref
        _mod = _model_ref
    return _mod


def mlp_tensors(pc):
    """The MLP parameters of a DecodeModel / AnchorModel in the order of RefModel.mlp_parameters():
    opacity, cov, color, appearance, feature bank (gaussian_model.cpp:62-96)."""
    out = []
    for s in (pc.mlp_opacity, pc.mlp_cov, pc.mlp_color, pc.mlp_apperance, pc.mlp_feature_bank):
        if s is not None:
            out += [p.detach() for p in s.parameters()]
    return out


def cfg_of(pc):
    g = lambda k, d: getattr(pc, k, getattr(getattr(pc, "cfg", None), k, d))
    return dict(appearance_dim=int(g("appearance_dim", 32)), use_feat_bank=bool(g("use_feat_bank", True)),
                add_opacity_dist=bool(g("add_opacity_dist", False)), add_cov_dist=bool(g("add_cov_dist", False)),
                add_color_dist=bool(g("add_color_dist", False)))


def from_model(pc, reference_ctor: bool = False, **model_kw):
    """A reference GaussianModel holding a copy of `pc`'s state (on CUDA when a device is present, else CPU).
    appearance_dim == 0: the reference still builds Linear(7, 0) (gaussian_model.cpp:84-86); it holds no numbers."""
    mr = load()
    cfg = cfg_of(pc)
    m = mr.RefModel(reference_ctor=reference_ctor, **cfg, **model_kw)
    A = pc._anchor.size(0)
    rot = getattr(pc, "_rotation", None)
    if rot is None:
        rot = torch.zeros(A, 4)
        rot[:, 0] = 1.0
    opa = getattr(pc, "_opacity", None)
    if opa is None:
        opa = torch.zeros(A, 1)
    m.set_state(pc._anchor.detach(), pc._offset.detach(), pc._anchor_feat.detach(), pc._scaling.detach(), rot.detach(),
                opa.detach())
    w = mlp_tensors(pc)
    ref_w = m.mlp_parameters()
    if cfg["appearance_dim"] == 0:                  # insert the empty Linear(7, 0) tensors the reference keeps
        w = w[:12] + [ref_w[12], ref_w[13]] + w[12:]
    m.load_mlp_parameters(w)
    return m
