"""Shared by tests/test_densify_{gpu,cpu}.py and tests/golden/make_densify_golden.py: seeded synthetic model states in
which anchor_growing adds anchors on several levels and prune_anchor removes some."""
import numpy as np
import torch

NAMES = ("_anchor", "_offset", "_anchor_feat", "_opacity", "_scaling", "_rotation", "opacity_accum", "anchor_demon",
         "offset_gradient_accum", "offset_denom")
MODEL = dict(n_offsets=10, update_depth=3, update_init_factor=16, update_hierachy_factor=4, voxel_size=0.01)
CASES = {"a2000": (2000, 31), "a257": (257, 5), "a6000": (6000, 77)}


def make_state(A: int, seed: int, k: int = 10, feat_dim: int = 32):
    """CPU tensors.  Gradients straddle the growing thresholds (0.0002 x {1, 2, 4}); about half of the offsets have been
    seen often enough (offset_denom > 40); a third of the anchors are old enough to be judged (anchor_demon > 80) and
    some of those have a low accumulated opacity."""
    rng = np.random.default_rng(seed)
    f = np.float32
    st = {
        "_anchor": rng.uniform(-1.0, 1.0, (A, 3)).astype(f),
        "_offset": rng.uniform(-1.0, 1.0, (A, k, 3)).astype(f),
        "_anchor_feat": rng.normal(0.0, 0.1, (A, feat_dim)).astype(f),
        "_opacity": rng.normal(0.0, 1.0, (A, 1)).astype(f),
        "_scaling": np.log(rng.uniform(0.01, 0.3, (A, 6))).astype(f),          # some log-scalings above 0.05's clamp? (log < 0)
        "_rotation": rng.normal(0.0, 1.0, (A, 4)).astype(f),
    }
    st["_scaling"][rng.uniform(size=A) < 0.1, 3:] = f(0.2)                     # rows prune_anchor's clamp(max = 0.05) changes
    denom = np.floor(rng.uniform(0.0, 100.0, (A * k, 1))).astype(f)
    grads = rng.uniform(0.0, 0.0012, (A * k, 1)).astype(f)
    st["offset_denom"] = denom
    st["offset_gradient_accum"] = (grads * denom).astype(f)
    demon = np.floor(rng.uniform(0.0, 130.0, (A, 1))).astype(f)
    st["anchor_demon"] = demon
    st["opacity_accum"] = (demon * rng.uniform(0.0, 0.02, (A, 1))).astype(f)
    grads_adam = [rng.normal(0.0, 1.0, st[n].shape).astype(f) for n in ("_anchor", "_offset", "_anchor_feat", "_scaling")]
    return {k_: torch.from_numpy(v) for k_, v in st.items()}, [torch.from_numpy(g) for g in grads_adam]


def reference_model(model_ref, st, grads_adam):
    """The reference's GaussianModel (compiled unmodified) holding `st`, with Adam moments from one zero-lr step."""
    mr = model_ref.load()
    m = mr.RefModel(voxel_size=MODEL["voxel_size"], update_depth=MODEL["update_depth"], update_init_factor=MODEL["update_init_factor"],
                    update_hierachy_factor=MODEL["update_hierachy_factor"], reference_ctor=True)
    m.set_state(st["_anchor"], st["_offset"], st["_anchor_feat"], st["_scaling"], st["_rotation"], st["_opacity"])
    m.training_setup()
    m.set_learning_rates([0.0] * m.n_param_groups())
    m.adam_step_with_grads(grads_adam)
    m.set_statistics(st["opacity_accum"], st["anchor_demon"], st["offset_gradient_accum"], st["offset_denom"])
    return m


def reference_state(m):
    """dict of the reference model's tensors + the moments of the four trainable anchor tensors."""
    # RefModel.state() order (oracle/model_ref_wrap.cpp)
    order = ("_anchor", "_offset", "_anchor_feat", "_scaling", "_rotation", "_opacity", "opacity_accum", "anchor_demon",
             "offset_gradient_accum", "offset_denom")
    out = dict(zip(order, m.state()))
    adam = m.adam_state()          # groups: anchor, offset, feat, opacity, scaling, rotation
    for name, g in zip(("_anchor", "_offset", "_anchor_feat", "_opacity", "_scaling", "_rotation"), adam):
        if len(g) == 3:
            out["m_" + name], out["v_" + name] = g[1], g[2]
    return out
