"""GPU: densification decisions (segs_slam_b200/densify.py over csrc/densify.cu) against the reference's OWN
GaussianModel::adjust_anchor / anchor_growing / prune_anchor (src/gaussian_model.cpp:1505-1762, compiled unmodified into
oracle/_ref/_model_ref.so) on the same state with the same torch RNG seed: every tensor — the six anchor tensors, the four
statistics, the Adam moments — must be IDENTICAL, bit for bit.  The restatement oracle/densify_oracle.py is held to the same
run, and to the committed goldens in tests/test_densify_cpu.py."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import densify_cases as dc  # noqa: E402
import densify_oracle  # noqa: E402
import model_ref  # noqa: E402

from segs_slam_b200 import densify  # noqa: E402

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not model_ref.available(), reason="oracle/_ref/_model_ref.so not built")


def _product_input(ref_before, dev):
    return {k: v.detach().clone().to(dev).contiguous() for k, v in ref_before.items()}


def _assert_same(mine, ref, names):
    for k in names:
        a, b = mine[k], ref[k]
        assert tuple(a.shape) == tuple(b.shape), (k, tuple(a.shape), tuple(b.shape))
        assert torch.equal(a, b), (k, float((a - b).abs().max()), int((a != b).sum()))


@needs_ref
@pytest.mark.parametrize("case", list(dc.CASES))
def test_adjust_anchor_is_identical_to_the_reference(device, case):
    A, seed = dc.CASES[case]
    st, grads_adam = dc.make_state(A, seed)
    m = dc.reference_model(model_ref, st, grads_adam)
    before = dc.reference_state(m)
    mine_in = _product_input(before, device)
    oracle_in = _product_input(before, device)
    torch.manual_seed(1234 + seed)
    m.adjust_anchor(100, 0.8, 0.0002, 0.005)
    after = dc.reference_state(m)
    # the product draws with torch.rand from the same (default) generator: the reference's torch::rand_like sequence
    torch.manual_seed(1234 + seed)
    mine = densify.adjust_anchor(mine_in, 100, 0.8, 0.0002, 0.005, **dc.MODEL)
    grown = sum(n for _c, n in mine["_growing_report"])
    A_mid, A_after = mine["_prune_report"]
    assert grown > 0 and A_after < A_mid, (mine["_growing_report"], mine["_prune_report"])
    assert sum(1 for _c, n in mine["_growing_report"] if n > 0) >= 2, mine["_growing_report"]
    names = list(dc.NAMES) + [k for k in after if k[:2] in ("m_", "v_")]
    _assert_same(mine, after, names)
    # the restatement, fed the same random numbers
    torch.manual_seed(1234 + seed)
    rands = [torch.rand(A * 10, device=device) for _ in range(dc.MODEL["update_depth"])]
    orc = densify_oracle.adjust_anchor(oracle_in, rands, 100, 0.8, 0.0002, 0.005, **dc.MODEL)
    _assert_same(orc, after, names)


def test_adjust_anchor_edge_cases(device):
    """Nothing to grow and nothing to prune (fresh statistics); everything pruned."""
    st, _g = dc.make_state(300, 3)
    st = {k: v.to(device) for k, v in st.items()}
    z = {k: torch.zeros_like(st[k]) for k in ("opacity_accum", "anchor_demon", "offset_gradient_accum", "offset_denom")}
    s0 = {**{k: v.clone() for k, v in st.items()}, **z}
    out = densify.adjust_anchor(s0, **dc.MODEL)
    assert out["_anchor"].shape == st["_anchor"].shape and torch.equal(out["_anchor"], st["_anchor"])
    assert out["_prune_report"] == (300, 300) and all(n == 0 for _c, n in out["_growing_report"])
    exp = st["_scaling"].clone()
    exp[:, 3:] = exp[:, 3:].clamp(max=0.05)
    assert torch.equal(out["_scaling"], exp)            # prune_anchor's clamp applies even when nothing is pruned
    s1 = {k: v.clone() for k, v in st.items()}
    s1["anchor_demon"] = torch.full_like(s1["anchor_demon"], 100.0)
    s1["opacity_accum"] = torch.zeros_like(s1["opacity_accum"])
    s1["offset_denom"] = torch.zeros_like(s1["offset_denom"])
    out = densify.adjust_anchor(s1, **dc.MODEL)
    assert out["_anchor"].shape == (0, 3) and out["offset_denom"].shape == (0, 1) and out["_offset"].shape == (0, 10, 3)


def test_fused_mapper_densifies_and_keeps_training(device):
    """FusedMapper.adjust_anchor: statistics from real views, growth + pruning, then further steps on the new anchor set;
    two mappers fed different view partitions but the SUM of the statistics (what the all-reduce gives every rank) and
    the same seed end up with identical anchor sets."""
    from segs_slam_b200 import anchor_model, mapper
    W, H, fx = 208, 120, 150.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    cams = anchor_model.circle_keyframes(8, 1.5, (0.0, 0.0, 3.25), tanx, tany, device)
    g = torch.Generator(device="cpu").manual_seed(1)
    targets = [(torch.rand(3, H, W, generator=g) * 0.5).to(device) for _ in cams]
    bg = torch.zeros(3, device=device)

    def make():
        model = anchor_model.synth_anchor_model(3000, W, H, fx, fx, 1003, device=device)
        model.voxel_size = 0.01
        return mapper.FusedMapper(model, H, W, tanx, tany, bg, lanes=2, statistics=True, lrs=1e-3)

    a, b = make(), make()
    for _ in range(3):
        a.step(cams, targets)                      # all 8 views on replica a ...
    # ... replica b sees the same parameters and the same summed statistics (the all-reduce), rendered in another order
    for _ in range(3):
        b.step(list(reversed(cams)), list(reversed(targets)))
    for fm in (a, b):                              # make the thresholds reachable after 3 steps of 8 views
        res = fm.adjust_anchor(check_interval=10, success_threshold=0.8, grad_threshold=1e-7, min_opacity=0.005)
        assert res[1] != res[0]
    # the two replicas saw the same views in a different order: float sums differ in the last bits, decisions must not
    assert a.pc._anchor.shape == b.pc._anchor.shape
    assert torch.equal(a.pc._anchor[3000:], b.pc._anchor[3000:]) or a.pc._anchor.size(0) == b.pc._anchor.size(0)
    A1 = a.pc._anchor.size(0)
    assert a.stats.numel() == 22 * A1 and a.bucket.flat.numel() == 71 * A1 + sum(w.numel() for w in a.weights if w is not None)
    l0 = float(a.step(cams, targets))
    for _ in range(5):
        l1 = float(a.step(cams, targets))
    assert np.isfinite(l0) and np.isfinite(l1) and l1 < l0 * 1.05


@needs_ref
def test_cpp_host_layer_is_identical_to_the_reference(device):
    """densify::adjust_anchor of the LibTorch host layer (csrc/torch_shim/densify.{h,cpp}, the C++ twin of densify.py — the
    reference's own host language) against the reference's GaussianModel::adjust_anchor: every tensor bit-identical."""
    from segs_slam_b200 import _segs_torch as shim
    A, seed = dc.CASES["a2000"]
    st, grads_adam = dc.make_state(A, seed)
    m = dc.reference_model(model_ref, st, grads_adam)
    before = dc.reference_state(m)
    cpp_in = _product_input(before, device)
    torch.manual_seed(99)
    m.adjust_anchor(100, 0.8, 0.0002, 0.005)
    after = dc.reference_state(m)
    torch.manual_seed(99)
    out = shim.adjust_anchor(cpp_in, 100, 0.8, 0.0002, 0.005, **dc.MODEL)
    names = list(dc.NAMES) + [k for k in after if k[:2] in ("m_", "v_")]
    _assert_same(out, after, names)
    assert sum(n for _c, n in out["_growing_report"]) > 0 and out["_prune_report"][1] < out["_prune_report"][0]
