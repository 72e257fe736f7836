"""GPU parity of the fused anchor decode (segs_decode_forward / segs_decode_backward through
segs_slam_b200.generate_neural_gaussians) against the oracle restatement of
GaussianRenderer::generate_neural_gaussians (oracle/decode_oracle.py;
/root/reference/src/gaussian_renderer.cpp:214-334), FP32 with TF32 off.

Tolerances: forward rows 1e-5 relative (+1e-6 absolute), gradients 1e-4 relative to
|ref| + mean|ref| (north_star).  The opacity > 0 mask is compared exactly except where
|neural_opacity| < 1e-6 (FP32 summation order differs between cuBLAS and the fused kernel); rows are
compared on the intersection of the two masks.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import decode_oracle as do  # noqa: E402

from segs_slam_b200 import generate_neural_gaussians  # noqa: E402

pytestmark = pytest.mark.gpu


class Cam:
    def __init__(self, dev, center=(0.1, -0.2, 0.05), t=(0.3, -0.1, 0.2), q=(0.98, 0.05, -0.1, 0.15)):
        self.camera_center_ = torch.tensor(center, dtype=torch.float32, device=dev)
        self.t_ = t
        self.R_quaternion_ = q


def _adapt(model):
    """Give the oracle's DecodeModel the reference GaussianModel's member names."""
    cfg = model.cfg
    model.get_anchor = lambda: model._anchor
    model.use_feat_bank, model.appearance_dim = cfg.use_feat_bank, cfg.appearance_dim
    model.add_opacity_dist, model.add_cov_dist, model.add_color_dist = cfg.add_opacity_dist, cfg.add_cov_dist, cfg.add_color_dist
    return model


def _run_pair(A, cfg, dev, vis_frac=None, seed=3):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = _adapt(do.synth_model(A, 1200, 680, 600.0, 600.0, seed, cfg, device=dev))
    cam = Cam(dev)
    vm = None
    if vis_frac is not None:
        g = torch.Generator(device="cpu").manual_seed(seed)
        vm = (torch.rand(A, generator=g) < vis_frac).to(dev)
    ref = do.generate_neural_gaussians(model, cam.camera_center_, cam.t_, cam.R_quaternion_, vm)
    mine = generate_neural_gaussians(cam, model, vm)
    return model, ref, mine


def _close(a, b, rtol, atol):
    return ((a - b).abs() <= atol + rtol * b.abs()).all().item()


def _compare_forward(ref, mine):
    names = ("xyz", "color", "opacity", "scaling", "rot")
    nop_r, nop_m = ref[5].view(-1), mine[5].view(-1)
    assert nop_r.shape == nop_m.shape
    assert _close(nop_m, nop_r, 1e-5, 1e-6), (nop_m - nop_r).abs().max().item()
    mask_r, mask_m = ref[6], mine[6]
    flips = mask_r != mask_m
    assert (nop_r[flips].abs() < 1e-6).all(), "mask differs away from the opacity = 0 threshold"
    assert mine[0].size(0) == int(mask_m.sum().item())
    both = mask_r & mask_m
    idx_r = (torch.cumsum(mask_r, 0) - 1)[both]
    idx_m = (torch.cumsum(mask_m, 0) - 1)[both]
    for k, n in enumerate(names):
        a, b = mine[k][idx_m], ref[k][idx_r]
        assert _close(a, b, 1e-5, 1e-6), (n, (a - b).abs().max().item())
    return both, idx_r, idx_m


@pytest.fixture(params=[2, 1], ids=["v2_both_layers_tcgen05", "v1_thread_per_anchor"])
def variant(request):
    """Run a test on both forward kernels of segs_decode_forward (include/segs_raster.h: segs_decode_set_variant)."""
    from segs_slam_b200 import _lib
    lib = _lib.load()
    before = lib.segs_decode_get_variant()
    _lib.check(lib.segs_decode_set_variant(request.param))
    yield request.param
    _lib.check(lib.segs_decode_set_variant(before))


CONFIGS = {
    "bank_app32": do.DecodeConfig(32, True, False, False, False),
    "plain": do.DecodeConfig(0, False, False, False, False),
    "app16_dist": do.DecodeConfig(16, False, True, True, True),
    "bank_dist_app1": do.DecodeConfig(1, True, True, False, True),
}


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("A,vis", [(1000, None), (5000, 0.6), (333, 0.1)])
def test_decode_forward_parity(device, variant, name, A, vis):
    _, ref, mine = _run_pair(A, CONFIGS[name], device, vis)
    _compare_forward(ref, mine)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_decode_backward_parity(device, variant, name):
    A = 4000
    model, ref, mine = _run_pair(A, CONFIGS[name], device, 0.7)
    both, idx_r, idx_m = _compare_forward(ref, mine)
    g = torch.Generator(device="cpu").manual_seed(11)
    n_slots = ref[6].numel()
    G = torch.randn(n_slots, 14, generator=g).to(device)          # one weight row per (anchor, offset) slot
    Gn = torch.randn(n_slots, generator=g).to(device) * 0.1
    slot = torch.nonzero(both).view(-1)

    def loss_of(out, idx):
        cols = torch.cat([out[0], out[1], out[2], out[3], out[4]], dim=1)       # [rows, 14]
        return (cols[idx] * G[slot]).sum() + (out[5].view(-1) * Gn).sum()

    params = [p for p in model.parameters()]
    g_ref = torch.autograd.grad(loss_of(ref, idx_r), params, allow_unused=True)
    g_mine = torch.autograd.grad(loss_of(mine, idx_m), params, allow_unused=True)
    names = [n for n, _ in model.named_parameters()]
    for n, a, b in zip(names, g_mine, g_ref):
        assert (a is None) == (b is None), n
        if a is None:
            continue
        tol = 1e-4 * (b.abs() + b.abs().mean() + 1e-30)
        bad = ((a - b).abs() > tol)
        assert bad.float().mean().item() <= 1e-4, (n, bad.sum().item(), ((a - b).abs() / tol).max().item())


@pytest.mark.parametrize("A,vis", [(1, None), (127, None), (128, None), (129, None), (1025, None), (2048, 0.5),
                                   (5000, 0.002), (40_000, 0.2), (1031, 1.1)])
def test_decode_ragged_tiles(device, variant, A, vis):
    """Tile edges of both kernels: one anchor, exactly / one over a 128-anchor tile and the 1024-anchor compaction tile,
    a handful of visible anchors spread over many tiles, a mapping-view visibility (20 %), a mask that is all true."""
    _, ref, mine = _run_pair(A, CONFIGS["bank_app32"], device, vis, seed=A)
    _compare_forward(ref, mine)


def test_decode_variants_agree(device):
    """The two forward kernels on the same input: same mask (away from opacity = 0), same rows to 1e-5, and the state
    they leave makes the backward produce the same gradients (the backward reads ordinals, row starts and masks)."""
    from segs_slam_b200 import _lib
    lib = _lib.load()
    before = lib.segs_decode_get_variant()
    outs = {}
    try:
        for v in (1, 2):
            _lib.check(lib.segs_decode_set_variant(v))
            model, _, mine = _run_pair(6000, CONFIGS["bank_app32"], device, 0.35, seed=21)
            # fixed random weights per output element (a quadratic loss on the unit quaternions would have a zero
            # gradient made of rounding noise)
            g = torch.Generator(device="cpu").manual_seed(5)
            loss = sum((t * torch.randn(t.shape, generator=g).to(device)).sum() for t in mine[:6])
            grads = torch.autograd.grad(loss, [p for p in model.parameters()], allow_unused=True)
            outs[v] = (mine, grads)
    finally:
        _lib.check(lib.segs_decode_set_variant(before))
    (m1, g1), (m2, g2) = outs[1], outs[2]
    _compare_forward(m1, m2)
    if torch.equal(m1[6], m2[6]):
        for a, b in zip(g1, g2):
            if a is None:
                assert b is None
                continue
            assert (a - b).abs().max().item() <= 1e-4 * (b.abs().max().item() + 1e-30)


def test_decode_empty_and_all_invisible(device, variant):
    cfg = CONFIGS["bank_app32"]
    model = _adapt(do.synth_model(256, 1200, 680, 600.0, 600.0, 5, cfg, device=device))
    cam = Cam(device)
    vm = torch.zeros(256, dtype=torch.bool, device=device)
    out = generate_neural_gaussians(cam, model, vm)
    assert out[0].shape == (0, 3) and out[5].numel() == 0 and out[6].numel() == 0


@pytest.mark.slow
def test_decode_C3_size(device, variant):
    """BASELINE config 3: 200k anchors x 10 offsets; row count, order and values at full size."""
    _, ref, mine = _run_pair(200_000, CONFIGS["bank_app32"], device, None, seed=1003)
    _compare_forward(ref, mine)
    frac = mine[6].float().mean().item()
    assert 0.2 < frac < 0.8, frac
