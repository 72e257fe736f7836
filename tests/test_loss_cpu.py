"""CPU suite: the loss / Adam oracle (oracle/loss_oracle.py) against the goldens produced by the reference's own
loss_utils.h and torch::optim::Adam (tests/golden/make_loss_golden.py)."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import loss_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
LOSS = sorted(glob.glob(os.path.join(GOLD, "loss_*.npz")))
ADAM = sorted(glob.glob(os.path.join(GOLD, "adam_*.npz")))


def test_fixtures_present():
    assert len(LOSS) >= 4 and len(ADAM) >= 3


@pytest.mark.parametrize("path", LOSS, ids=[os.path.basename(p) for p in LOSS])
def test_loss_oracle_matches_reference(path):
    g = np.load(path)
    o = loss_oracle.mapper_loss(g["image"], g["gt"], float(g["lambda_dssim"]), bool(g["apply_mask"]),
                                g["scaling"] if "scaling" in g.files else None)
    for k in ("l1", "ssim", "loss"):
        np.testing.assert_allclose(o[k], g[k], rtol=1e-6, atol=0)
    # where image == gt exactly the L1 term vanishes and the SSIM term is ~1e-13: rounding noise, hence the atol
    np.testing.assert_allclose(o["dL_dimage"], g["dL_dimage"], rtol=1e-5, atol=1e-5 * float(np.abs(g["dL_dimage"]).max()))
    if "scaling" in g.files:
        np.testing.assert_allclose(o["dL_dscaling"], g["dL_dscaling"], rtol=1e-6, atol=0)
    import torch
    ps = loss_oracle.psnr(torch.from_numpy(g["image"]), torch.from_numpy(g["gt"]))
    np.testing.assert_allclose(float(ps), float(g["psnr"]), rtol=1e-6)


@pytest.mark.parametrize("path", ADAM, ids=[os.path.basename(p) for p in ADAM])
def test_adam_oracle_matches_torch_optim_adam(path):
    g = np.load(path)
    p, _m, _v = loss_oracle.adam(g["param0"], g["grads"], float(g["lr"]), float(g["beta1"]), float(g["beta2"]),
                                 float(g["eps"]), float(g["weight_decay"]))
    # the update itself is what must agree: compare the parameter DELTAS to 1e-5 relative
    d_ref, d_mine = g["param"] - g["param0"], p - g["param0"]
    # the deltas are differences of FP32 parameters: they carry the parameters' own rounding (1-2 ulp of |p|)
    ulp = float(np.spacing(np.float32(np.abs(g["param0"]).max())))
    np.testing.assert_allclose(d_mine, d_ref, rtol=2e-5, atol=2.0 * ulp)
    if float(g["weight_decay"]) == 0.0:
        assert np.array_equal(p[:7], g["param0"][:7])      # zero gradient, zero moments: untouched


FREQ = sorted(glob.glob(os.path.join(GOLD, "freq_*.npz")))


@pytest.mark.parametrize("path", FREQ, ids=[os.path.basename(p) for p in FREQ])
def test_frequency_loss_oracle_matches_reference(path):
    """high_frequency_loss / low_freq_loss of the reference's own header (incl. its mask-indexing quirk)."""
    import torch
    g = np.load(path)
    x = torch.from_numpy(g["image"]).clone().requires_grad_(True)
    y = torch.from_numpy(g["gt"])
    hi = loss_oracle.high_frequency_loss(x, y)
    hi.backward()
    np.testing.assert_allclose(float(hi), float(g["high"]), rtol=1e-5)
    np.testing.assert_allclose(x.grad.numpy(), g["d_high"], rtol=1e-4, atol=1e-5 * float(np.abs(g["d_high"]).max()))
    np.testing.assert_allclose(float(loss_oracle.low_freq_loss(x.detach(), y)), float(g["low"]), rtol=1e-5)
    assert len(FREQ) >= 2
