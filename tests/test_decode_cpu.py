"""CPU: the decode oracle (oracle/decode_oracle.py) is held to outputs of the reference ITSELF —
tests/golden/decode_*.npz were written by GaussianRenderer::generate_neural_gaussians
(/root/reference/src/gaussian_renderer.cpp:214-334) compiled unmodified into oracle/_ref/_model_ref.so
(tests/golden/make_model_golden.py), forward and backward.  When that module has been built in this container
(`make -C oracle modelref`, done by __graft_entry__.build()) the oracle is also compared with it live on a second
seed."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import decode_oracle as do  # noqa: E402
import make_model_golden as mg  # noqa: E402
import model_ref  # noqa: E402

NAMES = ("xyz", "color", "opacity", "scaling", "rot", "neural_opacity", "mask")


def _oracle_run(cfg, A, seed):
    pc = do.synth_model(A, 1200, 680, 600.0, 600.0, seed, cfg)
    g, vis = mg.seeded_inputs(A, seed)
    out = do.generate_neural_gaussians(pc, torch.tensor(mg.CENTER), mg.T, mg.Q, vis)
    return pc, g, vis, out


@pytest.mark.parametrize("name", list(mg.CONFIGS))
def test_decode_oracle_matches_reference_golden(name):
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"decode_{name}.npz"))
    pc, g, vis, out = _oracle_run(mg.CONFIGS[name], mg.A, mg.SEED)
    assert np.array_equal(vis.numpy(), gold["visible"])
    assert np.array_equal(out[6].numpy(), gold["mask"])
    for k, t in zip(NAMES[:6], out[:6]):
        np.testing.assert_allclose(t.detach().numpy(), gold[k], rtol=2e-6, atol=1e-7, err_msg=k)
    G, Gn = torch.from_numpy(gold["G"]), torch.from_numpy(gold["Gn"])
    mg.functional(out, G, Gn).backward()
    for k, p in zip(("g_anchor", "g_offset", "g_anchor_feat", "g_scaling"), (pc._anchor, pc._offset, pc._anchor_feat, pc._scaling)):
        scale = float(np.abs(gold[k]).max()) + 1e-30
        np.testing.assert_allclose(p.grad.numpy(), gold[k], rtol=1e-5, atol=1e-6 * scale, err_msg=k)
    w = model_ref.mlp_tensors(pc)
    if mg.CONFIGS[name].appearance_dim == 0:
        w = w[:12] + [None, None] + w[12:]
    for i, p in enumerate(w):
        if p is None or p.numel() == 0 or i == 12:      # 12: see make_model_golden.py (dangling from_blob on CPU)
            continue
        gp = [q for q in pc.parameters() if q.data_ptr() == p.data_ptr()][0].grad
        ref = gold[f"g_w{i}"]
        scale = float(np.abs(ref).max()) + 1e-30
        np.testing.assert_allclose((gp if gp is not None else torch.zeros_like(p)).numpy(), ref, rtol=1e-5, atol=1e-6 * scale,
                                   err_msg=f"w{i}")


@pytest.mark.skipif(not model_ref.available(), reason="oracle/_ref/_model_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["bank_app32", "app16_dist"])
def test_decode_oracle_matches_reference_live(name):
    cfg = mg.CONFIGS[name]
    pc, g, vis, out = _oracle_run(cfg, 777, 5)
    m = model_ref.from_model(pc)
    if m.device() != "cpu":
        pytest.skip("reference model landed on a GPU")
    ref = m.generate_neural_gaussians(torch.eye(4), torch.eye(4), torch.tensor(mg.CENTER), list(mg.T), list(mg.Q), vis)
    for k, a, b in zip(NAMES, out, ref):
        assert a.shape == b.shape, k
        assert torch.equal(a, b), k
