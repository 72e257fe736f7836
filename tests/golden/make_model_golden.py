"""Generates tests/golden/decode_*.npz by running the reference's OWN GaussianRenderer::generate_neural_gaussians
(/root/reference/src/gaussian_renderer.cpp:214-334, compiled unmodified into oracle/_ref/_model_ref.so — see
oracle/Makefile `modelref`) on seeded inputs, forward and backward, on the CPU of the build container:

    make -C oracle modelref && python tests/golden/make_model_golden.py

Inputs are regenerated from the seed (oracle/decode_oracle.synth_model); the fixtures hold the visible mask, the seven
outputs and the gradients of a seeded linear functional of them w.r.t. every trainable tensor.  tests/test_decode_cpu.py
holds oracle/decode_oracle.py to them; tests/test_decode_gpu.py holds the CUDA decode to them on the GPU.
With --densify (GPU box only: the reference hard-codes torch::kCUDA there) it also writes densify_*.npz: the state of
the reference's GaussianModel before and after adjust_anchor (gaussian_model.cpp:1505-1762)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import decode_oracle as do  # noqa: E402
import model_ref  # noqa: E402

CONFIGS = {
    "bank_app32": do.DecodeConfig(32, True, False, False, False),
    "plain": do.DecodeConfig(0, False, False, False, False),
    "app16_dist": do.DecodeConfig(16, False, True, True, True),
    "bank_dist_app1": do.DecodeConfig(1, True, True, False, True),
}
A, SEED = 300, 23
CENTER, T, Q = (0.1, -0.2, 0.05), (0.3, -0.1, 0.2), (0.98, 0.05, -0.1, 0.15)


def functional(out, G, Gn):
    """The seeded linear functional whose gradients the fixtures hold: sum(cols * G[surviving slots]) + sum(nop * Gn)."""
    cols = torch.cat([out[0], out[1], out[2], out[3], out[4]], dim=1)
    slot = torch.nonzero(out[6]).view(-1)
    return (cols * G[slot]).sum() + (out[5].view(-1) * Gn).sum()


def seeded_inputs(A_, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    vis = torch.rand(A_, generator=g) < 0.7
    return g, vis


def decode_golden(name, cfg):
    pc = do.synth_model(A, 1200, 680, 600.0, 600.0, SEED, cfg)
    g, vis = seeded_inputs(A, SEED)
    m = model_ref.from_model(pc)
    out = m.generate_neural_gaussians(torch.eye(4), torch.eye(4), torch.tensor(CENTER), list(T), list(Q), vis)
    n_slots = out[6].numel()
    G = torch.randn(n_slots, 14, generator=g)
    Gn = torch.randn(n_slots, generator=g) * 0.1
    functional(out, G, Gn).backward()
    st = m.state()
    d = {"visible": vis.numpy(), "G": G.numpy(), "Gn": Gn.numpy()}
    for k, t in zip(("xyz", "color", "opacity", "scaling", "rot", "neural_opacity", "mask"), out):
        d[k] = t.detach().numpy()
    for k, t in zip(("g_anchor", "g_offset", "g_anchor_feat", "g_scaling"), st[:4]):
        d[k] = t.grad.numpy()
    for i, p in enumerate(m.mlp_parameters()):
        if i == 12:
            # mlp_apperance.weight: its gradient needs the pose row, which the reference builds with torch::from_blob over
            # a local std::vector (gaussian_renderer.cpp:258-266).  On CUDA `.to(device)` copies it; on the CPU of this
            # container it is a no-op, the tensor dangles after the function returns and backward reads freed memory.
            # The GPU test (tests/test_decode_gpu.py, live against _model_ref on CUDA) covers this tensor.
            continue
        d[f"g_w{i}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    path = os.path.join(HERE, f"decode_{name}.npz")
    np.savez_compressed(path, **d)
    print(path, {k: v.shape for k, v in d.items() if k in ("xyz", "neural_opacity")})


if __name__ == "__main__":
    for name, cfg in CONFIGS.items():
        decode_golden(name, cfg)
    if "--densify" in sys.argv:
        import make_densify_golden
        make_densify_golden.main()
