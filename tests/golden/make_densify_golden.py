"""Generates tests/golden/densify_*.npz ON THE GPU BOX (the reference hard-codes torch::kCUDA in anchor_growing): the
state of the reference's own GaussianModel before and after GaussianModel::adjust_anchor
(/root/reference/src/gaussian_model.cpp:1505-1762, compiled unmodified into oracle/_ref/_model_ref.so), plus the random
numbers torch::rand_like drew:

    gpurun -- 'python tests/golden/make_densify_golden.py gpurun_out/golden'     # then copy the .npz files here

tests/test_densify_cpu.py holds oracle/densify_oracle.py to them on the CPU."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import densify_cases as dc  # noqa: E402
import model_ref  # noqa: E402


def main(out_dir=HERE):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    for case in ("a257",):          # the larger cases are compared live on the GPU (tests/test_densify_gpu.py)
        A, seed = dc.CASES[case]
        st, grads_adam = dc.make_state(A, seed)
        m = dc.reference_model(model_ref, st, grads_adam)
        before = dc.reference_state(m)
        d = {"before/" + k: v.detach().cpu().numpy() for k, v in before.items()}
        torch.manual_seed(1234 + seed)
        rands = [torch.rand(A * 10, device=dev) for _ in range(dc.MODEL["update_depth"])]
        torch.manual_seed(1234 + seed)
        m.adjust_anchor(100, 0.8, 0.0002, 0.005)
        after = dc.reference_state(m)
        d.update({"after/" + k: v.detach().cpu().numpy() for k, v in after.items()})
        for i, r in enumerate(rands):
            d[f"rand/{i}"] = r.cpu().numpy()
        path = os.path.join(out_dir, f"densify_{case}.npz")
        np.savez_compressed(path, **d)
        print(path, "anchors", before["_anchor"].shape[0], "->", after["_anchor"].shape[0])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
