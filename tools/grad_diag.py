"""Dev tool: how do product-vs-reference gradient errors compare with the reference's own
run-to-run spread?  python tools/grad_diag.py C2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import common
from segs_slam_b200 import synth
import test_raster_parity_gpu as T

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda:0")
scene = T.scenes()[name] if name in ("tiny", "small_bg", "odd_rot", "C1") else synth.config(name)
t = scene.to_torch(dev)
a = common.scene_args(t, scene, dev)
m = common.run_mine(a, t["dL_dout"])
m2 = common.run_mine(a, t["dL_dout"])
refs = [common.run_ref(a, t["dL_dout"]) for _ in range(3)]
def stats(x, ref):
    x, ref = x.double().flatten(), ref.double().flatten()
    tol = 1e-4 * (ref.abs() + ref.abs().mean())
    ratio = (x - ref).abs() / tol
    return ratio.max().item(), (ratio > 1).double().mean().item(), ((x - ref).norm() / ref.norm()).item()
for k in m["grads"]:
    if m["grads"][k].numel() == 0: continue
    r0 = refs[0]["grads"][k]
    print(f"{k:14s} mine-ref max/frac>1/relL2 = %.2f %.2e %.2e | ref1-ref0 %.2f %.2e %.2e | ref2-ref0 %.2f %.2e %.2e | mine2-mine %.2f %.2e %.2e" % (
        *stats(m["grads"][k], r0), *stats(refs[1]["grads"][k], r0), *stats(refs[2]["grads"][k], r0), *stats(m2["grads"][k], m["grads"][k])))
