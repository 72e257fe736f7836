"""Run N forward+backward iterations of the product (or the reference with --ref) on one config.
Dev tool for ncu launch lists:  python tools/one_iter.py C2 3 [--ref]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
from segs_slam_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
use_ref = "--ref" in sys.argv
dev = torch.device("cuda:0")
scene = synth.config(name)
t = scene.to_torch(dev)
a = common.scene_args(t, scene, dev)
for i in range(n):
    out = (common.run_ref if use_ref else common.run_mine)(a, t["dL_dout"])
torch.cuda.synchronize()
print("ok", name, "R", out["R"])
