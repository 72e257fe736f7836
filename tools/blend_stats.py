"""Dev tool (GPU, library built with EXTRA=-DSEGS_BLEND_STATS): how much work the blend kernels do per view."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
from segs_slam_b200 import _lib, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda:0")
scene = synth.config(name)
t = scene.to_torch(dev)
a = common.scene_args(t, scene, dev)
lib = _lib.load()
out = (C.c_ulonglong * 8)()
common.run_mine(a, t["dL_dout"])
lib.segs_debug_blend_stats(out, 1)
m = common.run_mine(a, t["dL_dout"])
lib.segs_debug_blend_stats(out, 1)
R = m["R"]
names = ["records staged", "(G,subtile) evaluated", "(G,subtile) reaching a pixel", "blended (G,pixel) pairs"]
print(name, "R", R, "tiles", ((scene.W + 15) // 16) * ((scene.H + 15) // 16))
for k in range(2):
    print("forward" if k == 0 else "backward", {n: int(out[4 * k + i]) for i, n in enumerate(names)})
