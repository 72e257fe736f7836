"""Dev tool: where does the wall time per view go?  python tools/host_overhead.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import common
from segs_slam_b200 import synth, _lib, rasterize_points as rp
dev = torch.device("cuda:0")
scene = synth.config("C2"); t = scene.to_torch(dev); a = common.scene_args(t, scene, dev)
lib = _lib.load()
def fwd():
    return rp.RasterizeGaussiansCUDA(a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], 1.0,
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"], a["sh"], 0, a["campos"], False)
def bwd(st):
    return rp.RasterizeGaussiansBackwardCUDA(a["bg"], a["means3D"], st[2], a["colors"], a["scales"], a["rotations"], 1.0,
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], t["dL_dout"], a["sh"], 0, a["campos"], st[3], st[0], st[4], st[5])
def loop(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("fwd only        ms/view", loop(lambda: fwd()))
st = fwd()
print("bwd only        ms/view", loop(lambda: bwd(st)))
print("fwd+bwd         ms/view", loop(lambda: bwd(fwd())))
bucket = torch.zeros((scene.P, 17), device=dev)
def acc(g):
    col = 0
    for gi, w in zip((3, 0, 1, 2, 6, 7), (3, 3, 3, 1, 3, 4)):
        bucket[:, col:col + w].add_(g[gi].view(scene.P, w)); col += w
print("fwd+bwd+acc     ms/view", loop(lambda: acc(bwd(fwd()))))
g = bwd(st)
print("acc only        ms/view", loop(lambda: acc(g)))
lib.segs_profile_enable(1)
print("fwd+bwd profile ms/view", loop(lambda: bwd(fwd())))
lib.segs_profile_enable(0)
# host-side cost of the python wrapper alone: time the forward call with a CPU timer split
import ctypes as C
t0 = time.perf_counter(); n = 200
for _ in range(n): common.empty(dev)
print("empty() us", (time.perf_counter() - t0) / n * 1e6)
t0 = time.perf_counter()
for _ in range(n): rp._Grower(dev)
print("_Grower() us", (time.perf_counter() - t0) / n * 1e6)
t0 = time.perf_counter()
for _ in range(n): torch.empty((scene.P, 3), device=dev)
print("torch.empty us", (time.perf_counter() - t0) / n * 1e6)
