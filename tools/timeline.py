"""Dev tool: GPU timeline of the bench step (torch.profiler/CUPTI): busy time vs span, largest gaps."""
import os, sys, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
from torch.profiler import profile, ProfilerActivity
import common
from segs_slam_b200 import synth, rasterize_points as rp
dev = torch.device("cuda:0")
scene = synth.config(sys.argv[1] if len(sys.argv) > 1 else "C2"); t = scene.to_torch(dev); a = common.scene_args(t, scene, dev)
P = scene.P
from segs_slam_b200 import mapper
gb = mapper.GradBucket([torch.empty((P, w), device=dev) for w in (3, 3, 3, 1, 3, 4)])
def view(first):
    st = rp.RasterizeGaussiansCUDA(a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], 1.0,
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"], a["sh"], 0, a["campos"], False)
    g = rp.RasterizeGaussiansBackwardCUDA(a["bg"], a["means3D"], st[2], a["colors"], a["scales"], a["rotations"], 1.0,
        a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], t["dL_dout"], a["sh"], 0, a["campos"], st[3], st[0], st[4], st[5])
    if first: gb.zero_()
    gb.accumulate([g[gi].view(P, w) for gi, w in zip((3, 0, 1, 2, 6, 7), (3, 3, 3, 1, 3, 4))])
for i in range(30): view(i % 8 == 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(8): view(i == 0)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
busy = sum(e["dur"] for e in ev)
print(f"events {len(ev)} span {span/8:.1f} us/view busy {busy/8:.1f} us/view idle {(span-busy)/8:.1f} us/view")
agg = {}
for e in ev:
    k = e["name"][:50]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += e["dur"]
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"  {d/8:8.1f} us/view  n/view={n/8:5.1f}  {k}")
gaps = []
for x, y in zip(ev[:-1], ev[1:]):
    g = y["ts"] - (x["ts"] + x["dur"])
    if g > 0: gaps.append((g, x["name"][:36], y["name"][:36]))
gaps.sort(reverse=True)
gagg = {}
for g, x, y in gaps:
    gagg.setdefault((x, y), [0, 0.0]); gagg[(x, y)][0] += 1; gagg[(x, y)][1] += g
print("largest gap classes (us/view):")
for (x, y), (n, d) in sorted(gagg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {d/8:7.1f}  n/view={n/8:4.1f}  {x}  ->  {y}")
