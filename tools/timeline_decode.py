"""Dev tool: CUPTI timeline of decode forward+backward at C3 (kernels, memsets, gaps)."""
import os, sys, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from torch.profiler import profile, ProfilerActivity
import decode_oracle as do
import test_decode_gpu as td
from segs_slam_b200 import generate_neural_gaussians
A = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dev = torch.device("cuda:0")
model = td._adapt(do.synth_model(A, 1200, 680, 600.0, 600.0, 1003, do.DecodeConfig(), device=dev))
cam = td.Cam(dev); params = list(model.parameters())
def run():
    out = generate_neural_gaussians(cam, model, None)
    loss = out[0].sum() + out[1].sum() + out[2].sum() + out[3].sum() + out[4].sum()
    torch.autograd.grad(loss, params, allow_unused=True)
for _ in range(5): run()
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N): run()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace_decode.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]; busy = sum(e["dur"] for e in ev)
print(f"span {span/N:.1f} us/iter busy {busy/N:.1f} us/iter")
agg = {}
for e in ev:
    k = e["name"][:60]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += e["dur"]
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {d/N:8.1f} us/iter  n/iter={n/N:5.1f}  {k}")
