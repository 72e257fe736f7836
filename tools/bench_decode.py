"""Dev tool: time the fused anchor decode (forward, forward+backward) against the PyTorch-eager oracle
restatement of the reference on config C3 (200k anchors x 10 offsets).
    python tools/bench_decode.py [A] [--no-oracle]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import decode_oracle as do
import test_decode_gpu as td
from segs_slam_b200 import generate_neural_gaussians

A = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 200_000
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
model = td._adapt(do.synth_model(A, 1200, 680, 600.0, 600.0, 1003, do.DecodeConfig(), device=dev))
cam = td.Cam(dev)
params = list(model.parameters())

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def fwd_mine(): return generate_neural_gaussians(cam, model, None)
def fwd_ref(): return do.generate_neural_gaussians(model, cam.camera_center_, cam.t_, cam.R_quaternion_, None)
def fb(f):
    def run():
        out = f()
        loss = out[0].sum() + out[1].sum() + out[2].sum() + out[3].sum() + out[4].sum()
        torch.autograd.grad(loss, params, allow_unused=True)
    return run

out = fwd_mine()
res = {"A": A, "n_out": int(out[0].size(0)), "mine_fwd_ms": timeit(fwd_mine), "mine_fwd_bwd_ms": timeit(fb(fwd_mine))}
if "--no-oracle" not in sys.argv:
    res["oracle_fwd_ms"] = timeit(fwd_ref); res["oracle_fwd_bwd_ms"] = timeit(fb(fwd_ref))
# algorithmic bytes: 284 B per visible anchor in, 56 B per emitted Gaussian out
res["fwd_alg_GB"] = (284 * A + 56 * res["n_out"]) / 1e9
res["fwd_GBps"] = res["fwd_alg_GB"] / (res["mine_fwd_ms"] * 1e-3)
print(json.dumps(res))
