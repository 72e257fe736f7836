import os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import decode_oracle as do
import test_decode_gpu as td
from segs_slam_b200 import generate_neural_gaussians
dev = torch.device("cuda:0")
model = td._adapt(do.synth_model(200000, 1200, 680, 600.0, 600.0, 1003, do.DecodeConfig(), device=dev))
cam = td.Cam(dev); params = list(model.parameters())
def run():
    out = generate_neural_gaussians(cam, model, None)
    loss = out[0].sum() + out[1].sum() + out[2].sum() + out[3].sum() + out[4].sum()
    torch.autograd.grad(loss, params, allow_unused=True)
for i in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize()
    print(f"{(time.perf_counter()-t0)*1e3:.2f}", end=" ")
print()
print(torch.cuda.memory_reserved()/1e9, torch.cuda.memory_allocated()/1e9)
