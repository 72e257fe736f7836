"""Dev tool: device time of the fused L1+SSIM loss (forward + backward) at 1200x680 vs the ATen composition."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from segs_slam_b200 import loss_utils
dev = torch.device("cuda:0")
H, W = 680, 1200
g = torch.Generator(device="cpu").manual_seed(1)
x = torch.rand(3, H, W, generator=g).to(dev).requires_grad_(True)
y = torch.rand(3, H, W, generator=g).to(dev)

def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

def mine():
    x.grad = None
    loss_utils.l1_ssim_loss(x, y, 0.2)[0].backward()

def aten():
    import loss_oracle
    x.grad = None
    (0.8 * loss_oracle.l1_loss(x, y) + 0.2 * (1.0 - loss_oracle.ssim(x, y))).backward()

print("fused loss fwd+bwd ms", round(timeit(mine), 4))
if "--no-oracle" not in sys.argv:
    print("ATen composition (reference's op sequence) fwd+bwd ms", round(timeit(aten), 4))
