"""Dev tool for ncu: N views of the fused C4 mapping step on ONE lane (kernels in issue order).
    python tools/one_view_mapping.py [views]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from segs_slam_b200 import anchor_model, mapper
dev = torch.device("cuda:0")
NV = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W, H, fx = 1200, 680, 600.0
tanx, tany = W / (2 * fx), H / (2 * fx)
model = anchor_model.synth_anchor_model(200_000, W, H, fx, fx, 1003, device=dev)
cams = anchor_model.circle_keyframes(64, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)[:NV]
target = (torch.rand(3, H, W, generator=torch.Generator().manual_seed(1)) * 0.5).to(dev)
fm = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lrs=1e-4, lanes=1)
loss = fm.step(cams, [target] * NV)
torch.cuda.synchronize()
print("ok", float(loss), fm.last_result.n_gaussians, fm.last_result.num_rendered)
