"""Keyframe-batched mapping step (BASELINE.json config 4): 64 synthetic keyframes at 1200x680 of the
C3 anchor model (200k anchors x 10 offsets), per view  prefilter -> fused decode -> rasterize -> L1 ->
backward, gradients accumulated in the flat bucket, ONE all-reduce per step, Adam.  Strong scaling: the
64 views are partitioned across the ranks.

    python tools/bench_mapping.py [--steps K] [--views 64] [--anchors 200000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_mapping.py
"""
import argparse, datetime, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from segs_slam_b200 import anchor_model, mapper

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--views", type=int, default=64)
ap.add_argument("--anchors", type=int, default=200_000)
ap.add_argument("--lanes", type=int, default=2)
ap.add_argument("--path", default="fused", choices=["fused", "autograd", "autograd-l1"])
args = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
W, H, fx = 1200, 680, 600.0
tanx, tany = W / (2 * fx), H / (2 * fx)
model = anchor_model.synth_anchor_model(args.anchors, W, H, fx, fx, 1003, device=dev)
cams = anchor_model.circle_keyframes(args.views, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)
g = torch.Generator(device="cpu").manual_seed(1)
target = (torch.rand(3, H, W, generator=g) * 0.5).to(dev)
targets = [target] * args.views
bg = torch.zeros(3, device=dev)
if args.path == "fused":
    fm = mapper.FusedMapper(model, H, W, tanx, tany, bg, lrs=1e-4, lanes=args.lanes)
    class _B:  # noqa: E701
        flat = fm.bucket.flat
    def one_step(bucket):
        return fm.step(cams, targets), _B
else:
    render_loss = mapper.make_render_loss(model, cams, targets, H, W, tanx, tany, bg,
                                          loss="l1" if args.path == "autograd-l1" else "l1_ssim")
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    def one_step(bucket):
        return mapper.mapping_step(params, render_loss, args.views, opt, bucket)
bucket = None

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

loss, bucket = one_step(bucket)     # warm-up step
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
losses = []
for _ in range(args.steps):
    loss, bucket = one_step(bucket)
    losses.append(float(loss))
e1.record()
sync()
ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"metric": "mapping keyframes/s (C4: 64 keyframes, 1200x680, C3 anchor model)", "n_gpus": world,
                      "views_per_step": args.views, "path": args.path, "lanes": args.lanes, "steps": args.steps, "anchors": args.anchors,
                      "value": round(args.views * args.steps / (float(t.item()) * 1e-3), 2), "unit": "keyframes/s",
                      "ms_per_step": round(float(t.item()) / args.steps, 2), "scaling": "strong",
                      "bucket_MB": round(bucket.flat.numel() * 4 / 1e6, 1), "losses": [round(x, 5) for x in losses]}))
if world > 1:
    dist.destroy_process_group()
