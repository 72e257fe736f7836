"""Multi-GPU check (torchrun, NCCL): a sequence of 64-view keyframe-batched steps that DENSIFIES and stays
replica-identical.  Every rank renders its partition of the views, ONE all-reduce per step sums gradients + loss +
statistics delta, every rank applies the same fused Adam step and — every `--interval` steps — the same
adjust_anchor (shared-seed generator).  After every step the ranks compare a checksum of their whole state.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/densify_replicas.py --steps 6 --interval 2"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segs_slam_b200 import anchor_model, mapper  # noqa: E402


def checksum(fm):
    """Bit-level digest of parameters, moments and statistics: sum of the int32 views (wraps, order-independent)."""
    parts = list(fm.params) + [fm.optimizer.exp_avg, fm.optimizer.exp_avg_sq, fm.stats]
    acc = torch.zeros((), dtype=torch.int64, device=fm.stats.device)
    for t in parts:
        acc += t.detach().contiguous().view(torch.int32).to(torch.int64).sum()
    return torch.stack([acc, torch.tensor(fm.pc._anchor.size(0), device=acc.device, dtype=torch.int64)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--interval", type=int, default=2)
    ap.add_argument("--anchors", type=int, default=200_000)
    ap.add_argument("--views", type=int, default=64)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, fx = 1200, 680, 600.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    model = anchor_model.synth_anchor_model(a.anchors, W, H, fx, fx, 1003, device=dev)
    model.voxel_size = 0.01
    cams = anchor_model.circle_keyframes(a.views, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)
    g = torch.Generator(device="cpu").manual_seed(1)
    target = (torch.rand(3, H, W, generator=g) * 0.5).to(dev)
    targets = [target] * a.views
    fm = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lrs=1e-3, statistics=True)
    log = []
    for s in range(1, a.steps + 1):
        t0 = time.perf_counter()
        loss = float(fm.step(cams, targets))
        grown = None
        if s % a.interval == 0:
            # thresholds scaled to a handful of steps (the reference judges after 100 iterations of one view each)
            grown = fm.adjust_anchor(check_interval=a.interval * a.views // 8, success_threshold=0.8, grad_threshold=2e-6,
                                     min_opacity=0.005)
        torch.cuda.synchronize()
        cs = checksum(fm)
        same = True
        if world > 1:
            all_cs = [torch.zeros_like(cs) for _ in range(world)]
            dist.all_gather(all_cs, cs)
            same = all(torch.equal(c, all_cs[0]) for c in all_cs)
        log.append({"step": s, "loss": round(loss, 6), "anchors": int(cs[1]), "adjust": grown, "replicas_identical": bool(same),
                    "ms": round((time.perf_counter() - t0) * 1e3, 1)})
        if not same:
            break
    if rank == 0:
        print(json.dumps({"world": world, "views_per_step": a.views, "steps": log,
                          "all_identical": all(x["replicas_identical"] for x in log)}))
    if world > 1:
        dist.destroy_process_group()
    return 0 if all(x["replicas_identical"] for x in log) else 1


if __name__ == "__main__":
    sys.exit(main())
