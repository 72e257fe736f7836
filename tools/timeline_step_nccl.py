"""Multi-GPU dev tool (torchrun, NCCL): the 64-keyframe mapping step (C4) at N ranks — step time for several
(lanes, read-back wait mode) settings, and a CUPTI timeline (torch.profiler) of ONE step on rank 0 that splits the
step into: lane compute (first kernel .. last lane kernel), drain tail (last lane to finish vs the first), the NCCL
all-reduce kernel, the fused Adam launch, and host gaps between them.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
        tools/timeline_step_nccl.py"""
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segs_slam_b200 import _lib, anchor_model, mapper  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, fx = 1200, 680, 600.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    model = anchor_model.synth_anchor_model(200_000, W, H, fx, fx, 1003, device=dev)
    cams = anchor_model.circle_keyframes(64, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)
    g = torch.Generator(device="cpu").manual_seed(1)
    target = (torch.rand(3, H, W, generator=g) * 0.5).to(dev)
    targets = [target] * 64
    lib = _lib.load()
    out = {"world": world, "cores": os.cpu_count(), "settings": []}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(lanes, blocking, steps=12):
        lib.segs_set_blocking_sync(blocking)
        fm = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lrs=1e-4, lanes=lanes)
        for _ in range(3):
            fm.step(cams, targets)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fm.step(cams, targets)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return fm, float(t.item())

    best = None
    for lanes, blocking in ((4, 1), (4, 0), (2, 0)):
        fm, ms = measure(lanes, blocking, steps=8)
        out["settings"].append({"lanes": lanes, "blocking_sync": blocking, "ms_per_step": round(ms, 3), "keyframes_per_s": round(64e3 / ms, 1)})
        if best is None or ms < best[1]:
            best = ((lanes, blocking), ms)
        del fm
    (lanes, blocking), _ = best
    lib.segs_set_blocking_sync(blocking)
    fm = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lrs=1e-4, lanes=lanes)
    for _ in range(3):
        fm.step(cams, targets)
    barrier()
    # EVERY rank runs the profiled step (it contains the all-reduce); only rank 0 records it
    if rank != 0:
        fm.step(cams, targets)
        torch.cuda.synchronize()
    else:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            t0 = time.perf_counter()
            fm.step(cams, targets)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
        path = os.path.join(tempfile.gettempdir(), "trace_nccl.json")
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
        ev.sort(key=lambda e: e["ts"])
        t_first = ev[0]["ts"]
        end = lambda e: e["ts"] + e["dur"]
        nccl = [e for e in ev if "nccl" in e["name"].lower()]
        adam = [e for e in ev if "adam" in e["name"].lower()]
        lane = [e for e in ev if e not in nccl and e not in adam]
        by_stream = {}
        for e in lane:
            by_stream.setdefault(e["args"].get("stream"), []).append(e)
        lane_ends = sorted(end(v[-1]) for v in by_stream.values() if len(v) > 20)
        tl = {"picked": {"lanes": lanes, "blocking_sync": blocking}, "host_wall_ms": round(wall, 3),
              "gpu_span_ms": round((end(ev[-1]) - t_first) * 1e-3, 3),
              "lane_compute_ms": round((max(end(e) for e in lane) - t_first) * 1e-3, 3),
              "drain_tail_ms (last lane to finish - first lane to finish)": round((lane_ends[-1] - lane_ends[0]) * 1e-3, 3) if len(lane_ends) > 1 else 0.0,
              "lane_streams": len(lane_ends),
              "nccl_kernels": [{"name": e["name"][:60], "start_ms": round((e["ts"] - t_first) * 1e-3, 3), "dur_ms": round(e["dur"] * 1e-3, 3)} for e in nccl],
              "gap_last_lane_kernel_to_nccl_ms": round((nccl[0]["ts"] - max(end(e) for e in lane if e["ts"] < nccl[0]["ts"])) * 1e-3, 3) if nccl else None,
              "adam": [{"start_ms": round((e["ts"] - t_first) * 1e-3, 3), "dur_ms": round(e["dur"] * 1e-3, 3)} for e in adam],
              "gap_nccl_to_adam_ms": round((adam[0]["ts"] - end(nccl[-1])) * 1e-3, 3) if nccl and adam else None,
              "busy_ms_sum_over_streams": round(sum(e["dur"] for e in ev) * 1e-3, 3)}
        out["timeline_rank0"] = tl
    barrier()
    if rank == 0:
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
