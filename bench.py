#!/usr/bin/env python
"""bench.py — rasterizer forward+backward throughput on the BASELINE.json workload.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config C2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): rasterizer fwd+bwd iterations/s @ 1M Gaussians, 1200x680 (config C2 =
`synth(1_000_000, 1200, 680, 600, 600, 1002)`).  One iteration = one view through
segs_raster_forward + segs_raster_backward.  A *step* is VIEWS_PER_STEP (8) views of the replicated
Gaussian set from 8 keyframe poses per GPU, gradients accumulated over the views, followed — when
N > 1 — by one NCCL all-reduce of the accumulated gradients (the keyframe-batched mapping step of
north_star: 8 GPUs x 8 views = the 64-keyframe batch).  `value` = views processed by all ranks /
max-over-ranks device time, so it is directly the headline iterations/s at N = 1 and the mapping
keyframes/s at N > 1 (weak scaling: per-GPU work is fixed).

JSON keys follow the driver's contract; see DESIGN.md §"Measurement".  `--impl reference` times the
UNMODIFIED reference CUDA rasterizer (oracle/_ref/libsegs_ref.so, sm_100a recompile) on the same
GPU with the same harness — the denominator of north_star's ">= 3x" target; when that library is
absent it times the CPU oracle port instead.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VIEWS_PER_STEP = 8
METRIC = "rasterizer fwd+bwd iters/sec @1M Gaussians 1200x680; mapping keyframes/sec 1/2/4/8 GPU"
UNIT = "iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cpu"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-rank runs: do not pin each rank to its GPU's NUMA node")
    ap.add_argument("--blocking-sync", type=int, default=-1, help="-1 auto, 0 spin, 1 sleep in the read-back waits")
    ap.add_argument("--lanes", type=int, default=4, help="views in flight per GPU (concurrent lanes of the batch API)")
    ap.add_argument("--no-mapping", action="store_true", help="skip the keyframe-batched mapping measurement")
    ap.add_argument("--mapping-steps", type=int, default=3)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def view_poses(n_views, first):
    """Keyframe poses of the batched mapping step: the C2 camera shifted sideways by a few cm per
    view (disjoint views of the same replicated scene, near-identical cost per view).  View 0 is
    exactly the C2 camera of BASELINE.md §3."""
    import numpy as np
    out = []
    for v in range(first, first + n_views):
        k = v % 64
        t = np.array([0.02 * (k % 8), 0.015 * (k // 8), 0.0], dtype=np.float32)
        out.append((np.eye(3, dtype=np.float32), t))
    return out


# ---------------------------------------------------------------------------------------------
def algorithmic_bytes(P, R, N):
    """compulsory bytes per stage (each unique array touched once)."""
    return {
        "preprocess": 104 * P,
        "depth_order": 8 * P + 20 * P,                 # F2 scan + the per-Gaussian part of F4
        "binning": 12 * R + 24 * R + 8 * R,            # F4 emission + F5 sort + F6 ranges
        "blend_forward": 40 * R + 20 * N,
        "blend_backward": 40 * R + 20 * N + 44 * P,
        "preprocess_backward": 160 * P,
    }


STAGES = ["preprocess", "depth_order", "binning", "blend_forward", "blend_backward", "preprocess_backward"]


def main():
    args = parse()
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1

    if args.impl != "ours" and distributed and rank != 0:
        return 0          # the reference is single-GPU: rank 0 alone runs and prints it

    if args.impl == "reference-cpu" or (args.impl == "reference" and not _gpu_reference_available()):
        return reference_cpu_arm(args, world)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    numa = _bind_to_gpu_numa_node(local_rank) if (distributed and not args.no_numa_bind) else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed and args.impl == "ours":
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

    from segs_slam_b200 import synth
    import common

    scene0 = synth.config(args.config)
    P, W, H = scene0.P, scene0.W, scene0.H
    N = W * H
    poses = view_poses(VIEWS_PER_STEP, first=rank * VIEWS_PER_STEP)
    scenes = [synth.with_camera(scene0, Rm, t) if (t != 0).any() else scene0 for Rm, t in poses]
    base = scene0.to_torch(dev)
    cams = [{k: torch.from_numpy(np.ascontiguousarray(getattr(s, k))).to(dev)
             for k in ("viewmatrix", "projmatrix", "campos")} for s in scenes]
    dL = base["dL_dout"]

    if args.impl == "ours":
        from segs_slam_b200 import _lib, rasterize_points as rp
        lib = _lib.load()
        # more waiting host threads (ranks x lanes) than cores: sleep in the read-back waits instead of spinning
        blocking = args.blocking_sync if args.blocking_sync >= 0 else int(world * args.lanes * 2 > (os.cpu_count() or 1))
        lib.segs_set_blocking_sync(blocking)

        def fwd_bwd(cam, dL_dout, pr=None):
            pr = pr or base
            a = (base["bg"], pr["means3D"], pr["colors"], pr["opacities"], pr["scales"], pr["rotations"],
                 1.0, common.empty(dev), cam["viewmatrix"], cam["projmatrix"], scene0.tanfovx, scene0.tanfovy, H, W,
                 common.empty(dev), 0, cam["campos"], False)
            R, color, radii, g, b, i = rp.RasterizeGaussiansCUDA(*a)
            grads = rp.RasterizeGaussiansBackwardCUDA(
                base["bg"], pr["means3D"], radii, pr["colors"], pr["scales"], pr["rotations"], 1.0,
                common.empty(dev), cam["viewmatrix"], cam["projmatrix"], scene0.tanfovx, scene0.tanfovy, dL_dout,
                common.empty(dev), 0, cam["campos"], g, R, b, i)
            return R, color, grads
    else:
        import refimpl
        lib = None

        def fwd_bwd(cam, dL_dout, pr=None):
            pr = pr or base
            e = common.empty(dev)
            R, color, radii, g, b, i = refimpl.forward(
                base["bg"], pr["means3D"], pr["colors"], pr["opacities"], pr["scales"], pr["rotations"],
                1.0, e, cam["viewmatrix"], cam["projmatrix"], scene0.tanfovx, scene0.tanfovy, H, W, e, 0,
                cam["campos"])
            d = refimpl.backward(base["bg"], pr["means3D"], radii, pr["colors"], pr["scales"],
                                 pr["rotations"], 1.0, e, cam["viewmatrix"], cam["projmatrix"], scene0.tanfovx,
                                 scene0.tanfovy, dL_dout, e, 0, cam["campos"], g, R, b, i)
            grads = (d["dL_dmeans2D"], d["dL_dcolors"], d["dL_dopacity"], d["dL_dmeans3D"], d["dL_dcov3D"],
                     d["dL_dsh"], d["dL_dscales"], d["dL_drotations"])
            return R, color, grads

    # flat gradient bucket = what the mapper all-reduces (segs_slam_b200/mapper.py: one contiguous FP32
    # slice per tensor: means3D 3, means2D 3, colors 3, opacity 1, scales 3, rotations 4 floats per Gaussian)
    from segs_slam_b200 import mapper
    GRAD_IDX = (3, 0, 1, 2, 6, 7)
    widths = (3, 3, 3, 1, 3, 4)
    shapes = [torch.empty((P, w), dtype=torch.float32, device=dev) for w in widths]
    gb = mapper.GradBucket(shapes)
    bucket = gb.flat

    def accumulate(grads, first, into=None):
        b = gb if into is None else into
        if first:
            b.zero_()
        b.accumulate([grads[gi].view(P, w) for gi, w in zip(GRAD_IDX, widths)])

    R_seen = []
    LANES = args.lanes if args.impl == "ours" else 1
    if args.impl == "ours":
        # The step goes through the batch API (mapper.RasterBatch -> segs_raster_views): the views of the keyframe
        # batch are issued from C++ on LANES concurrent lanes (stream + host thread each), so the latency-bound
        # stages of one view (sorts, binning, the num_rendered read-back) overlap the issue-bound blend kernels of
        # another; gradients are accumulated straight into the flat bucket.
        rb = mapper.RasterBatch(dev, lanes=LANES)
        step_images = [torch.empty((3, H, W), dtype=torch.float32, device=dev) for _ in cams]
        dLs_same = [dL] * len(cams)

        def run_batch(pr, dLs, imgs, bkt, lanes=None):
            bkt.zero_()
            return rb.run(pr["means3D"], pr["colors"], pr["opacities"], pr["scales"], pr["rotations"], base["bg"], cams,
                          H, W, scene0.tanfovx, scene0.tanfovy, imgs, dLs, bkt.views, lanes=lanes)

        def step(lanes=None):
            R_seen.extend(run_batch(base, dLs_same, step_images, gb, lanes))
            if distributed:
                dist.all_reduce(bucket)
            return step_images[-1]
    else:
        def step(lanes=None):
            for v, cam in enumerate(cams):
                R, color, grads = fwd_bwd(cam, dL)
                accumulate(grads, v == 0)
                R_seen.append(R)
            return color

    def barrier():
        torch.cuda.synchronize()
        if distributed and args.impl == "ours":
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing -------------------------------------------------
    # warm-up: at least W (>= 3) steps, continued until the GPU has been busy for ~1.5 s so that
    # clocks and the caching allocator are in steady state (the count actually run is reported)
    warm = 0
    t_warm = time.perf_counter()
    for _ in range(max(3, args.warmup)):
        step()
        torch.cuda.synchronize()
        warm += 1
    # ... continued until ~1.5 s of GPU work; the extra count is agreed across ranks (the step
    # contains a collective, so every rank must run the same number of steps)
    t_two = time.perf_counter()
    for _ in range(2):                      # steady-state step time (the first steps pay for allocation)
        step()
        torch.cuda.synchronize()
        warm += 1
    per_step = (time.perf_counter() - t_two) / 2
    extra = torch.tensor([max(0, min(200 - warm, int(math.ceil(1.5 / max(per_step, 1e-4)))))], device=dev)
    if distributed and args.impl == "ours":
        dist.all_reduce(extra, op=dist.ReduceOp.MAX)
    for _ in range(int(extra.item())):
        step()
        torch.cuda.synchronize()
        warm += 1
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stage_ms = np.zeros(len(STAGES))
    launches0 = lib.segs_launch_count() if lib else 0
    import ctypes as C
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R_seen.clear()
    barrier()
    ev0.record()
    n_prof = 0
    for s in range(args.steps):
        # per-stage CUDA events are armed on every 10th step only; that step runs its views on ONE lane so that
        # each stage is timed on its own (under two lanes the stages of different views overlap); it stays inside
        # the timed region and costs it a few percent
        prof = lib is not None and s % 10 == 0
        if prof:
            lib.segs_profile_enable(1)
        step(lanes=1 if prof else None)
        if prof:
            lib.segs_profile_enable(0)
            ms = (C.c_float * len(STAGES))()
            lib.segs_profile_read(ms)          # stages of the step's last view
            stage_ms += np.array(list(ms)); n_prof += 1
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = (lib.segs_launch_count() - launches0) if lib else 0
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if distributed and args.impl == "ours":
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    n_ranks = world if args.impl == "ours" else 1
    views = VIEWS_PER_STEP * args.steps * n_ranks
    value = views / (elapsed_ms * 1e-3)
    R_mean = float(np.mean(R_seen)) if R_seen else 0.0

    # ---------------- end-to-end: host buffers in, host buffers out --------------------------
    # Per step: the Gaussian parameters come from pinned host memory, every view's dL_dout comes
    # from pinned host memory, every view's image and the step's accumulated gradients go back to
    # pinned host memory.  Copies run on two copy streams (H2D / D2H) and are double-buffered
    # against the compute stream with events, so PCIe traffic of view v+1 / step s+1 overlaps the
    # kernels of view v / step s; the host owns step s's results before step s+2 is queued, and
    # everything is drained inside the timed region.  (Identical harness for both arms.)
    PKEYS = ("means3D", "colors", "opacities", "scales", "rotations")
    host_in = {k: base[k].cpu().pin_memory() for k in PKEYS}
    host_dL = dL.cpu().pin_memory()
    IMG_RING = 8
    host_img = [torch.empty((3, H, W), dtype=torch.float32).pin_memory() for _ in range(IMG_RING)]
    host_grads = [torch.empty_like(bucket, device="cpu").pin_memory() for _ in range(2)]
    h2d = sum(v.numel() * 4 for v in host_in.values()) + VIEWS_PER_STEP * host_dL.numel() * 4
    d2h = VIEWS_PER_STEP * host_img[0].numel() * 4 + host_grads[0].numel() * 4
    cur = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dev_params = [{k: torch.empty_like(base[k]) for k in PKEYS} for _ in range(2)]
    dev_dL = [torch.empty_like(dL) for _ in range(2)]
    buckets = [gb, mapper.GradBucket(shapes)]
    Ev = torch.cuda.Event
    ev_par_ready, ev_par_free = [Ev(), Ev()], [Ev(), Ev()]
    ev_dL_ready, ev_dL_free = [Ev(), Ev()], [Ev(), Ev()]
    ev_img_done, ev_grads_done = [Ev() for _ in range(IMG_RING)], [Ev(), Ev()]
    for e_ in ev_par_free + ev_dL_free + ev_img_done + ev_grads_done:
        e_.record(cur)

    def upload_params(slot):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_par_free[slot])
            for k in PKEYS:
                dev_params[slot][k].copy_(host_in[k], non_blocking=True)
            ev_par_ready[slot].record(s_in)

    def upload_dL(slot):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_dL_free[slot])
            dev_dL[slot].copy_(host_dL, non_blocking=True)
            ev_dL_ready[slot].record(s_in)

    state = {"views": 0}

    def e2e_step(i, last):
        p = i & 1
        if i == 0:
            upload_params(0)
            upload_dL(0)
        if not last:
            upload_params(p ^ 1)                      # next step's inputs ride behind this step's kernels
        cur.wait_event(ev_par_ready[p])
        cur.wait_event(ev_grads_done[p])              # bucket p was downloaded two steps ago
        for v, cam in enumerate(cams):
            q = state["views"] & 1
            state["views"] += 1
            if not (last and v == len(cams) - 1):
                upload_dL(q ^ 1)                      # next view's dL_dout
            cur.wait_event(ev_dL_ready[q])
            R, color, grads = fwd_bwd(cam, dev_dL[q], dev_params[p])
            ev_dL_free[q].record(cur)
            accumulate(grads, v == 0, buckets[p])
            done = Ev()
            done.record(cur)
            r = (state["views"] - 1) % IMG_RING
            ev_img_done[r].synchronize()              # the image that used this host slot a ring ago has landed
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                host_img[r].copy_(color, non_blocking=True)
                color.record_stream(s_out)
                ev_img_done[r].record(s_out)
        ev_par_free[p].record(cur)
        if distributed and args.impl == "ours":
            dist.all_reduce(buckets[p].flat)
        done = Ev()
        done.record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            host_grads[p].copy_(buckets[p].flat, non_blocking=True)
            ev_grads_done[p].record(s_out)
        ev_grads_done[p ^ 1].synchronize()            # the host owns the previous step's results

    if args.impl == "ours":
        # batch API: the step's inputs (parameters + every view's dL_dout) are uploaded while the previous step
        # computes, its outputs (every view's image + the accumulated gradients) are downloaded while the next one
        # computes; same bytes per step as the per-view harness of the reference arm
        nv = len(cams)
        dev_dLs = [[torch.empty_like(dL) for _ in range(nv)] for _ in range(2)]
        dev_imgs = [[torch.empty((3, H, W), dtype=torch.float32, device=dev) for _ in range(nv)] for _ in range(2)]
        host_imgs = [[torch.empty((3, H, W), dtype=torch.float32).pin_memory() for _ in range(nv)] for _ in range(2)]
        ev_in_ready, ev_in_free, ev_out_done = [Ev(), Ev()], [Ev(), Ev()], [Ev(), Ev()]
        for e_ in ev_in_free + ev_out_done:
            e_.record(cur)

        def upload_step(slot):
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_in_free[slot])
                for k in PKEYS:
                    dev_params[slot][k].copy_(host_in[k], non_blocking=True)
                for v in range(nv):
                    dev_dLs[slot][v].copy_(host_dL, non_blocking=True)
                ev_in_ready[slot].record(s_in)

        def e2e_step(i, last):                      # noqa: F811  (replaces the per-view variant above)
            p = i & 1
            if i == 0:
                upload_step(0)
            if not last:
                upload_step(p ^ 1)                    # next step's inputs ride behind this step's kernels
            cur.wait_event(ev_in_ready[p])
            cur.wait_event(ev_out_done[p])            # slot p's images / bucket were downloaded two steps ago
            run_batch(dev_params[p], dev_dLs[p], dev_imgs[p], buckets[p])
            ev_in_free[p].record(cur)
            if distributed:
                dist.all_reduce(buckets[p].flat)
            done = Ev()
            done.record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for v in range(nv):
                    host_imgs[p][v].copy_(dev_imgs[p][v], non_blocking=True)
                host_grads[p].copy_(buckets[p].flat, non_blocking=True)
                ev_out_done[p].record(s_out)
            ev_out_done[p ^ 1].synchronize()          # the host owns the previous step's results
            state["views"] += nv

        ev_grads_done = ev_out_done                   # what e2e_run drains

    def e2e_run(n):
        state["views"] = 0
        for i in range(n):
            e2e_step(i, i == n - 1)
        for e_ in ev_grads_done + ev_img_done:        # drain: every result is in host memory
            e_.synchronize()
        torch.cuda.synchronize()

    e2e_steps = max(3, args.steps)
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # device events vs host wall clock
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if distributed and args.impl == "ours":
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = VIEWS_PER_STEP * e2e_steps * n_ranks / (float(t.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    mapping = None
    if not args.no_mapping:
        mapping = mapping_ours(args, dev, rank, n_ranks, distributed) if args.impl == "ours" else mapping_reference(args, dev)

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline of the dominant kernel ----------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    roofline, stages_out = None, {}
    if lib and n_prof:
        ms = stage_ms / n_prof
        ab = algorithmic_bytes(P, R_mean, N)
        for name, m in zip(STAGES, ms):
            stages_out[name] = {"ms": round(float(m), 4), "algorithmic_MB": round(ab[name] / 1e6, 2),
                                "GBps": round(ab[name] / (m * 1e-3) / 1e9, 1) if m > 0 else None}
        top = STAGES[int(np.argmax(ms))]
        achieved = ab[top] / (float(ms[STAGES.index(top)]) * 1e-3) / 1e9
        traffic, issue_pct = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(top)
            issue_pct = tj.get("_issue_active_pct", {}).get(top)
        except Exception:
            pass
        roofline = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak_gbs,
                    "unit": "GB/s", "frac": round(achieved / peak_gbs, 4), "traffic": traffic,
                    "peak_source": peak_src, "ms_per_launch": round(float(ms[STAGES.index(top)]), 4),
                    "note": "blend kernels are FP32-ALU/MUFU bound, not HBM bound (SURVEY 8d); "
                            "frac is algorithmic bytes / time / copy peak as the contract asks",
                    "issue_active_pct_ncu": issue_pct, "stages": stages_out, "step_share": round(float(ms[STAGES.index(top)] / ms.sum()), 3)}

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": round(elapsed_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: synth({P},{W},{H}) single-view rasterizer forward+backward "
                               f"(BASELINE.md section 3), colours precomputed, scale+quaternion",
                   "views_per_step_per_gpu": VIEWS_PER_STEP, "views_in_flight_per_gpu": LANES, "P": P, "W": W, "H": H,
                   "num_rendered_mean": round(R_mean), "parallelism": f"dp{n_ranks} (keyframe views)",
                   "numa_node_rank0": numa,
                   "l2": "working set per view (~0.45 GB state + gradients) exceeds the 126 MB L2; no flush"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "pipeline": "pinned host buffers; H2D and D2H on two copy streams, double-buffered against the "
                            "compute stream; drained inside the timed region" +
                            ("; batch API (mapper.RasterBatch), step-level double buffering" if args.impl == "ours" else "")},
        "gpu_launches": int(launches),
    }
    if mapping is not None:
        line["mapping"] = mapping
    if args.impl != "ours":
        line["impl"] = "reference"
        line["gpu_launches"] = 0
        line["config"]["reference"] = ("unmodified SEGS-SLAM cuda_rasterizer compiled for sm_100a "
                                       "(oracle/_ref), same GPU, same harness; single GPU")
        line["cpu_baseline"] = {"value": line["value"], "unit": UNIT, "cores": 0, "kind": "reference",
                                "sample": "the reference's own CUDA kernels on the GPU (it has no CPU path)"}
    else:
        line["roofline"] = roofline
        if not args.no_cpu_baseline and args.gpus == 1:
            line["cpu_baseline"] = cpu_baseline(args.config)
    print(json.dumps(line), flush=True)
    if distributed and args.impl == "ours":
        dist.destroy_process_group()
    return 0


MAP_VIEWS, MAP_ANCHORS = 64, 200_000


def _mapping_setup(dev):
    """BASELINE config 4: 64 keyframes at 1200x680 on a 1.5 m circle around the C3 anchor model."""
    import torch
    from segs_slam_b200 import anchor_model
    W, H, fx = 1200, 680, 600.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    model = anchor_model.synth_anchor_model(MAP_ANCHORS, W, H, fx, fx, 1003, device=dev)
    cams = anchor_model.circle_keyframes(MAP_VIEWS, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)
    g = torch.Generator(device="cpu").manual_seed(1)
    target = (torch.rand(3, H, W, generator=g) * 0.5).to(dev)
    return model, cams, [target] * MAP_VIEWS, (W, H, tanx, tany)


def mapping_ours(args, dev, rank, n_ranks, distributed):
    """Mapping keyframes/s (the second half of BASELINE.json's metric): the keyframe-batched step of
    segs_slam_b200.mapper.FusedMapper — per view prefilter -> fused decode -> rasterize -> L1+SSIM+scaling
    regulariser -> backward (segs_mapper_view), ONE NCCL all-reduce of the flat gradient bucket, ONE fused Adam
    launch.  Strong scaling: the 64 views are partitioned across the ranks."""
    import torch
    import torch.distributed as dist
    from segs_slam_b200 import _lib, mapper
    model, cams, targets, (W, H, tanx, tany) = _mapping_setup(dev)
    fm = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lambda_dssim=0.2, lrs=1e-4)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        fm.step(cams, targets)
    barrier()
    lib = _lib.load()
    l0 = lib.segs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    losses = [fm.step(cams, targets) for _ in range(args.mapping_steps)]
    e1.record()
    barrier()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"metric": "mapping keyframes/s", "value": round(MAP_VIEWS * args.mapping_steps / (ms * 1e-3), 2),
            "unit": "keyframes/s", "n_gpus": n_ranks, "views_per_step": MAP_VIEWS, "steps": args.mapping_steps,
            "ms_per_step": round(ms / args.mapping_steps, 3), "scaling": "strong",
            "config": {"workload": "C4: 64 keyframes 1200x680, C3 anchor model (200k anchors x 10 offsets, appearance "
                                   "embedding + feature bank)", "loss": "0.8 L1 + 0.2 (1 - SSIM) + 0.01 scaling regulariser",
                       "optimizer": "fused Adam, one step per 64-keyframe batch", "bucket_MB": round(fm.bucket.flat.numel() * 4 / 1e6, 1)},
            "losses": [round(float(x), 5) for x in losses],
            "gpu_launches_per_rank": int(lib.segs_launch_count() - l0)}


def mapping_reference(args, dev):
    """The reference's own mapping iteration (one keyframe, one optimizer step: gaussian_mapper.cpp:823-1032) over the
    unmodified reference rasterizer kernels and the ATen op sequences of its LibTorch host code (tests/ref_mapper.py).
    Single GPU (the reference has no multi-GPU mapper)."""
    import torch
    try:
        import ref_mapper
        import decode_oracle
    except Exception as exc:                                   # oracle/_ref absent
        return {"unavailable": f"{type(exc).__name__}: {exc}"}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True          # the reference arm at its best: cuDNN times its conv2d algorithms once
    model, cams, targets, (W, H, tanx, tany) = _mapping_setup(dev)
    # the decode oracle reads its configuration from `cfg`
    model.cfg = decode_oracle.DecodeConfig(appearance_dim=model.appearance_dim, use_feat_bank=model.use_feat_bank)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, eps=1e-15)
    bg = torch.zeros(3, device=dev)
    n = min(MAP_VIEWS, 16 * args.mapping_steps)
    for v in range(10):                                        # cuDNN picks its conv2d algorithms, the allocator settles
        ref_mapper.iteration(model, cams[v], targets[v], H, W, tanx, tany, bg, opt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    losses = [ref_mapper.iteration(model, cams[v % MAP_VIEWS], targets[v % MAP_VIEWS], H, W, tanx, tany, bg, opt) for v in range(n)]
    e1.record()
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    return {"metric": "mapping keyframes/s", "value": round(n / (ms * 1e-3), 2), "unit": "keyframes/s", "n_gpus": 1,
            "views_per_step": 1, "steps": n, "ms_per_step": round(ms / n, 3), "scaling": "strong",
            "config": {"workload": "C4 keyframes, C3 anchor model; the reference's iteration: one keyframe per optimizer step, "
                                   "unmodified reference rasterizer kernels + ATen decode/loss/Adam (tests/ref_mapper.py)",
                       "loss": "0.8 L1 + 0.2 (1 - SSIM) + 0.01 scaling regulariser"},
            "losses": [round(float(x), 5) for x in losses[:4]]}


def _bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: pin this rank's host threads (and therefore its pinned staging buffers, first-touch) to the
    NUMA node its GPU hangs off, like `numactl --cpunodebind --membind` per rank.  The end-to-end pipeline moves
    280 MB per step and rank over PCIe; with 8 ranks allocating on whatever node they start on, half of that crosses
    the socket interconnect.  Returns the node, or None when the platform does not expose one."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def _gpu_reference_available():
    try:
        import torch
        import refimpl
        return torch.cuda.is_available() and refimpl.available()
    except Exception:
        return False


def cpu_baseline(config):
    """CPU oracle (scalar C++ port, oracle/raster_oracle.cpp) on all host cores: one full
    forward+backward of the same workload."""
    import oracle_lib
    from segs_slam_b200 import synth
    cores = os.cpu_count() or 1
    scene = synth.config(config)
    t0 = time.perf_counter()
    f = oracle_lib.from_scene(scene, nthreads=cores)
    f.backward(scene.dL_dout)
    dt = time.perf_counter() - t0
    f.close()
    return {"value": round(1.0 / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 full forward+backward iteration of {config} ({scene.P} Gaussians, {scene.W}x{scene.H}) "
                      f"on {cores} threads, {dt:.2f} s"}


def reference_cpu_arm(args, world):
    """Fallback reference arm when oracle/_ref cannot run: the CPU oracle port."""
    cb = cpu_baseline(args.config)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": round(1e3 / cb["value"], 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.config} single-view rasterizer forward+backward, CPU oracle port"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
