#!/usr/bin/env python
"""bench.py — rasterizer forward+backward throughput on the BASELINE.json workload.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config C2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): rasterizer fwd+bwd iterations/s @ 1M Gaussians, 1200x680 (config C2 =
`synth(1_000_000, 1200, 680, 600, 600, 1002)`).

`value` is SURVEY §8(d)'s M1, identically in BOTH arms: views ONE AT A TIME through the two drop-in entry points
RasterizeGaussiansCUDA + RasterizeGaussiansBackwardCUDA (ours: segs_slam_b200.rasterize_points over the C ABI; reference
arm: the unmodified reference kernels of oracle/_ref).  A *step* is the 64-keyframe batch of north_star: VIEWS_PER_STEP
(64) keyframe views of the replicated Gaussian set per GPU, the gradients of every view accumulated into the flat FP32
bucket, and — when N > 1 — one NCCL all-reduce of the bucket (weak scaling: per-GPU work is fixed).  `value` = views
processed by all ranks / max-over-ranks device time.  K steps of 64 views put >= 1 s inside the timed region.

Named extras on the same line:
  `batch`   — the same step through the NEW batch API (segs_raster_views: 4 views in flight on concurrent lanes); the
              reference has no such API, its batch figure is its M1 figure;
  `e2e`     — M1 with HOST buffers (pinned) in and out inside the timed region; `e2e_batch` the batch API likewise;
  `mapping` — mapping keyframes/s (config C4) of the fused keyframe-batched step vs the reference's own iteration
              (compiled unmodified: oracle/_ref/_model_ref.so), per-keyframe and as accumulated baseline "B";
  `configs` — M1 at C1 and C5, decode+rasterize at C3, distCUDA2 at P = 1e3 / 1e4 / 1e5;
  `roofline`, `cpu_baseline`, `clocks`, `gpu_launches` as the contract asks.
See DESIGN.md §"Measurement".
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
sys.path.insert(0, os.path.join(ROOT, "oracle"))

VIEWS_PER_STEP = 64
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
    ap.add_argument("--lanes", type=int, default=4, help="views in flight per GPU in the batch-API figures")
    ap.add_argument("--no-mapping", action="store_true", help="skip the keyframe-batched mapping measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C5 / distCUDA2 sub-lines")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mapping-steps", type=int, default=20)
    ap.add_argument("--views", type=int, default=VIEWS_PER_STEP, help="views per step and GPU")
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


class Arm:
    """One implementation of the two drop-in entry points over a fixed scene: ours (the product) or the unmodified
    reference kernels.  fwd_bwd(cam, dL_dout, params) -> (num_rendered, image, 8 gradient tensors in the reference's
    tuple order)."""

    def __init__(self, impl, scene, base, dev):
        import common
        self.impl, self.scene, self.base, self.dev = impl, scene, base, dev
        self.e = common.empty(dev)
        self.lib = None
        if impl == "ours":
            from segs_slam_b200 import _lib, rasterize_points as rp
            self.lib, self.rp = _lib.load(), rp
        else:
            import refimpl
            self.ref = refimpl

    def fwd_bwd(self, cam, dL_dout, pr=None):
        s, b, e = self.scene, self.base, self.e
        pr = pr or b
        if self.impl == "ours":
            rp = self.rp
            R, color, radii, g, bn, im = rp.RasterizeGaussiansCUDA(
                b["bg"], pr["means3D"], pr["colors"], pr["opacities"], pr["scales"], pr["rotations"], 1.0, e,
                cam["viewmatrix"], cam["projmatrix"], s.tanfovx, s.tanfovy, s.H, s.W, e, 0, cam["campos"], False)
            grads = rp.RasterizeGaussiansBackwardCUDA(
                b["bg"], pr["means3D"], radii, pr["colors"], pr["scales"], pr["rotations"], 1.0, e, cam["viewmatrix"],
                cam["projmatrix"], s.tanfovx, s.tanfovy, dL_dout, e, 0, cam["campos"], g, R, bn, im)
            return R, color, grads
        ref = self.ref
        R, color, radii, g, bn, im = ref.forward(
            b["bg"], pr["means3D"], pr["colors"], pr["opacities"], pr["scales"], pr["rotations"], 1.0, e,
            cam["viewmatrix"], cam["projmatrix"], s.tanfovx, s.tanfovy, s.H, s.W, e, 0, cam["campos"])
        d = ref.backward(b["bg"], pr["means3D"], radii, pr["colors"], pr["scales"], pr["rotations"], 1.0, e,
                         cam["viewmatrix"], cam["projmatrix"], s.tanfovx, s.tanfovy, dL_dout, e, 0, cam["campos"], g, R, bn, im)
        grads = (d["dL_dmeans2D"], d["dL_dcolors"], d["dL_dopacity"], d["dL_dmeans3D"], d["dL_dcov3D"],
                 d["dL_dsh"], d["dL_dscales"], d["dL_drotations"])
        return R, color, grads


def timed(fn, n, sync):
    """-> list of n host-observed durations (ms) of fn(), each bracketed by a device synchronize."""
    out = []
    for _ in range(n):
        sync()
        t0 = time.perf_counter()
        fn()
        sync()
        out.append((time.perf_counter() - t0) * 1e3)
    return out


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
    ours = args.impl == "ours"
    use_dist = distributed and ours
    if use_dist:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))

    from segs_slam_b200 import synth, mapper
    import common

    NV = max(1, args.views)
    scene0 = synth.config(args.config)
    P, W, H = scene0.P, scene0.W, scene0.H
    N = W * H
    poses = view_poses(NV, first=rank * NV)
    scenes = [synth.with_camera(scene0, Rm, t) if (t != 0).any() else scene0 for Rm, t in poses]
    base = scene0.to_torch(dev)
    cams = [{k: torch.from_numpy(np.ascontiguousarray(getattr(s, k))).to(dev)
             for k in ("viewmatrix", "projmatrix", "campos")} for s in scenes]
    dL = base["dL_dout"]
    arm = Arm(args.impl, scene0, base, dev)
    lib = arm.lib
    def set_wait_mode(threads_per_rank):
        """Read-back waits spin unless there are more waiting host threads on the box than cores (then they sleep)."""
        if not ours:
            return
        blocking = args.blocking_sync if args.blocking_sync >= 0 else int(world * threads_per_rank > (os.cpu_count() or 1))
        lib.segs_set_blocking_sync(blocking)

    set_wait_mode(1)                    # M1: one host thread per rank

    # flat gradient bucket = what the mapper all-reduces (segs_slam_b200/mapper.py: one contiguous FP32
    # slice per tensor: means3D 3, means2D 3, colors 3, opacity 1, scales 3, rotations 4 floats per Gaussian)
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
    prof_state = {"ms": np.zeros(len(STAGES)), "n": 0, "on": False}

    def step_m1():
        """64 views, one at a time through the two drop-in entry points; gradients accumulated; one all-reduce."""
        for v, cam in enumerate(cams):
            prof = prof_state["on"] and lib is not None and v == len(cams) - 1
            if prof:
                lib.segs_profile_enable(1)
            Rn, color, grads = arm.fwd_bwd(cam, dL)
            if prof:
                import ctypes as C
                lib.segs_profile_enable(0)
                ms = (C.c_float * len(STAGES))()
                lib.segs_profile_read(ms)
                prof_state["ms"] += np.array(list(ms)); prof_state["n"] += 1
            accumulate(grads, v == 0)
            R_seen.append(Rn)
        if use_dist:
            dist.all_reduce(bucket)
        return color

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def device_timed(step_fn, steps, warmup):
        """W untimed warm-up steps, then EXACTLY `steps` steps between CUDA events, barrier + synchronize on both
        sides, max over ranks.  -> (elapsed ms, warm-up steps run)"""
        for _ in range(warmup):
            step_fn()
            torch.cuda.synchronize()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            step_fn()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if use_dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), warmup

    n_ranks = world if ours else 1
    warm = max(3, args.warmup)

    # ---------------- M1, device-resident: the headline `value` ---------------------------------
    sampler = ClockSampler(local_rank)
    launches0 = lib.segs_launch_count() if lib else 0
    for _ in range(warm):                      # warm-up outside the clock sampling window
        step_m1()
        torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    R_seen.clear()
    prof_state["on"] = True
    launches0 = lib.segs_launch_count() if lib else 0
    elapsed_ms, _ = device_timed(step_m1, args.steps, 0)
    prof_state["on"] = False
    launches = (lib.segs_launch_count() - launches0) if lib else 0
    views = NV * args.steps * n_ranks
    value = views / (elapsed_ms * 1e-3)
    R_mean = float(np.mean(R_seen)) if R_seen else 0.0

    # ---------------- the batch API (ours only): 4 views in flight --------------------------------
    batch = None
    LANES = args.lanes
    if ours:
        set_wait_mode(LANES)            # lane 0 is the calling thread; measured at 8 ranks x 4 lanes on 32 cores: spinning 6012, sleeping 5769 keyframes/s
        rb = mapper.RasterBatch(dev, lanes=LANES)
        step_images = [torch.empty((3, H, W), dtype=torch.float32, device=dev) for _ in cams]
        dLs_same = [dL] * len(cams)

        def run_batch(pr, dLs, imgs, bkt, lanes=None):
            bkt.zero_()
            return rb.run(pr["means3D"], pr["colors"], pr["opacities"], pr["scales"], pr["rotations"], base["bg"], cams,
                          H, W, scene0.tanfovx, scene0.tanfovy, imgs, dLs, bkt.views, lanes=lanes)

        def step_batch():
            run_batch(base, dLs_same, step_images, gb)
            if use_dist:
                dist.all_reduce(bucket)

        b_ms, _ = device_timed(step_batch, args.steps, warm)
        batch = {"value": round(NV * args.steps * n_ranks / (b_ms * 1e-3), 2), "unit": UNIT, "views_in_flight_per_gpu": LANES,
                 "ms_per_step": round(b_ms / args.steps, 4),
                 "what": "the same step through segs_raster_views (mapper.RasterBatch): the keyframe views of a step on "
                         "concurrent lanes; a capability the reference does not have"}
    else:
        batch = {"value": round(value, 2), "unit": UNIT, "views_in_flight_per_gpu": 1,
                 "what": "the reference has no batch API (legacy default stream, blocking cudaMemcpy): same as `value`"}

    # ---------------- end-to-end: host buffers in, host buffers out --------------------------
    # Per view: dL_dout comes from pinned host memory, the image goes back to pinned host memory.  Per step: the
    # Gaussian parameters come from pinned host memory (N > 1: rank 0 uploads them and broadcasts over NVLink), the
    # accumulated (all-reduced) gradient bucket goes back to pinned host memory (N > 1: on rank 0 only — the replicas
    # hold the same bucket).  Copies ride on two copy streams (H2D / D2H), double-buffered against the compute
    # stream, and are drained inside the timed region.  Identical harness for both arms.
    e2e = e2e_b = None
    set_wait_mode(1)
    if not args.no_e2e:
        e2e, e2e_b = e2e_measure(args, arm, gb, shapes, base, dL, cams, dev, P, W, H, NV, accumulate, barrier, use_dist,
                                 rank, n_ranks, (rb, run_batch) if ours else None, set_wait_mode)
    clocks = sampler.stop() if rank == 0 else None

    mapping = None
    set_wait_mode(LANES)
    if not args.no_mapping:
        mapping = mapping_ours(args, dev, rank, n_ranks, use_dist) if ours else mapping_reference(args, dev)
    configs = None
    if not args.no_configs and rank == 0 and args.gpus == 1:
        configs = sub_configs(args.impl, dev)

    if rank != 0:
        if use_dist:
            dist.barrier()
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
    if lib and prof_state["n"]:
        ms = prof_state["ms"] / prof_state["n"]
        ab = algorithmic_bytes(P, R_mean, N)
        tj = {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        for name, m in zip(STAGES, ms):
            stages_out[name] = {"ms": round(float(m), 4), "algorithmic_MB": round(ab[name] / 1e6, 2),
                                "GBps": round(ab[name] / (m * 1e-3) / 1e9, 1) if m > 0 else None,
                                "hbm_frac": round(ab[name] / (m * 1e-3) / 1e9 / peak_gbs, 4) if m > 0 else None,
                                "dram_traffic_MB_ncu": round(tj[name] / 1e6, 1) if name in tj else None}
        top = STAGES[int(np.argmax(ms))]
        t_top = float(ms[STAGES.index(top)]) * 1e-3
        hbm_achieved = ab[top] / t_top / 1e9
        sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        inst = (tj.get("_inst_executed") or {}).get(top)
        issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9                 # warp instructions / s (G), one per scheduler and cycle
        blend = top.startswith("blend")
        hbm_view = {"achieved": round(hbm_achieved, 1), "peak": peak_gbs, "unit": "GB/s", "frac": round(hbm_achieved / peak_gbs, 4),
                    "peak_source": peak_src, "algorithmic_MB": round(ab[top] / 1e6, 2)}
        if blend and inst:
            achieved = inst / t_top / 1e9
            roofline = {"kernel": top, "bound": "issue", "achieved": round(achieved, 1), "peak": round(issue_peak, 1),
                        "unit": "Gwarp-inst/s", "frac": round(achieved / issue_peak, 4), "traffic": tj.get(top),
                        "warp_instructions_per_launch_ncu": inst,
                        "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (median SM clock of this run)",
                        "pairs_per_s": None,
                        "note": "the blend kernels gather L2-resident 48-byte records and are bound by instruction issue, not "
                                "by HBM (DRAM traffic is a few percent of the algorithmic bytes); `hbm` restates the kernel "
                                "against the copy peak as the contract's byte formula would"}
        else:
            roofline = {"kernel": top, "bound": "hbm", "achieved": hbm_view["achieved"], "peak": peak_gbs, "unit": "GB/s",
                        "frac": hbm_view["frac"], "traffic": tj.get(top), "peak_source": peak_src}
        roofline.update({"hbm": hbm_view, "ms_per_launch": round(t_top * 1e3, 4), "stages": stages_out,
                         "issue_active_pct_ncu": (tj.get("_issue_active_pct") or {}).get(top),
                         "step_share": round(float(ms[STAGES.index(top)] / ms.sum()), 3),
                         "stage_sum_ms": round(float(ms.sum()), 4)})

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": round(elapsed_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: synth({P},{W},{H}) single-view rasterizer forward+backward "
                               f"(BASELINE.md section 3), colours precomputed, scale+quaternion; views one at a time through "
                               f"RasterizeGaussiansCUDA + RasterizeGaussiansBackwardCUDA (SURVEY 8d M1)",
                   "views_per_step_per_gpu": NV, "views_in_flight_per_gpu": 1, "P": P, "W": W, "H": H,
                   "num_rendered_mean": round(R_mean), "parallelism": f"dp{n_ranks} (keyframe views)",
                   "numa_node_rank0": numa,
                   "l2": "working set per view (~0.45 GB state + gradients) exceeds the 126 MB L2; no flush"},
        "ms_per_view": round(elapsed_ms / args.steps / NV, 4),
        "timed_region_s": round(elapsed_ms * 1e-3, 3),
        "clocks": clocks,
        "batch": batch,
        "gpu_launches": int(launches),
    }
    if e2e is not None:
        line["e2e"] = e2e
    if e2e_b is not None:
        line["e2e_batch"] = e2e_b
    if mapping is not None:
        line["mapping"] = mapping
    if configs is not None:
        line["configs"] = configs
    if not ours:
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
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
def e2e_measure(args, arm, gb, shapes, base, dL, cams, dev, P, W, H, NV, accumulate, barrier, use_dist, rank, n_ranks, batch_api,
                set_wait_mode):
    import torch
    import torch.distributed as dist
    from segs_slam_b200 import mapper
    PKEYS = ("means3D", "colors", "opacities", "scales", "rotations")
    bucket = gb.flat
    root = (rank == 0)
    host_in = {k: base[k].cpu().pin_memory() for k in PKEYS}
    host_dL = dL.cpu().pin_memory()
    IMG_RING = 8
    host_img = [torch.empty((3, H, W), dtype=torch.float32).pin_memory() for _ in range(IMG_RING)]
    host_grads = [torch.empty_like(bucket, device="cpu").pin_memory() for _ in range(2)]
    par_bytes = sum(v.numel() * 4 for v in host_in.values())
    # bytes moved over PCIe by THIS rank per step (rank 0 is the one the line reports)
    h2d = par_bytes + NV * host_dL.numel() * 4
    d2h = NV * host_img[0].numel() * 4 + host_grads[0].numel() * 4
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
        """Rank 0 uploads from pinned host memory; with N > 1 the other ranks receive the parameters over NVLink
        (ncclBroadcast on the copy stream) instead of pulling the same 44 MB through the host again."""
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_par_free[slot])
            for k in PKEYS:
                if root or not use_dist:
                    dev_params[slot][k].copy_(host_in[k], non_blocking=True)
                if use_dist:
                    dist.broadcast(dev_params[slot][k], src=0)
            ev_par_ready[slot].record(s_in)

    def upload_dL(slot):
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_dL_free[slot])
            dev_dL[slot].copy_(host_dL, non_blocking=True)
            ev_dL_ready[slot].record(s_in)

    def download_grads(p, done):
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            if root or not use_dist:                 # the all-reduced bucket is identical on every rank
                host_grads[p].copy_(buckets[p].flat, non_blocking=True)
            ev_grads_done[p].record(s_out)

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
            R, color, grads = arm.fwd_bwd(cam, dev_dL[q], dev_params[p])
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
        if use_dist:
            dist.all_reduce(buckets[p].flat)
        done = Ev()
        done.record(cur)
        download_grads(p, done)
        ev_grads_done[p ^ 1].synchronize()            # the host owns the previous step's results

    drain_events = [ev_grads_done, ev_img_done]

    def run(step_fn, n):
        state["views"] = 0
        for i in range(n):
            step_fn(i, i == n - 1)
        for group in drain_events:
            for e_ in group:
                e_.synchronize()
        torch.cuda.synchronize()

    def measure(step_fn, steps):
        run(step_fn, 2)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(step_fn, steps)
        e1.record()
        barrier()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # device events vs host wall clock
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if use_dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return NV * steps * n_ranks / (float(t.item()) * 1e-3)

    e2e_steps = max(3, args.steps)
    pipeline = ("pinned host buffers; per view dL_dout up / image down, per step parameters up / accumulated gradients down; "
                "H2D and D2H on two copy streams, double-buffered against the compute stream; drained inside the timed region")
    if use_dist:
        pipeline += "; N > 1: parameters uploaded by rank 0 and broadcast over NVLink, the all-reduced bucket downloaded by rank 0 only"
    e2e = {"value": round(measure(e2e_step, e2e_steps), 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "views_in_flight_per_gpu": 1,
           "api": "RasterizeGaussiansCUDA + RasterizeGaussiansBackwardCUDA, one view at a time (the M1 path)", "pipeline": pipeline}

    e2e_b = None
    if batch_api is not None:
        rb, run_batch = batch_api
        set_wait_mode(rb.lanes)
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
                    if root or not use_dist:
                        dev_params[slot][k].copy_(host_in[k], non_blocking=True)
                    if use_dist:
                        dist.broadcast(dev_params[slot][k], src=0)
                for v in range(nv):
                    dev_dLs[slot][v].copy_(host_dL, non_blocking=True)
                ev_in_ready[slot].record(s_in)

        def e2e_step_batch(i, last):
            p = i & 1
            if i == 0:
                upload_step(0)
            if not last:
                upload_step(p ^ 1)                    # next step's inputs ride behind this step's kernels
            cur.wait_event(ev_in_ready[p])
            cur.wait_event(ev_out_done[p])            # slot p's images / bucket were downloaded two steps ago
            run_batch(dev_params[p], dev_dLs[p], dev_imgs[p], buckets[p])
            ev_in_free[p].record(cur)
            if use_dist:
                dist.all_reduce(buckets[p].flat)
            done = Ev()
            done.record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                for v in range(nv):
                    host_imgs[p][v].copy_(dev_imgs[p][v], non_blocking=True)
                if root or not use_dist:
                    host_grads[p].copy_(buckets[p].flat, non_blocking=True)
                ev_out_done[p].record(s_out)
            ev_out_done[p ^ 1].synchronize()          # the host owns the previous step's results

        drain_events[:] = [ev_out_done]
        e2e_b = {"value": round(measure(e2e_step_batch, e2e_steps), 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                 "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "views_in_flight_per_gpu": rb.lanes,
                 "api": "mapper.RasterBatch (segs_raster_views), step-level double buffering"}
    return e2e, e2e_b


# ---------------------------------------------------------------------------------------------
MAP_VIEWS, MAP_ANCHORS = 64, 200_000


def _mapping_setup(dev, render_target):
    """BASELINE config 4: 64 keyframes at 1200x680 on a 1.5 m circle around the C3 anchor model.  Ground-truth images =
    renders of a PERTURBED copy of the model, made by `render_target(model_copy, cam)` of the arm under
    test (the two arms' renders agree to 1e-5)."""
    import copy
    import torch
    from segs_slam_b200 import anchor_model
    W, H, fx = 1200, 680, 600.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    model = anchor_model.synth_anchor_model(MAP_ANCHORS, W, H, fx, fx, 1003, device=dev)
    cams = anchor_model.circle_keyframes(MAP_VIEWS, 1.5, (0.0, 0.0, 3.25), tanx, tany, dev)
    gt_model = copy.deepcopy(model)
    g = torch.Generator(device="cpu").manual_seed(1)
    with torch.no_grad():
        gt_model._anchor_feat.add_((torch.randn(gt_model._anchor_feat.shape, generator=g) * 0.05).to(dev))
        gt_model._offset.add_((torch.randn(gt_model._offset.shape, generator=g) * 0.05).to(dev))
    targets = [render_target(gt_model, cam, (W, H, tanx, tany)).clamp_(0.0, 1.0) for cam in cams]
    del gt_model
    return model, cams, targets, (W, H, tanx, tany)


def _render_ours(model, cam, dims):
    import torch
    from segs_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer, generate_neural_gaussians
    from segs_slam_b200.rasterize_points import RasterizeGaussiansfilterCUDA
    W, H, tanx, tany = dims
    dev = model._anchor.device
    e = torch.empty(0, dtype=torch.float32, device=dev)
    with torch.no_grad():
        radii = RasterizeGaussiansfilterCUDA(model.get_anchor(), model.get_scaling()[:, :3].contiguous(), model.get_rotation(), 1.0, e,
                                             cam.world_view_transform_, cam.full_proj_transform_, tanx, tany, H, W, False)
        xyz, color, opacity, scaling, rots, _n, _m = generate_neural_gaussians(cam, model, radii > 0)
        settings = GaussianRasterizationSettings(H, W, tanx, tany, torch.zeros(3, device=dev), 1.0, cam.world_view_transform_,
                                                 cam.full_proj_transform_, 0, cam.camera_center_, False)
        image, _r = GaussianRasterizer(settings)(xyz, torch.zeros_like(xyz), opacity, False, True, True, True, False, e, color,
                                                 scaling, rots, e)
    return image.detach().clone()


def _stats(ms_list):
    s = sorted(ms_list)
    return {"min": round(s[0], 3), "median": round(s[len(s) // 2], 3), "max": round(s[-1], 3)}


def mapping_ours(args, dev, rank, n_ranks, distributed):
    """Mapping keyframes/s (the second half of BASELINE.json's metric): the keyframe-batched step of
    segs_slam_b200.mapper.FusedMapper — per view prefilter -> fused decode -> rasterize -> L1+SSIM+scaling
    regulariser -> backward (segs_mapper_view), ONE NCCL all-reduce of the flat gradient bucket, ONE fused Adam
    launch.  Strong scaling: the 64 views are partitioned across the ranks."""
    import torch
    import torch.distributed as dist
    from segs_slam_b200 import _lib, mapper
    model, cams, targets, (W, H, tanx, tany) = _mapping_setup(dev, _render_ours)
    fm = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lambda_dssim=0.2, lrs=1e-4, lanes=args.lanes)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        fm.step(cams, targets)
    barrier()
    lib = _lib.load()
    l0 = lib.segs_launch_count()
    steps = max(1, args.mapping_steps)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0 = time.perf_counter()
    evs[0].record()
    losses = []
    for s in range(steps):
        losses.append(fm.step(cams, targets))
        evs[s + 1].record()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    per_step = [evs[s].elapsed_time(evs[s + 1]) for s in range(steps)]
    ms = max(evs[0].elapsed_time(evs[steps]), wall)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # the Replica configuration of the loss: + 0.01 * multi_scale_loss over 3 scales (office0.yaml:140-146)
    fq = mapper.FusedMapper(model, H, W, tanx, tany, torch.zeros(3, device=dev), lambda_dssim=0.2, lrs=1e-4, lanes=args.lanes,
                            lambda_frequency_high=0.01, use_multi_resolution=True, freq_scale_num=3)
    for _ in range(2):
        fq.step(cams, targets)
    barrier()
    fsteps = max(2, steps // 4)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tf = time.perf_counter()
    f0.record()
    for _ in range(fsteps):
        fq.step(cams, targets)
    f1.record()
    barrier()
    fms = max(f0.elapsed_time(f1), (time.perf_counter() - tf) * 1e3)
    tfm = torch.tensor([fms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tfm, op=dist.ReduceOp.MAX)
    fms = float(tfm.item())
    with_frequency = {"value": round(MAP_VIEWS * fsteps / (fms * 1e-3), 2), "unit": "keyframes/s", "steps": fsteps,
                      "ms_per_step": round(fms / fsteps, 3),
                      "what": "use_frequency_regularization: 1 — loss + 0.01 * multi_scale_loss(scales 1, 1/2, 1/4) fused into the view "
                              "(csrc/freq.cu), target spectra cached per keyframe"}
    return {"metric": "mapping keyframes/s", "value": round(MAP_VIEWS * steps / (ms * 1e-3), 2), "with_frequency": with_frequency,
            "unit": "keyframes/s", "n_gpus": n_ranks, "views_per_step": MAP_VIEWS, "steps": steps,
            "ms_per_step": round(ms / steps, 3), "ms_per_step_rank0": _stats(per_step), "scaling": "strong",
            "config": {"workload": "C4: 64 keyframes 1200x680, C3 anchor model (200k anchors x 10 offsets, appearance "
                                   "embedding + feature bank); targets = renders of a perturbed copy of the model",
                       "loss": "0.8 L1 + 0.2 (1 - SSIM) + 0.01 scaling regulariser",
                       "optimizer": "fused Adam, one step per 64-keyframe batch", "bucket_MB": round(fm.bucket.flat.numel() * 4 / 1e6, 1),
                       "views_in_flight_per_gpu": fm.lanes},
            "losses": [round(float(x), 5) for x in losses[:4]] + [round(float(losses[-1]), 5)],
            "gpu_launches_per_rank": int(lib.segs_launch_count() - l0)}


def mapping_reference(args, dev):
    """The reference's own mapping iteration, compiled UNMODIFIED (oracle/_ref/_model_ref.so: GaussianRenderer::prefilter_voxel
    + render, GaussianRasterizer, the reference CUDA kernels, loss_utils.h, torch::optim::Adam; the veneer restates only
    the call site src/gaussian_mapper.cpp:870-1006).  Two figures: (A) one keyframe per optimizer step, as the reference
    runs; (B) BASELINE.md's like-for-like baseline: 64 views with gradient accumulation + ONE Adam step.  Single GPU."""
    import torch
    try:
        import model_ref
        if not model_ref.available():
            raise RuntimeError("oracle/_ref/_model_ref.so not built")
        mr = model_ref.load()
    except Exception as exc:
        return mapping_reference_python(args, dev, f"{type(exc).__name__}: {exc}")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True          # the reference arm at its best: cuDNN times its conv2d algorithms once
    bg = torch.zeros(3, device=dev)
    holder = {}

    def render_ref(model, cam, dims):
        W, H, tanx, tany = dims
        m = model_ref.from_model(model, reference_ctor=True)
        fx, fy = 2.0 * math.atan(tanx), 2.0 * math.atan(tany)
        return m.render_image(cam.world_view_transform_, cam.full_proj_transform_, cam.camera_center_, list(cam.t_),
                              list(cam.R_quaternion_), fx, fy, H, W, bg).clone()

    def render_ref_cached(model, cam, dims):
        if "m" not in holder:
            holder["m"] = model_ref.from_model(model, reference_ctor=True)
        W, H, tanx, tany = dims
        fx, fy = 2.0 * math.atan(tanx), 2.0 * math.atan(tany)
        return holder["m"].render_image(cam.world_view_transform_, cam.full_proj_transform_, cam.camera_center_, list(cam.t_),
                                        list(cam.R_quaternion_), fx, fy, H, W, bg).clone()

    model, cams, targets, (W, H, tanx, tany) = _mapping_setup(dev, render_ref_cached)
    holder.clear()
    fovx, fovy = 2.0 * math.atan(tanx), 2.0 * math.atan(tany)
    out = {"metric": "mapping keyframes/s", "unit": "keyframes/s", "n_gpus": 1, "scaling": "strong",
           "config": {"workload": "C4 keyframes, C3 anchor model; the reference's own iteration compiled unmodified "
                                  "(oracle/_ref/_model_ref.so); targets = renders of a perturbed copy of the model",
                      "loss": "0.8 L1 + 0.2 (1 - SSIM) + 0.01 scaling regulariser"}}

    def fresh():
        m = model_ref.from_model(model, reference_ctor=True)
        m.training_setup()
        m.set_learning_rates([1e-4] * m.n_param_groups())
        return m

    def args_of(v):
        cam = cams[v % MAP_VIEWS]
        return (cam.world_view_transform_, cam.full_proj_transform_, cam.camera_center_, list(cam.t_), list(cam.R_quaternion_),
                fovx, fovy, H, W, bg, targets[v % MAP_VIEWS], 0.2)

    # (A) the reference as it runs: one keyframe, one optimizer step, torch::cuda::synchronize() per iteration (:952)
    m = fresh()
    for v in range(10):
        m.train_iteration(*args_of(v), False, False, False, True)
    torch.cuda.synchronize()
    n = MAP_VIEWS * 2
    t0 = time.perf_counter()
    losses = [m.train_iteration(*args_of(v), False, False, False, True)[0] for v in range(n)]
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    out.update({"value": round(n / (ms * 1e-3), 2), "views_per_step": 1, "steps": n, "ms_per_step": round(ms / n, 3),
                "losses": [round(float(x), 5) for x in losses[:4]]})
    del m
    # (A') the same with the Replica configuration of the loss (use_frequency_regularization: 1, office0.yaml:140-146)
    m = fresh()
    m.set_frequency(0.01, True, 3)
    for v in range(5):
        m.train_iteration(*args_of(v), False, False, True, True)
    torch.cuda.synchronize()
    nf = MAP_VIEWS
    t0 = time.perf_counter()
    for v in range(nf):
        m.train_iteration(*args_of(v), False, False, True, True)
    torch.cuda.synchronize()
    msf = (time.perf_counter() - t0) * 1e3
    out["with_frequency"] = {"value": round(nf / (msf * 1e-3), 2), "unit": "keyframes/s", "steps": nf, "ms_per_step": round(msf / nf, 3),
                             "what": "use_frequency_regularization: 1 — the reference's own multi_scale_loss (3 scales, lambda 0.01)"}
    del m
    # (B) 64 reference views with gradient accumulation + ONE Adam step (the 1-GPU equivalent of the batched step)
    m = fresh()
    for v in range(4):
        m.backward_view(*args_of(v), 1.0 / MAP_VIEWS)
    m.optimizer_step()
    torch.cuda.synchronize()
    steps = max(1, min(3, args.mapping_steps))
    per = []
    for s in range(steps):
        t0 = time.perf_counter()
        for v in range(MAP_VIEWS):
            m.backward_view(*args_of(v), 1.0 / MAP_VIEWS)
        m.optimizer_step()
        torch.cuda.synchronize()
        per.append((time.perf_counter() - t0) * 1e3)
    tot = sum(per)
    out["baseline_B"] = {"value": round(MAP_VIEWS * steps / (tot * 1e-3), 2), "unit": "keyframes/s", "views_per_step": MAP_VIEWS,
                         "steps": steps, "ms_per_step": _stats(per),
                         "what": "64 reference views, gradients accumulated in .grad, ONE torch::optim::Adam step (BASELINE.md "
                                 "section 2 baseline B)"}
    return out


def mapping_reference_python(args, dev, why):
    """Fallback when _model_ref.so is absent: the reference's iteration restated over the reference kernels and the ATen
    sequences of its host code (tests/ref_mapper.py)."""
    import torch
    try:
        import ref_mapper
        import decode_oracle
    except Exception as exc:                                   # oracle/_ref absent
        return {"unavailable": f"{why}; {type(exc).__name__}: {exc}"}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    model, cams, targets, (W, H, tanx, tany) = _mapping_setup(dev, _render_ours)
    model.cfg = decode_oracle.DecodeConfig(appearance_dim=model.appearance_dim, use_feat_bank=model.use_feat_bank)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, eps=1e-15)
    bg = torch.zeros(3, device=dev)
    n = MAP_VIEWS
    for v in range(10):
        ref_mapper.iteration(model, cams[v], targets[v], H, W, tanx, tany, bg, opt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = [ref_mapper.iteration(model, cams[v % MAP_VIEWS], targets[v % MAP_VIEWS], H, W, tanx, tany, bg, opt) for v in range(n)]
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    return {"metric": "mapping keyframes/s", "value": round(n / (ms * 1e-3), 2), "unit": "keyframes/s", "n_gpus": 1,
            "views_per_step": 1, "steps": n, "ms_per_step": round(ms / n, 3), "scaling": "strong",
            "config": {"workload": "C4 keyframes, C3 anchor model; reference kernels + ATen restatement (tests/ref_mapper.py); "
                                   f"compiled reference unavailable: {why}"},
            "losses": [round(float(x), 5) for x in losses[:4]]}


# ---------------------------------------------------------------------------------------------
def sub_configs(impl, dev):
    """Per-config sub-lines (N = 1): M1 at C1 and C5, decode + rasterize forward+backward at C3 (BASELINE config 3),
    distCUDA2 at P = 1e3 / 1e4 / 1e5.  20 warm-up + timed iterations each, host-observed with a device synchronize on both
    sides (one view at a time, which is what the entry points are)."""
    import numpy as np
    import torch
    from segs_slam_b200 import synth
    out = {}
    sync = torch.cuda.synchronize
    for name, n_timed in (("C1", 100), ("C5", 20)):
        try:
            scene = synth.config(name)
            base = scene.to_torch(dev)
            cam = {k: base[k] for k in ("viewmatrix", "projmatrix", "campos")}
            arm = Arm(impl, scene, base, dev)
            R = 0
            for _ in range(20 if name == "C1" else 5):
                R = arm.fwd_bwd(cam, base["dL_dout"])[0]
            ms = timed(lambda: arm.fwd_bwd(cam, base["dL_dout"]), n_timed, sync)
            out[name] = {"workload": f"synth({scene.P},{scene.W},{scene.H}) single view fwd+bwd (M1)", "num_rendered": int(R),
                         "ms": _stats(ms), "mean_ms": round(float(np.mean(ms)), 4),
                         "iterations_per_s": round(1e3 / float(np.mean(ms)), 2)}
            del arm, base, scene
            torch.cuda.empty_cache()
        except Exception as exc:
            out[name] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    try:
        out["C3"] = _c3_decode_raster(impl, dev)
    except Exception as exc:
        out["C3"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    torch.cuda.empty_cache()
    knn = {}
    rng = np.random.default_rng(5)
    for Pn in (1000, 10_000, 100_000):
        pts = torch.from_numpy(rng.uniform(-2.0, 2.0, (Pn, 3)).astype(np.float32)).to(dev)
        try:
            if impl == "ours":
                from segs_slam_b200 import rasterize_points as rp
                fn = lambda: rp.distCUDA2(pts)
            else:
                import refimpl
                fn = lambda: refimpl.knn(pts)
            for _ in range(10):
                fn()
            ms = timed(fn, 50, sync)
            knn[str(Pn)] = {"ms": _stats(ms), "points_per_s": round(Pn / (float(np.median(ms)) * 1e-3))}
        except Exception as exc:
            knn[str(Pn)] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    out["distCUDA2"] = knn
    return out


def _c3_decode_raster(impl, dev):
    """BASELINE config 3: 200k anchors x 10 offsets with appearance embedding: prefilter -> decode -> rasterize and the
    backward of a fixed dL_dimage down to the anchor parameters and MLP weights."""
    import numpy as np
    import torch
    from segs_slam_b200 import anchor_model
    W, H, fx = 1200, 680, 600.0
    tanx, tany = W / (2 * fx), H / (2 * fx)
    model = anchor_model.synth_anchor_model(MAP_ANCHORS, W, H, fx, fx, 1003, device=dev)
    cam = anchor_model.Keyframe(np.eye(3), np.zeros(3), tanx, tany, dev)
    g = torch.Generator(device="cpu").manual_seed(3)
    dL = torch.randn(3, H, W, generator=g).to(dev)
    bg = torch.zeros(3, device=dev)
    sync = torch.cuda.synchronize
    info = {}
    if impl == "ours":
        from segs_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer, generate_neural_gaussians
        from segs_slam_b200.rasterize_points import RasterizeGaussiansfilterCUDA
        e = torch.empty(0, dtype=torch.float32, device=dev)
        params = [p for p in model.parameters() if p.requires_grad]

        def run():
            with torch.no_grad():
                radii = RasterizeGaussiansfilterCUDA(model.get_anchor(), model.get_scaling()[:, :3].contiguous(), model.get_rotation(),
                                                     1.0, e, cam.world_view_transform_, cam.full_proj_transform_, tanx, tany, H, W, False)
            xyz, color, opacity, scaling, rots, _n, _m = generate_neural_gaussians(cam, model, radii > 0)
            settings = GaussianRasterizationSettings(H, W, tanx, tany, bg, 1.0, cam.world_view_transform_, cam.full_proj_transform_, 0,
                                                     cam.camera_center_, False)
            means2D = torch.zeros_like(xyz, requires_grad=True)
            image, _r = GaussianRasterizer(settings)(xyz, means2D, opacity, False, True, True, True, False, e, color, scaling, rots, e)
            torch.autograd.grad(image, params, dL, allow_unused=True)
            info["P"] = int(xyz.size(0))
    else:
        import model_ref
        torch.backends.cuda.matmul.allow_tf32 = False
        m = model_ref.from_model(model, reference_ctor=True)
        fovx, fovy = 2.0 * math.atan(tanx), 2.0 * math.atan(tany)

        def run():
            m.render_fwd_bwd(cam.world_view_transform_, cam.full_proj_transform_, cam.camera_center_, list(cam.t_),
                             list(cam.R_quaternion_), fovx, fovy, H, W, bg, dL)
    for _ in range(5):
        run()
    ms = timed(run, 20, sync)
    out = {"workload": "200k anchors x 10 offsets (appearance 32, feature bank) at 1200x680: prefilter + decode + rasterize, "
                       "forward + backward (tensor-level API)", "gaussians": info.get("P"), "ms": _stats(ms),
           "mean_ms": round(float(np.mean(ms)), 4), "iterations_per_s": round(1e3 / float(np.mean(ms)), 2)}
    if impl == "ours":
        # the decode kernels of round 1 (thread = anchor, first layers on tcgen05 only) beside the default (variant 2: both
        # layers and the weight gradients on tcgen05, tiles of visible anchors); include/segs_raster.h: segs_decode_set_variant
        from segs_slam_b200 import _lib
        lib = _lib.load()
        out["decode_variant"] = int(lib.segs_decode_get_variant())
        try:
            _lib.check(lib.segs_decode_set_variant(1))
            for _ in range(3):
                run()
            ms1 = timed(run, 20, sync)
            out["decode_variant_1"] = {"ms": _stats(ms1), "mean_ms": round(float(np.mean(ms1)), 4)}
        finally:
            _lib.check(lib.segs_decode_set_variant(out["decode_variant"]))
    return out


def _bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: pin this rank's host threads (and therefore its pinned staging buffers, first-touch) to the
    NUMA node its GPU hangs off, like `numactl --cpunodebind --membind` per rank.  Returns the node, or None when the
    platform does not expose one."""
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
