"""Keyframe-batched, data-parallel mapping step (BASELINE.json config 4).

The reference optimises ONE keyframe per optimizer step on one GPU
(/root/reference/src/gaussian_mapper.cpp:823-1032: useOneRandomSlidingWindowKeyframe -> render ->
loss.backward() -> optimizer step).  The batched step here is the new construct north_star asks
for: a batch of B keyframe views is partitioned across the G ranks of one box (rank g renders views
g, g+G, ...), every rank accumulates the gradients of its views locally in FP32, ONE all-reduce
(NCCL over NVLink on the GPU box; gloo in the CPU tests) sums the flat gradient bucket, and every
rank applies the same optimizer step to its replica — the replicas stay bit-identical without any
parameter broadcast.  Its 1-GPU equivalent (the denominator of scaling efficiency) is the same
function with world_size 1: sequential rendering of all B views with gradient accumulation.

Single-view rendering does not shard (sort -> ranges -> blend of one image is one dependency
chain): replicas only.

Only host-side orchestration lives here.  Two ways to do the per-view work:
  * `mapping_step` + `make_render_loss`: the autograd composition of the tensor-level API (prefilter -> decode ->
    rasterize -> loss), reference-shaped, one interpreter-issued launch sequence per view;
  * `FusedMapper` (anchor model) / `RasterBatch` (explicit Gaussians): the same views issued from C++
    (`segs_mapper_views` / `segs_raster_views`, csrc/mapper_view.cu) on concurrent lanes, gradients accumulated straight
    into the flat bucket, one fused Adam launch per step — what bench.py measures.
"""
from __future__ import annotations

from typing import Callable, Iterable, Sequence

import torch
import torch.distributed as dist


def partition_views(n_views: int, world_size: int, rank: int) -> list[int]:
    """Views rendered by `rank`: rank, rank + G, rank + 2G, ...  (disjoint, covers 0..n_views-1)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_views, world_size))


class GradBucket:
    """Flat FP32 bucket over the trainable tensors, in a fixed order: what one all-reduce moves.

    At A = 200k anchors: anchor 3 + offset 30 + feat 32 + scaling 6 = 71 floats per anchor
    (56.8 MB) plus ~8.6k MLP floats."""

    def __init__(self, params: Sequence[torch.Tensor], tail: int = 0):
        """tail: extra floats behind the gradients in the same allocation (`self.tail`): the loss sum and the
        densification-statistics delta ride in the SAME all-reduce as the gradients (`self.everything`)."""
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.everything = torch.zeros(total + int(tail), dtype=torch.float32, device=dev)
        self.flat = self.everything[:total]
        self.tail = self.everything[total:]
        self.views = []
        off = 0
        for p, n in zip(self.params, self.sizes):
            self.views.append(self.flat[off:off + n].view(p.shape))
            off += n

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads: Iterable[torch.Tensor | None]):
        """flat += grads (None = no gradient for that tensor in this view)."""
        for v, g in zip(self.views, grads):
            if g is not None:
                v.add_(g)

    def all_reduce_mean(self, n_views_total: int, group=None):
        """Sum over ranks, divide by the batch size: the gradient of the mean loss over the batch."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / float(n_views_total))

    def scatter_to_grads(self):
        """Point every parameter's .grad at its slice of the bucket (no copies)."""
        for p, v in zip(self.params, self.views):
            p.grad = v


def mapping_step(params: Sequence[torch.Tensor], render_loss: Callable[[int], torch.Tensor], n_views: int,
                 optimizer: torch.optim.Optimizer | None = None, bucket: GradBucket | None = None, group=None):
    """One keyframe-batched optimisation step.  Returns (mean loss over the batch, bucket).

    render_loss(view_index) -> scalar loss of that keyframe view (differentiable w.r.t. `params`).
    """
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    bucket = bucket or GradBucket(params)
    bucket.zero_()
    loss_sum = torch.zeros((), dtype=torch.float32, device=bucket.flat.device)
    for v in partition_views(n_views, world, rank):
        loss = render_loss(v)
        grads = torch.autograd.grad(loss, bucket.params, allow_unused=True)
        bucket.accumulate(grads)
        loss_sum += loss.detach().float()
    if world > 1:
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)
    bucket.all_reduce_mean(n_views, group)
    if optimizer is not None:
        bucket.scatter_to_grads()
        optimizer.step()
    return loss_sum / float(n_views), bucket


def make_render_loss(pc, cameras, targets, image_height: int, image_width: int, tanfovx: float, tanfovy: float,
                     bg: torch.Tensor, loss: str = "l1", lambda_dssim: float = 0.2, scaling_reg_weight: float = 0.01,
                     row_masks=None):
    """Per-view work of the mapper on this package's kernels: anchor prefilter
    (RasterizeGaussiansfilterCUDA, gaussian_renderer.cpp:131-199) -> fused decode
    (generate_neural_gaussians, :214-334) -> rasterize (GaussianRasterizer, :40-127) -> loss against the
    keyframe image: loss="l1" is loss_utils::l1_loss alone; loss="l1_ssim" is the mapper's
    (1-l)*L1 + l*(1-SSIM) + 0.01*scaling.prod(1).mean() (gaussian_mapper.cpp:908-925) on the fused loss
    kernels (loss_utils.py).  The frequency terms (:927-942, off by default) are §8f 'next'.

    cameras[v] carries world_view_transform_, full_proj_transform_, camera_center_, t_, R_quaternion_."""
    from . import GaussianRasterizationSettings, GaussianRasterizer, generate_neural_gaussians
    from .rasterize_points import RasterizeGaussiansfilterCUDA
    from . import loss_utils
    if loss not in ("l1", "l1_ssim"):
        raise ValueError(f"unknown loss '{loss}'")

    def render_loss(v: int) -> torch.Tensor:
        cam = cameras[v]
        dev = bg.device
        e = torch.empty(0, dtype=torch.float32, device=dev)
        anchor = pc.get_anchor()
        with torch.no_grad():
            scal = pc.get_scaling()[:, :3].contiguous()
            rot = torch.nn.functional.normalize(pc._rotation) if hasattr(pc, "_rotation") else \
                torch.tensor([1.0, 0.0, 0.0, 0.0], device=dev).expand(anchor.size(0), 4).contiguous()
            radii = RasterizeGaussiansfilterCUDA(anchor, scal, rot, 1.0, e, cam.world_view_transform_,
                                                 cam.full_proj_transform_, tanfovx, tanfovy, image_height, image_width,
                                                 False)
            visible = radii > 0
        xyz, color, opacity, scaling, rots, _nop, _mask = generate_neural_gaussians(cam, pc, visible)
        settings = GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, bg, 1.0,
                                                 cam.world_view_transform_, cam.full_proj_transform_, 0,
                                                 cam.camera_center_, False)
        means2D = torch.zeros_like(xyz, requires_grad=True)
        image, _radii = GaussianRasterizer(settings)(xyz, means2D, opacity, False, True, True, True, False, e, color,
                                                     scaling, rots, e)
        if loss == "l1":
            return (image - targets[v]).abs().mean()
        photometric = loss_utils.l1_ssim_loss(image, targets[v], lambda_dssim, None if row_masks is None else row_masks[v])[0]
        return photometric + loss_utils.scaling_reg(scaling, scaling_reg_weight)

    return render_loss


def _one_allocation(*tensors):
    """The flat tensor covering `tensors` when they are consecutive slices of one allocation, else None."""
    ts = [t for t in tensors if t is not None]
    if not ts or any(not t.is_contiguous() for t in ts):
        return None
    base = ts[0]._base if ts[0]._base is not None else None
    if base is None or base.dim() != 1 or any(t._base is not base for t in ts):
        return None
    first = (ts[0].data_ptr() - base.data_ptr()) // 4
    end = first
    for t in ts:
        if t.dtype != torch.float32 or (t.data_ptr() - base.data_ptr()) // 4 != end:
            return None
        end += t.numel()
    return base[first:end]


def exchange_step(grad_flat: torch.Tensor, loss_accum: torch.Tensor, stats_delta: torch.Tensor | None = None,
                  stats: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """The exchange of one batched step between the ranks: the flat gradient bucket, the loss sum and —
    when densification statistics are kept — this step's statistics delta are summed over the ranks, and the delta is
    added to the running accumulators on every rank (so every replica holds the statistics of ALL views and takes the
    same densification decisions).  When the three live back to back in one allocation (FusedMapper: GradBucket with a
    tail) that is ONE all-reduce; otherwise one per tensor.  Device-agnostic host logic (NCCL on the GPU box, gloo in
    tests/test_mapper_cpu.py).  -> loss sum over all ranks (a new tensor; `loss_accum` then holds the sum too when fused)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        fused = _one_allocation(grad_flat, loss_accum.view(-1), stats_delta)
        if fused is not None:
            dist.all_reduce(fused, op=dist.ReduceOp.SUM, group=group)
            loss = loss_accum.clone()
        else:
            loss = loss_accum.clone()
            dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
            if stats_delta is not None:
                dist.all_reduce(stats_delta, op=dist.ReduceOp.SUM, group=group)
    else:
        loss = loss_accum.clone()
    if stats_delta is not None and stats is not None:
        stats.add_(stats_delta)
    return loss


class FusedMapper:
    """The keyframe-batched mapping step with the per-view work issued from C++ (`segs_mapper_view`,
    csrc/mapper_view.cu): prefilter -> decode -> rasterize -> L1+SSIM (+ scaling regulariser) -> backward, the
    parameter gradients accumulated straight into the flat bucket; then ONE all-reduce and ONE fused Adam launch
    (`segs_adam_step`) that also applies the 1/B scale and clears the bucket.

    Same mathematics as `mapping_step(params, make_render_loss(..., loss="l1_ssim"), ...)` — the autograd
    composition of the tensor-level API, which stays the reference-shaped interface and is what
    tests/test_mapper_gpu.py checks this class against — without ~150 interpreter-issued launches per view.

    `pc` carries the reference GaussianModel's members (anchor_model.AnchorModel).  Trainable tensors, in bucket
    order: _anchor, _offset, _anchor_feat, _scaling, then the MLP weights in segs_decode_params order
    (gaussian_model.cpp:620-872 gives each its own learning rate; `lrs` follows the same order)."""

    def __init__(self, pc, image_height: int, image_width: int, tanfovx: float, tanfovy: float, bg: torch.Tensor,
                 lambda_dssim: float = 0.2, scaling_reg_weight: float = 0.01, lrs=1e-4, eps: float = 1e-15, group=None,
                 lanes: int = 4, statistics: bool = False, lambda_frequency_high: float = 0.0,
                 use_multi_resolution: bool = True, freq_scale_num: int = 3, cache_freq_targets: int = 256):
        import ctypes as C
        from . import _lib
        from .gaussian_renderer import _weights
        from .optim import FusedAdam
        self._C, self._lib_mod, self.lib = C, _lib, _lib.load()
        self.pc, self.group = pc, group
        self.H, self.W, self.tanfovx, self.tanfovy = int(image_height), int(image_width), float(tanfovx), float(tanfovy)
        self.bg = bg.to(torch.float32).contiguous()
        self.lambda_dssim, self.scaling_reg_weight = float(lambda_dssim), float(scaling_reg_weight)
        # frequency regularisation (gaussian_mapper.cpp:930-945; Replica yamls: 0.01, multi-resolution, 3 scales); the
        # target magnitudes |fft2(D_s gt)| are per keyframe and cached (up to `cache_freq_targets` keyframes)
        self.lambda_frequency_high = float(lambda_frequency_high)
        self.use_multi_resolution, self.freq_scale_num = bool(use_multi_resolution), max(1, min(4, int(freq_scale_num)))
        self._freq_cache, self._freq_cache_max = {}, int(cache_freq_targets)
        if not pc._anchor.is_cuda:
            raise RuntimeError("segs_slam_b200 has no CPU path: the model must live on a CUDA device")
        self.weights = _weights(pc)                                   # 18 entries, None = absent
        self.params = [pc._anchor, pc._offset, pc._anchor_feat, pc._scaling] + [w for w in self.weights if w is not None]
        for p in self.params:
            if not p.is_contiguous():
                raise RuntimeError("FusedMapper: parameters must be contiguous")
        self.statistics = bool(statistics)
        # gradients | loss sum | statistics delta in ONE allocation: one all-reduce per step moves all three
        self.bucket = GradBucket(self.params, tail=1 + (22 * pc._anchor.size(0) if self.statistics else 0))
        self.optimizer = FusedAdam(self.bucket, lrs, eps=eps)
        v = iter(self.bucket.views[4:])
        self._wgrad_views = [next(v) if w is not None else None for w in self.weights]
        self.loss_accum = self.bucket.tail[:1].view(())
        # densification statistics (GaussianModel::training_statis, gaussian_model.cpp:1459-1503): running accumulators
        # opacity_accum [A,1], anchor_demon [A,1], offset_gradient_accum [A*10,1], offset_denom [A*10,1] and the per-step
        # delta the views add to; the delta is all-reduced (the second, small all-reduce of SURVEY §8e) so that the
        # replicas' statistics — and therefore their densification decisions — stay identical
        if self.statistics:
            A_ = pc._anchor.size(0)
            self.stats = torch.zeros(22 * A_, dtype=torch.float32, device=pc._anchor.device)
            self.stats_delta = self.bucket.tail[1:]
            cut = lambda t: (t[:A_].view(A_, 1), t[A_:2 * A_].view(A_, 1), t[2 * A_:12 * A_].view(10 * A_, 1), t[12 * A_:].view(10 * A_, 1))
            self.opacity_accum, self.anchor_demon, self.offset_gradient_accum, self.offset_denom = cut(self.stats)
            self._stat_delta_views = cut(self.stats_delta)
        # lanes: concurrent views per rank (segs_mapper_views): lane = workspace + stream + persistent host thread.
        # The sorts / decode / read-backs of one view overlap the issue-bound blend kernels of another; gradients
        # are then accumulated with RED.ADD (one more order-dependent sum on top of the rasterizer backward's own).
        self.lanes = max(1, int(lanes))
        self._wss = []
        for _ in range(self.lanes):
            ws = C.c_void_p()
            _lib.check(self.lib.segs_workspace_create(C.byref(ws)))
            self._wss.append(ws)
        self._ws = self._wss[0]
        self._streams = [torch.cuda.Stream(device=pc._anchor.device) for _ in range(self.lanes)] if self.lanes > 1 else []
        self.last_result = None
        self._dirty = False                                           # the bucket starts zeroed

    def __del__(self):
        try:
            for ws in getattr(self, "_wss", []):
                self.lib.segs_workspace_destroy(ws)
            self._wss, self._ws = [], None
        except Exception:
            pass

    def workspace_bytes(self) -> int:
        return sum(int(self.lib.segs_workspace_bytes(ws)) for ws in self._wss)

    def _prepare(self):
        """Per-step derived tensors (the parameters do not change between the views of one step)."""
        pc = self.pc
        with torch.no_grad():
            self._scaling = torch.exp(pc._scaling)                    # get_scaling
            self._fscales = self._scaling[:, :3].contiguous()
            if hasattr(pc, "_rotation"):
                self._frot = torch.nn.functional.normalize(pc._rotation).contiguous()
            else:
                self._frot = torch.tensor([1.0, 0.0, 0.0, 0.0], device=pc._anchor.device).repeat(pc._anchor.size(0), 1)
        C, L = self._C, self._lib_mod
        ptr = lambda t: None if t is None else t.data_ptr()
        cfg = (int(getattr(pc, "appearance_dim", 0)), int(bool(getattr(pc, "use_feat_bank", False))),
               int(bool(getattr(pc, "add_opacity_dist", False))), int(bool(getattr(pc, "add_cov_dist", False))),
               int(bool(getattr(pc, "add_color_dist", False))))
        self._dparams = L.DecodeParams(*[ptr(w) for w in self.weights], *cfg)
        self._dgrads = L.DecodeGrads(*[ptr(g) for g in self._wgrad_views])
        a = L.MapperViewArgs()
        a.A = pc._anchor.size(0)
        a.anchor, a.anchor_feat, a.offset = pc._anchor.data_ptr(), pc._anchor_feat.data_ptr(), pc._offset.data_ptr()
        a.scaling, a.scaling_is_log = self._scaling.data_ptr(), 1
        a.filter_scales, a.filter_rotations = self._fscales.data_ptr(), self._frot.data_ptr()
        a.params = C.pointer(self._dparams)
        a.width, a.height, a.tan_fovx, a.tan_fovy = self.W, self.H, self.tanfovx, self.tanfovy
        a.background = self.bg.data_ptr()
        a.lambda_dssim, a.scaling_reg_weight = self.lambda_dssim, self.scaling_reg_weight
        a.lambda_frequency_high, a.use_multi_resolution = self.lambda_frequency_high, int(self.use_multi_resolution)
        a.freq_scale_num = self.freq_scale_num
        v = self.bucket.views
        a.grad_anchor, a.grad_offset, a.grad_anchor_feat, a.grad_scaling = (v[0].data_ptr(), v[1].data_ptr(),
                                                                             v[2].data_ptr(), v[3].data_ptr())
        a.grad_params = C.pointer(self._dgrads)
        a.loss_accum = self.loss_accum.data_ptr()
        if self.statistics:
            (a.stat_opacity_accum, a.stat_anchor_demon, a.stat_offset_gradient_accum,
             a.stat_offset_denom) = [t.data_ptr() for t in self._stat_delta_views]
        self._args = a

    def _fill(self, a, cam, target, row_mask, keep):
        C = self._C
        C.memmove(C.byref(a), C.byref(self._args), C.sizeof(self._args))
        a.viewmatrix, a.projmatrix = cam.world_view_transform_.data_ptr(), cam.full_proj_transform_.data_ptr()
        a.campos = cam.camera_center_.data_ptr()
        t, q = cam.t_, cam.R_quaternion_
        pose = (C.c_float * 7)(float(t[0]), float(t[1]), float(t[2]), float(q[0]), float(q[1]), float(q[2]), float(q[3]))
        keep.append(pose)
        a.pose = pose
        a.gt_image = target.data_ptr()
        a.row_mask = None if row_mask is None else row_mask.data_ptr()
        a.gt_freq_mag = None
        if self.lambda_frequency_high != 0.0 and self._freq_cache_max > 0:
            a.gt_freq_mag = self._freq_target(target, row_mask).data_ptr()

    def _freq_target(self, target, row_mask):
        """|fft2(D_s (gt * mask))| of a keyframe image, computed once and kept while the image tensor lives at the same
        address (segs_freq_target); the views of later steps reuse it."""
        from . import loss_utils
        key = (target.data_ptr(), None if row_mask is None else row_mask.data_ptr(), target._version)
        hit = self._freq_cache.get(key)
        if hit is None:
            if len(self._freq_cache) >= self._freq_cache_max:
                self._freq_cache.pop(next(iter(self._freq_cache)))
            scales = [1.0 / 2 ** i for i in range(self.freq_scale_num)] if self.use_multi_resolution else [1.0]
            hit = loss_utils.freq_target(target, scales, row_mask)
            self._freq_cache[key] = hit
        return hit

    def render_views(self, cams, targets, row_masks=None):
        """A batch of views on `self.lanes` concurrent lanes: accumulates gradients and the loss."""
        C, L = self._C, self._lib_mod
        n = len(cams)
        if n == 0:
            return []
        arr, res, keep = (L.MapperViewArgs * n)(), (L.MapperViewResult * n)(), []
        for i in range(n):
            self._fill(arr[i], cams[i], targets[i], None if row_masks is None else row_masks[i], keep)
        dev = self.bucket.flat.device
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream().cuda_stream
            lanes = min(self.lanes, n)
            wss = (C.c_void_p * lanes)(*[w.value for w in self._wss[:lanes]])
            sts = (C.c_void_p * lanes)(*([main] if lanes == 1 else [s.cuda_stream for s in self._streams[:lanes]]))
            L.check(self.lib.segs_mapper_views(n, arr, res, lanes, wss, sts, main))
        self.last_result = res[n - 1]
        return list(res)

    def render_view(self, cam, target: torch.Tensor, row_mask: torch.Tensor | None = None, image_out=None,
                    loss_terms_out=None):
        """One view: accumulates gradients and the loss.  `_prepare()` must have run this step."""
        C, L = self._C, self._lib_mod
        a, keep = L.MapperViewArgs(), []         # a private copy: the optional outputs must not leak into later batches
        self._fill(a, cam, target, row_mask, keep)
        a.image_out = None if image_out is None else image_out.data_ptr()
        a.loss_terms_out = None if loss_terms_out is None else loss_terms_out.data_ptr()
        res = L.MapperViewResult()
        dev = self.bucket.flat.device
        with torch.cuda.device(dev):
            L.check(self.lib.segs_mapper_view(self._ws, C.byref(a), C.byref(res), torch.cuda.current_stream().cuda_stream))
        self.last_result = res
        return res

    def step(self, cameras, targets, row_masks=None, n_views: int | None = None, optimize: bool = True):
        """One batched step over `cameras` (all ranks pass the full list; each renders its partition).
        -> mean loss over the batch (device scalar)."""
        n_views = len(cameras) if n_views is None else n_views
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if world > 1 else 0
        self._prepare()
        self.bucket.tail.zero_()                                     # loss sum + this step's statistics delta
        if self._dirty:
            self.bucket.zero_()                                      # normally cleared by the previous Adam launch
        self._dirty = True
        mine = partition_views(n_views, world, rank)
        self.render_views([cameras[v] for v in mine], [targets[v] for v in mine],
                          None if row_masks is None else [row_masks[v] for v in mine])
        loss = exchange_step(self.bucket.flat, self.loss_accum, self.stats_delta if self.statistics else None,
                             self.stats if self.statistics else None, self.group)
        if optimize:
            self.optimizer.step(grad_scale=1.0 / float(n_views), zero_grad=True)
            self._dirty = False
        return loss / float(n_views)


    # ---- densification (GaussianModel::adjust_anchor, gaussian_model.cpp:1705-1762) -------------------------------------
    def adjust_anchor(self, check_interval: int = 100, success_threshold: float = 0.8, grad_threshold: float = 0.0002,
                      min_opacity: float = 0.005, rands=None, generator=None, **model):
        """Grow and prune the anchor set from the running statistics (`statistics=True`), exactly as the reference does
        every `update_interval` iterations (src/gaussian_mapper.cpp:966-971), then rebuild the flat gradient bucket, the
        fused Adam's moments (old rows keep theirs, new rows start from zero, the step count is kept — :1663-1681) and the
        statistics for the new anchor count.  The statistics are all-reduced sums and the random draw comes from
        `generator` (default: a generator every rank seeds identically), so every replica takes the same decisions
        without a broadcast.  `model`: n_offsets, update_depth, update_init_factor, update_hierachy_factor, voxel_size
        (defaults: GaussianModelParams, include/gaussian_parameters.h:35-41).  -> (anchors before, anchors after)"""
        from . import densify
        if not self.statistics:
            raise RuntimeError("FusedMapper.adjust_anchor needs statistics=True")
        pc, opt = self.pc, self.optimizer
        dev = pc._anchor.device
        A0 = pc._anchor.size(0)
        if generator is None and rands is None:
            if getattr(self, "_densify_generator", None) is None:
                self._densify_generator = torch.Generator(device=dev)
                self._densify_generator.manual_seed(int(getattr(self, "densify_seed", 20260101)))
            generator = self._densify_generator
        for k in ("n_offsets", "update_depth", "update_init_factor", "update_hierachy_factor", "voxel_size"):
            if k not in model and hasattr(pc, k):
                model[k] = getattr(pc, k)
        with torch.no_grad():
            st = {"_anchor": pc._anchor.detach(), "_offset": pc._offset.detach(), "_anchor_feat": pc._anchor_feat.detach(),
                  "_scaling": pc._scaling.detach(),
                  "_opacity": pc._opacity.detach() if hasattr(pc, "_opacity") else torch.zeros(A0, 1, device=dev),
                  "_rotation": pc._rotation.detach() if hasattr(pc, "_rotation") else
                  torch.tensor([1.0, 0.0, 0.0, 0.0], device=dev).repeat(A0, 1),
                  "opacity_accum": self.opacity_accum.clone(), "anchor_demon": self.anchor_demon.clone(),
                  "offset_gradient_accum": self.offset_gradient_accum.clone(), "offset_denom": self.offset_denom.clone()}
            off = 0
            for name, p, n in zip(("_anchor", "_offset", "_anchor_feat", "_scaling"), self.params[:4], self.bucket.sizes[:4]):
                st["m_" + name] = opt.exp_avg[off:off + n].view(p.shape).clone()
                st["v_" + name] = opt.exp_avg_sq[off:off + n].view(p.shape).clone()
                off += n
            tail_m, tail_v = opt.exp_avg[off:].clone(), opt.exp_avg_sq[off:].clone()          # the MLP tensors' moments
            densify.adjust_anchor(st, check_interval, success_threshold, grad_threshold, min_opacity, rands, generator, **model)
            for name in ("_anchor", "_offset", "_anchor_feat", "_scaling", "_opacity", "_rotation"):
                if hasattr(pc, name):
                    old = getattr(pc, name)
                    setattr(pc, name, torch.nn.Parameter(st[name].contiguous(), requires_grad=old.requires_grad))
            A1 = pc._anchor.size(0)
            self.params = [pc._anchor, pc._offset, pc._anchor_feat, pc._scaling] + [w for w in self.weights if w is not None]
            self.bucket = GradBucket(self.params, tail=1 + 22 * A1)
            opt.bucket = self.bucket
            self.loss_accum = self.bucket.tail[:1].view(())
            opt.exp_avg = torch.cat([st["m_" + n].reshape(-1) for n in ("_anchor", "_offset", "_anchor_feat", "_scaling")] + [tail_m])
            opt.exp_avg_sq = torch.cat([st["v_" + n].reshape(-1) for n in ("_anchor", "_offset", "_anchor_feat", "_scaling")] + [tail_v])
            v = iter(self.bucket.views[4:])
            self._wgrad_views = [next(v) if w is not None else None for w in self.weights]
            self.stats = torch.cat([st["opacity_accum"].reshape(-1), st["anchor_demon"].reshape(-1),
                                    st["offset_gradient_accum"].reshape(-1), st["offset_denom"].reshape(-1)]).contiguous()
            self.stats_delta = self.bucket.tail[1:]
            cut = lambda t: (t[:A1].view(A1, 1), t[A1:2 * A1].view(A1, 1), t[2 * A1:12 * A1].view(10 * A1, 1), t[12 * A1:].view(10 * A1, 1))
            self.opacity_accum, self.anchor_demon, self.offset_gradient_accum, self.offset_denom = cut(self.stats)
            self._stat_delta_views = cut(self.stats_delta)
            self._dirty = False                                                              # the new bucket starts zeroed
            self.last_densify = {"growing": st.get("_growing_report"), "prune": st.get("_prune_report")}
        return A0, A1


class RasterBatch:
    """A keyframe batch over EXPLICIT Gaussians (precomputed colours, scale + quaternion) on concurrent lanes
    (`segs_raster_views`, csrc/mapper_view.cu): per view the GaussianRasterizer forward
    (/root/reference/src/gaussian_renderer.cpp:86-127) and, when a dL_dout is given, its backward
    (src/gaussian_rasterizer.cpp:88-154), the parameter gradients ADDED to caller-owned accumulators — the
    rasterizer-only sibling of FusedMapper, and what bench.py times on BASELINE config 2."""

    def __init__(self, device, lanes: int = 4):
        import ctypes as C
        from . import _lib
        self._C, self._L, self.lib = C, _lib, _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("segs_slam_b200 has no CPU path: RasterBatch needs a CUDA device")
        self.lanes = max(1, int(lanes))
        self._wss = []
        for _ in range(self.lanes):
            ws = C.c_void_p()
            _lib.check(self.lib.segs_workspace_create(C.byref(ws)))
            self._wss.append(ws)
        self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)] if self.lanes > 1 else []

    def __del__(self):
        try:
            for ws in getattr(self, "_wss", []):
                self.lib.segs_workspace_destroy(ws)
            self._wss = []
        except Exception:
            pass

    def run(self, means3D, colors, opacities, scales, rotations, bg, cameras, image_height, image_width, tanfovx, tanfovy,
            images_out, dL_douts=None, grads=None, lanes: int | None = None):
        """cameras: objects/dicts with viewmatrix, projmatrix, campos (DEVICE tensors, kernel memory order).
        images_out[v]: [3,H,W] tensors written.  dL_douts[v] (optional): [3,H,W].  grads: 6 accumulators in the order
        (means3D [P,3], means2D [P,3], colors [P,3], opacity [P,1], scales [P,3], rotations [P,4]), entries may be None.
        -> list of num_rendered."""
        C, L = self._C, self._L
        n = len(cameras)
        if n == 0:
            return []
        for t in (means3D, colors, opacities, scales, rotations, bg):
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("RasterBatch: contiguous FP32 CUDA tensors expected (no CPU path)")
        get = (lambda c, k: c[k]) if isinstance(cameras[0], dict) else getattr
        arr, res = (L.RasterViewArgs * n)(), (L.MapperViewResult * n)()
        g = [None] * 6 if grads is None else [None if t is None else t.data_ptr() for t in grads]
        for v in range(n):
            a = arr[v]
            a.P = means3D.size(0)
            a.means3D, a.colors_precomp, a.opacities = means3D.data_ptr(), colors.data_ptr(), opacities.data_ptr()
            a.scales, a.rotations, a.background = scales.data_ptr(), rotations.data_ptr(), bg.data_ptr()
            a.width, a.height, a.tan_fovx, a.tan_fovy = int(image_width), int(image_height), float(tanfovx), float(tanfovy)
            a.viewmatrix, a.projmatrix = get(cameras[v], "viewmatrix").data_ptr(), get(cameras[v], "projmatrix").data_ptr()
            a.campos = get(cameras[v], "campos").data_ptr()
            a.dL_dout = None if dL_douts is None else dL_douts[v].data_ptr()
            a.image_out = images_out[v].data_ptr()
            (a.grad_means3D, a.grad_means2D, a.grad_colors, a.grad_opacity, a.grad_scales, a.grad_rotations) = g
        nl = min(self.lanes if lanes is None else max(1, min(int(lanes), self.lanes)), n)
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream().cuda_stream
            wss = (C.c_void_p * nl)(*[w.value for w in self._wss[:nl]])
            sts = (C.c_void_p * nl)(*([main] if nl == 1 else [s.cuda_stream for s in self._streams[:nl]]))
            L.check(self.lib.segs_raster_views(n, arr, res, nl, wss, sts, main))
        return [int(r.num_rendered) for r in res]
