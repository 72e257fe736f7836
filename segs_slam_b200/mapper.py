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

Only host-side orchestration lives here; `render_loss` is the per-view work (prefilter -> decode ->
rasterize -> loss, all on the GPU kernels of this package — `make_render_loss` below builds it).
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

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
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
                     bg: torch.Tensor):
    """Per-view work of the mapper on this package's kernels: anchor prefilter
    (RasterizeGaussiansfilterCUDA, gaussian_renderer.cpp:131-199) -> fused decode
    (generate_neural_gaussians, :214-334) -> rasterize (GaussianRasterizer, :40-127) -> L1 loss against
    the keyframe image (loss_utils::l1_loss; SSIM and the frequency terms are §8f 'next').

    cameras[v] carries world_view_transform_, full_proj_transform_, camera_center_, t_, R_quaternion_."""
    from . import GaussianRasterizationSettings, GaussianRasterizer, generate_neural_gaussians
    from .rasterize_points import RasterizeGaussiansfilterCUDA

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
        return (image - targets[v]).abs().mean()

    return render_loss
