"""Fused Adam over a flat gradient bucket (C ABI: segs_adam_step, segs_slam_b200/csrc/optim.cu).

Stands in for the torch::optim::Adam the reference builds in GaussianModel::trainingSetup
(/root/reference/src/gaussian_model.cpp:620-872: one parameter group per tensor, each with its own learning
rate, eps = 1e-15) and steps once per iteration (src/gaussian_mapper.cpp:1003-1006): every tensor is updated
by ONE kernel launch that also applies the 1/B scale of the batch-mean gradient and clears the bucket.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib
from .rasterize_points import _stream


class FusedAdam:
    """params: tensors in bucket order; grads live in `bucket.flat` (mapper.GradBucket).
    lrs: one learning rate per tensor (the reference's per-group lr); update with `set_lr`."""

    def __init__(self, bucket, lrs: Sequence[float] | float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-15,
                 weight_decay: float = 0.0):
        self.bucket = bucket
        n = len(bucket.params)
        self.lrs = [float(lrs)] * n if isinstance(lrs, (int, float)) else [float(x) for x in lrs]
        if len(self.lrs) != n:
            raise ValueError(f"{len(self.lrs)} learning rates for {n} tensors")
        for p in bucket.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam: parameters must be contiguous FP32 CUDA tensors (no CPU path)")
        self.betas, self.eps, self.weight_decay = (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.exp_avg = torch.zeros_like(bucket.flat)
        self.exp_avg_sq = torch.zeros_like(bucket.flat)
        self.step_count = 0

    def set_lr(self, index: int, lr: float):
        self.lrs[index] = float(lr)

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, zero_grad: bool = True):
        lib = _lib.load()
        self.step_count += 1
        b = self.bucket
        arr = (_lib.AdamTensor * len(b.params))()
        off = 0
        for k, (p, n) in enumerate(zip(b.params, b.sizes)):
            arr[k] = _lib.AdamTensor(p.data_ptr(), off, n, self.lrs[k], self.betas[0], self.betas[1], self.eps,
                                     self.weight_decay, self.step_count)
            off += n
        with torch.cuda.device(b.flat.device):
            _lib.check(lib.segs_adam_step(len(b.params), arr, b.flat.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), float(grad_scale), int(zero_grad), _stream()))


def get_expon_lr_func(step: int, lr_init: float, lr_final: float, lr_delay_mult: float = 1.0, max_steps: int = 1000000,
                      lr_delay_steps: int = 0) -> float:
    """GaussianModel::getExponLrFunc / exponLrFunc (/root/reference/src/gaussian_model.cpp:1368-1407): log-linear
    interpolation from lr_init to lr_final over max_steps with an optional sine warm-up; FP32 arithmetic like the
    reference.  Host scalar code (the per-group learning rates the mapper sets every iteration, :874-960)."""
    import math

    import numpy as np
    f = np.float32
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    if lr_delay_steps > 0:
        delay_rate = f(lr_delay_mult) + (f(1.0) - f(lr_delay_mult)) * f(math.sin(f(math.pi / 2) * f(np.clip(f(step) / f(lr_delay_steps), 0.0, 1.0))))
    else:
        delay_rate = f(1.0)
    t = f(np.clip(f(step) / f(max_steps), 0.0, 1.0))
    log_lerp = f(math.exp(f(math.log(f(lr_init))) * (f(1.0) - t) + f(math.log(f(lr_final))) * t))
    return float(f(delay_rate) * log_lerp)
