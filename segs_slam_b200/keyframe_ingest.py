"""Keyframe image ingest: what GaussianMapper does to every new keyframe image before training sees it
(/root/reference/src/gaussian_mapper.cpp:557-645, 1230-1309) on this library's kernels (csrc/ingest.cu):

    camera.undistortImage(img)                  include/camera.h:106-115  (cv::remap through the maps of :70-86)
    tensor_utils::cvMat2TorchTensor_Float32     include/tensor_utils.h:40-69  ([H,W,3] cv::Mat -> [3,H,W] CUDA tensor)
    cv::cuda::resize to the pyramid levels      src/gaussian_mapper.cpp:621-632
    the undistortion mask                       include/camera.h:88-103  (a white image through the same remap)

The per-camera maps are a one-off host computation (`init_undistort_rectify_map`, the pinhole + radial/tangential model of
cv::initUndistortRectifyMap with R = I); everything per keyframe runs on the device: ONE upload of the interleaved image
from pinned memory, ONE kernel that undistorts and writes the planar tensor, one kernel per pyramid level."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .rasterize_points import _stream


def init_undistort_rectify_map(K, dist_coeffs, new_K, width: int, height: int):
    """cv::initUndistortRectifyMap(K, dist, I, new_K, (width, height), CV_32FC1) -> (map_x, map_y) float32 [height, width].
    dist_coeffs = (k1, k2, p1, p2[, k3]) — the reference keeps four (include/camera.h:135).  Double arithmetic, rounded to
    FP32 at the end, like OpenCV."""
    K, new_K = np.asarray(K, dtype=np.float64).reshape(3, 3), np.asarray(new_K, dtype=np.float64).reshape(3, 3)
    d = np.zeros(5, dtype=np.float64)
    dc = np.asarray(dist_coeffs, dtype=np.float64).reshape(-1)
    d[:min(5, dc.size)] = dc[:5]
    k1, k2, p1, p2, k3 = d
    ir = np.linalg.inv(new_K)                              # (new_K * R)^-1 with R = I
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    _x = u * ir[0, 0] + v * ir[0, 1] + ir[0, 2]
    _y = u * ir[1, 0] + v * ir[1, 1] + ir[1, 2]
    _w = u * ir[2, 0] + v * ir[2, 1] + ir[2, 2]
    x, y = _x / _w, _y / _w
    x2, y2 = x * x, y * y
    r2, _2xy = x2 + y2, 2 * x * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
    map_x = K[0, 0] * xd + K[0, 2]
    map_y = K[1, 1] * yd + K[1, 2]
    return map_x.astype(np.float32), map_y.astype(np.float32)


class KeyframeIngest:
    """Per camera: the undistortion maps on the device, a pinned staging buffer, the pyramid sizes."""

    def __init__(self, device, height: int, width: int, K=None, dist_coeffs=None, new_K=None, src_height: int | None = None,
                 src_width: int | None = None, pyramid_sizes=()):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("segs_slam_b200 has no CPU path: KeyframeIngest needs a CUDA device")
        self.lib = _lib.load()
        self.H, self.W = int(height), int(width)
        self.src_H, self.src_W = int(src_height or height), int(src_width or width)
        self.map_x = self.map_y = None
        if K is not None and dist_coeffs is not None and np.any(np.asarray(dist_coeffs) != 0):
            mx, my = init_undistort_rectify_map(K, dist_coeffs, K if new_K is None else new_K, self.W, self.H)
            self.map_x, self.map_y = torch.from_numpy(mx).to(self.device), torch.from_numpy(my).to(self.device)
        elif (self.src_H, self.src_W) != (self.H, self.W):
            raise ValueError("without distortion coefficients the source must have the output size")
        self.pyramid_sizes = [(int(h), int(w)) for h, w in pyramid_sizes]      # (gaus_pyramid_height_[l], gaus_pyramid_width_[l])
        self._stage = None

    def _upload(self, image_hwc) -> torch.Tensor:
        if isinstance(image_hwc, torch.Tensor) and image_hwc.is_cuda:
            t = image_hwc
        else:
            a = image_hwc.numpy() if isinstance(image_hwc, torch.Tensor) else np.asarray(image_hwc)
            if a.dtype != np.float32 or a.ndim != 3:
                raise RuntimeError("ingest: an [H,W,C] float32 image is expected (cvMat2TorchTensor_Float32)")
            if self._stage is None or tuple(self._stage.shape) != a.shape:
                self._stage = torch.empty(a.shape, dtype=torch.float32).pin_memory()
            self._stage.copy_(torch.from_numpy(np.ascontiguousarray(a)))
            t = self._stage.to(self.device, non_blocking=True)
        if t.dtype != torch.float32 or tuple(t.shape[:2]) != (self.src_H, self.src_W):
            raise RuntimeError(f"ingest: expected a [{self.src_H},{self.src_W},C] float32 image, got {tuple(t.shape)} {t.dtype}")
        return t.contiguous()

    def ingest(self, image_hwc) -> torch.Tensor:
        """[src_H, src_W, C] float32 (host or device) -> undistorted [C, H, W] CUDA tensor (`original_image_`)."""
        src = self._upload(image_hwc)
        C_ = src.size(2)
        out = torch.empty((C_, self.H, self.W), dtype=torch.float32, device=self.device)
        mx = None if self.map_x is None else self.map_x.data_ptr()
        my = None if self.map_y is None else self.map_y.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.segs_ingest_image(self.H, self.W, C_, self.src_H, self.src_W, src.data_ptr(), mx, my, out.data_ptr(),
                                                  _stream()))
        return out

    def undistort_mask(self, channels: int = 3) -> torch.Tensor:
        """include/camera.h:88-103: a white image through the same remap ([C,H,W]; 1 inside, < 1 at the border, 0 outside)."""
        white = torch.ones((self.src_H, self.src_W, channels), dtype=torch.float32, device=self.device)
        return self.ingest(white)

    def pyramid(self, image_chw: torch.Tensor) -> list[torch.Tensor]:
        """The Gaussian-pyramid levels of an ingested image (`gaus_pyramid_original_image_`)."""
        return [resize_bilinear(image_chw, h, w) for h, w in self.pyramid_sizes]


def resize_bilinear(image_chw: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """cv::resize(..., INTER_LINEAR) of a planar [C,H,W] float32 CUDA image."""
    if not image_chw.is_cuda or image_chw.dtype != torch.float32 or image_chw.dim() != 3:
        raise RuntimeError("resize: a [C,H,W] float32 CUDA tensor is expected (no CPU path)")
    lib = _lib.load()
    src = image_chw.contiguous()
    C_, H, W = src.shape
    out = torch.empty((C_, int(height), int(width)), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(lib.segs_resize_bilinear(C_, H, W, src.data_ptr(), int(height), int(width), out.data_ptr(), _stream()))
    return out
