"""Tensor-level entry points, mirroring the reference's L6 interface one to one
(/root/reference/include/rasterize_points.h:18-102, src/rasterize_points.cu;
/root/reference/third_party/simple-knn/spatial.h:14) — same names, argument order, tuple
orders and error behaviour — on top of the C-ABI library.  torch is only used here for
device memory and the current stream.

The C++/LibTorch twin of this file (what a SEGS-SLAM maintainer links instead of the
reference's rasterize_points.cu) is segs_slam_b200/csrc/torch_shim/rasterize_points.cpp.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

NUM_CHANNELS = 3


def _ptr(t: torch.Tensor | None) -> int | None:
    """data_ptr() of a contiguous tensor; a 0-element tensor is the reference's "absent" signal
    (src/gaussian_rasterizer.cpp:183-193) and maps to NULL."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: torch.Tensor, like: torch.Tensor | None = None, dtype=torch.float32) -> torch.Tensor:
    """Contiguous tensor of the dtype the kernels read.  The reference's `data_ptr<float>()` throws on any other dtype
    (src/rasterize_points.cu:88-107); a silent reinterpretation would render garbage or read out of bounds."""
    if t.numel() != 0:
        if t.dtype != dtype:
            raise RuntimeError(f"expected scalar type {str(dtype).replace('torch.', '')} but found {str(t.dtype).replace('torch.', '')}")
        if not t.is_cuda or (like is not None and t.device != like.device):
            raise RuntimeError(f"expected a CUDA tensor on {like.device if like is not None else 'the device of means3D'}, got {t.device}")
    return t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Grower:
    """resizeFunctional (src/rasterize_points.cu:28-34): grows a byte tensor on request."""

    def __init__(self, device):
        self.t = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, _user, nbytes):
        self.t.resize_(int(nbytes))
        return self.t.data_ptr()

    def done(self):
        """Drop the ctypes thunk: it references the bound method, i.e. a reference cycle that would keep
        the buffer alive until the cyclic GC runs (hundreds of MB per call at C2)."""
        self.cb = None
        return self.t


def _check_means(means3D: torch.Tensor) -> None:
    if means3D.dim() != 2 or means3D.size(1) != 3:
        # AT_ERROR of src/rasterize_points.cu:57-59
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("segs_slam_b200 has no CPU path: tensors must live on a CUDA device")


def RasterizeGaussiansCUDA(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                           cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                           image_width, sh, degree, campos, prefiltered):
    """-> (rendered, out_color[3,H,W], radii[P], geomBuffer, binningBuffer, imgBuffer)
    (src/rasterize_points.cu:36-114)."""
    _check_means(means3D)
    lib = _lib.load()
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    dev = means3D.device
    # out_color/radii are fully written by the kernels for P > 0; P == 0 leaves zeros like the reference
    out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev) if P else \
        torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    geom, binning, img = _Grower(dev), _Grower(dev), _Grower(dev)
    rendered = C.c_int(0)
    if P != 0:
        M = sh.size(1) if sh.numel() != 0 else 0
        keep = [_f32c(x, means3D) for x in (background, means3D, sh, colors, opacity, scales, rotations,
                                   cov3D_precomp, viewmatrix, projmatrix, campos)]
        bg, m3, shc, col, opa, sca, rot, cov, view, proj, cam = keep
        with torch.cuda.device(dev):
            _lib.check(lib.segs_raster_forward(
                geom.cb, None, binning.cb, None, img.cb, None,
                P, int(degree), int(M),
                _ptr(bg), W, H,
                _ptr(m3), _ptr(shc), _ptr(col), _ptr(opa), _ptr(sca), float(scale_modifier), _ptr(rot),
                _ptr(cov), _ptr(view), _ptr(proj), _ptr(cam),
                float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
                _ptr(out_color), _ptr(radii), C.byref(rendered), _stream()))
    return rendered.value, out_color, radii, geom.done(), binning.done(), img.done()


def RasterizeGaussiansBackwardCUDA(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                   cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                   dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer,
                                   imageBuffer):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
    dL_drotations) (src/rasterize_points.cu:116-193)."""
    lib = _lib.load()
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if sh.numel() != 0 else 0
    o = dict(dtype=torch.float32, device=means3D.device)
    # the kernels write every element (zeros where nothing was rendered): no memsets needed
    dL_dmeans3D = torch.empty((P, 3), **o)
    dL_dmeans2D = torch.empty((P, 3), **o)
    dL_dcolors = torch.empty((P, NUM_CHANNELS), **o)
    dL_dopacity = torch.empty((P, 1), **o)
    dL_dcov3D = torch.empty((P, 6), **o)
    dL_dsh = torch.empty((P, M, 3), **o)
    dL_dscales = torch.empty((P, 3), **o)
    dL_drotations = torch.empty((P, 4), **o)
    if P != 0:
        keep = [_f32c(x, means3D) for x in (background, means3D, sh, colors, scales, rotations, cov3D_precomp,
                                   viewmatrix, projmatrix, campos, dL_dout_color)]
        bg, m3, shc, col, sca, rot, cov, view, proj, cam, dpix = keep
        rad = _f32c(radii, means3D, torch.int32)
        with torch.cuda.device(means3D.device):
            _lib.check(lib.segs_raster_backward(
                P, int(degree), int(M), int(R),
                _ptr(bg), W, H,
                _ptr(m3), _ptr(shc), _ptr(col), _ptr(sca), float(scale_modifier), _ptr(rot), _ptr(cov),
                _ptr(view), _ptr(proj), _ptr(cam),
                float(tan_fovx), float(tan_fovy), _ptr(rad),
                _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                _ptr(dpix), _ptr(dL_dmeans2D), None, _ptr(dL_dopacity), _ptr(dL_dcolors),
                _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations),
                _stream()))
    return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
            dL_drotations)


def markVisible(means3D, viewmatrix, projmatrix):
    """-> bool[P] (src/rasterize_points.cu:195-214)."""
    lib = _lib.load()
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        m3, view, proj = _f32c(means3D), _f32c(viewmatrix, means3D), _f32c(projmatrix, means3D)
        with torch.cuda.device(means3D.device):
            _lib.check(lib.segs_mark_visible(P, _ptr(m3), _ptr(view), _ptr(proj), present.data_ptr(), _stream()))
    return present


def RasterizeGaussiansfilterCUDA(means3D, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                                 projmatrix, tan_fovx, tan_fovy, image_height, image_width, prefiltered,
                                 debug=False):
    """-> radii[P] int32, > 0 where the anchor is visible (src/rasterize_points.cu:216-276)."""
    _check_means(means3D)
    lib = _lib.load()
    P = means3D.size(0)
    radii = torch.zeros((P,), dtype=torch.int32, device=means3D.device)
    if P != 0:
        keep = [_f32c(x, means3D) for x in (means3D, scales, rotations, cov3D_precomp, viewmatrix, projmatrix)]
        m3, sca, rot, cov, view, proj = keep
        with torch.cuda.device(means3D.device):
            _lib.check(lib.segs_visible_filter(
                P, 0, int(image_width), int(image_height), _ptr(m3), _ptr(sca), float(scale_modifier),
                _ptr(rot), _ptr(cov), _ptr(view), _ptr(proj), float(tan_fovx), float(tan_fovy),
                int(bool(prefiltered)), _ptr(radii), _stream()))
    return radii


def RasterizeGaussiansprojectCUDA(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                                  cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                                  image_width, sh, degree, campos, prefiltered):
    """-> (points_image[P,2], radii[P], out_color[P,3]) (src/rasterize_points.cu:278-362)."""
    _check_means(means3D)
    lib = _lib.load()
    P = means3D.size(0)
    dev = means3D.device
    out_color = torch.zeros((P, NUM_CHANNELS), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    points_image = torch.zeros((P, 2), dtype=torch.float32, device=dev)
    if P != 0:
        M = sh.size(1) if sh.numel() != 0 else 0
        keep = [_f32c(x, means3D) for x in (means3D, sh, colors, opacity, scales, rotations, cov3D_precomp,
                                   viewmatrix, projmatrix, campos)]
        m3, shc, col, opa, sca, rot, cov, view, proj, cam = keep
        with torch.cuda.device(dev):
            _lib.check(lib.segs_project(
                P, int(degree), int(M), int(image_width), int(image_height),
                _ptr(m3), _ptr(shc), _ptr(col), _ptr(opa), _ptr(sca), float(scale_modifier), _ptr(rot),
                _ptr(cov), _ptr(view), _ptr(proj), _ptr(cam), float(tan_fovx), float(tan_fovy),
                int(bool(prefiltered)), _ptr(out_color), _ptr(points_image), _ptr(radii), _stream()))
    return points_image, radii, out_color


def distCUDA2(points):
    """-> float[P]: mean squared distance to the 3 nearest neighbours
    (third_party/simple-knn/spatial.cu:16-25)."""
    lib = _lib.load()
    if not points.is_cuda:
        raise RuntimeError("segs_slam_b200 has no CPU path: tensors must live on a CUDA device")
    P = points.size(0)
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)
    if P != 0:
        pts = _f32c(points)
        scratch = _Grower(points.device)
        with torch.cuda.device(points.device):
            _lib.check(lib.segs_knn_mean_dist2(P, _ptr(pts), _ptr(means), scratch.cb, None, _stream()))
            scratch.done()
            # scratch is only referenced by work already queued on the current stream; torch's
            # caching allocator keeps the block stream-ordered, so dropping it here is safe.
    return means


def buffer_section(name, geomBuffer, binningBuffer, imageBuffer, P, R, W, H, dtype, shape=None):
    """View a named section of the opaque buffers as a tensor (parity tests / debugging)."""
    lib = _lib.load()
    ptr, nbytes = C.c_void_p(), C.c_size_t()
    _lib.check(lib.segs_buffer_section(name.encode(), _ptr(geomBuffer), _ptr(binningBuffer),
                                       _ptr(imageBuffer), int(P), int(R), int(W), int(H),
                                       C.byref(ptr), C.byref(nbytes)))
    owners = {"geom": geomBuffer, "binning": binningBuffer, "image": imageBuffer}
    for buf in owners.values():
        if buf is None or buf.numel() == 0:
            continue
        off = (ptr.value or 0) - buf.data_ptr()
        if 0 <= off and off + nbytes.value <= buf.numel():
            t = buf[off:off + nbytes.value].view(dtype)
            return t.view(shape) if shape is not None else t
    if nbytes.value == 0:
        return torch.empty(0, dtype=dtype, device=geomBuffer.device)
    raise RuntimeError(f"section {name} is not inside any buffer")
