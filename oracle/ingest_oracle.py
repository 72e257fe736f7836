"""TEST INFRASTRUCTURE — numpy restatement of the OpenCV routines the reference's keyframe ingest calls
(include/camera.h:70-115, src/gaussian_mapper.cpp:557-645): cv::remap (CV_32FC1 maps, INTER_LINEAR, BORDER_CONSTANT 0)
and cv::resize (INTER_LINEAR, CV_32F), following imgwarp.cpp / resize.cpp of OpenCV 4.x (the reference's Dockerfile builds
OpenCV 4.7 from source; OpenCV is a dependency outside /root/reference).  PINNED to the real library:
tests/test_ingest_cpu.py compares it with cv2 (4.13 in this image) where importable and with tests/golden/ingest_*.npz
(written by tests/golden/make_ingest_golden.py with cv2).  Only tests/ imports this file."""
import numpy as np

INTER_TAB_SIZE = 32


def remap_bilinear(src_hwc: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    """-> [H, W, C] float32.  Maps quantised to 1/32 pixel with round-half-even, FP32 bilinear table, zero border."""
    f32 = np.float32
    src = np.asarray(src_hwc, dtype=f32)
    sH, sW, _C = src.shape
    sx = np.rint(map_x.astype(f32) * f32(INTER_TAB_SIZE)).astype(np.int64)
    sy = np.rint(map_y.astype(f32) * f32(INTER_TAB_SIZE)).astype(np.int64)
    ix, iy = sx >> 5, sy >> 5
    fx = (sx & 31).astype(f32) * f32(1.0 / INTER_TAB_SIZE)
    fy = (sy & 31).astype(f32) * f32(1.0 / INTER_TAB_SIZE)
    w = [(f32(1) - fy) * (f32(1) - fx), (f32(1) - fy) * fx, fy * (f32(1) - fx), fy * fx]

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < sH) & (xx >= 0) & (xx < sW)
        v = src[np.clip(yy, 0, sH - 1), np.clip(xx, 0, sW - 1)]
        return np.where(ok[..., None], v, f32(0))

    out = tap(iy, ix) * w[0][..., None]
    out = out + tap(iy, ix + 1) * w[1][..., None]
    out = out + tap(iy + 1, ix) * w[2][..., None]
    out = out + tap(iy + 1, ix + 1) * w[3][..., None]
    return out.astype(f32)


def _resize_axis(d, scale, n):
    f = ((np.arange(d, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    lo = s < 0
    s[lo], f[lo] = 0, 0
    hi = s >= n - 1
    s[hi], f[hi] = n - 1, 0
    s1 = np.where(hi, s, s + 1)
    return s, s1, (np.float32(1) - f).astype(np.float32), f.astype(np.float32)


def resize_bilinear(src_chw: np.ndarray, h: int, w: int) -> np.ndarray:
    """cv::resize INTER_LINEAR on every plane of [C,H,W] float32: horizontal pass, then vertical."""
    src = np.asarray(src_chw, dtype=np.float32)
    _C, H, W = src.shape
    # resize.cpp: inv_scale = (double)dsize / ssize; scale = 1. / inv_scale
    x0, x1, a0, a1 = _resize_axis(w, 1.0 / (w / W), W)
    y0, y1, b0, b1 = _resize_axis(h, 1.0 / (h / H), H)
    rows = src[:, :, x0] * a0 + src[:, :, x1] * a1                      # [C,H,w]
    return (rows[:, y0, :] * b0[None, :, None] + rows[:, y1, :] * b1[None, :, None]).astype(np.float32)
