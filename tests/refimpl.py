"""TEST INFRASTRUCTURE: drive the UNMODIFIED reference CUDA rasterizer / simple-knn
(oracle/_ref/libsegs_ref.so, built by `make -C oracle ref` from the sources under
/root/reference — see oracle/Makefile, oracle/ref_wrap.cu) from Python, and parse the
reference's private opaque-buffer layout (GeometryState/BinningState/ImageState::fromChunk,
/root/reference/cuda_rasterizer/rasterizer_impl.cu:155-194: sections carved in declaration
order, each aligned to 128 bytes) so that parity tests can read its internals.

Used only by tests/, tests/golden/make_golden.py and bench.py's reference arm.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libsegs_ref.so")
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)
_lib = None


def available() -> bool:
    return os.path.exists(REF_PATH)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(REF_PATH)
        vp = C.c_void_p
        lib.ref_raster_forward.restype = C.c_int
        lib.ref_raster_forward.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp, C.c_int, C.c_int, C.c_int,
                                           vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_float, vp, vp, vp, vp,
                                           vp, C.c_float, C.c_float, C.c_int, vp, vp]
        lib.ref_raster_backward.restype = None
        lib.ref_raster_backward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp,
                                            vp, C.c_float, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp,
                                            vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.ref_visible_filter.restype = None
        lib.ref_visible_filter.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp, C.c_int, C.c_int, C.c_int,
                                           C.c_int, vp, vp, C.c_float, vp, vp, vp, vp, C.c_float, C.c_float,
                                           C.c_int, vp]
        lib.ref_mark_visible.restype = None
        lib.ref_mark_visible.argtypes = [C.c_int, vp, vp, vp, vp]
        lib.ref_knn_mean_dist2.restype = None
        lib.ref_knn_mean_dist2.argtypes = [C.c_int, vp, vp]
        lib.ref_sync.restype = C.c_int
        _lib = lib
    return _lib


def _ptr(t):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


class _Grower:
    def __init__(self, device):
        self.t = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = ALLOC_FN(self._alloc)

    def _alloc(self, _u, n):
        self.t.resize_(int(n))
        return self.t.data_ptr()

    def done(self):   # break the ctypes-thunk reference cycle (same as the product's binding)
        self.cb = None
        return self.t


def _empty(dev):
    return torch.empty(0, dtype=torch.float32, device=dev)


def forward(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
            projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, prefiltered=False):
    """RasterizeGaussiansCUDA of the reference, minus torch glue (zeros for outputs as
    src/rasterize_points.cu:69-70).  The reference launches on the legacy default stream, which
    is also torch's default stream, so no extra synchronisation is needed (or added)."""
    lib = load()
    dev = means3D.device
    P = means3D.size(0)
    out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    g, b, i = _Grower(dev), _Grower(dev), _Grower(dev)
    M = sh.size(1) if sh.numel() else 0
    rendered = 0
    if P:
        rendered = lib.ref_raster_forward(g.cb, None, b.cb, None, i.cb, None, P, int(degree), int(M),
                                          _ptr(bg), W, H, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity),
                                          _ptr(scales), float(scale_modifier), _ptr(rotations),
                                          _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                                          float(tan_fovx), float(tan_fovy), int(prefiltered), _ptr(out_color),
                                          _ptr(radii))
    return rendered, out_color, radii, g.done(), b.done(), i.done()


def backward(bg, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
             projmatrix, tan_fovx, tan_fovy, dL_dout, sh, degree, campos, geom, R, binning, img):
    lib = load()
    P = means3D.size(0)
    H, W = dL_dout.size(1), dL_dout.size(2)
    M = sh.size(1) if sh.numel() else 0
    o = dict(dtype=torch.float32, device=means3D.device)
    d = dict(dL_dmeans3D=torch.zeros((P, 3), **o), dL_dmeans2D=torch.zeros((P, 3), **o),
             dL_dcolors=torch.zeros((P, 3), **o), dL_dconic=torch.zeros((P, 2, 2), **o),
             dL_dopacity=torch.zeros((P, 1), **o), dL_dcov3D=torch.zeros((P, 6), **o),
             dL_dsh=torch.zeros((P, M, 3), **o), dL_dscales=torch.zeros((P, 3), **o),
             dL_drotations=torch.zeros((P, 4), **o))
    if P:
        lib.ref_raster_backward(P, int(degree), int(M), int(R), _ptr(bg), W, H, _ptr(means3D), _ptr(sh),
                                _ptr(colors), _ptr(scales), float(scale_modifier), _ptr(rotations),
                                _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                                float(tan_fovx), float(tan_fovy), _ptr(radii), _ptr(geom), _ptr(binning),
                                _ptr(img), _ptr(dL_dout), _ptr(d["dL_dmeans2D"]), _ptr(d["dL_dconic"]),
                                _ptr(d["dL_dopacity"]), _ptr(d["dL_dcolors"]), _ptr(d["dL_dmeans3D"]),
                                _ptr(d["dL_dcov3D"]), _ptr(d["dL_dsh"]), _ptr(d["dL_dscales"]),
                                _ptr(d["dL_drotations"]))
    return d


def visible_filter(means3D, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                   tan_fovx, tan_fovy, H, W):
    lib = load()
    dev = means3D.device
    P = means3D.size(0)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    g, b, i = _Grower(dev), _Grower(dev), _Grower(dev)
    if P:
        lib.ref_visible_filter(g.cb, None, b.cb, None, i.cb, None, P, 0, W, H, _ptr(means3D), _ptr(scales),
                               float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
                               _ptr(projmatrix), float(tan_fovx), float(tan_fovy), 0, _ptr(radii))
        lib.ref_sync()
    g.done(); b.done(); i.done()
    return radii


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = load()
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P:
        lib.ref_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), present.data_ptr())
        lib.ref_sync()
    return present


def knn(points):
    lib = load()
    P = points.size(0)
    out = torch.zeros((P,), dtype=torch.float32, device=points.device)
    if P:
        lib.ref_knn_mean_dist2(P, _ptr(points), _ptr(out))
        lib.ref_sync()
    return out


# ---- the reference's buffer layout ---------------------------------------------------------
def _carve(buf, specs):
    """specs: list of (name, dtype, count); returns {name: tensor view}; 128-byte alignment
    relative to the (>= 512-byte aligned) tensor base, like obtain() (rasterizer_impl.h:22-29)."""
    out, off = {}, 0
    base = buf.data_ptr()
    for name, dtype, count in specs:
        addr = (base + off + 127) & ~127
        off = addr - base
        nbytes = count * torch.empty(0, dtype=dtype).element_size()
        out[name] = buf[off:off + nbytes].view(dtype)
        off += nbytes
    return out


def parse_geom(geom, P):
    s = _carve(geom, [("depths", torch.float32, P), ("clamped", torch.uint8, 3 * P),
                      ("internal_radii", torch.int32, P), ("means2D", torch.float32, 2 * P),
                      ("cov3D", torch.float32, 6 * P), ("conic_opacity", torch.float32, 4 * P),
                      ("rgb", torch.float32, 3 * P), ("tiles_touched", torch.int32, P)])
    s["means2D"] = s["means2D"].view(P, 2)
    s["cov3D"] = s["cov3D"].view(P, 6)
    s["conic_opacity"] = s["conic_opacity"].view(P, 4)
    s["rgb"] = s["rgb"].view(P, 3)
    return s


def parse_binning(binning, R):
    return _carve(binning, [("point_list", torch.int32, R), ("point_list_unsorted", torch.int32, R),
                            ("point_list_keys", torch.int64, R), ("point_list_keys_unsorted", torch.int64, R)])


def parse_image(img, N, T):
    s = _carve(img, [("final_T", torch.float32, N), ("n_contrib", torch.int32, N),
                     ("ranges", torch.int32, 2 * N)])   # ranges is sized N, not T (rasterizer_impl.cu:177)
    s["ranges"] = s["ranges"].view(N, 2)[:T]
    return s
