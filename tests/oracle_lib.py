"""TEST INFRASTRUCTURE: ctypes wrapper of the CPU oracle (oracle/liboracle.so, built by
`make -C oracle oracle` from oracle/raster_oracle.cpp).  numpy in, numpy out."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "raster_oracle.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"],
                                  stdout=subprocess.DEVNULL)
        lib = C.CDLL(LIB)
        vp = C.c_void_p
        lib.oracle_forward.restype = vp
        lib.oracle_forward.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                       C.c_float, vp, vp, vp, vp, vp, C.c_float, C.c_float, C.c_int]
        lib.oracle_num_rendered.restype = C.c_int
        lib.oracle_num_rendered.argtypes = [vp]
        lib.oracle_get.restype = C.c_size_t
        lib.oracle_get.argtypes = [vp, C.c_char_p, vp]
        lib.oracle_free.restype = None
        lib.oracle_free.argtypes = [vp]
        lib.oracle_backward.restype = None
        lib.oracle_backward.argtypes = [vp] * 11
        lib.oracle_visible_filter.restype = None
        lib.oracle_visible_filter.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, C.c_float, vp, vp, vp, vp,
                                              C.c_float, C.c_float, vp, vp]
        lib.oracle_knn.restype = None
        lib.oracle_knn.argtypes = [C.c_int, vp, vp]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


_FIELDS = {"radii": np.int32, "tiles_touched": np.uint32, "point_offsets": np.uint32, "depths": np.float32,
           "means2D": np.float32, "conic_opacity": np.float32, "cov3D": np.float32, "rgb": np.float32,
           "clamped": np.uint8, "keys": np.uint64, "point_list": np.uint32, "ranges": np.uint32,
           "final_T": np.float32, "n_contrib": np.uint32, "out_color": np.float32}


class Forward:
    """One forward pass of the oracle; `.get(name)` reads its state, `.backward(dL)` the grads."""

    def __init__(self, *, bg, means3D, colors=None, opacities, scales=None, rotations=None, scale_modifier=1.0,
                 cov3D_precomp=None, viewmatrix, projmatrix, campos, tanfovx, tanfovy, W, H, sh=None, degree=0,
                 nthreads=1):
        lib = load()
        self.P = int(means3D.shape[0])
        self.W, self.H = int(W), int(H)
        self.M = 0 if sh is None else int(sh.shape[1])
        keep = [_f(x) for x in (bg, means3D, sh, colors, opacities, scales, rotations, cov3D_precomp,
                                viewmatrix, projmatrix, campos)]
        bg, m3, shc, col, opa, sca, rot, cov, view, proj, cam = keep
        self.h = lib.oracle_forward(self.P, int(degree), self.M, _p(bg), self.W, self.H, _p(m3), _p(shc), _p(col),
                                    _p(opa), _p(sca), float(scale_modifier), _p(rot), _p(cov), _p(view), _p(proj),
                                    _p(cam), float(tanfovx), float(tanfovy), int(nthreads))
        self.R = lib.oracle_num_rendered(self.h)

    def get(self, name):
        lib = load()
        n = lib.oracle_get(self.h, name.encode(), None)
        out = np.empty(n, dtype=_FIELDS[name])
        lib.oracle_get(self.h, name.encode(), _p(out))
        P, N = self.P, self.W * self.H
        shapes = {"means2D": (P, 2), "conic_opacity": (P, 4), "cov3D": (P, 6), "rgb": (P, 3), "clamped": (P, 3),
                  "ranges": (-1, 2), "out_color": (3, self.H, self.W)}
        return out.reshape(shapes[name]) if name in shapes else out

    def backward(self, dL_dout):
        lib = load()
        P, M = self.P, self.M
        dL = _f(dL_dout)
        g = dict(dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
                 dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
                 dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
                 dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
                 dL_drotations=np.zeros((P, 4), np.float32))
        lib.oracle_backward(self.h, _p(dL), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]),
                            _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
                            _p(g["dL_dsh"]) if M else None, _p(g["dL_dscales"]), _p(g["dL_drotations"]))
        return g

    def close(self):
        if self.h:
            load().oracle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def from_scene(scene, use_sh=False, nthreads=1, **kw):
    args = dict(bg=scene.bg, means3D=scene.means3D, opacities=scene.opacities, scales=scene.scales,
                rotations=scene.rotations, scale_modifier=scene.scale_modifier, viewmatrix=scene.viewmatrix,
                projmatrix=scene.projmatrix, campos=scene.campos, tanfovx=scene.tanfovx, tanfovy=scene.tanfovy,
                W=scene.W, H=scene.H, nthreads=nthreads)
    if use_sh:
        args.update(sh=scene.extras["sh"], degree=int(scene.extras["sh_degree"][0]))
    else:
        args.update(colors=scene.colors)
    args.update(kw)
    return Forward(**args)


def visible_filter(means3D, scales, rotations, viewmatrix, projmatrix, tanfovx, tanfovy, W, H, scale_modifier=1.0,
                   cov3D_precomp=None):
    lib = load()
    P = means3D.shape[0]
    radii = np.zeros(P, np.int32)
    present = np.zeros(P, np.uint8)
    m3, sca, rot, cov, view, proj = [_f(x) for x in (means3D, scales, rotations, cov3D_precomp, viewmatrix, projmatrix)]
    lib.oracle_visible_filter(P, int(W), int(H), _p(m3), _p(sca), float(scale_modifier), _p(rot), _p(cov), _p(view),
                              _p(proj), float(tanfovx), float(tanfovy), _p(radii), _p(present))
    return radii, present.astype(bool)


def knn(points):
    lib = load()
    pts = _f(points)
    out = np.zeros(pts.shape[0], np.float32)
    lib.oracle_knn(pts.shape[0], _p(pts), _p(out))
    return out
