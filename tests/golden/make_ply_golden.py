"""tests/golden/anchors_tinyply.ply: a 7-anchor checkpoint written by the reference's OWN PLY library (tinyply, compiled from
/root/reference/third_party/tinyply by `make -C oracle plyref`) with savePly's call sequence, plus the arrays that went in
(anchors_tinyply.npz).  Pins segs_slam_b200/checkpoint.py byte for byte.

    python tests/golden/make_ply_golden.py"""
import ctypes as C
import os

import numpy as np
import torch  # noqa: F401  (same libstdc++ resolution as the loss veneer)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libply_ref.so"))
fp = C.POINTER(C.c_float)
lib.ref_save_ply.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp]
lib.ref_load_ply.argtypes = [C.c_char_p, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp]
p = lambda a: a.ctypes.data_as(fp)

rng = np.random.default_rng(42)
A, F, K = 7, 32, 10
d = dict(anchor=rng.normal(0, 2, (A, 3)), feat=rng.normal(0, 0.1, (A, F)), offset=rng.uniform(-1, 1, (A, K, 3)),
         opacity=rng.normal(0, 1, (A, 1)), scale=np.log(rng.uniform(0.005, 0.03, (A, 6))), rot=rng.normal(0, 1, (A, 4)))
d = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in d.items()}
flat = np.ascontiguousarray(d["offset"].transpose(0, 2, 1).reshape(A, 3 * K))       # _offset.transpose(1, 2).flatten(1)
path = os.path.join(HERE, "anchors_tinyply.ply")
assert lib.ref_save_ply(path.encode(), A, F, 3 * K, p(d["anchor"]), p(d["feat"]), p(flat), p(d["opacity"]), p(d["scale"]), p(d["rot"])) == 0
np.savez(os.path.join(HERE, "anchors_tinyply.npz"), **d)
# and tinyply reads its own file back
back = {k: np.zeros_like(v) for k, v in d.items()}
bflat = np.zeros_like(flat)
n = lib.ref_load_ply(path.encode(), F, 3 * K, p(back["anchor"]), p(back["feat"]), p(bflat), p(back["opacity"]), p(back["scale"]), p(back["rot"]))
assert n == A and np.array_equal(bflat, flat) and np.array_equal(back["rot"], d["rot"])
print("wrote", path, os.path.getsize(path), "bytes")

# sparse points (saveSparsePointsPly): float xyz + normals, uchar rgb
lib.ref_save_sparse_ply.argtypes = [C.c_char_p, C.c_int, fp, C.POINTER(C.c_ubyte)]
xyz = np.ascontiguousarray(rng.normal(0, 1, (5, 3)), dtype=np.float32)
col = np.ascontiguousarray(rng.uniform(0, 1, (5, 3)), dtype=np.float32)
rgb = np.ascontiguousarray((torch.from_numpy(col) * 255.0).to(torch.uint8).numpy())       # toType(kUInt8), :1323
spath = os.path.join(HERE, "sparse_tinyply.ply")
assert lib.ref_save_sparse_ply(spath.encode(), 5, p(xyz), rgb.ctypes.data_as(C.POINTER(C.c_ubyte))) == 0
np.savez(os.path.join(HERE, "sparse_tinyply.npz"), xyz=xyz, color=col)
print("wrote", spath, os.path.getsize(spath), "bytes")
