"""GPU: a pure C++ program (examples/raster_views_capi.cpp: cudaMalloc + the C ABI, no Python, no LibTorch) renders
and back-propagates a batch of views through segs_raster_views; its images / num_rendered must be bit-identical to the
tensor-level API's and its accumulated gradients within 1e-4."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

import common
from segs_slam_b200 import rasterize_points as rp, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "examples", "raster_views_capi")


@pytest.mark.parametrize("lanes", [1, 2])
def test_cpp_consumer_of_the_c_abi(device, tmp_path, lanes):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")])
    scene = synth.config("small")
    P, W, H = scene.P, scene.W, scene.H
    views = [synth.with_camera(scene, np.eye(3, dtype=np.float32), np.array([0.02 * v, 0.01 * v, 0.0], np.float32)) for v in range(3)]
    rng = np.random.default_rng(7)
    dLs = [rng.normal(0, 1, (3, H, W)).astype(np.float32) for _ in views]
    path_in, path_out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(path_in, "wb") as f:
        f.write(struct.pack("<4i2f", P, W, H, len(views), scene.tanfovx, scene.tanfovy))
        for a in (scene.means3D, scene.colors, scene.opacities, scene.scales, scene.rotations, scene.bg):
            f.write(np.ascontiguousarray(a, dtype=np.float32).tobytes())
        for s, dL in zip(views, dLs):
            for a in (s.viewmatrix, s.projmatrix, s.campos, dL):
                f.write(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    out = subprocess.run([EXE, path_in, path_out, str(lanes)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("ok views=3")
    raw = open(path_out, "rb").read()
    N = 3 * H * W
    off = 0
    t = scene.to_torch(device)
    e = common.empty(device)
    ref_sum = None
    for s, dL in zip(views, dLs):
        R_cpp = struct.unpack_from("<i", raw, off)[0]
        off += 4
        img_cpp = np.frombuffer(raw, dtype=np.float32, count=N, offset=off).reshape(3, H, W)
        off += 4 * N
        cam = {k: torch.from_numpy(np.ascontiguousarray(getattr(s, k))).to(device) for k in ("viewmatrix", "projmatrix", "campos")}
        R, color, radii, g, b, i = rp.RasterizeGaussiansCUDA(t["bg"], t["means3D"], t["colors"], t["opacities"], t["scales"],
                                                             t["rotations"], 1.0, e, cam["viewmatrix"], cam["projmatrix"],
                                                             scene.tanfovx, scene.tanfovy, H, W, e, 0, cam["campos"], False)
        gr = rp.RasterizeGaussiansBackwardCUDA(t["bg"], t["means3D"], radii, t["colors"], t["scales"], t["rotations"], 1.0, e,
                                               cam["viewmatrix"], cam["projmatrix"], scene.tanfovx, scene.tanfovy,
                                               torch.from_numpy(dL).to(device), e, 0, cam["campos"], g, R, b, i)
        assert R_cpp == R
        assert np.array_equal(img_cpp, color.cpu().numpy())
        six = [gr[3], gr[0], gr[1], gr[2], gr[6], gr[7]]
        ref_sum = [x.clone() for x in six] if ref_sum is None else [a + b_ for a, b_ in zip(ref_sum, six)]
    grads = np.frombuffer(raw, dtype=np.float32, count=17 * P, offset=off)
    o = 0
    for w, ref in zip((3, 3, 3, 1, 3, 4), ref_sum):
        mine = torch.from_numpy(grads[o:o + w * P].copy()).to(device)
        o += w * P
        ok, why = common.grad_close(mine, ref.reshape(-1), None, rel=1e-4)
        assert ok, why
