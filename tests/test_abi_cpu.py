"""CPU suite: the C-ABI library loads and exports every symbol include/segs_raster.h declares
(no compute calls without a GPU), host-side argument validation mirrors the reference."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "segs_raster.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(segs_[a-z0-9_]+)\s*\(", txt)) - {"segs_alloc_fn"})


def test_library_exports_every_declared_symbol():
    from segs_slam_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/segs_raster.h but not exported"
    # and every binding the Python host side uses is declared in the header
    assert set(_lib.exported_symbols()) <= set(names)
    assert lib.segs_version() == 1


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    import torch
    from segs_slam_b200 import rasterize_points as rp
    e = torch.empty(0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rp.RasterizeGaussiansCUDA(torch.zeros(3), torch.zeros((4, 3)), e, e, e, e, 1.0, e, torch.eye(4), torch.eye(4),
                                  1.0, 1.0, 16, 16, e, 0, torch.zeros(3), False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rp.distCUDA2(torch.zeros((4, 3)))
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        rp.RasterizeGaussiansCUDA(torch.zeros(3), torch.zeros((4, 4)), e, e, e, e, 1.0, e, torch.eye(4), torch.eye(4),
                                  1.0, 1.0, 16, 16, e, 0, torch.zeros(3), False)


def test_rasterizer_argument_validation():
    """XOR checks of GaussianRasterizer::forward (src/gaussian_rasterizer.cpp:176-182)."""
    import torch
    from segs_slam_b200 import GaussianRasterizationSettings, GaussianRasterizer
    s = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                      torch.zeros(3), False)
    r = GaussianRasterizer(s)
    e = torch.empty(0)
    x = torch.zeros((4, 3))
    with pytest.raises(RuntimeError, match="excatly one of either SHs or precomputed colors"):
        r(x, x, x[:, :1], True, True, True, True, False, e, x, x, torch.zeros((4, 4)), e)
    with pytest.raises(RuntimeError, match="exactly one of either scale/rotation pair"):
        r(x, x, x[:, :1], False, True, False, True, False, e, x, x, torch.zeros((4, 4)), e)


def test_synth_scene_is_deterministic():
    from segs_slam_b200 import synth
    import numpy as np
    a, b = synth.config("tiny"), synth.config("tiny")
    for k in ("means3D", "scales", "rotations", "opacities", "colors", "dL_dout", "projmatrix"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k))
    assert abs(a.tanfovx - 64 / (2 * 60.0)) < 1e-6


def test_libtorch_shim_exports_reference_entry_points():
    """The C++/LibTorch host layer (csrc/torch_shim) builds, loads against libsegs_raster.so and
    exposes the six entry points of include/rasterize_points.h:18-102 + spatial.h:14 under the
    reference's names; shape errors are raised before any device work."""
    import torch
    from segs_slam_b200 import _segs_torch as shim
    for n in ("RasterizeGaussiansCUDA", "RasterizeGaussiansBackwardCUDA", "markVisible",
              "RasterizeGaussiansfilterCUDA", "RasterizeGaussiansprojectCUDA", "distCUDA2"):
        assert hasattr(shim, n), n
    e = torch.empty(0)
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        shim.RasterizeGaussiansCUDA(torch.zeros(3), torch.zeros((4, 4)), e, e, e, e, 1.0, e, torch.eye(4),
                                    torch.eye(4), 1.0, 1.0, 16, 16, e, 0, torch.zeros(3), False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        shim.distCUDA2(torch.zeros((4, 3)))
    x = torch.zeros((4, 3))
    with pytest.raises(RuntimeError, match="excatly one of either SHs or precomputed colors"):
        shim.rasterizer_forward(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3),
                                False, x, x, x[:, :1], e, e, x, torch.zeros((4, 4)), e)
    with pytest.raises(RuntimeError, match="exactly one of either scale/rotation pair"):
        shim.rasterizer_forward(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3),
                                False, x, x, x[:, :1], e, x, x, e, e)


def test_mapper_view_rejects_null_arguments():
    """Host-side validation of the fused mapper entry points needs no GPU."""
    import ctypes as C
    from segs_slam_b200 import _lib
    lib = _lib.load()
    ws = C.c_void_p()
    assert lib.segs_workspace_create(C.byref(ws)) == 0
    assert lib.segs_workspace_bytes(ws) == 0
    args, res = _lib.MapperViewArgs(), _lib.MapperViewResult()
    assert lib.segs_mapper_view(ws, C.byref(args), C.byref(res), None) != 0
    assert b"invalid sizes" in lib.segs_last_error()
    args.A, args.width, args.height = 10, 16, 16
    assert lib.segs_mapper_view(ws, C.byref(args), C.byref(res), None) != 0
    assert b"NULL required pointer" in lib.segs_last_error()
    assert lib.segs_workspace_destroy(ws) == 0
    assert lib.segs_loss_state_bytes(3, 680, 1200) > 3 * 3 * 680 * 1200 * 4
    assert lib.segs_adam_step(1, None, None, None, None, 1.0, 0, None) != 0


def test_expon_lr_schedule_matches_reference_formula():
    """getExponLrFunc (gaussian_model.cpp:1390-1407): end points, geometric midpoint, clamping, the zero shortcut."""
    import math
    from segs_slam_b200.optim import get_expon_lr_func as lr
    assert lr(-1, 1e-2, 1e-4) == 0.0 and lr(10, 0.0, 0.0) == 0.0
    assert abs(lr(0, 1e-2, 1e-4, max_steps=1000) - 1e-2) < 1e-8
    assert abs(lr(1000, 1e-2, 1e-4, max_steps=1000) - 1e-4) < 1e-10
    assert abs(lr(5000, 1e-2, 1e-4, max_steps=1000) - 1e-4) < 1e-10                      # clamped
    assert abs(lr(500, 1e-2, 1e-4, max_steps=1000) - 1e-3) < 1e-8                        # log-linear: geometric mean
    # sine warm-up: delay_mult at step 0, 1 at lr_delay_steps
    assert abs(lr(0, 1e-2, 1e-4, 0.01, 1000, lr_delay_steps=100) - 1e-4) < 1e-9
    mid = lr(50, 1e-2, 1e-2, 0.01, 1000, lr_delay_steps=100)
    assert abs(mid - 1e-2 * (0.01 + 0.99 * math.sin(math.pi / 4))) < 1e-7


def test_header_is_plain_c(tmp_path):
    """include/segs_raster.h is a C ABI: it must compile as C99 (no C++-isms, no torch types) and link by name."""
    import subprocess
    src = tmp_path / "c_abi.c"
    src.write_text('#include "segs_raster.h"\n'
                   'int main(void) {\n'
                   '    segs_mapper_view_args a; segs_raster_view_args r; segs_adam_tensor t; segs_decode_params p;\n'
                   '    (void)a; (void)r; (void)t; (void)p;\n'
                   '    return segs_version() == SEGS_ABI_VERSION ? 0 : 1;\n'
                   '}\n')
    inc = os.path.join(ROOT, "include")
    lib = os.path.join(ROOT, "segs_slam_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", f"-I{inc}", str(src)])
    exe = tmp_path / "c_abi"
    subprocess.check_call(["gcc", "-std=c99", f"-I{inc}", str(src), "-o", str(exe), f"-L{lib}", "-lsegs_raster",
                           f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    assert subprocess.run([str(exe)]).returncode == 0          # segs_version() needs no GPU


def test_decode_variant_switch():
    """segs_decode_set_variant / segs_decode_get_variant (include/segs_raster.h): 2 is the default unless the environment
    says otherwise, only 1 and 2 are accepted (no compute call: runs without a GPU)."""
    from segs_slam_b200 import _lib
    lib = _lib.load()
    before = lib.segs_decode_get_variant()
    assert before in (1, 2)
    if "SEGS_DECODE_VARIANT" not in os.environ:
        assert before == 2
    try:
        assert lib.segs_decode_set_variant(1) == 0 and lib.segs_decode_get_variant() == 1
        assert lib.segs_decode_set_variant(2) == 0 and lib.segs_decode_get_variant() == 2
        assert lib.segs_decode_set_variant(3) != 0 and lib.segs_decode_get_variant() == 2
        assert lib.segs_decode_state_bytes(1000) >= 1000 * (12 + 4 * (96 + 110))      # the state carries the activations
    finally:
        lib.segs_decode_set_variant(before)
