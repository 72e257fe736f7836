"""CPU: the reference's anchor checkpoint format (segs_slam_b200/checkpoint.py) against a file written by the reference's
own PLY library (tests/golden/anchors_tinyply.ply, made by tests/golden/make_ply_golden.py), and — where oracle/_ref is
built — read back by it."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from segs_slam_b200 import checkpoint

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class _PC:
    pass


def _model_from(d):
    pc = _PC()
    pc._anchor, pc._anchor_feat, pc._offset = (torch.from_numpy(d[k].copy()) for k in ("anchor", "feat", "offset"))
    pc._opacity, pc._scaling, pc._rotation = (torch.from_numpy(d[k].copy()) for k in ("opacity", "scale", "rot"))
    return pc


def test_save_ply_is_byte_identical_to_tinyply(tmp_path):
    d = np.load(os.path.join(GOLD, "anchors_tinyply.npz"))
    out = str(tmp_path / "mine.ply")
    checkpoint.save_ply(_model_from(d), out)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "anchors_tinyply.ply"), "rb").read()


def test_load_ply_reads_the_tinyply_file():
    d = np.load(os.path.join(GOLD, "anchors_tinyply.npz"))
    t = checkpoint.load_ply(os.path.join(GOLD, "anchors_tinyply.ply"))
    for mine, ref in (("_anchor", "anchor"), ("_anchor_feat", "feat"), ("_offset", "offset"), ("_opacity", "opacity"),
                      ("_scaling", "scale"), ("_rotation", "rot")):
        assert np.array_equal(t[mine].numpy(), d[ref]), mine
    assert t["_offset"].shape == (7, 10, 3)


def test_scaffold_names_round_trip_and_errors(tmp_path):
    d = np.load(os.path.join(GOLD, "anchors_tinyply.npz"))
    out = str(tmp_path / "scaffold.ply")
    checkpoint.save_ply(_model_from(d), out, scaffold_names=True)            # the names the reference's loadPly asks for
    assert b"property float f_offset_29\n" in open(out, "rb").read()
    t = checkpoint.load_ply(out)
    assert np.array_equal(t["_offset"].numpy(), d["offset"])
    bad = str(tmp_path / "bad.ply")
    open(bad, "wb").write(open(out, "rb").read()[:-10])
    with pytest.raises(ValueError, match="truncated"):
        checkpoint.load_ply(bad)
    open(bad, "wb").write(b"plx\n")
    with pytest.raises(ValueError, match="not a PLY"):
        checkpoint.load_ply(bad)


def test_tinyply_reads_what_save_ply_writes(tmp_path):
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libply_ref.so")
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libply_ref.so not built (make -C oracle plyref)")
    lib = C.CDLL(lib_path)
    fp = C.POINTER(C.c_float)
    lib.ref_load_ply.argtypes = [C.c_char_p, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp]
    d = np.load(os.path.join(GOLD, "anchors_tinyply.npz"))
    out = str(tmp_path / "mine.ply")
    checkpoint.save_ply(_model_from(d), out)
    back = {k: np.zeros_like(d[k]) for k in ("anchor", "feat", "opacity", "scale", "rot")}
    flat = np.zeros((7, 30), np.float32)
    p = lambda a: a.ctypes.data_as(fp)
    n = lib.ref_load_ply(out.encode(), 32, 30, p(back["anchor"]), p(back["feat"]), p(flat), p(back["opacity"]), p(back["scale"]),
                         p(back["rot"]))
    assert n == 7
    for k in back:
        assert np.array_equal(back[k], d[k]), k
    assert np.array_equal(flat.reshape(7, 3, 10).transpose(0, 2, 1), d["offset"])


def test_mlp_text_checkpoints_round_trip(tmp_path):
    from segs_slam_b200 import anchor_model
    torch.manual_seed(3)
    a, b = anchor_model.AnchorModel(4), anchor_model.AnchorModel(4)
    checkpoint.save_mlp_checkpoints(a, str(tmp_path))
    first = open(tmp_path / "opacity_weight1.txt").readline().split()
    assert len(first) == 35 and all(len(x.split(".")[1]) == 5 for x in first)        # `%.5f`, blank-separated
    checkpoint.load_mlp_checkpoints(b, str(tmp_path))
    for (n1, p1), (_n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        if n1.startswith(("mlp_opacity", "mlp_cov", "mlp_color", "mlp_feature_bank")):
            assert torch.allclose(p1, p2, atol=5.1e-6), n1                          # 5 decimals


def test_sparse_points_ply_is_byte_identical_to_tinyply(tmp_path):
    d = np.load(os.path.join(GOLD, "sparse_tinyply.npz"))
    out = str(tmp_path / "sparse.ply")
    checkpoint.save_sparse_points_ply(torch.from_numpy(d["xyz"]), torch.from_numpy(d["color"]), out)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "sparse_tinyply.ply"), "rb").read()
