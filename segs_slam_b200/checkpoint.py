"""The reference's anchor checkpoint format: GaussianModel::savePly / loadPly
(/root/reference/src/gaussian_model.cpp:1179-1256, 1040-1177) and save_mlp_checkpoints (:1257-1317).

* Anchors: a binary little-endian PLY written through tinyply — one `vertex` element of float properties
  x y z | nx ny nz (zeros) | anchor_feat_0..31 | offset_0..29 | opacity | scale_0..5 | rot_0..3, one row per anchor
  (79 floats).  `offset_*` is `_offset.transpose(1, 2).flatten(1)`: component-major, i.e. offset_{c*10+o} = _offset[a, o, c]
  (:1185, and the inverse reshape/transposition at :1153-1157).  The reference's own loader asks for the Scaffold-GS
  names `f_anchor_feat_*` / `f_offset_*` (:1085-1101) while its saver writes `anchor_feat_*` / `offset_*`; `load_ply`
  accepts both, `save_ply(..., scaffold_names=True)` writes the former.
* MLPs: one text file per tensor, rows of `%.5f` separated by blanks (saveTensorToTxt, :262-286).  Biases are 1-D and the
  reference indexes `sizes[1]` on them (out of range); they are written here as ONE row.

Host-side I/O only: numpy + torch, no kernels."""
from __future__ import annotations

import os

import numpy as np
import torch



def _property_names(feat_dim: int, n_offsets: int, scaffold_names: bool) -> list[str]:
    pre = "f_" if scaffold_names else ""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"{pre}anchor_feat_{i}" for i in range(feat_dim)]
    names += [f"{pre}offset_{i}" for i in range(3 * n_offsets)]
    names += ["opacity"] + [f"scale_{i}" for i in range(6)] + [f"rot_{i}" for i in range(4)]
    return names


def save_ply(pc, path: str, scaffold_names: bool = False) -> None:
    """GaussianModel::savePly.  `pc` carries _anchor [A,3], _anchor_feat [A,F], _offset [A,k,3], _scaling [A,6],
    _rotation [A,4] and optionally _opacity [A,1] (zeros otherwise)."""
    f32 = lambda t: t.detach().to("cpu", torch.float32).contiguous().numpy()
    anchor = f32(pc._anchor)
    A = anchor.shape[0]
    feat = f32(pc._anchor_feat)
    offset = f32(pc._offset.detach().transpose(1, 2).flatten(1))
    opacity = f32(pc._opacity) if hasattr(pc, "_opacity") else np.zeros((A, 1), np.float32)
    rows = np.concatenate([anchor, np.zeros_like(anchor), feat, offset, opacity.reshape(A, 1), f32(pc._scaling),
                           f32(pc._rotation)], axis=1).astype("<f4")
    names = _property_names(feat.shape[1], pc._offset.size(1), scaffold_names)
    assert rows.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {A}\n" + \
        "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rows.tobytes())


def load_ply(path: str) -> dict[str, torch.Tensor]:
    """GaussianModel::loadPly: -> {_anchor, _anchor_feat, _offset [A,k,3], _opacity, _scaling, _rotation} (CPU tensors)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, names, in_vertex = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
                elif count is None:
                    raise ValueError(f"{path}: elements before `vertex` are not supported")
            elif tok[0] == "property" and in_vertex:
                if tok[1] not in ("float", "float32"):
                    raise ValueError(f"{path}: vertex property {tok[-1]} is {tok[1]}, expected float")
                names.append(tok[-1])
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian" or count is None:
            raise ValueError(f"{path}: expected a binary_little_endian PLY with a vertex element")
        raw = f.read(4 * count * len(names))
    if len(raw) != 4 * count * len(names):
        raise ValueError(f"{path}: truncated vertex data")
    data = np.frombuffer(raw, dtype="<f4")
    rows = data.reshape(count, len(names))
    col = {n: i for i, n in enumerate(names)}

    def take(cands):
        for c in cands:
            if all(n in col for n in c):
                return torch.from_numpy(np.ascontiguousarray(rows[:, [col[n] for n in c]]))
        raise ValueError(f"{path}: missing properties {cands[0][:2]}...")

    n_feat = sum(1 for n in names if n.startswith("anchor_feat_") or n.startswith("f_anchor_feat_"))
    n_off = sum(1 for n in names if n.startswith("offset_") or n.startswith("f_offset_"))
    off = take([[f"offset_{i}" for i in range(n_off)], [f"f_offset_{i}" for i in range(n_off)]])
    return {
        "_anchor": take([["x", "y", "z"]]),
        "_anchor_feat": take([[f"anchor_feat_{i}" for i in range(n_feat)], [f"f_anchor_feat_{i}" for i in range(n_feat)]]),
        "_offset": off.reshape(count, 3, -1).transpose(1, 2).contiguous(),          # gaussian_model.cpp:1153-1157
        "_opacity": take([["opacity"]]),
        "_scaling": take([[f"scale_{i}" for i in range(6)]]),
        "_rotation": take([[f"rot_{i}" for i in range(4)]]),
    }


def save_sparse_points_ply(xyz: torch.Tensor, color: torch.Tensor, path: str) -> None:
    """GaussianModel::saveSparsePointsPly (gaussian_model.cpp:1319-1352): x y z | nx ny nz (zeros) as float, then
    red green blue as uchar = (color * 255) truncated like `toType(torch::kUInt8)`; one packed 27-byte row per point."""
    p = xyz.detach().to("cpu", torch.float32).contiguous().numpy()
    c = (color.detach().to("cpu", torch.float32) * 255.0).to(torch.uint8).contiguous().numpy()
    n = p.shape[0]
    rows = np.zeros(n, dtype=[("p", "<f4", 3), ("n", "<f4", 3), ("c", "u1", 3)])
    rows["p"], rows["c"] = p, c
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n" + \
        "".join(f"property float {k}\n" for k in ("x", "y", "z", "nx", "ny", "nz")) + \
        "".join(f"property uchar {k}\n" for k in ("red", "green", "blue")) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rows.tobytes())


def load_into(pc, tensors: dict[str, torch.Tensor]) -> None:
    """Copy a `load_ply` result into a model with the reference's member names (shapes must match)."""
    with torch.no_grad():
        for k, v in tensors.items():
            if hasattr(pc, k):
                getattr(pc, k).copy_(v.to(getattr(pc, k).device))


_MLP_FILES = (("mlp_opacity", "opacity"), ("mlp_cov", "cov"), ("mlp_color", "color"), ("mlp_feature_bank", "feat"))
# Extension: the reference's save_mlp_checkpoints (gaussian_model.cpp:1262-1317) does NOT write mlp_apperance — the
# Linear(7, appearance_dim) that this repo's decode trains — so its own round trip loses it.  It is written here under an
# extra pair of files the reference ignores; loading tolerates their absence (a checkpoint written by the reference).
_EXTRA_MLP_FILES = (("mlp_apperance", "appearance"),)


def _linears(seq):
    return [m for m in seq if isinstance(m, torch.nn.Linear)]


def save_tensor_txt(t: torch.Tensor, path: str) -> None:
    """saveTensorToTxt (gaussian_model.cpp:262-286): `%.5f`, blank-separated, one row per line."""
    a = t.detach().to("cpu", torch.float32).numpy()
    a = a.reshape(1, -1) if a.ndim == 1 else a
    with open(path, "w") as f:
        for row in a:
            f.write(" ".join(f"{float(v):.5f}" for v in row) + "\n")


def load_tensor_txt(path: str) -> torch.Tensor:
    rows = [[float(x) for x in line.split()] for line in open(path) if line.strip()]
    return torch.tensor(rows, dtype=torch.float32)


def save_mlp_checkpoints(pc, result_path: str) -> None:
    """GaussianModel::save_mlp_checkpoints: <name>_weight{1,2}.txt / <name>_bias{1,2}.txt for the opacity, cov, color
    and (when used) feature-bank MLPs."""
    os.makedirs(result_path, exist_ok=True)
    for attr, name in _MLP_FILES + _EXTRA_MLP_FILES:
        seq = getattr(pc, attr, None)
        if seq is None:
            continue
        for i, lin in enumerate(_linears(seq), start=1):
            save_tensor_txt(lin.weight, os.path.join(result_path, f"{name}_weight{i}.txt"))
            save_tensor_txt(lin.bias, os.path.join(result_path, f"{name}_bias{i}.txt"))
    emb = getattr(pc, "embedding_appearance", None)             # :1312-1315 (nn.Embedding weight, when the model has one)
    if emb is not None and getattr(pc, "appearance_dim", 0) > 0:
        w = emb.weight if hasattr(emb, "weight") else emb.get_embedding().weight
        save_tensor_txt(w, os.path.join(result_path, "embedding_weight.txt"))


def load_mlp_checkpoints(pc, result_path: str) -> None:
    with torch.no_grad():
        for attr, name in _MLP_FILES + _EXTRA_MLP_FILES:
            seq = getattr(pc, attr, None)
            if seq is None:
                continue
            for i, lin in enumerate(_linears(seq), start=1):
                if (attr, name) in _EXTRA_MLP_FILES and not os.path.exists(os.path.join(result_path, f"{name}_weight{i}.txt")):
                    continue                                    # written by the reference: it has no such file
                w = load_tensor_txt(os.path.join(result_path, f"{name}_weight{i}.txt"))
                b = load_tensor_txt(os.path.join(result_path, f"{name}_bias{i}.txt")).reshape(-1)
                lin.weight.copy_(w.to(lin.weight.device).view_as(lin.weight))
                lin.bias.copy_(b.to(lin.bias.device))
        emb = getattr(pc, "embedding_appearance", None)
        path = os.path.join(result_path, "embedding_weight.txt")
        if emb is not None and os.path.exists(path):
            w = emb.weight if hasattr(emb, "weight") else emb.get_embedding().weight
            w.copy_(load_tensor_txt(path).to(w.device).view_as(w))
