"""Extract the first keyframes of the reference's own transform dump (/root/reference/check_colmap.md — the
output of GaussianKeyframe::logger() for a real sequence: FoV, world_view_transform_, projection_matrix_,
full_proj_transform_, camera_center_, 4 decimals) into tests/golden/keyframe_transforms.json.

    python tests/golden/make_keyframe_golden.py

These are the only golden values the reference repository holds for anything near the hot path;
they pin the per-view constants of row T0 (segs_slam_b200/synth.py:camera_matrices / projection_matrix)."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/check_colmap.md"
N = 16

num = r"-?\d+\.\d+"
txt = open(SRC).read()
entries = txt.split("!KEY!")[1:]
out = []
for e in entries:
    m = re.search(r"fid_: (\d+), camera_id_ = (\d+), FoVx_ = ([\d.]+), FoVy_ = ([\d.]+), image_width_ = (\d+), image_height_ = (\d+)", e)
    if not m:
        continue
    blocks = {}
    for name in ("world_view_transform_", "projection_matrix_", "full_proj_transform_", "camera_center_"):
        seg = e.split(name + " =", 1)[1].split("[ CUDAFloatType", 1)[0]
        blocks[name] = [float(x) for x in re.findall(num, seg)]
    if len(blocks["world_view_transform_"]) != 16 or len(blocks["camera_center_"]) != 3:
        continue
    out.append(dict(fid=int(m.group(1)), FoVx=float(m.group(3)), FoVy=float(m.group(4)), width=int(m.group(5)),
                    height=int(m.group(6)), **blocks))
    if len(out) == N:
        break
json.dump(out, open(os.path.join(HERE, "keyframe_transforms.json"), "w"), indent=0)
print(len(out), "keyframes")
