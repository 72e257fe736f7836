"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference CUDA
rasterizer and simple-knn (oracle/_ref/libsegs_ref.so, built from /root/reference by
`make -C oracle ref`) on small seeded scenes ON A GPU.

    python tests/golden/make_golden.py [outdir]        # on the B200 box: outdir=gpurun_out/golden

The fixtures pin the CPU oracle (oracle/raster_oracle.cpp) in the `-m "not gpu"` suite and are
an independent check of the product in the `-m gpu` suite.  Each .npz holds the inputs and
every observable of the reference: radii, tiles_touched, depths, means2D, conic_opacity,
cov3D, num_rendered, point_list, point_list_keys, ranges, n_contrib, final_T, colour and
all gradients (two reference runs, to record the reference's own atomic nondeterminism).
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
import refimpl  # noqa: E402
from segs_slam_b200 import synth  # noqa: E402


def golden_scenes():
    out = {"tiny": (synth.config("tiny"), {})}
    s = synth.synth(3000, 75, 53, 70.0, 66.0, 33, bg=(0.3, 0.6, 0.1))
    s.scales *= 3.0
    a = 0.3
    R = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]], dtype=np.float32)
    out["rot_bg"] = (synth.with_camera(s, R, np.array([0.2, -0.1, 0.3], dtype=np.float32)), {})
    out["sh3"] = (synth.sh_variant(synth.synth(2000, 64, 48, 60.0, 60.0, 44), 3), {"use_sh": True})
    return out


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, (scene, kw) in golden_scenes().items():
        t = scene.to_torch(dev)
        a = common.scene_args(t, scene, dev, **kw)
        r = common.run_ref(a, t["dL_dout"])
        r2 = common.run_ref(a, t["dL_dout"])
        torch.cuda.synchronize()
        P, W, H = scene.P, scene.W, scene.H
        N, T = W * H, ((W + 15) // 16) * ((H + 15) // 16)
        g = refimpl.parse_geom(r["geom"], P)
        b = refimpl.parse_binning(r["binning"], r["R"])
        i = refimpl.parse_image(r["img"], N, T)
        vis = (r["radii"] > 0)
        z = lambda x: torch.where(vis.view(-1, *([1] * (x.dim() - 1))), x, torch.zeros_like(x))
        d = dict(
            P=P, W=W, H=H, tanfovx=np.float32(scene.tanfovx), tanfovy=np.float32(scene.tanfovy),
            means3D=scene.means3D, scales=scene.scales, rotations=scene.rotations, opacities=scene.opacities,
            colors=scene.colors, viewmatrix=scene.viewmatrix, projmatrix=scene.projmatrix, campos=scene.campos,
            bg=scene.bg, dL_dout=scene.dL_dout,
            R=r["R"], radii=common.to_np(r["radii"]), tiles_touched=common.to_np(g["tiles_touched"]),
            depths=common.to_np(z(g["depths"])), means2D=common.to_np(z(g["means2D"])),
            conic_opacity=common.to_np(z(g["conic_opacity"])), cov3D=common.to_np(z(g["cov3D"])),
            rgb=common.to_np(z(g["rgb"])) if kw.get("use_sh") else np.zeros((0,), np.float32),
            point_list=common.to_np(b["point_list"]), point_list_keys=common.to_np(b["point_list_keys"]),
            ranges=common.to_np(i["ranges"]), n_contrib=common.to_np(i["n_contrib"]),
            final_T=common.to_np(i["final_T"]), color=common.to_np(r["color"]),
        )
        if kw.get("use_sh"):
            d["sh"] = scene.extras["sh"]
            d["sh_degree"] = scene.extras["sh_degree"]
        for k, v in r["grads"].items():
            d["g_" + k] = common.to_np(v)
            d["g2_" + k] = common.to_np(r2["grads"][k])
        np.savez_compressed(os.path.join(outdir, f"raster_{name}.npz"), **d)
        print(name, "P", P, "R", r["R"], "visible", int(vis.sum()))

    # anchor prefilter on the tiny scene with some points behind the camera
    scene = synth.config("tiny")
    scene.means3D[::5, 2] *= -1.0
    t = scene.to_torch(dev)
    e = common.empty(dev)
    radii = refimpl.visible_filter(t["means3D"], t["scales"], t["rotations"], 1.0, e, t["viewmatrix"],
                                   t["projmatrix"], scene.tanfovx, scene.tanfovy, scene.H, scene.W)
    present = refimpl.mark_visible(t["means3D"], t["viewmatrix"], t["projmatrix"])
    np.savez_compressed(os.path.join(outdir, "filter_tiny.npz"), means3D=scene.means3D, scales=scene.scales,
                        rotations=scene.rotations, viewmatrix=scene.viewmatrix, projmatrix=scene.projmatrix,
                        tanfovx=np.float32(scene.tanfovx), tanfovy=np.float32(scene.tanfovy), W=scene.W, H=scene.H,
                        radii=common.to_np(radii), present=common.to_np(present))

    # kNN
    rng = np.random.default_rng(17)
    pts = np.concatenate([rng.uniform(-1, 1, (2500, 3)), rng.normal(0.5, 0.02, (700, 3))]).astype(np.float32)
    d2 = refimpl.knn(torch.from_numpy(pts).to(dev))
    np.savez_compressed(os.path.join(outdir, "knn_3200.npz"), points=pts, dist2=common.to_np(d2))
    print("golden fixtures written to", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
