"""Quick device timing of product vs reference (fwd, bwd) on one config. Dev tool."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common, refimpl
from segs_slam_b200 import synth

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts)//2], ts[0]

def main():
    names = sys.argv[1:] or ["C1", "C2"]
    dev = torch.device("cuda:0")
    for name in names:
        scene = synth.config(name)
        t = scene.to_torch(dev)
        a = common.scene_args(t, scene, dev)
        m = common.run_mine(a, t["dL_dout"])
        print(name, "P", scene.P, "R", m["R"], "visible", int((m["radii"] > 0).sum()))
        from segs_slam_b200 import rasterize_points as rp
        def mine_fwd():
            return rp.RasterizeGaussiansCUDA(a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], 1.0,
                a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"], a["sh"], 0, a["campos"], False)
        st = mine_fwd()
        def mine_bwd():
            return rp.RasterizeGaussiansBackwardCUDA(a["bg"], a["means3D"], st[2], a["colors"], a["scales"], a["rotations"], 1.0,
                a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], t["dL_dout"], a["sh"], 0, a["campos"], st[3], st[0], st[4], st[5])
        print("  mine fwd  med/min ms", timeit(mine_fwd))
        print("  mine bwd  med/min ms", timeit(mine_bwd))
        print("  mine f+b  med/min ms", timeit(lambda: (mine_fwd(), mine_bwd())))
        if refimpl.available():
            def ref_fwd():
                return refimpl.forward(a["bg"], a["means3D"], a["colors"], a["opacity"], a["scales"], a["rotations"], 1.0,
                    a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], a["H"], a["W"], a["sh"], 0, a["campos"])
            rs = ref_fwd()
            def ref_bwd():
                return refimpl.backward(a["bg"], a["means3D"], rs[2], a["colors"], a["scales"], a["rotations"], 1.0,
                    a["cov3D_precomp"], a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], t["dL_dout"], a["sh"], 0, a["campos"], rs[3], rs[0], rs[4], rs[5])
            print("  ref  fwd  med/min ms", timeit(ref_fwd))
            print("  ref  bwd  med/min ms", timeit(ref_bwd))
            print("  ref  f+b  med/min ms", timeit(lambda: (ref_fwd(), ref_bwd())))

if __name__ == "__main__":
    main()
