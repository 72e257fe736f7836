"""Synthetic Gaussian scenes of BASELINE.md §3 / SURVEY.md §8(d): `synth(P,W,H,fx,fy,seed)`.

numpy only (so the CPU oracle tests can use it without a GPU); `Scene.to_torch(device)` moves
a scene onto a device.  The camera model restates GaussianKeyframe::getProjectionMatrix
(/root/reference/src/gaussian_keyframe.cpp:251-279) with znear = 0.01, zfar = 100; view = I.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# name -> (P, W, H, fx, fy, seed)   (BASELINE.json configs made concrete)
CONFIGS = {
    "C1": (100_000, 640, 480, 517.3, 516.5, 1001),    # TUM fr1 intrinsics
    "C2": (1_000_000, 1200, 680, 600.0, 600.0, 1002),  # Replica
    "C5": (5_000_000, 1920, 1080, 960.0, 960.0, 1005),  # ScanNet++-scale stress
    "tiny": (4096, 64, 48, 60.0, 60.0, 7),
    "small": (20_000, 200, 120, 150.0, 150.0, 11),
}


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """gaussian_keyframe.cpp:251-279 (row-major P applied to column vectors)."""
    f = np.float32
    tan_y = f(math.tan(fovy / 2))
    tan_x = f(math.tan(fovx / 2))
    top = tan_y * f(znear)
    bottom = -top
    right = tan_x * f(znear)
    left = -right
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


@dataclass
class Scene:
    P: int
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    means3D: np.ndarray          # [P,3]
    scales: np.ndarray           # [P,3]
    rotations: np.ndarray        # [P,4] (r,x,y,z)
    opacities: np.ndarray        # [P,1]
    colors: np.ndarray           # [P,3]
    viewmatrix: np.ndarray       # [4,4] memory order m[4*col+row]
    projmatrix: np.ndarray       # [4,4] memory order m[4*col+row]
    campos: np.ndarray           # [3]
    bg: np.ndarray               # [3]
    dL_dout: np.ndarray          # [3,H,W]
    scale_modifier: float = 1.0
    extras: dict = field(default_factory=dict)

    def to_torch(self, device):
        import torch
        out = {}
        for k in ("means3D", "scales", "rotations", "opacities", "colors", "viewmatrix", "projmatrix",
                  "campos", "bg", "dL_dout"):
            out[k] = torch.from_numpy(np.ascontiguousarray(getattr(self, k))).to(device)
        for k, v in self.extras.items():
            out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device)
        return out


def synth(P: int, W: int, H: int, fx: float, fy: float, seed: int, bg=(0.0, 0.0, 0.0)) -> Scene:
    rng = np.random.default_rng(seed)
    f32 = np.float32
    fovx = 2.0 * math.atan(W / (2.0 * fx))
    fovy = 2.0 * math.atan(H / (2.0 * fy))
    tanfovx = math.tan(fovx * 0.5)
    tanfovy = math.tan(fovy * 0.5)
    view = np.eye(4, dtype=f32)
    proj = projection_matrix(0.01, 100.0, fovx, fovy)
    full = (view @ proj.T).astype(f32)     # world_view_transform (= I^T) @ projection_matrix^T

    z = rng.uniform(0.5, 6.0, P).astype(f32)
    x_ndc = rng.uniform(-1.1, 1.1, P).astype(f32)
    y_ndc = rng.uniform(-1.1, 1.1, P).astype(f32)
    means = np.stack([x_ndc * f32(tanfovx) * z, y_ndc * f32(tanfovy) * z, z], axis=1).astype(f32)
    scales = np.exp(rng.normal(math.log(0.01), 0.5, (P, 3))).astype(f32)
    rot = rng.normal(0.0, 1.0, (P, 4))
    rot = (rot / np.linalg.norm(rot, axis=1, keepdims=True)).astype(f32)
    opacity = rng.uniform(0.05, 1.0, (P, 1)).astype(f32)
    colors = rng.uniform(0.0, 1.0, (P, 3)).astype(f32)
    dL = np.random.default_rng(seed + 1).normal(0.0, 1.0, (3, H, W)).astype(f32)
    return Scene(P=P, W=W, H=H, tanfovx=float(f32(tanfovx)), tanfovy=float(f32(tanfovy)),
                 means3D=means, scales=scales, rotations=rot, opacities=opacity, colors=colors,
                 viewmatrix=view, projmatrix=full, campos=np.zeros(3, dtype=f32),
                 bg=np.asarray(bg, dtype=f32), dL_dout=dL)


def config(name: str, **kw) -> Scene:
    P, W, H, fx, fy, seed = CONFIGS[name]
    return synth(P, W, H, fx, fy, seed, **kw)


def camera_matrices(R: np.ndarray, t: np.ndarray, tanfovx: float, tanfovy: float):
    """(world_view_transform_, full_proj_transform_, camera_center_) of a world->camera pose (R, t) in
    the memory order the kernels read (m[4*col+row]), as GaussianKeyframe::computeTransformTensors
    builds them (gaussian_keyframe.cpp:151-184); znear 0.01, zfar 100."""
    f32 = np.float32
    Rt = np.eye(4, dtype=f32)
    Rt[:3, :3] = R
    Rt[:3, 3] = t
    fovx = 2.0 * math.atan(tanfovx)
    fovy = 2.0 * math.atan(tanfovy)
    proj = projection_matrix(0.01, 100.0, fovx, fovy)
    wvt = Rt.T.copy()                       # world_view_transform_ (memory m[4*col+row])
    full = (wvt @ proj.T).astype(f32)
    campos = (-np.asarray(R, dtype=f32).T @ np.asarray(t, dtype=f32)).astype(f32)
    return wvt.astype(f32), full, campos


def with_camera(scene: Scene, R: np.ndarray, t: np.ndarray) -> Scene:
    """Same Gaussians seen from a world->camera pose (R,t): view = [[R,t],[0,1]]."""
    wvt, full, campos = camera_matrices(R, t, scene.tanfovx, scene.tanfovy)
    import dataclasses
    return dataclasses.replace(scene, viewmatrix=wvt, projmatrix=full, campos=campos)


def sh_variant(scene: Scene, degree: int, seed: int = 5) -> Scene:
    """Replace precomputed colours by SH coefficients [P,16,3] (exercises the SH branch)."""
    rng = np.random.default_rng(seed)
    sh = (rng.normal(0.0, 0.3, (scene.P, 16, 3))).astype(np.float32)
    sh[:, 0, :] += 1.0
    import dataclasses
    s = dataclasses.replace(scene, extras=dict(scene.extras, sh=sh))
    s.extras["sh_degree"] = np.asarray([degree], dtype=np.int32)
    return s
