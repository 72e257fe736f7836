"""Writes tests/golden/ingest_*.npz with the REAL OpenCV (cv2): the maps of cv2.initUndistortRectifyMap, cv2.remap
(INTER_LINEAR) of a seeded image through them, and cv2.resize (INTER_LINEAR) to two pyramid sizes — the calls of the
reference's keyframe ingest (include/camera.h:70-115, src/gaussian_mapper.cpp:621-632).

    python tests/golden/make_ingest_golden.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, H, W, K, dist, seed, sizes):
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 1, (H, W, 3)).astype(np.float32)
    K = np.asarray(K, dtype=np.float64)
    m1, m2 = cv2.initUndistortRectifyMap(K, np.asarray(dist, dtype=np.float64), np.eye(3), K, (W, H), cv2.CV_32F)
    und = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
    mask = cv2.remap(np.ones((H, W, 3), np.float32), m1, m2, cv2.INTER_LINEAR)
    d = dict(image=img, K=K, dist=np.asarray(dist, np.float64), map_x=m1, map_y=m2, undistorted=und, mask=mask)
    for i, (h, w) in enumerate(sizes):
        d[f"resized_{i}"] = cv2.resize(und, (w, h), interpolation=cv2.INTER_LINEAR)
        d[f"size_{i}"] = np.asarray([h, w])
    np.savez_compressed(os.path.join(HERE, f"ingest_{name}.npz"), **d)
    print(name, cv2.__version__, und.shape)


if __name__ == "__main__":
    # TUM fr1-like intrinsics / distortion (cfg/ORB_SLAM3/RGB-D/TUM/tum_freiburg1_desk.yaml), reduced image size
    case("tum_96x128", 96, 128, [[103.5, 0, 63.7], [0, 103.3, 51.1], [0, 0, 1]], [0.2624, -0.9531, -0.0054, 0.0026], 3, [(48, 64), (33, 47)])
    case("wide_60x80", 60, 80, [[40.0, 0, 39.5], [0, 41.0, 29.5], [0, 0, 1]], [-0.28, 0.07, 0.001, -0.0007], 4, [(30, 40), (45, 60)])
