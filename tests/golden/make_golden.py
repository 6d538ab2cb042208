"""Generates tests/golden/cv2_golden.npz from OpenCV (cv2 4.13 in this image): the third-party arithmetic the
reference calls on this path (cv::pyrDown at frame.cc:119, cv::FAST at extra/fast_detector.cc:95, cv::undistort at
camera.cc:102).
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402

load_pkg()
import importlib  # noqa: E402

sw = importlib.import_module("slam_sdvl_b200.synthworld")

out = {"cv2_version": np.array(cv2.__version__)}
rng = np.random.default_rng(2024)
# --- pyrDown chains: a synthetic crop, noise, odd sizes
cfg, poses, imgs = sw.sequence("C1", 3, 1)
crop = np.ascontiguousarray(imgs[0][100:228, 200:392])        # 192 x 128
noise = rng.integers(0, 256, (67, 135), dtype=np.uint8)         # odd sizes (135 -> 67 -> 33)
for name, img, levels in [("crop", crop, 4), ("noise", noise, 3)]:
    out[f"pyr_{name}_0"] = img
    cur = img
    for l in range(1, levels):
        cur = cv2.pyrDown(cur, dstsize=(cur.shape[1] // 2, cur.shape[0] // 2))
        out[f"pyr_{name}_{l}"] = cur
# --- FAST(10, nonmax) on cell-sized ROIs of the crop and of blurred noise
fd = cv2.FastFeatureDetector_create(10, True)
blur = cv2.GaussianBlur(rng.integers(0, 256, (96, 96), dtype=np.uint8), (3, 3), 0.8)
full = imgs[0]
picked = []
for cy in range(0, full.shape[0] - 32, 32):
    for cx in range(0, full.shape[1] - 32, 32):
        if len(fd.detect(np.ascontiguousarray(full[cy:cy + 32, cx:cx + 32]))) >= 3 and len(picked) < 3:
            picked.append((full, cx, cy, 32 if len(picked) else 27, 32 if len(picked) else 27))
for k, (src, x0, y0, cw, ch) in enumerate(picked + [(blur, 0, 0, 32, 32), (blur, 40, 50, 11, 19), (blur, 7, 3, 7, 7)]):
    roi = np.ascontiguousarray(src[y0:y0 + ch, x0:x0 + cw])
    kps = fd.detect(roi)
    arr = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in kps], np.int32).reshape(-1, 3)
    out[f"fast_roi_{k}"] = roi
    out[f"fast_kps_{k}"] = arr
out["n_rois"] = np.array(6)
# --- cv::undistort (Camera::UndistortImage, camera.cc:100-105): EuRoC cam0 coefficients on a crop-sized camera, and a
# strongly distorted small camera whose corners map outside the source (BORDER_CONSTANT)
und_cases = [("euroc", 188, 120, 114.66, 114.32, 91.8, 62.1, (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)),
             ("strong", 135, 67, 90.0, 85.0, 70.3, 31.9, (-0.45, 0.21, 0.004, -0.003, -0.05)),
             ("barrel", 96, 96, 80.0, 80.0, 47.5, 47.5, (0.35, -0.2, -0.002, 0.001, 0.3))]
for name, w, h, fx, fy, u0, v0, D in und_cases:
    src = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 1.2)
    K = np.array([[fx, 0, u0], [0, fy, v0], [0, 0, 1.0]])
    out[f"und_{name}_src"] = src
    out[f"und_{name}_cam"] = np.array([w, h, fx, fy, u0, v0])
    out[f"und_{name}_D"] = np.array(D)
    out[f"und_{name}_dst"] = cv2.undistort(src, K, np.array(D))
out["und_names"] = np.array([c[0] for c in und_cases])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cv2_golden.npz"), **out)
print("written", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
