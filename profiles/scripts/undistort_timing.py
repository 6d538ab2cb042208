"""Cost of the fused undistortion: frame batches (64 device-resident 752x480 images, no FAST) built with and without
distortion set; the difference per batch is upload-to-scratch + undistort_kernel.
  python profiles/scripts/undistort_timing.py"""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.load_pkg()
sw = importlib.import_module("slam_sdvl_b200.synthworld")
binding = importlib.import_module("slam_sdvl_b200.binding")
cfg, poses, imgs = sw.sequence("C2", 0, 4)
n = 64
t = torch.from_numpy(np.ascontiguousarray(np.concatenate([imgs] * 16)[:n])).cuda()
stride = imgs.shape[1] * imgs.shape[2]
L = binding.load()
D = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
for label, dist in (("plain", (0, 0, 0, 0, 0)), ("undistort", D), ("plain", (0, 0, 0, 0, 0)), ("undistort", D)):
    ctx = binding.Context(cfg["params"], cfg["cam"])
    ctx.set_distortion(dist)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() + i * stride for i in range(n)])
    out = (C.c_void_p * n)()

    def batch():
        assert L.sdvlb_frames_submit(C.c_void_p(ctx.h), ptrs, n, 1, 0, 1000, out) == 0
        assert L.sdvlb_frames_wait(C.c_void_p(ctx.h), out, n) == 0
        for i in range(n):
            L.sdvlb_frame_destroy(C.c_void_p(ctx.h), C.c_void_p(out[i]))

    for _ in range(5):
        batch()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 40
    for _ in range(reps):
        batch()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label:10s}: {dt * 1e6:8.1f} us per batch of {n} frames (upload d2d + pyramid{' + undistort' if label != 'plain' else ''}), synchronous")
    ctx.close()
