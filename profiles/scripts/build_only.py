"""Frame-batch throughput alone (upload + pyramid + FAST + selection, no tracking): how fast can frames get from pinned
host memory / HBM into built frames?  python profiles/scripts/build_only.py"""
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
cfg = sw.config("C2")
w, h = cfg["w"], cfg["h"]
L = binding.load()
S, R = 64, 40
poses = sw.trajectory(cfg, 0, 8)
imgs = sw.render(cfg, poses)
host = torch.from_numpy(np.ascontiguousarray(np.tile(imgs, (S // 8 + 1, 1, 1))[:S])).pin_memory()
dev = host.cuda()
for G in (1, 8):
    n = S // G
    ctxs = [binding.Context(cfg["params"], cfg["cam"]) for _ in range(G)]
    for c in ctxs:
        L.sdvlb_ctx_reserve_frames(C.c_void_p(c.h), 4 * n)
    for name, base, loc in (("pinned/kernel", host.data_ptr(), 2), ("pinned/dma", host.data_ptr(), 0), ("hbm", dev.data_ptr(), 1)):
        def submit(g):
            ptrs = (C.c_void_p * n)(*[base + (g * n + i) * w * h for i in range(n)])
            out = (C.c_void_p * n)()
            rc = L.sdvlb_frames_submit(C.c_void_p(ctxs[g].h), ptrs, n, loc, 1, cfg["params"].num_features, out)
            assert rc == 0
            return out
        def finish(g, out):
            assert L.sdvlb_frames_wait(C.c_void_p(ctxs[g].h), out, n) == 0
            for i in range(n):
                L.sdvlb_frame_destroy(C.c_void_p(ctxs[g].h), C.c_void_p(out[i]))
        pend = [[submit(g), submit(g)] for g in range(G)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in range(R):
            for g in range(G):
                finish(g, pend[g].pop(0))
                pend[g].append(submit(g))
        for g in range(G):
            for o in pend[g]:
                finish(g, o)
        dt = time.perf_counter() - t0
        print(f"groups={G} {name:14s}: {S * (R + 2) / dt:9.0f} frames/s  ({S * (R + 2) * w * h / dt / 1e9:.1f} GB/s of level-0 pixels)", flush=True)
    for c in ctxs:
        c.close()
