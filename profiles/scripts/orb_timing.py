"""API-level latency of the ORB entry points on one B200 (host buffers, copies and synchronisation included):
descriptors for every corner of a C2 frame, and SearchPoint with descriptor scoring vs ZMSSD scoring for the same 500
candidates.  Run: python profiles/scripts/orb_timing.py > gpurun_out/orb_timing.txt"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402

conftest.load_pkg()
sw = importlib.import_module("slam_sdvl_b200.synthworld")
scenes = importlib.import_module("slam_sdvl_b200.scenes")
binding = importlib.import_module("slam_sdvl_b200.binding")


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


cfg, poses, imgs = sw.sequence("C2", 0, 4)
P, cam = cfg["params"], cfg["cam"]
for orb in (True, False):
    ctx = binding.Context(P, cam)
    ctx.set_orb(orb)
    ref = ctx.frame(imgs[0], corners=True)
    cur = ctx.frame(imgs[3], corners=True)
    xyl, _ = ref.corners()
    pts = scenes.seed_points(cfg, xyl, poses[0], max_points=500, one_per_cell=False, margin=20)
    cg = scenes.candidates(pts, poses[0], ref.h, fixed=True, project=True)
    if orb:
        pos = np.concatenate([(pts["px"] / (1 << pts["level"])[:, None]).astype(np.int32), pts["level"][:, None].astype(np.int32)], axis=1)
        q, _ = ref.orb_descriptors(pos)
        t_desc = timed(lambda: ref.orb_descriptors(xyl))
        t_s = timed(lambda: ctx.search_points_orb(cur, cg, poses[3], q))
        print(f"ORB descriptors of {len(xyl)} corners (upload positions, kernel, read back 32 B each): {t_desc:.1f} us per call "
              f"= {t_desc / len(xyl) * 1e3:.1f} ns per descriptor")
        print(f"sdvlb_search_points_orb, {len(cg)} candidates (descriptors of all {len(xyl)} corners of the current frame + scoring + LK): {t_s:.1f} us per call")
    else:
        t_s = timed(lambda: ctx.search_points(cur, cg, poses[3]))
        print(f"sdvlb_search_points (ZMSSD), {len(cg)} candidates, {len(xyl)} corners: {t_s:.1f} us per call")
    ref.destroy(); cur.destroy()
    ctx.close()
