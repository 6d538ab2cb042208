"""Seeded synthetic world (textured plane + trajectories) — Python face of synth/synth.cc.

Configs follow BASELINE.json / SURVEY.md §8: C1 640x480 TUM intrinsics, 4 levels, ~100 features; C2 752x480 EuRoC
intrinsics, 5 levels, 200 features; C3 640x480 fast motion; C5 1920x1080, 2000 features."""
import ctypes as C
import os
import subprocess
import numpy as np
from .abi import Camera, default_params, ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
TEXTURE_SEED = 0x5D71
TEXEL_M = 0.0025
PLANE = np.array([0.0, 0.0, 1.0, 0.0])  # n.X = d : z = 0


def build():
    src = os.path.join(_HERE, "synth", "synth.cc")
    out = os.path.join(_HERE, "synth", "libsdvl_synth.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-pthread", src,
                               "-o", out])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


_TEX = {}


def texture(size=4096, seed=TEXTURE_SEED, n_rects=5000):
    key = (size, seed, n_rects)
    if key not in _TEX:
        t = np.zeros((size, size), np.uint8)
        lib().synth_texture(ptr(t), size, C.c_uint32(seed), n_rects)
        _TEX[key] = t
    return _TEX[key]


CONFIGS = {
    # name: (w, h, fx, fy, u0, v0, levels, max_align_level, n_feat, num_features, traj_kind, fps)
    "C1": (640, 480, 517.3, 516.5, 318.6, 255.3, 4, 3, 100, 1000, 0, 30.0),
    "C2": (752, 480, 458.654, 457.296, 367.215, 248.375, 5, 4, 200, 1000, 0, 20.0),
    "C3": (640, 480, 517.3, 516.5, 318.6, 255.3, 5, 4, 200, 1000, 1, 30.0),
    "C5": (1920, 1080, 893.39, 898.33, 951.13, 555.13, 5, 4, 2000, 2000, 0, 20.0),
}


def config(name):
    w, h, fx, fy, u0, v0, levels, mal, nfeat, numf, kind, fps = CONFIGS[name]
    cam = Camera(w, h, fx, fy, u0, v0)
    p = default_params()
    p.pyramid_levels = levels
    p.max_align_level = mal
    p.max_matches = nfeat
    p.num_features = numf
    return dict(name=name, w=w, h=h, cam=cam, params=p, n_feat=nfeat, traj_kind=kind, fps=fps)


def trajectory(cfg, seed, n, height=2.0):
    poses = np.zeros((n, 7))
    lib().synth_trajectory(cfg["traj_kind"], C.c_uint32(seed), n, C.c_double(cfg["fps"]), C.c_double(height),
                           ptr(poses))
    return poses


def render(cfg, poses, threads=None, tex=None, out=None):
    tex = texture() if tex is None else tex
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
    n = poses.shape[0]
    if out is None:
        out = np.zeros((n, cfg["h"], cfg["w"]), np.uint8)
    cam = np.array([cfg["cam"].width, cfg["cam"].height, cfg["cam"].fx, cfg["cam"].fy, cfg["cam"].u0, cfg["cam"].v0])
    lib().synth_render_batch(ptr(tex), tex.shape[0], C.c_double(TEXEL_M), ptr(cam), ptr(poses), n, ptr(out),
                             cfg["w"], cfg["h"], threads or min(16, os.cpu_count() or 1))
    return out


def sequence(cfg_name, seed, n, threads=None):
    cfg = config(cfg_name)
    poses = trajectory(cfg, seed, n)
    return cfg, poses, render(cfg, poses, threads)


# ---- small SE3 helpers on {q0,q1,q2,q3,tx,ty,tz} world->camera poses (numpy, for tests/bench bookkeeping) ----
def quat_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def cam_center(T):
    return -quat_R(T[:4]).T @ T[4:7]


def ate(est, gt):
    """RMS distance between camera centres (no alignment: both are in the same world frame)."""
    d = np.array([cam_center(a) - cam_center(b) for a, b in zip(est, gt)])
    return float(np.sqrt((d ** 2).sum(1).mean()))
