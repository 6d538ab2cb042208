"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs)."""
import contextlib
import ctypes as C
import importlib.util
import os
import subprocess
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def _pkg():
    name = "slam_sdvl_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_ROOT, "slam-sdvl_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(_ROOT, "slam-sdvl_b200")])
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


abi = importlib.import_module(_pkg().__name__ + ".abi")
ptr = abi.ptr
_LIBS = {}
_SUF = ""   # "" = default build, "_strict" = -ffp-contract=off (oracle/Makefile)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "-j8"])
    return os.path.join(_HERE, "liboracle.so")


def build_strict():
    subprocess.check_call(["make", "-s", "-C", _HERE, "-j8", "SUF=_strict",
                           "CXXFLAGS=-O3 -march=x86-64-v3 -std=c++14 -fPIC -ffp-contract=off", "all"])
    return os.path.join(_HERE, "liboracle_strict.so")


def lib():
    if _SUF not in _LIBS:
        path = os.path.join(_HERE, "liboracle%s.so" % _SUF)
        if not os.path.exists(path):
            build_strict() if _SUF else build()
        l = C.CDLL(path)
        l.orc_pyramid.restype = C.c_int64
        l.orc_tracker_create.restype = C.c_void_p
        l.orc_tracker_run.restype = C.c_double
        _LIBS[_SUF] = l
    return _LIBS[_SUF]


@contextlib.contextmanager
def strict():
    """Within the block every call goes to liboracle_strict.so (no FMA contraction): bit-comparable with
    ref_py.strict()."""
    global _SUF
    old, _SUF = _SUF, "_strict"
    try:
        yield
    finally:
        _SUF = old


def pyramid(img, levels):
    h, w = img.shape
    img = np.ascontiguousarray(img)
    n = sum((w >> l) * (h >> l) for l in range(levels))
    out = np.empty(n, np.uint8)
    assert lib().orc_pyramid(ptr(img), w, h, levels, ptr(out)) == n
    res, off = [], 0
    for _ in range(levels):
        res.append(out[off:off + w * h].reshape(h, w))
        off += w * h
        w //= 2
        h //= 2
    return res


def fast_roi(img, x0, y0, cols, rows, thr=10):
    img = np.ascontiguousarray(img)
    cap = cols * rows
    xy = np.zeros((cap, 2), np.int32)
    sc = np.zeros(cap, np.int32)
    base = img.ctypes.data + y0 * img.shape[1] + x0
    n = lib().orc_fast_roi(C.c_void_p(base), img.shape[1], cols, rows, thr, ptr(xy), ptr(sc), cap)
    return xy[:n].copy(), sc[:n].copy()


def retain_best(xyr, keep):
    a = np.ascontiguousarray(xyr, np.float32).copy()
    n = lib().orc_retain_best(ptr(a), a.shape[0], keep)
    return a[:n]


def detect(params, img, nfeatures):
    img = np.ascontiguousarray(img)
    h, w = img.shape
    cap = 16 * max(nfeatures, 64) + 4096
    xyl = np.zeros((cap, 3), np.int32)
    sc = np.zeros(cap, np.int32)
    n = lib().orc_detect(C.byref(params), ptr(img), w, h, nfeatures, ptr(xyl), ptr(sc), cap)
    assert n <= cap
    return xyl[:n].copy(), sc[:n].copy()


def image_align(params, cam, ref_img, cur_img, feats, pos3, T_ref, T_cur, fast=False, trace_cap=256):
    ref_img = np.ascontiguousarray(ref_img)
    cur_img = np.ascontiguousarray(cur_img)
    h, w = ref_img.shape
    feats = np.ascontiguousarray(feats)
    pos3 = np.ascontiguousarray(pos3, np.float64)
    T_ref = np.ascontiguousarray(T_ref, np.float64)
    T_out = np.array(T_cur, np.float64)
    trace = np.zeros(trace_cap, abi.GN_ITER_DT)
    nt, tn, err = C.c_int(0), C.c_int(0), C.c_double(0)
    lib().orc_image_align(C.byref(params), C.byref(cam), ptr(ref_img), ptr(cur_img), w, h, ptr(feats), ptr(pos3),
                          feats.shape[0], ptr(T_ref), ptr(T_out), int(fast), C.byref(nt), C.byref(err), ptr(trace),
                          trace_cap, C.byref(tn))
    return T_out, nt.value, err.value, trace[:min(tn.value, trace_cap)].copy()


def search_points(params, cam, cur_img, T_cur, ref_imgs, cands):
    cur_img = np.ascontiguousarray(cur_img)
    h, w = cur_img.shape
    refs = [np.ascontiguousarray(r) for r in ref_imgs]
    arr = (C.c_void_p * len(refs))(*[r.ctypes.data for r in refs])
    cands = np.ascontiguousarray(cands)
    out = np.zeros(cands.shape[0], abi.MATCH_DT)
    T_cur = np.ascontiguousarray(T_cur, np.float64)
    rc = lib().orc_search_points(C.byref(params), C.byref(cam), ptr(cur_img), w, h, ptr(T_cur), arr, len(refs),
                                 ptr(cands), cands.shape[0], ptr(out))
    assert rc == 0
    return out


def update_candidates(params, cam, cur_img, T_cur, ref_imgs, seeds, depth_mean, min_kf_id=-1000, map_scale=1.0,
                      scale_min_dist=0.25, mode=0):
    """Oracle Map::UpdateCandidates loop body. seeds: abi.SEED_DT array whose ref_frame fields index ref_imgs."""
    cur_img = np.ascontiguousarray(cur_img)
    h, w = cur_img.shape
    refs = [np.ascontiguousarray(r) for r in ref_imgs]
    arr = (C.c_void_p * len(refs))(*[r.ctypes.data for r in refs])
    seeds = np.ascontiguousarray(seeds).copy()
    assert seeds.dtype == abi.SEED_DT
    T_cur = np.ascontiguousarray(T_cur, np.float64)
    sp = abi.SeedParams(depth_mean, map_scale, scale_min_dist, min_kf_id, mode)
    rc = lib().orc_update_candidates(C.byref(params), C.byref(cam), ptr(cur_img), w, h, ptr(T_cur), arr, len(refs),
                                     ptr(seeds), seeds.shape[0], C.byref(sp))
    assert rc == 0
    return seeds


def undistort(cam, dist, img):
    """Oracle Camera::UndistortImage (cv::undistort with D = (k1, k2, p1, p2, k3))."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    d = np.ascontiguousarray(dist, np.float64)
    assert d.shape == (5,)
    out = np.zeros_like(img)
    lib().orc_undistort(C.byref(cam), ptr(d), ptr(img), w, h, ptr(out))
    return out


def filter_corners(params, img, nfeatures, locked, min_feature_score=50):
    """Oracle Frame::FilterCorners: indices into the frame's corner list, one per free cell."""
    img = np.ascontiguousarray(img, np.uint8)
    locked = np.ascontiguousarray(locked, np.float64).reshape(-1, 2)
    out = np.zeros(8192, np.int32)
    n = lib().orc_filter_corners(C.byref(params), ptr(img), img.shape[1], img.shape[0], nfeatures, ptr(locked),
                                 locked.shape[0], min_feature_score, ptr(out), out.shape[0])
    return out[:n].copy()


def shi_tomasi(img, px, py):
    img = np.ascontiguousarray(img, np.uint8)
    lib().orc_shi_tomasi.restype = C.c_double
    return lib().orc_shi_tomasi(ptr(img), img.shape[1], img.shape[0], int(px), int(py))


def pose_refine(params, cam, obs, T, rng=None, mode=0):
    """Oracle FeatureAlign::SelectInliers (mode 0, rng = abi.Rand advanced in place) or OptimizePose (mode 1).
    Returns (obs with flags, pose)."""
    obs = np.ascontiguousarray(obs).copy()
    assert obs.dtype == abi.POSE_OBS_DT
    T = np.array(T, np.float64)
    lib().orc_pose_refine(C.byref(params), C.byref(cam), ptr(obs), obs.shape[0], ptr(T),
                          C.byref(rng) if rng is not None else None, mode)
    return obs, T


def relocalize(params, cam, kf_img, cur_img, feats, pos3, levels, T_kf):
    """SDVL::Relocalize's body for one keyframe: (pose after the fast ImageAlign, GetError(), [matches (-1 = rejected by
    the error gate), attempts, features gained by the current frame, sum of the points' scores])."""
    kf_img = np.ascontiguousarray(kf_img)
    cur_img = np.ascontiguousarray(cur_img)
    h, w = kf_img.shape
    feats = np.ascontiguousarray(feats)
    pos3 = np.ascontiguousarray(pos3, np.float64)
    levels = np.ascontiguousarray(levels, np.int32)
    T_kf = np.ascontiguousarray(T_kf, np.float64)
    T_out, err, out = np.zeros(7), C.c_double(0), np.zeros(4, np.int32)
    rc = lib().orc_relocalize(C.byref(params), C.byref(cam), ptr(kf_img), ptr(cur_img), w, h, ptr(feats), ptr(pos3),
                             ptr(levels), feats.shape[0], ptr(T_kf), ptr(T_out), C.byref(err), ptr(out))
    assert rc == 0
    return T_out, err.value, out


class Tracker:
    def __init__(self, params, cam, plane, max_points, kf_every):
        plane = np.ascontiguousarray(plane, np.float64)
        self.lib = lib()   # the variant this tracker lives in (default or strict)
        self.h = self.lib.orc_tracker_create(C.byref(params), C.byref(cam), ptr(plane), max_points, kf_every)

    def run(self, imgs, gt_poses):
        imgs = np.ascontiguousarray(imgs)
        n, h, w = imgs.shape
        gt = np.ascontiguousarray(gt_poses, np.float64)
        est = np.zeros((n, 7))
        stats = np.zeros((n, 8), np.int32)
        sec = self.lib.orc_tracker_run(C.c_void_p(self.h), ptr(imgs), n, w, h, ptr(gt), ptr(est), ptr(stats))
        return est, stats, sec

    def close(self):
        if self.h:
            self.lib.orc_tracker_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()
