"""ctypes binding of oracle/_ref/libsdvlref.so: the REFERENCE's own sources (compiled unmodified from /root/reference
against the stand-in headers of oracle/ref_shim) behind the same function names as oracle_py (TEST INFRASTRUCTURE:
import only from tests/, tests/golden generators and bench.py's cpu_baseline / --impl reference legs).

The library is built by `make -C oracle ref` where /root/reference exists; elsewhere the prebuilt .so (git-ignored,
shipped with the gpurun snapshot) is used.  `available()` says whether it can be loaded."""
import contextlib
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle_py

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
abi = oracle_py.abi
ptr = oracle_py.ptr
_LIBS = {}
_SUF = ""   # "" = default build, "_strict" = -ffp-contract=off (oracle/Makefile)


def path(suf=""):
    return os.path.join(_HERE, "_ref", "libsdvlref%s.so" % suf)


def build():
    """Compiles the reference's sources where they lie (never copied) -- only possible where the tree exists.
    Builds both variants (default flags and strict = no FMA contraction), and the oracle objects they link."""
    if not os.path.isdir(REFERENCE):
        return None
    subprocess.check_call(["make", "-s", "-C", _HERE, "-j8", "all", "ref", "REF=" + REFERENCE])
    subprocess.check_call(["make", "-s", "-C", _HERE, "-j8", "strict", "REF=" + REFERENCE])
    return path()


def available():
    return os.path.exists(path()) or os.path.isdir(REFERENCE)


def lib():
    if _SUF not in _LIBS:
        if os.path.isdir(REFERENCE):
            build()                      # make: no-op when up to date
        if not os.path.exists(path(_SUF)):
            raise RuntimeError("%s is missing and /root/reference is not here to build it" % path(_SUF))
        l = C.CDLL(path(_SUF))
        l.ref_pyramid.restype = C.c_int64
        l.ref_tracker_create.restype = C.c_void_p
        l.ref_tracker_run.restype = C.c_double
        l.ref_shi_tomasi.restype = C.c_double
        l.ref_interpolate8u.restype = C.c_float
        _LIBS[_SUF] = l
    return _LIBS[_SUF]


@contextlib.contextmanager
def strict():
    """Within the block every call goes to libsdvlref_strict.so (no FMA contraction)."""
    global _SUF
    old, _SUF = _SUF, "_strict"
    try:
        yield
    finally:
        _SUF = old


def read_config(filename=None):
    """Config::ReadParameters(filename) (or the constructor defaults when None) -> (params, camera)."""
    P, cam = abi.Params(), abi.Camera()
    rc = lib().ref_read_config(filename.encode() if filename else None, C.byref(P), C.byref(cam))
    assert rc == 0
    return P, cam


def pyramid(img, levels):
    h, w = img.shape
    img = np.ascontiguousarray(img)
    n = sum((w >> l) * (h >> l) for l in range(levels))
    out = np.empty(n, np.uint8)
    assert lib().ref_pyramid(ptr(img), w, h, levels, ptr(out)) == n
    res, off = [], 0
    for _ in range(levels):
        res.append(out[off:off + w * h].reshape(h, w))
        off += w * h
        w //= 2
        h //= 2
    return res


def detect(params, img, nfeatures):
    img = np.ascontiguousarray(img)
    h, w = img.shape
    cap = 16 * max(nfeatures, 64) + 4096
    xyl = np.zeros((cap, 3), np.int32)
    n = lib().ref_detect(C.byref(params), ptr(img), w, h, nfeatures, ptr(xyl), cap)
    assert n <= cap
    return xyl[:n].copy(), None


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
    rc = lib().ref_image_align(C.byref(params), C.byref(cam), ptr(ref_img), ptr(cur_img), w, h, ptr(feats), ptr(pos3),
                               feats.shape[0], ptr(T_ref), ptr(T_out), int(fast), C.byref(nt), C.byref(err),
                               ptr(trace), trace_cap, C.byref(tn))
    assert rc == 0, "trace replay diverged from ImageAlign::ComputePose (rc %d)" % rc
    return T_out, nt.value, err.value, trace[:min(tn.value, trace_cap)].copy()


def search_points(params, cam, cur_img, T_cur, ref_imgs, cands):
    cur_img = np.ascontiguousarray(cur_img)
    h, w = cur_img.shape
    refs = [np.ascontiguousarray(r) for r in ref_imgs]
    arr = (C.c_void_p * len(refs))(*[r.ctypes.data for r in refs])
    cands = np.ascontiguousarray(cands)
    out = np.zeros(cands.shape[0], abi.MATCH_DT)
    T_cur = np.ascontiguousarray(T_cur, np.float64)
    rc = lib().ref_search_points(C.byref(params), C.byref(cam), ptr(cur_img), w, h, ptr(T_cur), arr, len(refs),
                                 ptr(cands), cands.shape[0], ptr(out))
    assert rc == 0
    return out


def update_candidates(params, cam, cur_img, T_cur, ref_imgs, seeds, depth_mean, min_kf_id=-1000, map_scale=1.0,
                      scale_min_dist=0.25):
    """The reference's Map::UpdateCandidates on candidates built from `seeds` (abi.SEED_DT).  status out:
    SEED_CONVERGED, SEED_DELETE_OLD (= handed to DeletePoint) or -1 (still a candidate)."""
    cur_img = np.ascontiguousarray(cur_img)
    h, w = cur_img.shape
    refs = [np.ascontiguousarray(r) for r in ref_imgs]
    arr = (C.c_void_p * len(refs))(*[r.ctypes.data for r in refs])
    seeds = np.ascontiguousarray(seeds).copy()
    T_cur = np.ascontiguousarray(T_cur, np.float64)
    sp = abi.SeedParams(depth_mean, map_scale, scale_min_dist, min_kf_id, 0)
    rc = lib().ref_update_candidates(C.byref(params), C.byref(cam), ptr(cur_img), w, h, ptr(T_cur), arr, len(refs),
                                     ptr(seeds), seeds.shape[0], C.byref(sp))
    assert rc == 0, rc
    return seeds


def init_candidates(params, cam, new_img, T_new, old_img, T_old, depth_mean, map_scale=1.0, scale_min_dist=0.25,
                    min_feature_score=50):
    """The reference's Map::InitCandidates on keyframe `new_img` connected to keyframe `old_img`.  Returns a dict:
    ref_px, ref_level (init features), depth (InitCandidate), px, level (match in the old keyframe), listed."""
    new_img = np.ascontiguousarray(new_img)
    old_img = np.ascontiguousarray(old_img)
    h, w = new_img.shape
    cap = 4096
    ref_px, px2, depth = np.zeros((cap, 2)), np.zeros((cap, 2)), np.zeros(cap)
    ref_level, level2 = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    listed = C.c_int(0)
    T_new = np.ascontiguousarray(T_new, np.float64)
    T_old = np.ascontiguousarray(T_old, np.float64)
    n = lib().ref_init_candidates(C.byref(params), C.byref(cam), ptr(new_img), ptr(T_new), ptr(old_img), ptr(T_old), w, h,
                                  C.c_double(depth_mean), C.c_double(map_scale), C.c_double(scale_min_dist),
                                  min_feature_score, ptr(ref_px), ptr(ref_level), ptr(depth), ptr(px2), ptr(level2), cap,
                                  C.byref(listed))
    assert 0 <= n <= cap, n
    return dict(ref_px=ref_px[:n], ref_level=ref_level[:n], depth=depth[:n], px=px2[:n], level=level2[:n],
                listed=listed.value)


def add_connections_points(params, cam, cur_img, T_cur, kf_img, T_kf, cands):
    """The reference's Map::AddConnectionsPoints: new keyframe `cur_img`, one connected keyframe `kf_img` whose points
    `cands` (abi.CANDIDATE_DT) describe.  Returns abi.MATCH_DT records: FOUND + px + level for the points it linked."""
    cur_img = np.ascontiguousarray(cur_img)
    kf_img = np.ascontiguousarray(kf_img)
    h, w = cur_img.shape
    cands = np.ascontiguousarray(cands)
    out = np.zeros(cands.shape[0], abi.MATCH_DT)
    T_cur = np.ascontiguousarray(T_cur, np.float64)
    T_kf = np.ascontiguousarray(T_kf, np.float64)
    rc = lib().ref_add_connections_points(C.byref(params), C.byref(cam), ptr(cur_img), ptr(T_cur), ptr(kf_img), ptr(T_kf),
                                          w, h, ptr(cands), cands.shape[0], ptr(out))
    assert rc >= 0, rc
    return out


def align_patch(params, img, border_patch, px):
    img = np.ascontiguousarray(img, np.uint8)
    bp = np.ascontiguousarray(border_patch, np.uint8)
    p = np.array(px, np.float64)
    ok = lib().ref_align_patch(C.byref(params), ptr(img), img.shape[1], img.shape[0], ptr(bp), ptr(p))
    return bool(ok), p


def undistort(cam, dist, img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    d = np.ascontiguousarray(dist, np.float64)
    out = np.zeros_like(img)
    lib().ref_undistort(C.byref(cam), ptr(d), ptr(img), w, h, ptr(out))
    return out


def filter_corners(params, img, nfeatures, locked, min_feature_score=50):
    img = np.ascontiguousarray(img, np.uint8)
    locked = np.ascontiguousarray(locked, np.float64).reshape(-1, 2)
    out = np.zeros(8192, np.int32)
    n = lib().ref_filter_corners(C.byref(params), ptr(img), img.shape[1], img.shape[0], nfeatures, ptr(locked),
                                 locked.shape[0], min_feature_score, ptr(out), out.shape[0])
    return out[:n].copy()


def shi_tomasi(img, px, py):
    img = np.ascontiguousarray(img, np.uint8)
    return lib().ref_shi_tomasi(ptr(img), img.shape[1], img.shape[0], int(px), int(py))


def pose_refine(params, cam, obs, T, seed=1, mode=0):
    """FeatureAlign::SelectInliers with rand() = srand(seed) (mode 0) or OptimizePose (mode 1)."""
    obs = np.ascontiguousarray(obs).copy()
    T = np.array(T, np.float64)
    lib().ref_pose_refine(C.byref(params), C.byref(cam), ptr(obs), obs.shape[0], ptr(T), C.c_uint(seed), mode)
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
    rc = lib().ref_relocalize(C.byref(params), C.byref(cam), ptr(kf_img), ptr(cur_img), w, h, ptr(feats), ptr(pos3),
                             ptr(levels), feats.shape[0], ptr(T_kf), ptr(T_out), C.byref(err), ptr(out))
    assert rc == 0
    return T_out, err.value, out


class Tracker:
    def __init__(self, params, cam, plane, max_points, kf_every):
        plane = np.ascontiguousarray(plane, np.float64)
        self.lib = lib()   # the variant this tracker lives in (default or strict)
        self.h = self.lib.ref_tracker_create(C.byref(params), C.byref(cam), ptr(plane), max_points, kf_every)

    def run(self, imgs, gt_poses):
        imgs = np.ascontiguousarray(imgs)
        n, h, w = imgs.shape
        gt = np.ascontiguousarray(gt_poses, np.float64)
        est = np.zeros((n, 7))
        stats = np.zeros((n, 8), np.int32)
        sec = self.lib.ref_tracker_run(C.c_void_p(self.h), ptr(imgs), n, w, h, ptr(gt), ptr(est), ptr(stats))
        return est, stats, sec

    def close(self):
        if self.h:
            self.lib.ref_tracker_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()
