"""ctypes binding of libsdvl_b200.so (the C-ABI in include/sdvl_b200.h).

There is no CPU fallback: importing works anywhere, but `load()` raises if the CUDA library has not been built, and
`Context()` raises if no CUDA device is usable."""
import ctypes as C
import os
import subprocess
import numpy as np
from . import abi
from .abi import ptr

# A tracker drives many short kernels from several stream pairs at once; with the default of 8 hardware work queues
# unrelated streams share a queue and serialise behind each other.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdvl_b200.so")
_LIB = None

K_NAMES = ("pyramid", "fast", "select", "align", "search", "orb", "pose")


class SdvlbError(RuntimeError):
    pass


class TrackJob(C.Structure):
    _fields_ = [("image", C.c_void_p), ("image_on_device", C.c_int32), ("want_corners", C.c_int32),
                ("nfeatures", C.c_int32), ("n_feats", C.c_int32), ("n_cands", C.c_int32), ("n_tracked", C.c_int32),
                ("gn_iters", C.c_int32), ("pad_", C.c_int32), ("ref", C.c_void_p), ("cur", C.c_void_p), ("feats", C.c_void_p), ("cands", C.c_void_p),
                ("matches", C.c_void_p), ("T_ref", C.c_double * 7), ("T_cur", C.c_double * 7), ("error", C.c_double),
                ("cand_desc", C.c_void_p)]


def build(verbose=False):
    """Compiles the sm_100a kernels + C-ABI in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SdvlbError(f"{LIB_PATH} is missing: build it with slam_sdvl_b200.binding.build() "
                             "(there is no CPU fallback for this path)")
        L = C.CDLL(LIB_PATH)
        L.sdvlb_last_error.restype = C.c_char_p
        L.sdvlb_ctx_stream.restype = C.c_void_p
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise SdvlbError(f"sdvlb status {rc}: {load().sdvlb_last_error().decode()}")


EXPORTS = [
    "sdvlb_params_default", "sdvlb_ctx_create", "sdvlb_ctx_destroy", "sdvlb_ctx_sync", "sdvlb_ctx_stream",
    "sdvlb_last_error", "sdvlb_host_alloc", "sdvlb_host_free", "sdvlb_dev_alloc", "sdvlb_dev_free",
    "sdvlb_dev_upload", "sdvlb_ctx_counters", "sdvlb_timing_enable", "sdvlb_timing_read", "sdvlb_frame_create", "sdvlb_frame_detect",
    "sdvlb_frame_level", "sdvlb_frame_corners", "sdvlb_frame_destroy", "sdvlb_image_align", "sdvlb_search_points",
    "sdvlb_track_batch", "sdvlb_frames_submit", "sdvlb_frames_wait", "sdvlb_track_submit", "sdvlb_track_poll",
    "sdvlb_track_collect", "sdvlb_ctx_reserve_frames",
    "sdvlb_rand_seed", "sdvlb_rand_next", "sdvlb_rand_shuffle", "sdvlb_select_inliers", "sdvlb_optimize_pose",
    "sdvlb_seq_create", "sdvlb_seq_destroy", "sdvlb_seq_reset", "sdvlb_seq_add_points", "sdvlb_seq_track_submit",
    "sdvlb_seq_track_poll", "sdvlb_seq_track_collect", "sdvlb_seq_track_inflight", "sdvlb_seq_set_policy",
    "sdvlb_seq_release", "sdvlb_frame_filter_corners", "sdvlb_update_candidates",
    "sdvlb_ctx_set_distortion", "sdvlb_undistort",
    "sdvlb_ctx_set_orb", "sdvlb_frame_orb_descriptors", "sdvlb_search_points_orb", "sdvlb_frame_descriptors",
]


class Frame:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def level(self, l):
        data, w, h = C.c_void_p(), C.c_int(), C.c_int()
        _check(load().sdvlb_frame_level(C.c_void_p(self.h), l, C.byref(data), C.byref(w), C.byref(h)))
        buf = (C.c_uint8 * (w.value * h.value)).from_address(data.value)
        return np.frombuffer(buf, np.uint8).reshape(h.value, w.value).copy()

    def corners(self):
        xyls, n = C.c_void_p(), C.c_int()
        _check(load().sdvlb_frame_corners(C.c_void_p(self.h), C.byref(xyls), C.byref(n)))
        if n.value == 0:
            return np.zeros((0, 3), np.int32), np.zeros(0, np.int32)
        a = np.frombuffer((C.c_int32 * (4 * n.value)).from_address(xyls.value), np.int32).reshape(-1, 4)
        return a[:, :3].copy(), a[:, 3].copy()

    def filter_corners(self, locked_px, min_feature_score=50):
        """Frame::FilterCorners: indices into corners(), one per free cell with a good enough Shi-Tomasi score."""
        locked = np.ascontiguousarray(locked_px, np.float64).reshape(-1, 2)
        out = np.zeros(8192, np.int32)
        n = C.c_int()
        _check(load().sdvlb_frame_filter_corners(C.c_void_p(self.ctx.h), C.c_void_p(self.h), ptr(locked), locked.shape[0],
                                                 min_feature_score, ptr(out), out.shape[0], C.byref(n)))
        return out[:n.value].copy()

    def orb_descriptors(self, xyl):
        """ORBDetector::GetDescriptor / GetOrientation at positions xyl = n x (x, y, level): (n x 32 bytes, degrees)."""
        xyl = np.ascontiguousarray(xyl, np.int32).reshape(-1, 3)
        desc = np.zeros((xyl.shape[0], 32), np.uint8)
        ang = np.zeros(xyl.shape[0], np.float32)
        _check(load().sdvlb_frame_orb_descriptors(C.c_void_p(self.ctx.h), C.c_void_p(self.h), ptr(xyl), xyl.shape[0],
                                                  ptr(desc), ptr(ang)))
        return desc, ang

    def descriptors(self):
        """Frame::GetDescriptors(): (n_corners x 32) bytes, corners() order (ORB mode)."""
        n = C.c_int()
        _check(load().sdvlb_frame_descriptors(C.c_void_p(self.ctx.h), C.c_void_p(self.h), None, 0, C.byref(n)))
        d = np.zeros((n.value, 32), np.uint8)
        if n.value:
            _check(load().sdvlb_frame_descriptors(C.c_void_p(self.ctx.h), C.c_void_p(self.h), ptr(d), n.value, C.byref(n)))
        return d

    def detect(self, nfeatures):
        _check(load().sdvlb_frame_detect(C.c_void_p(self.ctx.h), C.c_void_p(self.h), nfeatures))

    def destroy(self):
        if self.h:
            load().sdvlb_frame_destroy(C.c_void_p(self.ctx.h), C.c_void_p(self.h))
            self.h = None


class Sequence:
    """A resident sequence (sdvlb_seq_*): the tracked state lives on the device."""

    def __init__(self, ctx, max_feats=0):
        self.ctx = ctx
        h = C.c_void_p()
        _check(load().sdvlb_seq_create(C.c_void_p(ctx.h), int(max_feats), C.byref(h)))
        self.h = h.value

    def reset(self, frame, T):
        T = np.ascontiguousarray(T, np.float64)
        _check(load().sdvlb_seq_reset(C.c_void_p(self.ctx.h), C.c_void_p(self.h), C.c_void_p(frame.h), ptr(T)))

    def add_points(self, kf_frame, T_kf, pts):
        """pts: array of abi.SEQ_POINT_DT.  Returns the keyframe slot."""
        from .abi import SEQ_POINT_DT
        pts = np.ascontiguousarray(pts, SEQ_POINT_DT)
        T_kf = np.ascontiguousarray(T_kf, np.float64)
        slot = C.c_int(-1)
        _check(load().sdvlb_seq_add_points(C.c_void_p(self.ctx.h), C.c_void_p(self.h), C.c_void_p(kf_frame.h), ptr(T_kf),
                                           ptr(pts) if len(pts) else None, len(pts), C.byref(slot)))
        return slot.value

    def set_policy(self, keyframe_rule=0, min_keyframe_its=0, lost_ratio=0.7, tracking_quality=0):
        from .abi import SeqPolicy
        pol = SeqPolicy(int(keyframe_rule), int(min_keyframe_its), float(lost_ratio), int(tracking_quality), 0)
        _check(load().sdvlb_seq_set_policy(C.c_void_p(self.ctx.h), C.c_void_p(self.h), C.byref(pol)))

    def release(self):
        _check(load().sdvlb_seq_release(C.c_void_p(self.ctx.h), C.c_void_p(self.h)))

    def destroy(self):
        if self.h:
            load().sdvlb_seq_destroy(C.c_void_p(self.ctx.h), C.c_void_p(self.h))
            self.h = None


class Context:
    def __init__(self, params, cam, device=0):
        self.params, self.cam = params, cam
        h = C.c_void_p()
        _check(load().sdvlb_ctx_create(device, C.byref(params), C.byref(cam), C.byref(h)))
        self.h = h.value
        self.w, self.hh = int(cam.width), int(cam.height)

    def close(self):
        if self.h:
            load().sdvlb_ctx_destroy(C.c_void_p(self.h))
            self.h = None

    def stream(self):
        return load().sdvlb_ctx_stream(C.c_void_p(self.h))

    # ---- resident sequences
    def seq_submit(self, seqs, frames):
        n = len(seqs)
        sa = (C.c_void_p * n)(*[s.h for s in seqs])
        fa = (C.c_void_p * n)(*[f.h for f in frames])
        self._seq_n = getattr(self, "_seq_n", []) + [n]
        rc = load().sdvlb_seq_track_submit(C.c_void_p(self.h), sa, fa, n)
        if rc:
            self._seq_n.pop()
        _check(rc)

    def seq_inflight(self):
        return load().sdvlb_seq_track_inflight(C.c_void_p(self.h))

    def seq_collect(self):
        """Results of the oldest submission in flight: a list of dicts (the feature list copied out of pinned memory)."""
        from .abi import SeqResult, SEQ_FEAT_DT
        n = self._seq_n.pop(0)
        res = (SeqResult * n)()
        _check(load().sdvlb_seq_track_collect(C.c_void_p(self.h), res))
        out = []
        for r in res:
            d = {k: getattr(r, k) for k in ("n_tracked", "matches", "attempts", "inliers", "outliers", "n_points", "gn_iters",
                                            "n_feats", "status", "quality", "need_keyframe", "lost_frames")}
            d["pose"] = np.array(r.pose)
            d["kf_live"] = np.array(r.kf_live)
            if r.status == 0 and r.n_feats > 0:
                buf = (C.c_uint8 * (r.n_feats * SEQ_FEAT_DT.itemsize)).from_address(r.feats)
                d["feats"] = np.frombuffer(buf, SEQ_FEAT_DT).copy()
            else:
                d["feats"] = np.zeros(0, SEQ_FEAT_DT)
            out.append(d)
        return out

    def sync(self):
        _check(load().sdvlb_ctx_sync(C.c_void_p(self.h)))

    def timing(self, on):
        load().sdvlb_timing_enable(C.c_void_p(self.h), int(on))

    def timing_read(self, reset=True):
        ms = (C.c_double * len(K_NAMES))()
        n = (C.c_int64 * len(K_NAMES))()
        _check(load().sdvlb_timing_read(C.c_void_p(self.h), ms, n, int(reset)))
        return {k: (ms[i], n[i]) for i, k in enumerate(K_NAMES)}

    def frame(self, img, corners=True, nfeatures=None):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = C.c_void_p()
        nf = self.params.num_features if nfeatures is None else nfeatures
        _check(load().sdvlb_frame_create(C.c_void_p(self.h), ptr(img), w, h, w, int(corners), nf, C.byref(out)))
        return Frame(self, out.value)

    def image_align(self, ref, cur, feats, T_ref, T_cur, fast=False, trace_cap=256, forced=None):
        feats = np.ascontiguousarray(feats)
        assert feats.dtype == abi.ALIGN_FEAT_DT
        T_ref = np.ascontiguousarray(T_ref, np.float64)
        T_out = np.array(T_cur, np.float64)
        trace = np.zeros(max(trace_cap, 1), abi.GN_ITER_DT)
        nt, tn, err = C.c_int(0), C.c_int(0), C.c_double(0)
        fptr = None
        keep = None
        if forced is not None:
            fT = np.ascontiguousarray(forced[0], np.float64)
            fi = np.ascontiguousarray(forced[1], np.int32)
            keep = (fT, fi)
            fs = abi.GnForced(fT.ctypes.data, fi.ctypes.data, fT.shape[0], 0)
            fptr = C.byref(fs)
        _check(load().sdvlb_image_align(C.c_void_p(self.h), C.c_void_p(ref.h), C.c_void_p(cur.h), ptr(feats),
                                        feats.shape[0], ptr(T_ref), ptr(T_out), int(fast), C.byref(nt),
                                        C.byref(err), ptr(trace) if trace_cap else None, trace_cap, C.byref(tn), fptr))
        return T_out, nt.value, err.value, trace[:min(tn.value, trace_cap)].copy()

    def search_points(self, cur, cands, T_cur):
        cands = np.ascontiguousarray(cands)
        assert cands.dtype == abi.CANDIDATE_DT
        out = np.zeros(cands.shape[0], abi.MATCH_DT)
        Tp = None
        if T_cur is not None:
            T_cur = np.ascontiguousarray(T_cur, np.float64)
            Tp = ptr(T_cur)
        _check(load().sdvlb_search_points(C.c_void_p(self.h), C.c_void_p(cur.h), ptr(cands), cands.shape[0], Tp,
                                          ptr(out)))
        return out

    def set_orb(self, on=True):
        """Config::UseORB(): ORB border margins for FAST / FilterCorners of the frames built from now on."""
        _check(load().sdvlb_ctx_set_orb(C.c_void_p(self.h), int(on)))

    def search_points_orb(self, cur, cands, T_cur, cand_desc):
        """Matcher::SearchPoint with Config::UseORB(): cand_desc = n x 32 bytes (the candidates' feature descriptors)."""
        cands = np.ascontiguousarray(cands)
        assert cands.dtype == abi.CANDIDATE_DT
        cand_desc = np.ascontiguousarray(cand_desc, np.uint8).reshape(-1, 32)
        assert cand_desc.shape[0] == cands.shape[0]
        out = np.zeros(cands.shape[0], abi.MATCH_DT)
        T_cur = np.ascontiguousarray(T_cur, np.float64)
        _check(load().sdvlb_search_points_orb(C.c_void_p(self.h), C.c_void_p(cur.h), ptr(cands), cands.shape[0], ptr(T_cur),
                                              ptr(cand_desc), ptr(out)))
        return out

    def set_distortion(self, d):
        """Camera::SetDistortions: d = (k1, k2, p1, p2, k3); zeros switch the undistortion of incoming images off."""
        d = np.ascontiguousarray(d, np.float64)
        assert d.shape == (5,)
        _check(load().sdvlb_ctx_set_distortion(C.c_void_p(self.h), ptr(d)))

    def undistort(self, img):
        """Camera::UndistortImage on the device (host image in, host image out)."""
        img = np.ascontiguousarray(img, np.uint8)
        assert img.shape == (self.hh, self.w)
        out = np.zeros_like(img)
        _check(load().sdvlb_undistort(C.c_void_p(self.h), ptr(img), ptr(out)))
        return out

    def update_candidates(self, cur, T_cur, seeds, depth_mean, min_kf_id=-1000, map_scale=1.0, scale_min_dist=0.25,
                          mode=0):
        """Map::UpdateCandidates loop body on the device. seeds: abi.SEED_DT array whose ref_frame fields hold Frame
        handles; returns the updated copy."""
        seeds = np.ascontiguousarray(seeds).copy()
        assert seeds.dtype == abi.SEED_DT
        T = np.ascontiguousarray(T_cur, np.float64)
        sp = abi.SeedParams(depth_mean, map_scale, scale_min_dist, min_kf_id, mode)
        _check(load().sdvlb_update_candidates(C.c_void_p(self.h), C.c_void_p(cur.h), ptr(T), ptr(seeds), seeds.shape[0],
                                              C.byref(sp)))
        return seeds

    def select_inliers(self, obs, T_frame, rng):
        """FeatureAlign::SelectInliers on the device. obs: POSE_OBS_DT array (flags overwritten); rng: abi.Rand (advanced)."""
        obs = np.ascontiguousarray(obs)
        assert obs.dtype == abi.POSE_OBS_DT
        T = np.ascontiguousarray(T_frame, np.float64)
        _check(load().sdvlb_select_inliers(C.c_void_p(self.h), ptr(obs), obs.shape[0], ptr(T), C.byref(rng)))
        return obs

    def optimize_pose(self, obs, T_frame):
        """FeatureAlign::OptimizePose(frame) on the device (without RemoveOutliers). Returns (obs, refined pose)."""
        obs = np.ascontiguousarray(obs)
        assert obs.dtype == abi.POSE_OBS_DT
        T = np.array(T_frame, np.float64)
        _check(load().sdvlb_optimize_pose(C.c_void_p(self.h), ptr(obs), obs.shape[0], ptr(T)))
        return obs, T

    def track_batch(self, jobs, mirror=1):
        """jobs: ctypes array of TrackJob."""
        _check(load().sdvlb_track_batch(C.c_void_p(self.h), jobs, len(jobs), self.w, self.hh, mirror))


# ---------------------------------------------------------------------------------------------- host mirror library
HOST_LIB_PATH = os.path.join(_HERE, "libsdvl_b200_host.so")
_HLIB = None

HOST_EXPORTS = ["sdvlh_last_error", "sdvlh_config_set", "sdvlh_tracker_create", "sdvlh_tracker_destroy",
                "sdvlh_tracker_step", "sdvlh_tracker_timing_read", "sdvlh_tracker_counters", "sdvlh_tracker_phases",
                "sdvlh_tracker_ctx", "sdvlh_tracker_groups", "sdvlh_tracker_threads", "sdvlh_tracker_run",
                "sdvlh_tracker_create2", "sdvlh_tracker_post_cycles", "sdvlh_tracker_set_prefetch", "sdvlh_tracker_set_depth", "sdvlh_tracker_slowest_cycles",
                "sdvlh_map_update_candidates", "sdvlh_map_init_candidates", "sdvlh_camera_undistort",
                "sdvlh_config_set_orb"]


def build_host(verbose=False):
    cmd = ["make", "-C", os.path.join(_HERE, "host"), "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return HOST_LIB_PATH


def load_host():
    global _HLIB
    if _HLIB is None:
        load()
        if not os.path.exists(HOST_LIB_PATH):
            raise SdvlbError(f"{HOST_LIB_PATH} is missing: build it with slam_sdvl_b200.binding.build_host()")
        H = C.CDLL(HOST_LIB_PATH)
        H.sdvlh_last_error.restype = C.c_char_p
        H.sdvlh_tracker_create.restype = C.c_void_p
        H.sdvlh_tracker_create2.restype = C.c_void_p
        H.sdvlh_tracker_ctx.restype = C.c_void_p
        _HLIB = H
    return _HLIB


def host_map_update_candidates(params, cam, ref_img, ref_T, cur_imgs, cur_poses, seeds, depth_mean, min_kf_id=-1000):
    """sdvl::Map::UpdateCandidates of the C++ host mirror over a list of frames (test hook). Returns (seeds, n_left)."""
    lib = load_host()
    lib.sdvlh_config_set(C.byref(params), C.byref(cam))
    ref_img = np.ascontiguousarray(ref_img, np.uint8)
    h, w = ref_img.shape
    curs = [np.ascontiguousarray(i, np.uint8) for i in cur_imgs]
    arr = (C.c_void_p * len(curs))(*[c.ctypes.data for c in curs])
    poses = np.ascontiguousarray(cur_poses, np.float64).reshape(-1, 7)
    ref_T = np.ascontiguousarray(ref_T, np.float64)
    seeds = np.ascontiguousarray(seeds).copy()
    assert seeds.dtype == abi.SEED_DT
    left = C.c_int(0)
    lib.sdvlh_map_update_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    rc = lib.sdvlh_map_update_candidates(ptr(ref_img), ptr(ref_T), arr, ptr(poses), len(curs), w, h, ptr(seeds),
                                         seeds.shape[0], depth_mean, min_kf_id, C.byref(left))
    if rc:
        raise RuntimeError(lib.sdvlh_last_error().decode())
    return seeds, left.value


def host_camera_undistort(params, cam, dist, img):
    """sdvl::Camera::SetDistortions + UndistortImage of the C++ host mirror (test hook)."""
    lib = load_host()
    lib.sdvlh_config_set(C.byref(params), C.byref(cam))
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    d = np.ascontiguousarray(dist, np.float64)
    out = np.zeros_like(img)
    if lib.sdvlh_camera_undistort(ptr(d), ptr(img), w, h, ptr(out)):
        raise RuntimeError(lib.sdvlh_last_error().decode())
    return out


def host_map_init_candidates(params, cam, img_new, T_new, img_old, T_old, upd_imgs, upd_T, img_last, T_last, depth_mean,
                             cap=4096):
    """sdvl::Map::InitCandidates + UpdateCandidates + AddConnectionsPoints of the C++ host mirror (test hook)."""
    lib = load_host()
    lib.sdvlh_config_set(C.byref(params), C.byref(cam))
    img_new = np.ascontiguousarray(img_new, np.uint8)
    img_old = np.ascontiguousarray(img_old, np.uint8)
    img_last = np.ascontiguousarray(img_last, np.uint8)
    h, w = img_new.shape
    upd = [np.ascontiguousarray(i, np.uint8) for i in upd_imgs]
    arr = (C.c_void_p * max(1, len(upd)))(*[u.ctypes.data for u in upd])
    upd_T = np.ascontiguousarray(upd_T, np.float64).reshape(-1, 7)
    T_new, T_old, T_last = (np.ascontiguousarray(t, np.float64) for t in (T_new, T_old, T_last))
    out = np.zeros(5, np.int32)
    px = np.zeros((cap, 2))
    rho = np.zeros(cap)
    fixed = np.zeros(cap, np.int32)
    lib.sdvlh_map_init_candidates.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                               C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                               C.c_int]
    rc = lib.sdvlh_map_init_candidates(ptr(img_new), ptr(T_new), ptr(img_old), ptr(T_old), arr, ptr(upd_T), len(upd),
                                       ptr(img_last), ptr(T_last), w, h, depth_mean, ptr(out), ptr(px), ptr(rho),
                                       ptr(fixed), cap)
    if rc:
        raise RuntimeError(lib.sdvlh_last_error().decode())
    n = int(out[0])
    return dict(created=n, listed=int(out[1]), left=int(out[2]), fixed=int(out[3]), linked=int(out[4]),
                px=px[:n], rho=rho[:n], is_fixed=fixed[:n])


class HostTracker:
    """Multi-sequence tracker over the C++ host mirror (Frame / ImageAlign / FeatureAlign classes).

    step(images) takes one image per sequence; `classic=True` drives the reference's per-call class API instead of
    the batched submission."""

    def __init__(self, params, cam, plane, max_points, kf_every, n_seq, n_groups=1, device=0, timing=False,
                 n_threads=0, resident=False, use_orb=False):
        """resident=True keeps the sequences on the device (sdvlb_seq_*): no host marshalling / match replay.
        use_orb=True: Config::UseORB() for this tracker's contexts (the flag is process-wide, as in the reference)."""
        H = load_host()
        H.sdvlh_config_set(C.byref(params), C.byref(cam))
        H.sdvlh_config_set_orb(int(use_orb))
        plane = np.ascontiguousarray(plane, np.float64)
        self.h = H.sdvlh_tracker_create2(ptr(plane), max_points, kf_every, n_seq, n_groups, n_threads, device,
                                         int(timing), int(resident))
        if not self.h:
            raise SdvlbError("sdvlh_tracker_create failed: " + H.sdvlh_last_error().decode())
        self.n_seq = n_seq
        self.w, self.hh = int(cam.width), int(cam.height)
        self._ptrs = (C.c_void_p * n_seq)()
        self.est = np.zeros((n_seq, 7))
        self.stats = np.zeros((n_seq, 8), np.int32)

    def step_ptrs(self, ptrs, gt, on_device=False, classic=False):
        """ptrs: sequence of n_seq raw addresses of w*h u8 images (host, or device when on_device)."""
        for i, p in enumerate(ptrs):
            self._ptrs[i] = p
        gt = np.ascontiguousarray(gt, np.float64)
        rc = load_host().sdvlh_tracker_step(C.c_void_p(self.h), self._ptrs, int(on_device), int(classic), ptr(gt),
                                            ptr(self.est), ptr(self.stats))
        if rc:
            raise SdvlbError("sdvlh_tracker_step failed: " + load_host().sdvlh_last_error().decode())
        return self.est, self.stats

    def step(self, images, gt, classic=False):
        """images: array (n_seq, h, w) u8 (host)."""
        images = np.ascontiguousarray(images, np.uint8)
        stride = images.shape[1] * images.shape[2]
        return self.step_ptrs([images.ctypes.data + i * stride for i in range(self.n_seq)], gt, False, classic)

    def run_ptrs(self, ptrs, gt, on_device=False):
        """Pipelined run. ptrs: (n_seq, n_steps) array of raw image addresses; gt: (n_seq, n_steps, 7).
        Returns est (n_seq, n_steps, 7) and stats (n_seq, n_steps, 8)."""
        ptrs = np.ascontiguousarray(ptrs, np.uint64)
        n_steps = ptrs.shape[1]
        assert ptrs.shape[0] == self.n_seq
        gt = np.ascontiguousarray(gt, np.float64)
        est = np.zeros((self.n_seq, n_steps, 7))
        stats = np.zeros((self.n_seq, n_steps, 8), np.int32)
        rc = load_host().sdvlh_tracker_run(C.c_void_p(self.h), ptr(ptrs), int(on_device), n_steps, ptr(gt), ptr(est),
                                           ptr(stats))
        if rc:
            raise SdvlbError("sdvlh_tracker_run failed: " + load_host().sdvlh_last_error().decode())
        return est, stats

    def run(self, images, gt):
        """images: (n_seq, n_steps, h, w) u8 host array."""
        images = np.ascontiguousarray(images, np.uint8)
        s0, s1 = images.strides[0], images.strides[1]
        ptrs = images.ctypes.data + np.arange(images.shape[0], dtype=np.uint64)[:, None] * s0 + \
            np.arange(images.shape[1], dtype=np.uint64)[None, :] * s1
        return self.run_ptrs(ptrs, gt)

    def threads(self):
        return load_host().sdvlh_tracker_threads(C.c_void_p(self.h))

    def timing_read(self, reset=True):
        ms = (C.c_double * len(K_NAMES))()
        n = (C.c_int64 * len(K_NAMES))()
        load_host().sdvlh_tracker_timing_read(C.c_void_p(self.h), ms, n, int(reset))
        return {k: (ms[i], n[i]) for i, k in enumerate(K_NAMES)}

    def counters(self, reset=True):
        """(kernel launches, host->device bytes, device->host bytes) summed over the tracker's contexts."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        load_host().sdvlh_tracker_counters(C.c_void_p(self.h), C.byref(a), C.byref(b), C.byref(c), int(reset))
        return a.value, b.value, c.value

    def phases(self, reset=True):
        """Thread-seconds in host marshalling / GPU submission+wait / host replay, summed over groups."""
        a = (C.c_double * 8)()
        load_host().sdvlh_tracker_phases(C.c_void_p(self.h), a, int(reset))
        return {"marshal_s": a[0], "gpu_submit_wait_s": a[1], "replay_s": a[2], "replay_apply_matches_s": a[3],
                "replay_ransac_s": a[4], "replay_finish_frame_s": a[5], "gpu_wait_only_s": a[6], "idle_poll_s": a[7]}

    def set_prefetch(self, depth):
        load_host().sdvlh_tracker_set_prefetch(C.c_void_p(self.h), int(depth))

    def set_depth(self, depth):
        """Tracking submissions kept in flight per group (resident sequences), 1..SDVLB_SEQ_DEPTH."""
        load_host().sdvlh_tracker_set_depth(C.c_void_p(self.h), int(depth))

    def post_cycles(self, reset=True):
        """Average SM cycles per tracked frame in the phases of the device-side FeatureAlign kernel."""
        a = (C.c_double * 13)()
        load_host().sdvlh_tracker_post_cycles(C.c_void_p(self.h), a, int(reset))
        n = max(1.0, a[8])
        return {"align_precompute": a[9] / n, "align_residuals": a[10] / n, "align_reduce": a[11] / n,
                "align_solve": a[12] / n, "cell_ranks": a[0] / n, "select_points": a[1] / n, "ransac_hypotheses": a[2] / n,
                "startup": a[3] / n, "ransac_replay": a[4] / n, "optimize_pose": a[5] / n, "finish": a[6] / n,
                "shuffle_warp_done_since_entry": a[7] / n}

    def slowest_cycles(self, reset=True):
        """The same phases for the slowest sequence of each submission (a step lasts as long as its slowest CTA)."""
        a = (C.c_double * 13)()
        load_host().sdvlh_tracker_slowest_cycles(C.c_void_p(self.h), a, int(reset))
        n = max(1.0, a[8])
        return {"align_precompute": a[9] / n, "align_residuals": a[10] / n, "align_reduce": a[11] / n,
                "align_solve": a[12] / n, "cell_ranks": a[0] / n, "select_points": a[1] / n, "ransac_hypotheses": a[2] / n,
                "startup": a[3] / n, "ransac_replay": a[4] / n, "optimize_pose": a[5] / n, "finish": a[6] / n,
                "shuffle_warp_done_since_entry": a[7] / n}

    def ctx_handle(self):
        return load_host().sdvlh_tracker_ctx(C.c_void_p(self.h))

    def groups(self):
        return load_host().sdvlh_tracker_groups(C.c_void_p(self.h))

    def close(self):
        if self.h:
            load_host().sdvlh_tracker_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
