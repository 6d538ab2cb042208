"""ctypes mirrors of the POD structs in include/sdvl_b200.h (shared by the product binding, the synthetic world
and the oracle's test binding)."""
import ctypes as C
import numpy as np


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "pyramid_levels", "cell_size", "max_matches", "max_align_level", "min_align_level", "max_img_align_its",
        "align_patch_size", "patch_size", "max_align_its", "search_size", "max_fast_levels", "fast_threshold",
        "num_features", "max_failed", "max_optim_pose_its", "max_ransac_points", "max_ransac_its", "min_matches")] + \
        [("inlier_error_threshold", C.c_double)]


class Camera(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("width", "height", "fx", "fy", "u0", "v0")]


class AlignFeat(C.Structure):
    _fields_ = [("px", C.c_double * 2), ("v", C.c_double * 3), ("depth", C.c_double), ("valid", C.c_int32),
                ("pad_", C.c_int32)]


class GnIter(C.Structure):
    _fields_ = [("level", C.c_int32), ("iter", C.c_int32), ("n_meas", C.c_int32), ("flags", C.c_int32),
                ("T_in", C.c_double * 7), ("H", C.c_double * 36), ("b", C.c_double * 6), ("x", C.c_double * 6),
                ("chi2", C.c_double)]


class GnForced(C.Structure):
    _fields_ = [("T", C.c_void_p), ("iters", C.c_void_p), ("n_total", C.c_int32), ("pad_", C.c_int32)]


class Candidate(C.Structure):
    _fields_ = [("ref_frame", C.c_void_p), ("ref_T", C.c_double * 7), ("ref_px", C.c_double * 2),
                ("ref_v", C.c_double * 3), ("idepth", C.c_double), ("idepth_std", C.c_double),
                ("px", C.c_double * 2), ("pos", C.c_double * 3), ("ref_level", C.c_int32), ("flags", C.c_int32)]


class Match(C.Structure):
    _fields_ = [("px", C.c_double * 2), ("proj", C.c_double * 2), ("level", C.c_int32), ("status", C.c_int32),
                ("zmssd", C.c_int32), ("n_in_range", C.c_int32)]


class PoseObs(C.Structure):
    _fields_ = [("v", C.c_double * 3), ("pos", C.c_double * 3), ("level", C.c_int32), ("flags", C.c_int32)]


class Rand(C.Structure):
    _fields_ = [("r", C.c_uint32 * 34), ("n", C.c_int32), ("pad_", C.c_int32)]


OBS_INLIER, OBS_OUTLIER = 1, 2
POSE_OBS_DT = np.dtype([("v", "f8", 3), ("pos", "f8", 3), ("level", "i4"), ("flags", "i4")])
assert POSE_OBS_DT.itemsize == C.sizeof(PoseObs)

CAND_FIXED, CAND_PROJECT = 1, 2
MATCH_UNSEEN, MATCH_NOT_FOUND, MATCH_FOUND = 0, 1, 2

# numpy views of the same layouts (for bulk construction)
ALIGN_FEAT_DT = np.dtype([("px", "f8", 2), ("v", "f8", 3), ("depth", "f8"), ("valid", "i4"), ("pad_", "i4")])
GN_ITER_DT = np.dtype([("level", "i4"), ("iter", "i4"), ("n_meas", "i4"), ("flags", "i4"), ("T_in", "f8", 7),
                       ("H", "f8", 36), ("b", "f8", 6), ("x", "f8", 6), ("chi2", "f8")])
CANDIDATE_DT = np.dtype([("ref_frame", "u8"), ("ref_T", "f8", 7), ("ref_px", "f8", 2), ("ref_v", "f8", 3),
                         ("idepth", "f8"), ("idepth_std", "f8"), ("px", "f8", 2), ("pos", "f8", 3),
                         ("ref_level", "i4"), ("flags", "i4")])
MATCH_DT = np.dtype([("px", "f8", 2), ("proj", "f8", 2), ("level", "i4"), ("status", "i4"), ("zmssd", "i4"),
                     ("n_in_range", "i4")])
# sdvlb_seed / sdvlb_seed_params (Map::UpdateCandidates)
SEED_DT = np.dtype([("ref_frame", "u8"), ("ref_T", "f8", 7), ("ref_px", "f8", 2), ("ref_v", "f8", 3), ("rho", "f8"),
                    ("sigma2", "f8"), ("a", "f8"), ("b", "f8"), ("z_range", "f8"), ("cos_alpha", "f8"),
                    ("last_distance", "f8"), ("p3d", "f8", 3), ("depth", "f8"), ("px", "f8", 2), ("ref_level", "i4"),
                    ("n_failed", "i4"), ("last_kf_id", "i4"), ("status", "i4"), ("level", "i4"), ("pad_", "i4")])
(SEED_NOT_VISIBLE, SEED_DELETE_OLD, SEED_SHORT_BASELINE, SEED_NOT_FOUND, SEED_DELETE_FAILED, SEED_NO_DEPTH,
 SEED_NO_PARALLAX, SEED_TOO_CLOSE, SEED_UPDATED, SEED_CONVERGED) = range(10)
SEEDS_UPDATE, SEEDS_INIT = 0, 1


class SeedParams(C.Structure):
    _fields_ = [("depth_mean", C.c_double), ("map_scale", C.c_double), ("scale_min_dist", C.c_double),
                ("min_kf_id", C.c_int32), ("mode", C.c_int32)]


# resident sequences (sdvlb_seq_*)
SEQ_KF_CAP, SEQ_DEPTH = 64, 4
SEQ_POINT_DT = np.dtype([("pos", "f8", 3), ("ref_px", "f8", 2), ("cur_px", "f8", 2), ("idepth", "f8"), ("idepth_std", "f8"),
                         ("user_id", "i8"), ("ref_level", "i4"), ("cur_level", "i4"), ("flags", "i4"),
                         ("n_successful", "i4"), ("n_failed", "i4"), ("pad_", "i4")])
SEQ_FEAT_DT = np.dtype([("px", "f8", 2), ("user_id", "i8"), ("level", "i4"), ("flags", "i4")])
SEQ_TRACKED, SEQ_HELD, SEQ_IDLE = 0, 1, 2
TRACKING_GOOD, TRACKING_INSUFFICIENT, TRACKING_BAD = 0, 1, 2
FEAT_HAS_POINT = 1


class SeqPolicy(C.Structure):
    _fields_ = [("keyframe_rule", C.c_int32), ("min_keyframe_its", C.c_int32), ("lost_ratio", C.c_double),
                ("tracking_quality", C.c_int32), ("pad_", C.c_int32)]


class SeqResult(C.Structure):
    _fields_ = [("pose", C.c_double * 7), ("n_tracked", C.c_int32), ("matches", C.c_int32), ("attempts", C.c_int32),
                ("inliers", C.c_int32), ("outliers", C.c_int32), ("n_points", C.c_int32), ("gn_iters", C.c_int32),
                ("n_feats", C.c_int32), ("feats", C.c_void_p), ("status", C.c_int32), ("quality", C.c_int32),
                ("need_keyframe", C.c_int32), ("lost_frames", C.c_int32), ("kf_live", C.c_int32 * SEQ_KF_CAP),
                ("phase_cycles", C.c_int32 * 8), ("align_cycles", C.c_int32 * 4)]


assert SEED_DT.itemsize == 232   # sizeof(sdvlb_seed)
assert SEQ_POINT_DT.itemsize == 104 and SEQ_FEAT_DT.itemsize == 32
assert ALIGN_FEAT_DT.itemsize == C.sizeof(AlignFeat)
assert GN_ITER_DT.itemsize == C.sizeof(GnIter)
assert CANDIDATE_DT.itemsize == C.sizeof(Candidate)
assert MATCH_DT.itemsize == C.sizeof(Match)


def default_params():
    """config.cc:55-85 of the reference."""
    p = Params()
    (p.pyramid_levels, p.cell_size, p.max_matches, p.max_align_level, p.min_align_level, p.max_img_align_its,
     p.align_patch_size, p.patch_size, p.max_align_its, p.search_size, p.max_fast_levels, p.fast_threshold,
     p.num_features, p.max_failed, p.max_optim_pose_its, p.max_ransac_points, p.max_ransac_its, p.min_matches) = \
        (5, 32, 150, 4, 2, 30, 4, 8, 10, 6, 3, 10, 1000, 15, 10, 5, 100, 20)
    p.inlier_error_threshold = 2.0
    return p


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
