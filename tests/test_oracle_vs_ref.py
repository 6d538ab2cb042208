"""Pins the CPU oracle against the REFERENCE'S OWN CODE (CPU only, no GPU).

oracle/_ref/libsdvlref.so is the reference's hot-path sources (image_align.cc, matcher.cc, feature_align.cc, frame.cc,
camera.cc, feature.cc, point.cc, map.cc, config.cc, extra/{se3,utils,fast_detector,orb_detector}.cc) compiled
UNMODIFIED from /root/reference against the stand-in Eigen / OpenCV headers of oracle/ref_shim (oracle/Makefile,
target `ref`), behind the flat entry points of oracle/ref_harness.cc.  Every test pushes the same seeded inputs
through that library and through the oracle's restatement.

Two builds of each side are compared:
  * strict (-ffp-contract=off on both): the same arithmetic in the same order must give the SAME BITS -- traces of
    H, b, x, chi2 per Gauss-Newton iteration, matched pixel positions, refined poses and whole trajectories are
    required to be bit-identical;
  * default flags (what bench.py times): the compiler may contract a*b+c into FMAs differently in the two code bases,
    so results agree to rounding (1e-9 relative on H/b, 1e-4 px (float LK), 1e-4 m on a 45-frame trajectory, identical counts).

Skipped where neither the prebuilt library nor /root/reference exists; tests/golden/ref_golden.npz (made from the same
library by tests/golden/make_ref_golden.py) keeps the oracle pinned there, see test_oracle_matches_reference_golden."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import ref_py as R

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference not present")
STAT_COLS = [0, 1, 2, 3, 4, 5, 7]   # gn_iters (6) is not observable on the reference's ImageAlign


def _both(O, strict):
    """Context managers selecting the same build variant on both sides."""
    import contextlib
    st = contextlib.ExitStack()
    if strict:
        st.enter_context(O.strict())
        st.enter_context(R.strict())
    return st


def _scene(O, sw, scenes, name, seed, n, max_points):
    cfg, poses, imgs = sw.sequence(name, seed, n)
    xyl, _ = O.detect(cfg["params"], imgs[0], cfg["params"].num_features)
    pts = scenes.seed_points(cfg, xyl, poses[0], max_points=max_points)
    return cfg, poses, imgs, pts


# ------------------------------------------------------------------------------------------------ Config
@needs_ref
def test_reference_config_defaults_and_cfg_files(O, sw, abi):
    """Config::Config() defaults (config.cc:33-86) == orc_params_default / sdvlb_params_default, and the intrinsics the
    synthetic configs use are the ones the reference's own cfg files carry (parsed by Config::ReadParameters)."""
    P0, _ = R.read_config()
    Pd = abi.Params()
    O.lib().orc_params_default(C.byref(Pd))
    for f, _ in Pd._fields_:
        assert getattr(Pd, f) == getattr(P0, f), f
    if not os.path.isdir(R.REFERENCE):
        pytest.skip("cfg files live in /root/reference")
    for cfg_name, fname in (("C2", "config/config_euroc.cfg"), ("C1", "config_example.cfg")):
        P, cam = R.read_config(os.path.join(R.REFERENCE, fname))
        mine = sw.config(cfg_name)["cam"]
        assert (cam.width, cam.height) == (mine.width, mine.height)
        assert (cam.fx, cam.fy, cam.u0, cam.v0) == (mine.fx, mine.fy, mine.u0, mine.v0)
    R.read_config(os.path.join(R.REFERENCE, "config.cfg"))
    # restore the defaults for whoever runs next in this process
    R.lib().ref_read_config(None, C.byref(P0), None)


# ------------------------------------------------------------------------------------------------ Frame
@needs_ref
@pytest.mark.parametrize("name,nfeat", [("C1", 1000), ("C2", 1000), ("C2", 2000), ("C3", 1000), ("C5", 2000)])
def test_pyramid_and_corners_identical(O, sw, name, nfeat):
    """Frame::CreatePyramid + Frame::CreateCorners -> FastDetector::DetectPyramid / SelectPixels (cells, ROI margins,
    water-filling quotas, per-cell and global retainBest, level budgets): same bytes, same corner list, same order."""
    cfg, poses, imgs = sw.sequence(name, 2, 2)
    P = cfg["params"]
    for img in imgs:
        a, b = O.pyramid(img, P.pyramid_levels), R.pyramid(img, P.pyramid_levels)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
        xo, _ = O.detect(P, img, nfeat)
        xr, _ = R.detect(P, img, nfeat)
        assert len(xo) > 300 and np.array_equal(xo, xr)


@needs_ref
def test_filter_corners_and_shi_tomasi_identical(O, sw):
    cfg, poses, imgs = sw.sequence("C2", 4, 1)
    P = cfg["params"]
    xyl, _ = O.detect(P, imgs[0], 1000)
    rng = np.random.default_rng(1)
    locked = np.stack([rng.uniform(0, 752, 60), rng.uniform(0, 480, 60)], 1)
    for strict in (True, False):
        with _both(O, strict):
            for lk, score in ((np.zeros((0, 2)), 50), (locked, 50), (locked, 400)):
                io, ir = O.filter_corners(P, imgs[0], 1000, lk, score), R.filter_corners(P, imgs[0], 1000, lk, score)
                assert len(io) > 20 and np.array_equal(io, ir)
            for (x, y, l) in xyl[xyl[:, 2] == 0][:200]:
                a, b = O.shi_tomasi(imgs[0], x, y), R.shi_tomasi(imgs[0], x, y)
                # float arithmetic: the same bits without FMA contraction, a few float ulps (amplified by the cancellation in tr - root) with it
                assert a == b if strict else abs(a - b) <= 2e-6 * abs(b), (x, y, a, b)
            assert O.shi_tomasi(imgs[0], 2, 2) == R.shi_tomasi(imgs[0], 2, 2) == 0.0


@needs_ref
def test_undistort_through_camera_identical(O, sw):
    """Camera::SetDistortions + UndistortImage: the reference's own mapping of Camera.d1..d5 onto cv::undistort's
    (k1, k2, p1, p2, k3) and its K matrix (camera.cc:38-67) give the oracle's image."""
    cfg, poses, imgs = sw.sequence("C1", 1, 1)
    d = np.array([0.2624, -0.9531, -0.0054, 0.0026, 1.1633])   # config_example.cfg
    assert np.array_equal(O.undistort(cfg["cam"], d, imgs[0]), R.undistort(cfg["cam"], d, imgs[0]))
    assert np.array_equal(R.undistort(cfg["cam"], np.zeros(5), imgs[0]), imgs[0])   # distortion off: plain copy


# ------------------------------------------------------------------------------------------------ SE3
@needs_ref
def test_se3_primitives_bit_identical(O):
    rng = np.random.default_rng(5)
    with _both(O, True):
        lo, lr = O.lib(), R.lib()
        for k in range(200):
            u = rng.normal(0, [0.3, 0.3, 0.3, 1.0, 1.0, 1.0][k % 6], 6) * (1e-9 if k % 17 == 0 else 1.0)
            v = rng.normal(0, 0.5, 6)
            p = rng.normal(0, 2, 3)
            A, B = np.zeros((2, 7)), np.zeros((2, 7))
            lo.orc_se3_exp(O.ptr(u), O.ptr(A[0])); lr.ref_se3_exp(O.ptr(u), O.ptr(A[1]))
            lo.orc_se3_exp(O.ptr(v), O.ptr(B[0])); lr.ref_se3_exp(O.ptr(v), O.ptr(B[1]))
            assert np.array_equal(A[0], A[1]) and np.array_equal(B[0], B[1])
            M, I, L, Q = np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 6)), np.zeros((2, 3))
            lo.orc_se3_mul(O.ptr(A[0]), O.ptr(B[0]), O.ptr(M[0])); lr.ref_se3_mul(O.ptr(A[0]), O.ptr(B[0]), O.ptr(M[1]))
            lo.orc_se3_inv(O.ptr(M[0]), O.ptr(I[0])); lr.ref_se3_inv(O.ptr(M[0]), O.ptr(I[1]))
            lo.orc_se3_log(O.ptr(M[0]), O.ptr(L[0])); lr.ref_se3_log(O.ptr(M[0]), O.ptr(L[1]))
            lo.orc_se3_apply(O.ptr(M[0]), O.ptr(p), O.ptr(Q[0])); lr.ref_se3_apply(O.ptr(M[0]), O.ptr(p), O.ptr(Q[1]))
            for a in (M, I, L, Q):
                assert np.array_equal(a[0], a[1])


# ------------------------------------------------------------------------------------------------ ImageAlign
ALIGN_CASES = [("C2", 0, 1, 0, False), ("C2", 0, 4, 7, False), ("C1", 3, 2, 0, False), ("C3", 5, 1, 5, False),
               ("C3", 5, 3, 0, True), ("C2", 0, 5, 0, True)]


@needs_ref
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name,seed,gap,invalid_every,fast", ALIGN_CASES)
def test_image_align_trace_vs_reference(O, sw, scenes, name, seed, gap, invalid_every, fast, strict):
    """ImageAlign::ComputePose: every Gauss-Newton iteration of every level (T_in, H, b, x, chi2, n_meas, accept /
    rollback flags), the final pose, the return value and GetError()."""
    with _both(O, strict):
        cfg, poses, imgs, pts = _scene(O, sw, scenes, name, seed, gap + 1, cfg_feats(name))
        P, cam = cfg["params"], cfg["cam"]
        feats = scenes.align_feats(pts, poses[0], invalid_every=invalid_every)
        a = O.image_align(P, cam, imgs[0], imgs[gap], feats, pts["pos"], poses[0], poses[0], fast=fast)
        b = R.image_align(P, cam, imgs[0], imgs[gap], feats, pts["pos"], poses[0], poses[0], fast=fast)
    (To, nto, eo, tro), (Tr, ntr, er, trr) = a, b
    assert len(tro) == len(trr) >= 3 and nto == ntr
    for k in ("level", "iter", "n_meas", "flags"):
        assert np.array_equal(tro[k], trr[k]), k
    if strict:
        assert np.array_equal(To, Tr) and eo == er
        for k in ("T_in", "H", "b", "x", "chi2"):
            assert np.array_equal(tro[k], trr[k]), k
    else:
        assert np.abs(To - Tr).max() < 1e-9 and abs(eo - er) <= 1e-6 * abs(er)
        for k, tol in (("T_in", 1e-9), ("H", 1e-9), ("b", 1e-7), ("chi2", 1e-9)):
            assert np.abs(tro[k] - trr[k]).max() <= tol * np.abs(trr[k]).max(), k
        assert np.abs(tro["x"] - trr["x"]).max() <= 1e-6 * np.abs(trr["x"]).max() + 1e-12


def cfg_feats(name):
    return {"C1": 100, "C2": 200, "C3": 200, "C5": 2000}[name]


@needs_ref
def test_image_align_iteration_budgets_and_empty(O, sw, scenes):
    """max_img_align_its = 1, 2, 3 (pose after k iterations per level) and the no-feature early return."""
    with _both(O, True):
        cfg, poses, imgs, pts = _scene(O, sw, scenes, "C2", 7, 3, 200)
        cam = cfg["cam"]
        feats = scenes.align_feats(pts, poses[0])
        for its in (1, 2, 3):
            P = type(cfg["params"])()
            C.memmove(C.byref(P), C.byref(cfg["params"]), C.sizeof(P))
            P.max_img_align_its = its
            To, nto, eo, tro = O.image_align(P, cam, imgs[0], imgs[2], feats, pts["pos"], poses[0], poses[0])
            Tr, ntr, er, trr = R.image_align(P, cam, imgs[0], imgs[2], feats, pts["pos"], poses[0], poses[0])
            assert len(tro) == len(trr) and np.array_equal(To, Tr) and (nto, eo) == (ntr, er)
        To, nto, _, _ = O.image_align(cfg["params"], cam, imgs[0], imgs[1], feats[:0], pts["pos"][:0], poses[0], poses[1])
        Tr, ntr, _, _ = R.image_align(cfg["params"], cam, imgs[0], imgs[1], feats[:0], pts["pos"][:0], poses[0], poses[1])
        assert nto == ntr == 0 and np.array_equal(To, Tr) and np.array_equal(To, poses[1])


@needs_ref
def test_image_align_degenerate_inputs_vs_reference(O, sw, scenes):
    """Edge cases of ImageAlign: a textureless current image (H singular: Eigen's LDLT with its pseudo-inverse of D),
    a single feature (rank-deficient H), every feature without a point, a prior so far off that nothing projects into
    the image (n_meas = 0 sets the sticky stop_ flag, image_align.cc:98-99), and features on the image border."""
    with _both(O, True):
        cfg, poses, imgs, pts = _scene(O, sw, scenes, "C2", 3, 3, 200)
        P, cam = cfg["params"], cfg["cam"]
        feats = scenes.align_feats(pts, poses[0])
        flat = np.full_like(imgs[0], 127)
        far = poses[0].copy()
        far[4:] += [5.0, -3.0, 0.5]
        none = feats.copy()
        none["valid"] = 0
        edge = feats.copy()
        edge["px"][::3] = [2.0, 3.0]
        cases = [("flat current image", imgs[0], flat, feats, poses[0]),
                 ("flat reference image", flat, imgs[1], feats, poses[0]),
                 ("single feature", imgs[0], imgs[1], feats[:1], poses[0]),
                 ("two features", imgs[0], imgs[1], feats[:2], poses[0]),
                 ("no feature has a point", imgs[0], imgs[1], none, poses[0]),
                 ("prior far away", imgs[0], imgs[2], feats, far),
                 ("features on the border", imgs[0], imgs[1], edge, poses[0])]
        for what, ref_img, cur_img, f, prior in cases:
            pos = pts["pos"][:len(f)]
            To, nto, eo, tro = O.image_align(P, cam, ref_img, cur_img, f, pos, poses[0], prior)
            Tr, ntr, er, trr = R.image_align(P, cam, ref_img, cur_img, f, pos, poses[0], prior)
            assert len(tro) == len(trr) and nto == ntr, what
            for k in ("level", "iter", "n_meas", "flags"):
                assert np.array_equal(tro[k], trr[k]), (what, k)
            for k in ("T_in", "H", "b", "x", "chi2"):
                assert np.array_equal(tro[k], trr[k], equal_nan=True), (what, k)
            assert np.array_equal(To, Tr, equal_nan=True) and (eo == er or (np.isnan(eo) and np.isnan(er))), what


# ------------------------------------------------------------------------------------------------ Matcher
@needs_ref
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name,seed", [("C2", 0), ("C3", 5), ("C1", 3)])
def test_search_point_vs_reference(O, sw, scenes, abi, name, seed, strict):
    """Matcher::SearchPoint for fixed points (circle around the projection) and for depth-filter candidates (capsule
    around the epipolar segment, wide and narrow depth ranges): found / not found / unseen, level, refined position."""
    with _both(O, strict):
        cfg, poses, imgs = sw.sequence(name, seed, 7)
        P, cam = cfg["params"], cfg["cam"]
        xyl, _ = O.detect(P, imgs[0], P.num_features)
        pts = scenes.seed_points(cfg, xyl, poses[0], max_points=400, one_per_cell=False)
        rng = np.random.default_rng(seed)
        n_found = 0
        for gap in (1, 3, 6):
            for fixed, std_frac in ((True, 0.05), (False, 0.05), (False, 1.5)):
                c = scenes.candidates(pts, poses[0], 0, fixed=fixed, project=True, std_frac=std_frac)
                if not fixed:
                    c["idepth"] *= 1.0 + rng.uniform(-0.1, 0.1, len(c))
                mo = O.search_points(P, cam, imgs[gap], poses[gap], [imgs[0]], c)
                mr = R.search_points(P, cam, imgs[gap], poses[gap], [imgs[0]], c)
                assert np.array_equal(mo["status"], mr["status"]) and np.array_equal(mo["level"], mr["level"])
                found = mo["status"] == abi.MATCH_FOUND
                n_found += int(found.sum())
                if strict:
                    assert np.array_equal(mo["px"], mr["px"]) and np.array_equal(mo["proj"], mr["proj"])
                else:
                    assert np.abs(mo["px"] - mr["px"]).max() < 1e-4 and np.abs(mo["proj"] - mr["proj"]).max() < 1e-9
        assert n_found > 500


@needs_ref
def test_search_point_degenerate_candidates_vs_reference(O, sw, scenes, abi):
    """Edge cases of Matcher::SearchPoint where a restatement could drift from the code: zero depth uncertainty on an
    epipolar candidate (zero-length segment: Eigen's guarded normalize, u = 0/0), a depth range reaching behind the
    camera, the 1e-8 inverse-depth clamp, reference patches that fail the margin test at their level, points projected
    outside the image, and a camera looking away from the points."""
    cfg, poses, imgs = sw.sequence("C2", 0, 4)
    P, cam = cfg["params"], cfg["cam"]
    with _both(O, True):
        xyl, _ = O.detect(P, imgs[0], P.num_features)
        pts = scenes.seed_points(cfg, xyl, poses[0], max_points=240, one_per_cell=False, margin=0)
        n = len(pts["px"])
        c = scenes.candidates(pts, poses[0], 0, fixed=False, project=False, std_frac=0.05)
        k = np.arange(n) % 6
        c["idepth_std"][k == 0] = 0.0                        # pxa == pxb
        c["idepth_std"][k == 1] = 10.0 * c["idepth"][k == 1]  # idepth - 2 std < 0: clamp to 1e-8, far point
        c["idepth"][k == 2] *= 40.0                          # 2.5 % of the true depth: projects far from the feature
        c["flags"][k == 3] |= abi.CAND_FIXED                 # fixed, searched around a prediction that is off
        c["px"] = pts["px"] + np.where((k == 3)[:, None], 4.0, 0.0)
        c["flags"][k == 4] |= abi.CAND_PROJECT
        c["pos"][k == 4] *= np.array([1.0, 1.0, -1.0])       # mirrored through the plane: other side of the camera? no, below it
        c["pos"][k == 5] += np.array([30.0, 0.0, 0.0])       # far outside the image
        c["flags"][k == 5] |= abi.CAND_PROJECT
        res = []
        for T in (poses[3], poses[0]):
            res.append((O.search_points(P, cam, imgs[3], T, [imgs[0]], c), R.search_points(P, cam, imgs[3], T, [imgs[0]], c)))
        # a camera turned away from the scene: every projection is behind it
        Tb = poses[3].copy()
        Tb[:4] = [0.0, 1.0, 0.0, 0.0]   # half a turn about x on top of nothing: looks along -z
        res.append((O.search_points(P, cam, imgs[3], Tb, [imgs[0]], c), R.search_points(P, cam, imgs[3], Tb, [imgs[0]], c)))
    seen = np.zeros(3, int)
    for mo, mr in res:
        assert np.array_equal(mo["status"], mr["status"])
        assert np.array_equal(mo["level"], mr["level"]) and np.array_equal(mo["px"], mr["px"])
        seen += np.bincount(mo["status"], minlength=3)[:3]
    assert (seen > 0).all(), seen      # unseen, not found and found all occurred


@needs_ref
def test_align_patch_identical(O, abi, seq_c2):
    """Matcher::AlignPatch alone (private in the reference, reached through the harness): convergence flag and position,
    including starts that run out of the image (the `break` of matcher.cc:402-403)."""
    cfg, poses, imgs = seq_c2
    P, img = cfg["params"], imgs[0]
    xyl, _ = O.detect(P, img, P.num_features)
    rng = np.random.default_rng(2)
    n_ok = 0
    with _both(O, True):
        for (x, y, l) in xyl[xyl[:, 2] == 0][:120]:
            if x < 8 or y < 8 or x > img.shape[1] - 8 or y > img.shape[0] - 8:
                continue
            bp = np.ascontiguousarray(img[y - 5:y + 5, x - 5:x + 5])
            start = np.array([x, y], np.float64) + rng.uniform(-2.5, 2.5, 2)
            if rng.uniform() < 0.1:
                start = np.array([2.0, 3.0])          # too close to the border
            po = start.copy()
            ro = O.lib().orc_align_patch(C.byref(P), O.ptr(img), img.shape[1], img.shape[0], O.ptr(bp), O.ptr(po))
            rr, pr = R.align_patch(P, img, bp, start)
            assert bool(ro) == rr and np.array_equal(po, pr)
            n_ok += rr
    assert n_ok > 30


# ------------------------------------------------------------------------------------------------ FeatureAlign
def _pose_obs(sw, abi, n, n_bad, seed):
    cfg = sw.config("C2")
    cam = cfg["cam"]
    rng = np.random.default_rng(seed)
    T_true = sw.trajectory(cfg, 4 + seed, 3)[2]
    Rm = sw.quat_R(T_true[:4])
    obs = np.zeros(n, abi.POSE_OBS_DT)
    bad = set(rng.choice(n, n_bad, replace=False).tolist())
    for i in range(n):
        u, v, depth = rng.uniform(20, cam.width - 20), rng.uniform(20, cam.height - 20), rng.uniform(1, 4)
        ray = np.array([(u - cam.u0) / cam.fx, (v - cam.v0) / cam.fy, 1.0])
        obs["pos"][i] = Rm.T @ (ray * depth - T_true[4:])
        u += rng.normal(0, 0.3)
        v += rng.normal(0, 0.3)
        if i in bad:
            u += rng.choice([-1, 1]) * rng.uniform(6, 40)
        b = np.array([(u - cam.u0) / cam.fx, (v - cam.v0) / cam.fy, 1.0])
        obs["v"][i] = b / np.linalg.norm(b)
        obs["level"][i] = rng.integers(0, 3)
    return cfg, obs, T_true


@needs_ref
@pytest.mark.parametrize("n,n_bad,seed", [(80, 9, 1), (200, 40, 2), (30, 12, 3), (6, 1, 4), (4, 0, 5), (1, 0, 6), (2, 1, 7),
                                            (3, 0, 8), (12, 11, 9)])
def test_select_inliers_and_optimize_pose_vs_reference(O, sw, abi, n, n_bad, seed):
    """FeatureAlign::SelectInliers (RANSAC on glibc rand(): same draws, same hypotheses, same supporters) and
    OptimizePose (MAD-scaled Tukey IRLS, RescueOutliers, RemoveOutliers) on noisy observations with gross outliers,
    down to fewer observations than RANSAC points."""
    cfg, obs, T_true = _pose_obs(sw, abi, n, n_bad, seed)
    P, cam = cfg["params"], cfg["cam"]
    du = np.array([0.004, -0.003, 0.002, 0.001, -0.002, 0.001]) * seed
    dT, T0 = np.zeros(7), np.zeros(7)
    O.lib().orc_se3_exp(O.ptr(du), O.ptr(dT))
    O.lib().orc_se3_mul(O.ptr(dT), O.ptr(np.ascontiguousarray(T_true)), O.ptr(T0))
    with _both(O, True):
        r = abi.Rand()
        O.lib().orc_rand_state(seed, C.byref(r))
        oo, _ = O.pose_refine(P, cam, obs, T0, r, mode=0)
        orr, _ = R.pose_refine(P, cam, obs, T0, seed=seed, mode=0)
        assert np.array_equal(oo["flags"], orr["flags"])
        o2, To = O.pose_refine(P, cam, oo, T0, None, mode=1)
        r2, Tr = R.pose_refine(P, cam, orr, T0, mode=1)
        assert np.array_equal(o2["flags"], r2["flags"]) and np.array_equal(To, Tr)
    if n >= 30:
        assert (o2["flags"] == abi.OBS_INLIER).sum() >= n - n_bad - 3


# ------------------------------------------------------------------------------------------------ ORB descriptor mode
@needs_ref
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name,seed", [("C2", 0), ("C1", 3), ("C5", 1)])
def test_orb_mode_vs_reference(O, sw, scenes, abi, name, seed, strict):
    """Config::UseORB(): FAST / FilterCorners / GetCornersInRange margins of 4 + orb_size/2, ORBDetector::GetOrientation
    and GetDescriptor at every corner, ORBDetector::Distance, and Matcher::SearchPoint scoring by descriptor distance
    (threshold 100) -- the reference's own extra/orb_detector.cc and matcher.cc against the oracle."""
    cfg, poses, imgs = sw.sequence(name, seed, 5)
    P, cam = cfg["params"], cfg["cam"]
    h, w = imgs[0].shape
    with _both(O, strict):
        O.lib().orc_set_orb(1)
        R.lib().ref_set_orb(1)
        try:
            xo, _ = O.detect(P, imgs[0], P.num_features)
            xr, _ = R.detect(P, imgs[0], P.num_features)
            assert len(xo) > 300 and np.array_equal(xo, xr) and xo[:, 0].min() >= 19 and xo[:, 1].min() >= 19
            io, ir = O.filter_corners(P, imgs[0], P.num_features, np.zeros((0, 2))), R.filter_corners(P, imgs[0], P.num_features, np.zeros((0, 2)))
            assert len(io) > 20 and np.array_equal(io, ir)
            xyl = np.ascontiguousarray(xo)
            do, dr = np.zeros((len(xyl), 32), np.uint8), np.zeros((len(xyl), 32), np.uint8)
            ao, ar = np.zeros(len(xyl), np.float32), np.zeros(len(xyl), np.float32)
            assert O.lib().orc_orb_descriptors(C.byref(P), O.ptr(imgs[0]), w, h, O.ptr(xyl), len(xyl), O.ptr(do), O.ptr(ao)) == 0
            assert R.lib().ref_orb_descriptors(C.byref(P), O.ptr(imgs[0]), w, h, O.ptr(xyl), len(xyl), O.ptr(dr), O.ptr(ar)) == 0
            assert np.array_equal(ao, ar)
            if strict:
                assert np.array_equal(do, dr)
            else:   # FMA contraction in the reference's build can move a sample to the neighbouring pixel
                assert (do == dr).all(1).mean() > 0.98
            for i in range(0, len(xyl) - 1, 7):
                d = int(np.unpackbits(do[i] ^ do[i + 1]).sum())
                assert R.lib().ref_orb_distance(O.ptr(do[i]), O.ptr(do[i + 1])) == d
            pts = scenes.seed_points(cfg, xo, poses[0], max_points=400, one_per_cell=False, margin=20)
            n_found = 0
            for gap in (1, 4):
                for fixed, std_frac in ((True, 0.05), (False, 0.5)):
                    c = scenes.candidates(pts, poses[0], 0, fixed=fixed, project=True, std_frac=std_frac)
                    mo = O.search_points(P, cam, imgs[gap], poses[gap], [imgs[0]], c)
                    mr = R.search_points(P, cam, imgs[gap], poses[gap], [imgs[0]], c)
                    assert np.array_equal(mo["status"], mr["status"]) and np.array_equal(mo["level"], mr["level"])
                    n_found += int((mo["status"] == abi.MATCH_FOUND).sum())
                    if strict:
                        assert np.array_equal(mo["px"], mr["px"])
                    else:
                        assert np.abs(mo["px"] - mr["px"]).max() < 1e-4
            assert n_found > 400
        finally:
            O.lib().orc_set_orb(0)
            R.lib().ref_set_orb(0)


@needs_ref
@pytest.mark.parametrize("name,seed,n", [("C2", 9, 30), ("C3", 5, 30)])
def test_orb_mode_trajectory_vs_reference(O, sw, name, seed, n):
    """Whole sequences with Config::UseORB(): FAST with the ORB margin, init features carrying descriptors, every
    SearchPoint of FeatureAlign::Reproject scored by descriptor distance -- reference code vs oracle, bit for bit."""
    cfg, poses, imgs = sw.sequence(name, seed, n)
    with _both(O, True):
        O.lib().orc_set_orb(1)
        R.lib().ref_set_orb(1)
        try:
            t = O.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20)
            eo, so, _ = t.run(imgs, poses)
            t.close()
            t = R.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20)
            er, sr, _ = t.run(imgs, poses)
            t.close()
        finally:
            O.lib().orc_set_orb(0)
            R.lib().ref_set_orb(0)
    assert np.array_equal(so[:, STAT_COLS], sr[:, STAT_COLS]) and np.array_equal(eo, er)
    assert so[1:, 1].mean() > 60 and sw.ate(er, poses) < 1e-3


# ------------------------------------------------------------------------------------------------ Map (mapping thread)
@needs_ref
@pytest.mark.parametrize("orb", [False, True])
def test_update_candidates_vs_reference(O, sw, scenes, abi, orb):
    """Map::UpdateCandidates run by the reference on its own candidates_ list: which candidates converge, which are
    handed to DeletePoint (too old and out of view / too many failures), and the depth-filter state of every other.
    orb: Config::UseORB() -- the epipolar SearchPoint is scored by the distance to the init feature's descriptor."""
    cfg, poses, imgs = sw.sequence("C2", 0, 25)
    P, cam = cfg["params"], cfg["cam"]
    with _both(O, True):
        O.lib().orc_set_orb(int(orb))
        R.lib().ref_set_orb(int(orb))
        try:
            xyl, _ = O.detect(P, imgs[0], P.num_features)
            pts = scenes.seed_points(cfg, xyl, poses[0], one_per_cell=True, margin=12)
            n = len(pts["px"])
            rng = np.random.default_rng(0)
            s = np.zeros(n, abi.SEED_DT)
            s["ref_frame"] = 0
            s["ref_T"] = poses[0]
            s["ref_px"] = pts["px"]; s["ref_v"] = pts["v"]; s["ref_level"] = pts["level"]
            s["rho"] = 1.0 / (pts["depth"] * (1.0 + rng.uniform(-0.1, 0.1, n)))
            s["sigma2"] = 1.0; s["a"] = 10.0; s["b"] = 10.0; s["z_range"] = 6.0; s["cos_alpha"] = 1.0
            s["last_distance"] = 1.0 / s["rho"]
            s["n_failed"] = rng.integers(0, 16, n)
            s["last_kf_id"] = rng.integers(0, 10, n)
            depth_mean = float(np.median(pts["depth"]))
            so, sr = s.copy(), s.copy()
            live = np.ones(n, bool)
            seen = np.zeros(10, int)
            for k in range(1, 25, 3):
                so[live] = O.update_candidates(P, cam, imgs[k], poses[k], [imgs[0]], so[live], depth_mean, min_kf_id=5)
                sr[live] = R.update_candidates(P, cam, imgs[k], poses[k], [imgs[0]], sr[live], depth_mean, min_kf_id=5)
                seen += np.bincount(so["status"][live], minlength=10)
                conv_o, conv_r = so["status"] == abi.SEED_CONVERGED, sr["status"] == abi.SEED_CONVERGED
                del_o = np.isin(so["status"], (abi.SEED_DELETE_OLD, abi.SEED_DELETE_FAILED))
                del_r = sr["status"] == abi.SEED_DELETE_OLD
                assert np.array_equal(conv_o, conv_r) and np.array_equal(del_o, del_r)
                assert np.array_equal(so["n_failed"][live], sr["n_failed"][live])
                for f in ("rho", "sigma2", "a", "b", "cos_alpha", "last_distance"):
                    assert np.allclose(so[f][live], sr[f][live], rtol=1e-10, atol=0), f
                if conv_o.any():
                    assert np.abs(so["p3d"][conv_o] - sr["p3d"][conv_o]).max() < 1e-12
                live &= ~(conv_o | del_o)
        finally:
            O.lib().orc_set_orb(0)
            R.lib().ref_set_orb(0)
    # the run exercised every branch the two lists can show
    for st in (abi.SEED_NOT_VISIBLE, abi.SEED_DELETE_OLD, abi.SEED_SHORT_BASELINE, abi.SEED_NOT_FOUND,
               abi.SEED_DELETE_FAILED, abi.SEED_UPDATED) + (() if orb else (abi.SEED_CONVERGED,)):
        assert seen[st] > 0, st


@needs_ref
@pytest.mark.parametrize("name,seed", [("C2", 2), ("C3", 1), ("C1", 3)])
def test_init_candidates_vs_reference(O, sw, scenes, abi, name, seed):
    """Map::InitCandidates run by the reference on a new keyframe connected to an older one: which filtered corners
    become candidates (SearchPoint along the whole epipolar range, triangulation, parallax and minimum-distance
    screens), with which depth and which match, and the candidates_ list holding every one of them twice
    (map.cc:384,392) -- against the oracle's per-corner pass (SDVLB_SEEDS_INIT), which is what the device runs."""
    cfg, poses, imgs = sw.sequence(name, seed, 10)
    P, cam = cfg["params"], cfg["cam"]
    k_old, k_new = 0, 9
    with _both(O, True):
        xyl, _ = O.detect(P, imgs[k_new], P.num_features)
        depth_mean = float(np.median(scenes.seed_points(cfg, xyl, poses[k_new])["depth"]))
        # the frame's only feature (the depth anchor on the optical axis) locks its cell, as in the reference harness
        idx = O.filter_corners(P, imgs[k_new], P.num_features, np.array([[cam.u0, cam.v0]]))
        s = np.zeros(len(idx), abi.SEED_DT)
        c = xyl[idx].astype(np.float64)
        px = c[:, :2] * (1 << xyl[idx, 2])[:, None]
        v = np.stack([(px[:, 0] - cam.u0) / cam.fx, (px[:, 1] - cam.v0) / cam.fy, np.ones(len(idx))], axis=1)
        v /= np.sqrt((v * v).sum(1))[:, None]
        s["ref_frame"] = 0
        s["ref_T"] = poses[k_new]
        s["ref_px"] = px; s["ref_v"] = v; s["ref_level"] = xyl[idx, 2]
        s["rho"] = 1.0 / depth_mean
        s["sigma2"] = 1.0; s["a"] = 10; s["b"] = 10; s["z_range"] = 6; s["cos_alpha"] = 1; s["last_distance"] = depth_mean
        exp = O.update_candidates(P, cam, imgs[k_old], poses[k_old], [imgs[k_new]], s, depth_mean, mode=abi.SEEDS_INIT)
        got = R.init_candidates(P, cam, imgs[k_new], poses[k_new], imgs[k_old], poses[k_old], depth_mean)
    ok = exp["status"] == abi.SEED_UPDATED
    assert ok.sum() > 40
    assert got["listed"] == 2 * ok.sum() and len(got["depth"]) == ok.sum()
    assert np.array_equal(got["ref_px"], exp["ref_px"][ok]) and np.array_equal(got["ref_level"], exp["ref_level"][ok])
    assert np.array_equal(got["level"], exp["level"][ok]) and np.array_equal(got["px"], exp["px"][ok])
    assert np.allclose(got["depth"], exp["depth"][ok], rtol=1e-12, atol=0)


@needs_ref
@pytest.mark.parametrize("name,seed,fixed", [("C2", 2, True), ("C2", 2, False), ("C3", 1, True)])
def test_add_connections_points_vs_reference(O, sw, scenes, abi, name, seed, fixed):
    """Map::AddConnectionsPoints run by the reference (projection into the new keyframe, the patch_size margin,
    SearchPoint, one new Feature per found point pushed to the front of the point's list) against the oracle's
    SearchPoint with SDVLB_CAND_PROJECT, which is how the device serves that loop."""
    cfg, poses, imgs = sw.sequence(name, seed, 10)
    P, cam = cfg["params"], cfg["cam"]
    with _both(O, True):
        xyl, _ = O.detect(P, imgs[0], P.num_features)
        pts = scenes.seed_points(cfg, xyl, poses[0], max_points=300, one_per_cell=False)
        c = scenes.candidates(pts, poses[0], 0, fixed=fixed, project=True, std_frac=0.05 if fixed else 0.02)
        exp = O.search_points(P, cam, imgs[9], poses[9], [imgs[0]], c)
        got = R.add_connections_points(P, cam, imgs[9], poses[9], imgs[0], poses[0], c)
    f = exp["status"] == abi.MATCH_FOUND
    assert f.sum() > 60 and (~f).sum() > 10
    assert np.array_equal(got["status"] == abi.MATCH_FOUND, f)
    assert np.array_equal(got["level"][f], exp["level"][f])
    if fixed:
        assert np.array_equal(got["px"][f], exp["px"][f])
    else:   # the reference recomputes the position from the inverse depth: the projection differs in the last bits
        assert np.abs(got["px"][f] - exp["px"][f]).max() < 1e-4


# ------------------------------------------------------------------------------------------------ relocalisation
@needs_ref
@pytest.mark.parametrize("name,seed,gap", [("C2", 0, 1), ("C2", 0, 12), ("C3", 5, 2), ("C3", 5, 25)])
def test_relocalize_vs_reference(O, sw, scenes, name, seed, gap):
    """SDVL::Relocalize's body for one keyframe: ImageAlign::ComputePose(fast = true) from the keyframe's pose (the
    early exit after the coarsest level when its error is large), the GetError() gate, and
    FeatureAlign::Reproject(reloc = true), which must neither promote points nor add features to the frame."""
    with _both(O, True):
        cfg, poses, imgs, pts = _scene(O, sw, scenes, name, seed, gap + 1, cfg_feats(name))
        P, cam = cfg["params"], cfg["cam"]
        feats = scenes.align_feats(pts, poses[0])
        a = O.relocalize(P, cam, imgs[0], imgs[gap], feats, pts["pos"], pts["level"], poses[0])
        b = R.relocalize(P, cam, imgs[0], imgs[gap], feats, pts["pos"], pts["level"], poses[0])
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2])
    assert a[2][2] == 0 and a[2][3] == 0                     # reloc: no features created, no point promoted
    if gap <= 2:
        assert a[2][0] > 50 and a[1] < 0.001                 # a nearby frame relocalises against the keyframe
    else:
        assert a[2][0] == -1 or a[2][0] < a[2][1]            # a far one is rejected by the error gate or matches poorly


# ------------------------------------------------------------------------------------------------ whole trajectories
@needs_ref
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name,seed,n", [("C2", 9, 45), ("C1", 3, 30), ("C3", 5, 45), ("C5", 1, 12)])
def test_trajectory_vs_reference(O, sw, name, seed, n, strict):
    """The reference's ImageAlign + FeatureAlign::Reproject + OptimizePose + motion model + Map::NeedKeyframe /
    EmptyTrash driven frame after frame (oracle/ref_harness.cc) against oracle/tracker.cc, same seeded map: per-frame
    tracked / matches / attempts / inliers / outliers / feature counts and keyframe decisions identical; poses
    bit-identical (strict) or within 0.1 mm (default flags: rounding differences feed back through 45 frames of RANSAC and
    IRLS; north_star allows 1 mm ATE)."""
    cfg, poses, imgs = sw.sequence(name, seed, n)
    with _both(O, strict):
        t = O.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20)
        eo, so, _ = t.run(imgs, poses)
        t.close()
        t = R.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20)
        er, sr, _ = t.run(imgs, poses)
        t.close()
    assert np.array_equal(so[:, STAT_COLS], sr[:, STAT_COLS])
    assert so[:, 7].sum() >= (2 if n >= 30 else 1) and so[1:, 1].mean() > 60   # keyframes were inserted, points were matched
    if strict:
        assert np.array_equal(eo, er)
    else:
        assert np.abs(eo - er).max() < 1e-4 and sw.ate(eo, er) < 1e-4
    assert sw.ate(er, poses) < 1e-3


# ------------------------------------------------------------------------------------------------ committed fixtures
def test_oracle_matches_reference_golden(O, sw, scenes, abi):
    """tests/golden/ref_golden.npz holds outputs of the reference library (strict build) on seeded scenes, written
    by tests/golden/make_ref_golden.py; the oracle must reproduce them wherever it is built (no reference needed)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz")
    g = np.load(path)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(os.path.dirname(path), "make_ref_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    with O.strict():
        mine = m.compute(O, sw, scenes, abi)
    assert set(mine) == set(g.files)
    for k in g.files:
        a, b = mine[k], g[k]
        assert a.shape == b.shape, k
        if a.dtype.kind in "iub":
            assert np.array_equal(a, b), k
        else:
            assert np.allclose(a, b, rtol=1e-9, atol=1e-12), (k, np.abs(a - b).max())
