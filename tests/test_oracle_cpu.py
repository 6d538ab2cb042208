"""CPU suite, part 1: pins the oracle (the parity checker) against everything available without a GPU:
OpenCV via committed golden vectors and live cv2, glibc/libstdc++ randomness, and analytic known-answer tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "cv2_golden.npz"))


# ------------------------------------------------------------------ third-party arithmetic: golden vectors
@pytest.mark.parametrize("name,levels", [("crop", 4), ("noise", 3)])
def test_pyrdown_golden(O, name, levels):
    pyr = O.pyramid(GOLD[f"pyr_{name}_0"], levels)
    for l in range(levels):
        assert np.array_equal(pyr[l], GOLD[f"pyr_{name}_{l}"]), f"{name} level {l}"


def test_fast_golden(O):
    for k in range(int(GOLD["n_rois"])):
        roi = np.ascontiguousarray(GOLD[f"fast_roi_{k}"])
        xy, sc = O.fast_roi(roi, 0, 0, roi.shape[1], roi.shape[0], 10)
        got = np.c_[xy, sc].reshape(-1, 3)
        assert np.array_equal(got, GOLD[f"fast_kps_{k}"]), f"roi {k}"


# ------------------------------------------------------------------ third-party arithmetic: live cv2 (same image)
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("w,h,levels", [(752, 480, 5), (640, 480, 4), (1920, 1080, 5), (47, 30, 2), (135, 67, 2)])
def test_pyrdown_vs_cv2(O, w, h, levels):
    rng = np.random.default_rng(w * 7 + h)
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (5, 5), 1.5)
    ref = [img]
    for _ in range(1, levels):
        ref.append(cv2.pyrDown(ref[-1], dstsize=(ref[-1].shape[1] // 2, ref[-1].shape[0] // 2)))
    for a, b in zip(O.pyramid(img, levels), ref):
        assert np.array_equal(a, b)


def test_fast_vs_cv2_rois(O):
    rng = np.random.default_rng(3)
    fd = cv2.FastFeatureDetector_create(10, True)
    n_kp = 0
    for t in range(120):
        cols, rows = int(rng.integers(7, 40)), int(rng.integers(7, 40))
        big = rng.integers(0, 256, (64, 64), dtype=np.uint8)
        if t % 2:
            big = cv2.GaussianBlur(big, (3, 3), 0.8)
        x0, y0 = int(rng.integers(0, 64 - cols)), int(rng.integers(0, 64 - rows))
        ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in fd.detect(np.ascontiguousarray(big[y0:y0 + rows, x0:x0 + cols]))]
        xy, sc = O.fast_roi(big, x0, y0, cols, rows, 10)
        got = [(int(a), int(b), int(c)) for (a, b), c in zip(xy, sc)]
        assert got == ref, f"roi {t}"
        n_kp += len(ref)
    assert n_kp > 500


def test_detect_pyramid_against_cv2_composition(O, abi, seq_c2):
    """FastDetector::SelectPixels restated in numpy on top of cv2.FAST per cell ROI: kept SET must equal the oracle's."""
    cfg, poses, imgs = seq_c2
    P = cfg["params"]
    img = imgs[0]
    xyl, sc = O.detect(P, img, P.num_features)
    pyr = O.pyramid(img, P.pyramid_levels)
    fd = cv2.FastFeatureDetector_create(P.fast_threshold, True)
    val = sum(1.2 ** -i for i in range(P.max_fast_levels))
    lf = int(P.num_features / val)
    expect = set()
    for level in range(P.max_fast_levels):
        src = pyr[level]
        m = 1 + P.patch_size // 2
        hc, wc = int(np.ceil(src.shape[0] / 32)), int(np.ceil(src.shape[1] / 32))
        cells, nleft = {}, {}
        nempty = 0
        for i in range(hc):
            iy, my = max(m, 32 * i), min(src.shape[0] - m, 32 * i + 32)
            if my <= iy:
                continue
            for j in range(wc):
                ix, mx = max(m, 32 * j), min(src.shape[1] - m, 32 * j + 32)
                if mx <= ix:
                    continue
                k = [(int(p.pt[0]) + ix, int(p.pt[1]) + iy, int(p.response)) for p in fd.detect(np.ascontiguousarray(src[iy:my, ix:mx]))]
                cells[(i, j)] = k
                nleft[(i, j)] = len(k)
                nempty += len(k) == 0
        nsel = {c: 0 for c in cells}
        selected, cells_left = 0, hc * wc - nempty
        while lf - selected > 0 and cells_left > 0:
            per = int(np.ceil((lf - selected) / cells_left))
            cells_left = 0
            for i in range(hc):
                for j in range(wc):
                    c = (i, j)
                    if c in nleft and nleft[c] > 0:
                        if nleft[c] > per:
                            nsel[c] += per; selected += per; nleft[c] -= per; cells_left += 1
                        else:
                            nsel[c] += nleft[c]; selected += nleft[c]; nleft[c] = 0
        kept = []
        for i in range(hc):
            for j in range(wc):
                k = cells.get((i, j), [])
                n = nsel.get((i, j), 0)
                if len(k) > n:
                    if n == 0:
                        k = []
                    else:
                        thr = sorted((r for _, _, r in k), reverse=True)[n - 1]
                        k = [e for e in k if e[2] >= thr]      # retainBest keeps boundary ties
                kept += k
        if len(kept) > lf:
            thr = sorted((r for _, _, r in kept), reverse=True)[lf - 1]
            kept = [e for e in kept if e[2] >= thr]
        expect |= {(x, y, level, r) for x, y, r in kept}
        lf = int(lf / 1.2)
    got = {(int(a), int(b), int(c), int(s)) for (a, b, c), s in zip(xyl, sc)}
    assert got == expect


# ------------------------------------------------------------------ libc / libstdc++ randomness
def test_rand_matches_glibc(O):
    libc = C.CDLL("libc.so.6")
    for seed in (1, 7, 12345):
        libc.srand(seed)
        ref = [libc.rand() for _ in range(3000)]
        out = np.zeros(3000, np.int32)
        O.lib().orc_rand_seq(seed, 3000, O.ptr(out))
        assert out.tolist() == ref


def test_shuffle_matches_libstdcxx(O, tmp_path):
    so = tmp_path / "libshuffle_shim.so"
    subprocess.check_call(["g++", "-O1", "-std=c++14", "-fPIC", "-shared", "-Wno-deprecated-declarations",
                           os.path.join(ROOT, "tests", "host_shims", "shuffle_shim.cc"), "-o", str(so)])
    S = C.CDLL(str(so))
    for n in (1, 2, 17, 360):
        a = np.arange(n, dtype=np.int32)
        b = a.copy()
        S.shim_std_shuffle(n, O.ptr(a), 1, 1)
        O.lib().orc_shuffle(n, O.ptr(b), 1)
        assert np.array_equal(a, b)


# ------------------------------------------------------------------ analytic known-answer tests
def test_se3_exp_log_roundtrip(O):
    rng = np.random.default_rng(0)
    for _ in range(50):
        u = rng.normal(0, 0.3, 6)
        T = np.zeros(7)
        v = np.zeros(6)
        O.lib().orc_se3_exp(O.ptr(u), O.ptr(T))
        O.lib().orc_se3_log(O.ptr(T), O.ptr(v))
        assert np.allclose(u, v, atol=1e-12)
        Ti, I = np.zeros(7), np.zeros(7)
        O.lib().orc_se3_inv(O.ptr(T), O.ptr(Ti))
        O.lib().orc_se3_mul(O.ptr(T), O.ptr(Ti), O.ptr(I))
        assert np.allclose(I, [1, 0, 0, 0, 0, 0, 0], atol=1e-12)


def test_ldlt_matches_numpy(O):
    rng = np.random.default_rng(1)
    for _ in range(50):
        A = rng.normal(size=(12, 6))
        H = A.T @ A
        b = rng.normal(size=6)
        x = np.zeros(6)
        O.lib().orc_ldlt6(O.ptr(np.ascontiguousarray(H)), O.ptr(b), O.ptr(x))
        assert np.allclose(x, np.linalg.solve(H, b), rtol=1e-9, atol=1e-12)
    x = np.ones(6)
    O.lib().orc_ldlt6(O.ptr(np.zeros((6, 6))), O.ptr(np.ones(6)), O.ptr(x))
    assert np.array_equal(x, np.zeros(6))      # Eigen's pseudo-inverse of D: zero matrix -> zero solution, no NaN


def test_image_align_identity_and_recovery(O, sw, scenes, seq_c2):
    cfg, poses, imgs = seq_c2
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = O.detect(P, imgs[0], P.num_features)
    pts = scenes.seed_points(cfg, xyl, poses[0], max_points=200)
    feats = scenes.align_feats(pts, poses[0])
    # identity: b = 0, x = 0, pose unchanged
    T, nt, err, tr = O.image_align(P, cam, imgs[0], imgs[0], feats, pts["pos"], poses[0], poses[0])
    assert nt > 150 and np.abs(tr[0]["x"]).max() < 1e-9 and np.allclose(T, poses[0], atol=1e-9)
    # recovery: starting from the previous pose the alignment lands within 2 mm of ground truth
    T, nt, err, tr = O.image_align(P, cam, imgs[0], imgs[1], feats, pts["pos"], poses[0], poses[0])
    d0 = np.linalg.norm(sw.cam_center(poses[0]) - sw.cam_center(poses[1]))
    d1 = np.linalg.norm(sw.cam_center(T) - sw.cam_center(poses[1]))
    assert d0 > 5e-3 and d1 < 2e-3, (d0, d1)


def test_lk_recovers_integer_shift(O, abi, seq_c2):
    cfg, poses, imgs = seq_c2
    P = cfg["params"]
    img = imgs[0]
    xyl, _ = O.detect(P, img, P.num_features)
    ok = 0
    for (x, y, l) in xyl[xyl[:, 2] == 0][:40]:
        if x < 20 or y < 20 or x > img.shape[1] - 20 or y > img.shape[0] - 20:
            continue
        bp = np.ascontiguousarray(img[y - 5:y + 5, x - 5:x + 5])          # 10x10 template centred like CreatePatch
        px = np.array([x + 1.0, y - 1.0])                                  # start one pixel off
        r = O.lib().orc_align_patch(C.byref(P), O.ptr(img), img.shape[1], img.shape[0], O.ptr(bp), O.ptr(px))
        if r:
            assert abs(px[0] - x) < 0.01 and abs(px[1] - y) < 0.01, (px, x, y)
            ok += 1
    assert ok >= 10


def test_tracker_follows_ground_truth(O, sw):
    cfg, poses, imgs = sw.sequence("C2", 9, 25)
    tr = O.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20)
    est, stats, _ = tr.run(imgs, poses)
    tr.close()
    assert sw.ate(est, poses) < 1e-3
    assert stats[1:, 1].mean() > 80


def test_pose_refine_recovers_pose_and_rejects_outliers(O, sw, abi):
    """Oracle FeatureAlign::SelectInliers + OptimizePose on exact observations of a known pose with a few gross
    outliers: RANSAC flags exactly the corrupted ones and the refined pose is the true one (analytic KAT)."""
    cfg = sw.config("C2")
    cam = cfg["cam"]
    rng = np.random.default_rng(0)
    T_true = sw.trajectory(cfg, 4, 3)[2]
    R = sw.quat_R(T_true[:4])
    n = 80
    obs = np.zeros(n, abi.POSE_OBS_DT)
    bad = set(range(0, n, 9))
    for i in range(n):
        u, v, depth = rng.uniform(20, cam.width - 20), rng.uniform(20, cam.height - 20), rng.uniform(1, 4)
        ray = np.array([(u - cam.u0) / cam.fx, (v - cam.v0) / cam.fy, 1.0])
        obs["pos"][i] = R.T @ (ray * depth - T_true[4:])
        if i in bad:
            u += 35.0
        b = np.array([(u - cam.u0) / cam.fx, (v - cam.v0) / cam.fy, 1.0])
        obs["v"][i] = b / np.linalg.norm(b)
    du = np.array([0.004, -0.003, 0.002, 0.001, -0.002, 0.001])
    dT, T0 = np.zeros(7), np.zeros(7)
    O.lib().orc_se3_exp(O.ptr(du), O.ptr(dT))
    O.lib().orc_se3_mul(O.ptr(dT), O.ptr(np.ascontiguousarray(T_true)), O.ptr(T0))
    r = abi.Rand()
    st = np.zeros(1, np.int32)
    O.lib().orc_rand_state(1, C.byref(r))   # srand(1)
    o, _ = O.pose_refine(cfg["params"], cam, obs, T0, r, mode=0)
    assert set(np.flatnonzero(o["flags"] == abi.OBS_OUTLIER)) == bad
    assert r.n > 344   # RANSAC drew from the stream
    o2, T = O.pose_refine(cfg["params"], cam, o, T0, None, mode=1)
    assert set(np.flatnonzero(o2["flags"] == abi.OBS_OUTLIER)) == bad
    assert np.abs(T[4:] - T_true[4:]).max() < 1e-9
    assert min(np.abs(T[:4] - T_true[:4]).max(), np.abs(T[:4] + T_true[:4]).max()) < 1e-9


def test_shi_tomasi_known_answers(O):
    """FindShiTomasiScoreAtPoint: flat patch -> 0, vertical step edge -> 0 (one zero eigenvalue), checkerboard corner
    -> the closed-form smaller eigenvalue; too close to the border -> 0."""
    img = np.full((40, 40), 90, np.uint8)
    assert O.shi_tomasi(img, 20, 20) == 0.0
    edge = img.copy()
    edge[:, 20:] = 200
    assert abs(O.shi_tomasi(edge, 20, 20)) < 1e-3
    rng = np.random.default_rng(1)
    tex = rng.integers(0, 256, (40, 40)).astype(np.uint8)
    x0, y0 = 20, 18
    dx = tex[y0 - 4:y0 + 4, x0 - 3:x0 + 5].astype(np.float64) - tex[y0 - 4:y0 + 4, x0 - 5:x0 + 3]
    dy = tex[y0 - 3:y0 + 5, x0 - 4:x0 + 4].astype(np.float64) - tex[y0 - 5:y0 + 3, x0 - 4:x0 + 4]
    a, b, c = (dx * dx).sum() / 128, (dy * dy).sum() / 128, (dx * dy).sum() / 128
    expect = 0.5 * (a + b - np.sqrt((a + b) ** 2 - 4 * (a * b - c * c)))
    got = O.shi_tomasi(tex, x0, y0)
    assert abs(got - expect) <= 1e-3 * expect
    assert O.shi_tomasi(tex, 4, 20) == 0.0


def test_depth_filter_converges_to_the_plane(O, sw, scenes, abi):
    """Known answer for Map::UpdateCandidates + Point::Update: candidates started 10 % off converge onto the synthetic
    plane z = 0 when updated with frames of known pose."""
    cfg, poses, imgs = sw.sequence("C2", 0, 25)
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = O.detect(P, imgs[0], P.num_features)
    pts = scenes.seed_points(cfg, xyl, poses[0], one_per_cell=True, margin=12)
    n = len(pts["px"])
    assert n > 150
    rng = np.random.default_rng(0)
    s = np.zeros(n, abi.SEED_DT)
    s["ref_frame"] = 0
    s["ref_T"] = poses[0]
    s["ref_px"] = pts["px"]; s["ref_v"] = pts["v"]; s["ref_level"] = pts["level"]
    s["rho"] = 1.0 / (pts["depth"] * (1.0 + rng.uniform(-0.1, 0.1, n)))
    s["sigma2"] = 1.0; s["a"] = 10.0; s["b"] = 10.0; s["z_range"] = 6.0; s["cos_alpha"] = 1.0
    s["last_distance"] = 1.0 / s["rho"]
    depth_mean = float(np.median(pts["depth"]))
    live = np.ones(n, bool)
    sig0 = s["sigma2"].copy()
    for k in range(3, 25, 3):
        s[live] = O.update_candidates(P, cam, imgs[k], poses[k], [imgs[0]], s[live], depth_mean)
        live &= ~np.isin(s["status"], (abi.SEED_CONVERGED, abi.SEED_DELETE_OLD, abi.SEED_DELETE_FAILED))
    conv = s["status"] == abi.SEED_CONVERGED
    assert conv.sum() > 0.4 * n, f"only {conv.sum()} of {n} candidates converged"
    z = np.abs(s["p3d"][conv][:, 2])
    assert np.median(z) < 0.01 and np.percentile(z, 90) < 0.05, (np.median(z), np.percentile(z, 90))
    # the variance of every updated seed shrank, and converged depths are within 1 % of the truth
    upd = s["status"] >= abi.SEED_UPDATED
    assert (s["sigma2"][upd] < sig0[upd]).all()
    err = np.abs(1.0 / s["rho"][conv] - pts["depth"][conv]) / pts["depth"][conv]
    assert np.median(err) < 0.01


# ------------------------------------------------------------------ Camera::UndistortImage = cv::undistort
def test_undistort_golden(O, abi):
    for name in GOLD["und_names"]:
        w, h, fx, fy, u0, v0 = GOLD[f"und_{name}_cam"]
        got = O.undistort(abi.Camera(w, h, fx, fy, u0, v0), GOLD[f"und_{name}_D"], GOLD[f"und_{name}_src"])
        exp = GOLD[f"und_{name}_dst"]
        assert np.array_equal(got, exp), f"{name}: {(got != exp).sum()} pixels differ from cv2.undistort"
    # taps outside the source read 0 (BORDER_CONSTANT): the case with positive k1 maps its corners outside the source
    assert (GOLD["und_barrel_dst"] == 0).sum() > 1000


@pytest.mark.parametrize("w,h,fx,fy,u0,v0,D", [
    (752, 480, 458.654, 457.296, 367.215, 248.375, (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)),   # EuRoC cam0
    (640, 480, 517.3, 516.5, 318.6, 255.3, (0.2624, -0.9531, -0.0054, 0.0026, 1.1633)),                            # TUM fr1
    (333, 201, 300.0, 310.0, 160.2, 99.7, (-0.3, 0.1, 0.001, -0.002, 0.01))])
def test_undistort_vs_cv2(O, abi, w, h, fx, fy, u0, v0, D):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(w)
    src = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), 1.5)
    K = np.array([[fx, 0, u0], [0, fy, v0], [0, 0, 1.0]])
    exp = cv2.undistort(src, K, np.array(D))
    got = O.undistort(abi.Camera(w, h, fx, fy, u0, v0), D, src)
    assert np.array_equal(got, exp)


# ------------------------------------------------------------------ ORB descriptor mode: third-party pins
def test_fast_atan2_vs_cv2(O):
    """cv::fastAtan2 (the orientation of ORBDetector::GetOrientation, extra/orb_detector.cc:436): the oracle's
    restatement of OpenCV's polynomial returns cv2.fastAtan2's bits."""
    cv2 = pytest.importorskip("cv2")
    L = O.lib()
    L.orc_fast_atan2.restype = C.c_float
    L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
    rng = np.random.default_rng(0)
    for i in range(20000):
        y, x = (rng.normal(0, 10 ** rng.uniform(-3, 6), 2)).astype(np.float32)
        if i % 50 == 0:
            y = np.float32(0)
        if i % 77 == 0:
            x = np.float32(0)
        assert L.orc_fast_atan2(float(y), float(x)) == cv2.fastAtan2(float(y), float(x)), (y, x)
    for y, x in ((0, 0), (1, 0), (-1, 0), (0, -1), (5, 5), (-5, 5), (3e5, -2e-3)):
        assert L.orc_fast_atan2(float(y), float(x)) == cv2.fastAtan2(float(y), float(x))


def test_orb_pattern_against_cv2(O, sw):
    """The 256 learned tests of csrc/orb_pattern.h are OpenCV's: cv2.ORB descriptors (computed by OpenCV on its own
    7x7 sigma-2 blur, at the orientation handed to it) against the oracle's GetDescriptor on the same blurred image.
    A wrong table would differ in ~128 of 256 bits; the residue (well under one bit per descriptor) is a sample that
    rounds to the neighbouring pixel in OpenCV's build."""
    cv2 = pytest.importorskip("cv2")
    cfg, poses, imgs = sw.sequence("C2", 0, 1)
    P, img = cfg["params"], imgs[0]
    O.lib().orc_set_orb(1)
    try:
        xyl, _ = O.detect(P, img, 1000)
    finally:
        O.lib().orc_set_orb(0)
    assert xyl[:, 0].min() >= 19 and xyl[:, 1].min() >= 19      # ORB mode: FAST margin 4 + 31/2
    xyl = np.ascontiguousarray(xyl[(xyl[:, 2] == 0) & (xyl[:, 0] > 40) & (xyl[:, 0] < 712) & (xyl[:, 1] > 40) & (xyl[:, 1] < 440)])
    assert len(xyl) > 200
    blur = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    P1 = type(P)()
    C.memmove(C.byref(P1), C.byref(P), C.sizeof(P))
    P1.pyramid_levels = 1
    d = np.zeros((len(xyl), 32), np.uint8)
    ang = np.zeros(len(xyl), np.float32)
    assert O.lib().orc_orb_descriptors(C.byref(P1), O.ptr(blur), img.shape[1], img.shape[0], O.ptr(xyl), len(xyl), O.ptr(d), O.ptr(ang)) == 0
    kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, 0) for (x, y, l), a in zip(xyl, ang)]
    orb = cv2.ORB_create(nfeatures=5000, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0, WTA_K=2, patchSize=31)
    kps2, desc = orb.compute(img, kps)
    idx = {(int(x), int(y)): i for i, (x, y, l) in enumerate(xyl)}
    sel = [idx[(int(k.pt[0]), int(k.pt[1]))] for k in kps2]
    assert len(sel) > 200
    bits = np.unpackbits(d[sel] ^ desc, axis=1).sum(1)
    assert bits.mean() < 0.5 and (bits == 0).mean() > 0.8, (bits.mean(), (bits == 0).mean())
