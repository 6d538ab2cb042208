"""GPU parity against the REFERENCE'S OWN CODE: the sm_100a path, called through the C-ABI, against
tests/golden/ref_golden.npz -- outputs of oracle/_ref (the reference's sources compiled unmodified from
/root/reference, strict build) on seeded synthetic scenes, written by tests/golden/make_ref_golden.py.  Nothing here
reads /root/reference or loads the oracle: the fixture travels with the repo.

Bars (BASELINE.json north_star): FAST corner lists bit-exact; H, b, x per Gauss-Newton iteration within 1e-4 relative
(teacher-forced on the reference's own iterates); LK positions within 0.01 px; trajectory within 1 mm ATE.
Measured on B200 (gpurun_out -> profiles/r01_gpu_vs_reference_golden.txt): corner lists identical; H 2e-15, x 5e-5,
chi2 2e-6; refined pixels <= 0.0098 px with identical found / level decisions; RANSAC / IRLS sets identical, pose
4e-16; ATE 0.007 mm (C2) and 0.04 mm (C3) against the reference's own trajectories."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-4
LK_PX = 0.01
ATE_MM = 1.0
STAT_COLS = [0, 1, 2, 3, 4, 5, 7]


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))


@pytest.fixture(scope="module")
def GD():
    """The same reference code built with its default flags (FMA contraction at the compiler's discretion)."""
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden_default_flags.npz"))


@pytest.mark.parametrize("name,seed,gap", [("C1", 3, 2), ("C2", 0, 3)])
def test_corners_and_image_align_vs_reference(binding, sw, scenes, G, name, seed, gap):
    """Frame::CreateCorners: the reference's corner list, same order.  ImageAlign: every iteration the reference ran,
    evaluated at the reference's own T_in, reproduces its n_meas exactly and its H, b, x, chi2 within 1e-4; the
    free-running alignment ends on the reference's pose."""
    k = "align_%s_" % name
    cfg, poses, imgs = sw.sequence(name, seed, gap + 1)
    P = cfg["params"]
    ctx = binding.Context(P, cfg["cam"])
    try:
        ref = ctx.frame(imgs[0], corners=True)
        cur = ctx.frame(imgs[gap], corners=False)
        xyl, _ = ref.corners()
        assert np.array_equal(xyl, G[k + "corners"]), "FAST corner list differs from the reference's"
        pts = scenes.seed_points(cfg, xyl, poses[0], max_points=cfg["n_feat"])
        feats = scenes.align_feats(pts, poses[0], invalid_every=9)
        iters = np.zeros(8, np.int32)
        for l in G[k + "level"]:
            iters[l] += 1
        T_g, nt_g, err_g, tr = ctx.image_align(ref, cur, feats, poses[0], poses[0],
                                               forced=(np.ascontiguousarray(G[k + "T_in"]), iters))
        n = len(G[k + "level"])
        assert len(tr) == n >= 3
        assert np.array_equal(tr["level"], G[k + "level"]) and np.array_equal(tr["iter"], G[k + "iter"])
        assert np.array_equal(tr["n_meas"], G[k + "n_meas"])
        worst = {}
        for key in ("H", "b", "x"):
            den = np.abs(G[k + key]).reshape(n, -1).max(axis=1)
            err = np.abs(tr[key] - G[k + key]).reshape(n, -1).max(axis=1)
            worst[key] = float((err[den > 0] / den[den > 0]).max())
        worst["chi2"] = float((np.abs(tr["chi2"] - G[k + "chi2"]) / np.maximum(np.abs(G[k + "chi2"]), 1e-30)).max())
        # b = -sum J res vanishes as a level converges while its rounding error (f32 residuals, as in the reference)
        # does not: iterations whose |b| has dropped below 0.1 % of the level's first are held to the same ABSOLUTE error
        # (1e-4 of the level's first |b|), the others to 1e-4 relative to their own |b|.
        bden = np.abs(G[k + "b"]).max(axis=1)
        berr = np.abs(tr["b"] - G[k + "b"]).max(axis=1)
        first = np.array([bden[np.flatnonzero(G[k + "level"] == l)[0]] for l in G[k + "level"]])
        live = bden >= 1e-3 * first
        worst["b_live"] = float((berr[live] / bden[live]).max())
        worst["b_abs_over_first"] = float((berr / first).max())
        print(f"{name}: {n} reference GN iterations, worst relative error {worst}")
        assert worst["H"] <= REL and worst["x"] <= REL and worst["chi2"] <= REL, worst
        assert worst["b_live"] <= REL and worst["b_abs_over_first"] <= REL, worst
        T_f, nt_f, err_f, tr_f = ctx.image_align(ref, cur, feats, poses[0], poses[0])
        dC = np.linalg.norm(sw.cam_center(T_f) - sw.cam_center(G[k + "T"]))
        print(f"{name} free run: {len(tr_f)} iterations (reference {n}), |dC| = {dC:.2e} m, tracked {nt_f} vs {int(G[k + 'n_tracked_err'][0])}")
        assert dC < 1e-4 and abs(nt_f - int(G[k + "n_tracked_err"][0])) <= 2
        ref.destroy(); cur.destroy()
    finally:
        ctx.close()


def test_search_point_vs_reference(binding, sw, scenes, abi, G, GD):
    """Matcher::SearchPoint, fixed (circle) and epipolar (capsule) candidates: the reference's found / not-found /
    unseen decisions and search levels, refined positions within 0.01 px -- against BOTH builds of the reference's own
    code: strict (-ffp-contract=off, ref_golden.npz) and default flags (ref_golden_default_flags.npz).  The two builds
    of the reference differ from each other by up to 0.0098 px on these candidates (the float LK stops on
    |update|^2 < 9e-4, so one more or one fewer iteration moves the result by a few thousandths of a pixel;
    make_ref_golden.py prints the figure); the device lands on the default-flag build to ~2e-5 px, so its distance to the
    strict build is that compiler-flag spread of the reference itself."""
    cfg, poses, imgs = sw.sequence("C2", 0, 5)
    P = cfg["params"]
    ctx = binding.Context(P, cfg["cam"])
    try:
        ref = ctx.frame(imgs[0], corners=True)
        cur = ctx.frame(imgs[4], corners=True)
        xyl, _ = ref.corners()
        pts = scenes.seed_points(cfg, xyl, poses[0], max_points=300, one_per_cell=False)
        for fixed, std_frac in ((True, 0.05), (False, 0.5)):
            k = "search_%s_" % ("fixed" if fixed else "epipolar")
            c = scenes.candidates(pts, poses[0], ref.h, fixed=fixed, project=True, std_frac=std_frac)
            got = ctx.search_points(cur, c, poses[4])
            assert np.array_equal(got["status"], G[k + "status"]), \
                f"status differs for {(got['status'] != G[k + 'status']).sum()} of {len(c)} candidates"
            f = G[k + "status"] == abi.MATCH_FOUND
            assert f.sum() > 80
            assert np.array_equal(got["level"][f], G[k + "level"][f])
            d = np.abs(got["px"][f] - G[k + "px"][f]).max()
            assert np.array_equal(got["status"], GD[k + "status"]) and np.array_equal(got["level"][f], GD[k + "level"][f])
            dd = np.abs(got["px"][f] - GD[k + "px"][f]).max()
            print(f"fixed={fixed}: {f.sum()} found of {len(c)}, max |dpx| vs the reference: strict build {d:.2e}, "
                  f"default-flag build {dd:.2e}")
            assert d <= LK_PX and dd <= LK_PX
        ref.destroy(); cur.destroy()
    finally:
        ctx.close()


def test_pose_refinement_vs_reference(binding, sw, abi, G):
    """FeatureAlign::SelectInliers (rand() stream of srand(1)) and OptimizePose: the reference's inlier / outlier sets
    and its refined pose."""
    cfg, poses, imgs = sw.sequence("C2", 0, 2)   # only for the pose array shape the fixture generator used
    cam = cfg["cam"]
    rng = np.random.default_rng(11)
    T_true = sw.trajectory(cfg, 4, 3)[2]
    Rm = sw.quat_R(T_true[:4])
    n = 120
    obs = np.zeros(n, abi.POSE_OBS_DT)
    for i in range(n):
        u, v, depth = rng.uniform(20, cam.width - 20), rng.uniform(20, cam.height - 20), rng.uniform(1, 4)
        ray = np.array([(u - cam.u0) / cam.fx, (v - cam.v0) / cam.fy, 1.0])
        obs["pos"][i] = Rm.T @ (ray * depth - T_true[4:])
        u += rng.normal(0, 0.3) + (25.0 if i % 7 == 0 else 0.0)
        v += rng.normal(0, 0.3)
        b = np.array([(u - cam.u0) / cam.fx, (v - cam.v0) / cam.fy, 1.0])
        obs["v"][i] = b / np.linalg.norm(b)
        obs["level"][i] = i % 3
    T0 = T_true.copy()
    T0[4:] += [0.004, -0.003, 0.002]
    ctx = binding.Context(cfg["params"], cam)
    try:
        r = abi.Rand()
        binding.load().sdvlb_rand_seed(C.byref(r), 1)
        g = ctx.select_inliers(obs.copy(), T0, r)
        assert np.array_equal(g["flags"], G["refine_ransac_flags"])
        g2, T = ctx.optimize_pose(g.copy(), T0)
        assert np.array_equal(g2["flags"], G["refine_final_flags"])
        d = max(np.abs(T[4:] - G["refine_T"][4:]).max(),
                min(np.abs(T[:4] - G["refine_T"][:4]).max(), np.abs(T[:4] + G["refine_T"][:4]).max()))
        print(f"pose refinement vs the reference: {int((g2['flags'] == abi.OBS_INLIER).sum())} inliers, pose diff {d:.2e}")
        assert d < 1e-9
    finally:
        ctx.close()


@pytest.mark.parametrize("name,seed,n", [("C2", 9, 30), ("C3", 5, 30)])
@pytest.mark.parametrize("resident", [False, True])
def test_trajectory_vs_reference(binding, sw, G, name, seed, n, resident):
    """Whole sequences: the reference's ImageAlign + FeatureAlign + motion model + keyframe test, frame after frame,
    against the class-API path and the resident-sequence path: ATE within 1 mm of the reference's trajectory, and
    the per-frame match / inlier / keyframe counts it reported."""
    cfg, poses, imgs = sw.sequence(name, seed, n)
    t = binding.HostTracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20, 1, 1, resident=resident)
    est = np.zeros((n, 7))
    stats = np.zeros((n, 8), np.int32)
    for k in range(n):
        e, st = t.step(imgs[k:k + 1], poses[k:k + 1], classic=not resident)
        est[k], stats[k] = e[0], st[0]
    t.close()
    ref_poses, ref_stats = G["traj_%s_poses" % name], G["traj_%s_stats" % name]
    d = np.array([np.linalg.norm(sw.cam_center(a) - sw.cam_center(b)) for a, b in zip(est, ref_poses)])
    ate = float(np.sqrt((d ** 2).mean())) * 1e3
    same = float((stats[:, STAT_COLS][:, 1:] == ref_stats[:, 1:]).all(axis=1).mean())
    print(f"{name} resident={resident}: ATE vs the reference's trajectory {ate:.5f} mm (max {d.max() * 1e3:.5f}), "
          f"frames with the reference's match/inlier/keyframe counts {same:.0%}")
    assert ate <= ATE_MM
    assert np.array_equal(stats[:, 7], ref_stats[:, 6]), "keyframe decisions differ from the reference's"
    # counts: identical on C2; on the fast C3 motion a match near its acceptance threshold flips now and then (f32 LK /
    # ZMSSD gate on poses that differ by micrometres), so the bar there is the mean, not every frame
    dm = np.abs(stats[1:, 1].astype(float) - ref_stats[1:, 1]).mean()
    print(f"{name}: mean |matches - reference's| per frame {dm:.2f} of {ref_stats[1:, 1].mean():.0f}")
    assert dm <= 0.02 * ref_stats[1:, 1].mean()
    if name == "C2":
        assert same >= 0.9
