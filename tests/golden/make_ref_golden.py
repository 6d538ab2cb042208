"""Writes tests/golden/ref_golden.npz: outputs of the REFERENCE'S OWN CODE (oracle/_ref/libsdvlref_strict.so, i.e. the
reference's sources compiled unmodified from /root/reference, see oracle/Makefile target `ref`) on seeded synthetic
scenes.  The scenes come from the repo's deterministic generator (slam-sdvl_b200/synthworld.py), so only outputs are
stored.  tests/test_oracle_vs_ref.py::test_oracle_matches_reference_golden recomputes the same dictionary with the
oracle -- on machines where the reference tree does not exist this fixture is what pins the oracle.

Run here (needs /root/reference):   python tests/golden/make_ref_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

STAT_COLS = [0, 1, 2, 3, 4, 5, 7]


def compute(impl, sw, scenes, abi):
    """impl = oracle.oracle_py or oracle.ref_py (same function names)."""
    is_ref = impl.__name__.endswith("ref_py")
    out = {}
    # ---- ImageAlign traces
    for name, seed, gap in (("C1", 3, 2), ("C2", 0, 3)):
        cfg, poses, imgs = sw.sequence(name, seed, gap + 1)
        P, cam = cfg["params"], cfg["cam"]
        xyl, _ = impl.detect(P, imgs[0], P.num_features)
        pts = scenes.seed_points(cfg, xyl, poses[0], max_points=cfg["n_feat"])
        feats = scenes.align_feats(pts, poses[0], invalid_every=9)
        T, nt, err, tr = impl.image_align(P, cam, imgs[0], imgs[gap], feats, pts["pos"], poses[0], poses[0])
        k = "align_%s_" % name
        out[k + "corners"] = xyl
        out[k + "T"] = T
        out[k + "n_tracked_err"] = np.array([nt, err], np.float64)
        for f in ("level", "iter", "n_meas", "flags", "T_in", "H", "b", "x", "chi2"):
            out[k + f] = np.ascontiguousarray(tr[f])
    # ---- Matcher::SearchPoint, fixed and epipolar
    cfg, poses, imgs = sw.sequence("C2", 0, 5)
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = impl.detect(P, imgs[0], P.num_features)
    pts = scenes.seed_points(cfg, xyl, poses[0], max_points=300, one_per_cell=False)
    for fixed, std_frac in ((True, 0.05), (False, 0.5)):
        c = scenes.candidates(pts, poses[0], 0, fixed=fixed, project=True, std_frac=std_frac)
        m = impl.search_points(P, cam, imgs[4], poses[4], [imgs[0]], c)
        k = "search_%s_" % ("fixed" if fixed else "epipolar")
        out[k + "status"] = np.ascontiguousarray(m["status"])
        out[k + "level"] = np.ascontiguousarray(m["level"])
        out[k + "px"] = np.ascontiguousarray(m["px"])
    # ---- FeatureAlign::SelectInliers / OptimizePose
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
    T0 = poses[1].copy()
    T0[:] = T_true
    T0[4:] += [0.004, -0.003, 0.002]
    if is_ref:
        o1, _ = impl.pose_refine(P, cam, obs, T0, seed=1, mode=0)
        o2, T = impl.pose_refine(P, cam, o1, T0, mode=1)
    else:
        r = abi.Rand()
        impl.lib().orc_rand_state(1, C.byref(r))
        o1, _ = impl.pose_refine(P, cam, obs, T0, r, mode=0)
        o2, T = impl.pose_refine(P, cam, o1, T0, None, mode=1)
    out["refine_ransac_flags"] = np.ascontiguousarray(o1["flags"])
    out["refine_final_flags"] = np.ascontiguousarray(o2["flags"])
    out["refine_T"] = T
    # ---- whole trajectories
    for name, seed, n in (("C2", 9, 30), ("C3", 5, 30)):
        cfg, poses, imgs = sw.sequence(name, seed, n)
        t = impl.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20)
        est, stats, _ = t.run(imgs, poses)
        t.close()
        out["traj_%s_poses" % name] = est
        out["traj_%s_stats" % name] = np.ascontiguousarray(stats[:, STAT_COLS])
    # ---- ORB descriptor mode (Config::UseORB()): corners with the ORB margin, the descriptor of every corner of a
    # frame (Frame::descriptors_), whole trajectories with SearchPoint scored by descriptor distance
    set_orb = impl.lib().ref_set_orb if is_ref else impl.lib().orc_set_orb
    orb_desc = impl.lib().ref_orb_descriptors if is_ref else impl.lib().orc_orb_descriptors
    set_orb(1)
    try:
        cfg, poses, imgs = sw.sequence("C2", 0, 1)
        P = cfg["params"]
        h, w = imgs[0].shape
        xyl, _ = impl.detect(P, imgs[0], P.num_features)
        xc = np.ascontiguousarray(xyl, np.int32)
        d = np.zeros((len(xc), 32), np.uint8)
        ang = np.zeros(len(xc), np.float32)
        assert orb_desc(C.byref(P), abi.ptr(imgs[0]), w, h, abi.ptr(xc), len(xc), abi.ptr(d), abi.ptr(ang)) == 0
        out["orb_C2_corners"] = xc
        out["orb_C2_desc"] = d
        out["orb_C2_angle"] = ang
        for name, seed, n in (("C2", 9, 30), ("C3", 5, 30)):
            cfg, poses, imgs = sw.sequence(name, seed, n)
            t = impl.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20)
            est, stats, _ = t.run(imgs, poses)
            t.close()
            out["orbtraj_%s_poses" % name] = est
            out["orbtraj_%s_stats" % name] = np.ascontiguousarray(stats[:, STAT_COLS])
    finally:
        set_orb(0)
    return out


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import importlib
    import conftest
    conftest.load_pkg()
    sw = importlib.import_module("slam_sdvl_b200.synthworld")
    scenes = importlib.import_module("slam_sdvl_b200.scenes")
    abi = importlib.import_module("slam_sdvl_b200.abi")
    from oracle import ref_py
    with ref_py.strict():
        g = compute(ref_py, sw, scenes, abi)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes,", len(g), "arrays")
    # The same reference code built with its DEFAULT flags (the compiler may contract a*b+c into FMAs): the float
    # Lucas-Kanade loop of Matcher::AlignPatch stops on |update|^2 < 9e-4, so the two builds of the reference itself end
    # a few thousandths of a pixel apart.  The GPU test holds the device to 0.01 px against BOTH.
    gd = compute(ref_py, sw, scenes, abi)
    gd = {k: v for k, v in gd.items() if k.startswith("search_")}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_default_flags.npz")
    np.savez_compressed(path, **gd)
    print("wrote", path, os.path.getsize(path), "bytes,", len(gd), "arrays")
    for k in gd:
        if k.endswith("_px"):
            f = (g[k.replace("_px", "_status")] == abi.MATCH_FOUND) & (gd[k.replace("_px", "_status")] == abi.MATCH_FOUND)
            print(k, "strict vs default build of the reference: max |dpx| =", float(np.abs(g[k][f] - gd[k][f]).max()))
