"""Map::UpdateCandidates (map.cc:397-498) on the device vs the CPU oracle: depth-filter seeds searched along their
epipolar segments, triangulated and updated, frame after frame."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_seeds(cfg, scenes, abi, xyl, T_ref, ref_handle, depth_mean, seed=0):
    """Candidates the way Map::InitCandidates leaves them (Point::InitCandidate, point.cc:48-61): a = b = 10,
    sigma2 = 1, z_range = 6, rho = 1 / (a first triangulated depth: here truth +-10 %)."""
    pts = scenes.seed_points(cfg, xyl, T_ref, one_per_cell=False, margin=0)
    n = len(pts["px"])
    rng = np.random.default_rng(seed)
    s = np.zeros(n, abi.SEED_DT)
    s["ref_frame"] = ref_handle
    s["ref_T"] = np.asarray(T_ref)
    s["ref_px"] = pts["px"]
    s["ref_v"] = pts["v"]
    s["rho"] = 1.0 / (pts["depth"] * (1.0 + rng.uniform(-0.1, 0.1, n)))
    s["sigma2"] = 1.0
    s["a"] = 10.0
    s["b"] = 10.0
    s["z_range"] = 6.0
    s["cos_alpha"] = 1.0
    s["last_distance"] = 1.0 / s["rho"]
    s["ref_level"] = pts["level"]
    s["last_kf_id"] = 0
    return s, pts


LIVE = lambda abi: (abi.SEED_NOT_VISIBLE, abi.SEED_SHORT_BASELINE, abi.SEED_NOT_FOUND, abi.SEED_NO_DEPTH,
                    abi.SEED_NO_PARALLAX, abi.SEED_TOO_CLOSE, abi.SEED_UPDATED)


@pytest.mark.parametrize("name,seed,step", [("C2", 0, 3), ("C1", 3, 3), ("C3", 2, 1)])
def test_update_candidates_vs_oracle(binding, sw, scenes, abi, O, name, seed, step):
    cfg, poses, imgs = sw.sequence(name, seed, 1 + 8 * step)
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = O.detect(P, imgs[0], P.num_features)
    depth_mean = float(np.median(scenes.seed_points(cfg, xyl, poses[0])["depth"]))
    ctx = binding.Context(P, cam)
    try:
        ref = ctx.frame(imgs[0], corners=False)
        s0, pts = make_seeds(cfg, scenes, abi, xyl, poses[0], ref.h, depth_mean, seed)
        assert len(s0) > 400
        so = s0.copy()
        so["ref_frame"] = 0
        sg_free = s0.copy()
        live_o = np.ones(len(so), bool)
        live_g = np.ones(len(so), bool)
        worst = dict(rho=0.0, sigma2=0.0, a=0.0, b=0.0, px=0.0)
        n_status = n_cmp = 0
        for k in range(step, len(imgs), step):
            cur = ctx.frame(imgs[k], corners=True)
            # teacher forcing: the device starts every frame from the oracle's state
            sg_in = so.copy()
            sg_in["ref_frame"] = ref.h
            got = ctx.update_candidates(cur, poses[k], sg_in[live_o], depth_mean)
            so[live_o] = O.update_candidates(P, cam, imgs[k], poses[k], [imgs[0]], so[live_o], depth_mean)
            exp = so[live_o]
            same = got["status"] == exp["status"]
            n_status += int((~same).sum())
            n_cmp += len(exp)
            upd = same & (exp["status"] >= abi.SEED_UPDATED)
            for key in ("rho", "sigma2", "a", "b"):
                if upd.any():
                    worst[key] = max(worst[key], float(np.max(np.abs(got[key][upd] - exp[key][upd]) /
                                                              np.maximum(np.abs(exp[key][upd]), 1e-12))))
            found = same & (exp["status"] >= abi.SEED_NO_DEPTH)
            if found.any():
                worst["px"] = max(worst["px"], float(np.abs(got["px"][found] - exp["px"][found]).max()))
            nf = same & np.isin(exp["status"], (abi.SEED_NOT_FOUND, abi.SEED_DELETE_FAILED))
            assert np.array_equal(got["n_failed"][nf], exp["n_failed"][nf]) and np.array_equal(got["b"][nf], exp["b"][nf])
            # free run of the device, its own state from frame to frame
            sg_free[live_g] = ctx.update_candidates(cur, poses[k], sg_free[live_g], depth_mean)
            live_o &= np.isin(so["status"], LIVE(abi))
            live_g &= np.isin(sg_free["status"], LIVE(abi))
            cur.destroy()
        conv_o = so["status"] == abi.SEED_CONVERGED
        conv_g = sg_free["status"] == abi.SEED_CONVERGED
        print(f"{name}: {len(so)} seeds, {n_status} of {n_cmp} statuses differ, worst relative {worst}, "
              f"converged oracle {conv_o.sum()} device {conv_g.sum()}")
        assert n_status <= max(2, n_cmp // 500)
        assert worst["px"] <= 0.01                       # FeatureAlign / Matcher bar of BASELINE.json
        assert worst["rho"] <= 1e-4 and worst["sigma2"] <= 1e-3 and worst["a"] <= 1e-3 and worst["b"] <= 1e-3
        assert conv_o.sum() >= 20
        assert abs(int(conv_o.sum()) - int(conv_g.sum())) <= max(3, len(so) // 100)
        both = conv_o & conv_g
        assert np.abs(so["p3d"][both] - sg_free["p3d"][both]).max() < 1e-3      # 1 mm
        # converged points lie on the plane z = 0 of the synthetic world
        assert np.median(np.abs(sg_free["p3d"][conv_g][:, 2])) < 0.02
        ref.destroy()
    finally:
        ctx.close()


def test_update_candidates_edge_cases(binding, sw, scenes, abi, O):
    cfg, poses, imgs = sw.sequence("C2", 1, 7)
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = O.detect(P, imgs[0], P.num_features)
    ctx = binding.Context(P, cam)
    try:
        ref = ctx.frame(imgs[0], corners=False)
        cur = ctx.frame(imgs[6], corners=True)
        s, _ = make_seeds(cfg, scenes, abi, xyl, poses[0], ref.h, 2.0)
        s = s[:64].copy()
        # empty batch
        assert len(ctx.update_candidates(cur, poses[6], s[:0], 2.0)) == 0
        # behind the camera / far outside the image: not visible; old ones are deleted
        s["rho"][:8] = -0.5
        s["last_kf_id"][:4] = -50
        # same pose as the reference: baseline too short
        got0 = ctx.update_candidates(cur, poses[0], s, 2.0, min_kf_id=-10)
        so = s.copy(); so["ref_frame"] = 0
        exp0 = O.update_candidates(P, cam, imgs[6], poses[0], [imgs[0]], so, 2.0, min_kf_id=-10)
        assert np.array_equal(got0["status"], exp0["status"])
        assert (got0["status"][:4] == abi.SEED_DELETE_OLD).all() and (got0["status"][4:8] == abi.SEED_NOT_VISIBLE).all()
        assert (got0["status"][8:] == abi.SEED_SHORT_BASELINE).all()
        # a failing seed is deleted once n_failed exceeds max_failed
        s2 = s[8:16].copy()
        s2["ref_px"] = [[3.0, 3.0]] * 8          # reference patch leaves the image: SearchPoint fails
        s2["n_failed"] = P.max_failed
        got = ctx.update_candidates(cur, poses[6], s2, 2.0)
        s2o = s2.copy(); s2o["ref_frame"] = 0
        exp = O.update_candidates(P, cam, imgs[6], poses[6], [imgs[0]], s2o, 2.0)
        assert np.array_equal(got["status"], exp["status"])
        vis = exp["status"] != abi.SEED_NOT_VISIBLE
        assert (got["status"][vis] == abi.SEED_DELETE_FAILED).all() and vis.any()
        assert np.array_equal(got["b"], exp["b"]) and np.array_equal(got["n_failed"], exp["n_failed"])
        ref.destroy(); cur.destroy()
    finally:
        ctx.close()


def test_host_map_update_candidates_equals_c_abi(binding, sw, scenes, abi, O):
    """sdvl::Map::UpdateCandidates (C++ host mirror: one device call per frame + the reference's list surgery) ends in
    the same candidate states as driving sdvlb_update_candidates by hand."""
    cfg, poses, imgs = sw.sequence("C2", 4, 19)
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = O.detect(P, imgs[0], P.num_features)
    depth_mean = float(np.median(scenes.seed_points(cfg, xyl, poses[0])["depth"]))
    ks = list(range(3, 19, 3))
    ctx = binding.Context(P, cam)
    try:
        ref = ctx.frame(imgs[0], corners=False)
        s0, _ = make_seeds(cfg, scenes, abi, xyl, poses[0], ref.h, depth_mean, 4)
        s0 = s0[:300].copy()
        s0["last_kf_id"][:5] = -100          # will be deleted as too old once they leave the image
        s = s0.copy()
        s["status"] = -1
        live = np.ones(len(s), bool)
        for k in ks:
            cur = ctx.frame(imgs[k], corners=True)
            s[live] = ctx.update_candidates(cur, poses[k], s[live], depth_mean, min_kf_id=-10)
            live &= ~np.isin(s["status"], (abi.SEED_CONVERGED, abi.SEED_DELETE_OLD))
            # map.cc:449-452: a candidate that failed too often is handed to DeletePoint; the trash is emptied after
            # the pass and the next pass erases it
            live &= s["status"] != abi.SEED_DELETE_FAILED
            cur.destroy()
        ref.destroy()
    finally:
        ctx.close()
    hs, n_left = binding.host_map_update_candidates(P, cam, imgs[0], poses[0], [imgs[k] for k in ks],
                                                    [poses[k] for k in ks], s0, depth_mean, min_kf_id=-10)
    assert np.array_equal(hs["status"], s["status"])
    assert n_left == int(live.sum())
    for key in ("rho", "sigma2", "a", "b", "cos_alpha", "last_distance"):
        np.testing.assert_allclose(hs[key], s[key], rtol=1e-9, atol=0)
    assert np.array_equal(hs["n_failed"], s["n_failed"])
    print(f"host Map::UpdateCandidates: {len(s)} candidates, {n_left} left after {len(ks)} frames, "
          f"{(s['status'] == abi.SEED_CONVERGED).sum()} converged")


def _init_seeds(cfg, abi, xyl, idx, T_new, handle, depth_mean):
    s = np.zeros(len(idx), abi.SEED_DT)
    cam = cfg["cam"]
    c = xyl[idx].astype(np.float64)
    px = c[:, :2] * (1 << xyl[idx, 2])[:, None]
    v = np.stack([(px[:, 0] - cam.u0) / cam.fx, (px[:, 1] - cam.v0) / cam.fy, np.ones(len(idx))], axis=1)
    v /= np.sqrt((v * v).sum(1))[:, None]
    s["ref_frame"] = handle
    s["ref_T"] = np.asarray(T_new)
    s["ref_px"] = px
    s["ref_v"] = v
    s["rho"] = 1.0 / depth_mean
    s["sigma2"] = 1.0
    s["a"] = 10; s["b"] = 10; s["z_range"] = 6; s["cos_alpha"] = 1; s["last_distance"] = depth_mean
    s["ref_level"] = xyl[idx, 2]
    return s


@pytest.mark.parametrize("name,seed,orb", [("C2", 2, False), ("C3", 1, False), ("C2", 2, True)])
def test_init_candidates_vs_oracle_and_host_map(binding, sw, scenes, abi, O, name, seed, orb):
    """The per-corner part of Map::InitCandidates (SDVLB_SEEDS_INIT) against the oracle, and the host mirror's
    InitCandidates -> UpdateCandidates (every candidate listed twice, as the reference does) -> AddConnectionsPoints.
    orb: the same with Config::UseORB() (descriptors of the filtered corners saved into the new features, map.cc:319-323;
    every SearchPoint scored by descriptor distance)."""
    cfg, poses, imgs = sw.sequence(name, seed, 28)
    P, cam = cfg["params"], cfg["cam"]
    k_old, k_new = 0, 9
    ctx = binding.Context(P, cam)
    O.lib().orc_set_orb(int(orb))
    binding.load_host().sdvlh_config_set_orb(int(orb))
    try:
        if orb:
            ctx.set_orb(True)
        kf_new = ctx.frame(imgs[k_new], corners=True)
        kf_old = ctx.frame(imgs[k_old], corners=True)
        xyl, _ = kf_new.corners()
        depth_mean = float(np.median(scenes.seed_points(cfg, xyl, poses[k_new])["depth"]))
        idx = kf_new.filter_corners(np.zeros((0, 2)))
        assert len(idx) > 150
        s0 = _init_seeds(cfg, abi, xyl, idx, poses[k_new], kf_new.h, depth_mean)
        got = ctx.update_candidates(kf_old, poses[k_old], s0, depth_mean, mode=abi.SEEDS_INIT)
        so = s0.copy(); so["ref_frame"] = 0
        exp = O.update_candidates(P, cam, imgs[k_old], poses[k_old], [imgs[k_new]], so, depth_mean, mode=abi.SEEDS_INIT)
        assert np.array_equal(got["status"], exp["status"])
        ok = exp["status"] == abi.SEED_UPDATED
        found = exp["status"] >= abi.SEED_NO_DEPTH
        assert ok.sum() > 50
        assert np.array_equal(got["level"][found], exp["level"][found])
        assert np.abs(got["px"][found] - exp["px"][found]).max() <= 0.01
        np.testing.assert_allclose(got["depth"][ok], exp["depth"][ok], rtol=1e-4)
        # nothing but status / depth / px / level is touched in this mode
        for key in ("rho", "sigma2", "a", "b", "n_failed"):
            assert np.array_equal(got[key], s0[key])
        # triangulated depths are the plane's
        true_depth = scenes.seed_points(cfg, xyl[idx][ok], poses[k_new], one_per_cell=False, margin=0)
        print(f"{name}: {len(idx)} filtered corners, {found.sum()} found in the old keyframe, {ok.sum()} initialisable")

        # host mirror: same candidates, then the depth filter on the following frames
        ks = list(range(k_new + 3, 27, 3))
        res = binding.host_map_init_candidates(P, cam, imgs[k_new], poses[k_new], imgs[k_old], poses[k_old],
                                               [imgs[k] for k in ks], [poses[k] for k in ks], imgs[27], poses[27], depth_mean)
        assert res["created"] == int(ok.sum()) and res["listed"] == 2 * res["created"]
        np.testing.assert_allclose(res["px"], got["ref_px"][ok], rtol=0, atol=1e-9)
        # replay of the reference's walk: every candidate is listed twice, so each frame updates it twice in a row
        st = got[ok].copy()
        st["rho"] = 1.0 / st["depth"]
        st["last_distance"] = st["depth"]
        fixed = np.zeros(len(st), bool)
        alive = np.ones((len(st), 2), bool)          # the two list entries of every candidate
        for k in ks:
            cur = ctx.frame(imgs[k], corners=True)
            n_entries = alive.sum(1)
            first = np.where(alive[:, 0], 0, 1)       # which entry comes first in the list
            for occ in range(2):
                m = n_entries > occ                   # k-th occurrences form the k-th batch
                entry = first if occ == 0 else np.ones(len(st), int)
                was_fixed = fixed.copy()
                st[m] = ctx.update_candidates(cur, poses[k], st[m], depth_mean)
                conv = m & (st["status"] == abi.SEED_CONVERGED)
                gone = conv | (m & was_fixed & (st["status"] == abi.SEED_UPDATED)) | (m & (st["status"] == abi.SEED_DELETE_OLD))
                fixed |= conv
                alive[np.where(gone)[0], entry[gone]] = False
            cur.destroy()
        listed = alive.sum(1)
        assert res["fixed"] == int(fixed.sum())
        assert res["left"] == int(listed.sum())
        np.testing.assert_allclose(res["rho"], st["rho"], rtol=1e-9)
        assert res["linked"] > 0.5 * res["fixed"]
        print(f"{name}: host Map created {res['created']} candidates, {res['fixed']} converged, {res['left']} entries left, "
              f"{res['linked']} linked into the last frame")
        kf_new.destroy(); kf_old.destroy()
    finally:
        ctx.close()
        O.lib().orc_set_orb(0)
        binding.load_host().sdvlh_config_set_orb(0)
