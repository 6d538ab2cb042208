"""GPU parity of the ORB descriptor mode (Config::UseORB()) against the CPU oracle, whose
ORB restatement is pinned bit for bit against the reference's own extra/orb_detector.cc / matcher.cc
(tests/test_oracle_vs_ref.py::test_orb_mode_vs_reference) and against OpenCV (fastAtan2, the learned pattern).
Bars: FAST corner lists with the ORB margin, orientations and descriptors bit-exact; SearchPoint decisions, levels and
descriptor distances identical, refined positions within 0.01 px.  (File name: runs after every other GPU test.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def orb_oracle(O):
    O.lib().orc_set_orb(1)
    yield O
    O.lib().orc_set_orb(0)


@pytest.mark.parametrize("name,seed", [("C2", 0), ("C1", 3), ("C5", 1)])
def test_orb_corners_descriptors_and_search(binding, sw, scenes, abi, orb_oracle, name, seed):
    O = orb_oracle
    cfg, poses, imgs = sw.sequence(name, seed, 4)
    P, cam = cfg["params"], cfg["cam"]
    h, w = imgs[0].shape
    ctx = binding.Context(P, cam)
    try:
        ctx.set_orb(True)
        ref = ctx.frame(imgs[0], corners=True)
        cur = ctx.frame(imgs[3], corners=True)
        # FAST with the ORB border margin (extra/fast_detector.cc:63-64), FilterCorners likewise (:183-184)
        xyl, sc = ref.corners()
        xo, so = O.detect(P, imgs[0], P.num_features)
        assert np.array_equal(xyl, xo) and np.array_equal(sc, so)
        assert xyl[:, 0].min() >= 19 and xyl[:, 1].min() >= 19
        assert np.array_equal(ref.filter_corners(np.zeros((0, 2))), O.filter_corners(P, imgs[0], P.num_features, np.zeros((0, 2))))
        # orientation + descriptor at every corner
        d_g, a_g = ref.orb_descriptors(xyl)
        d_o, a_o = np.zeros((len(xyl), 32), np.uint8), np.zeros(len(xyl), np.float32)
        xc = np.ascontiguousarray(xyl)
        assert O.lib().orc_orb_descriptors(C.byref(P), O.ptr(imgs[0]), w, h, O.ptr(xc), len(xc), O.ptr(d_o), O.ptr(a_o)) == 0
        same = (d_g == d_o).all(1)
        print(f"{name}: {len(xyl)} corners, orientations identical {np.array_equal(a_g, a_o)}, descriptors identical {same.mean():.4%}")
        assert np.array_equal(a_g, a_o)
        assert np.array_equal(d_g, d_o)
        # Matcher::SearchPoint scored by descriptor distance
        pts = scenes.seed_points(cfg, xyl, poses[0], max_points=500, one_per_cell=False, margin=20)
        pos = np.concatenate([(pts["px"] / (1 << pts["level"])[:, None]).astype(np.int32), pts["level"][:, None].astype(np.int32)], axis=1)
        q_desc, _ = ref.orb_descriptors(pos)
        for fixed, std_frac in ((True, 0.05), (False, 0.5)):
            cg = scenes.candidates(pts, poses[0], ref.h, fixed=fixed, project=True, std_frac=std_frac)
            co = cg.copy()
            co["ref_frame"] = 0
            got = ctx.search_points_orb(cur, cg, poses[3], q_desc)
            exp = O.search_points(P, cam, imgs[3], poses[3], [imgs[0]], co)
            assert np.array_equal(got["status"], exp["status"]), \
                f"status differs for {(got['status'] != exp['status']).sum()} of {len(exp)} candidates"
            assert np.array_equal(got["n_in_range"], exp["n_in_range"])
            assert np.array_equal(got["zmssd"], exp["zmssd"])          # descriptor distance of the best corner
            f = exp["status"] == abi.MATCH_FOUND
            assert f.sum() > 100
            assert np.array_equal(got["level"][f], exp["level"][f])
            d = np.abs(got["px"][f] - exp["px"][f]).max()
            print(f"{name} fixed={fixed}: {f.sum()} found of {len(exp)}, max |dpx| = {d:.2e}")
            assert d <= 0.01
        ref.destroy(); cur.destroy()
        # back to the ZMSSD mode: the margin of 5 returns
        ctx.set_orb(False)
        f2 = ctx.frame(imgs[0], corners=True)
        O.lib().orc_set_orb(0)
        x2, _ = f2.corners()
        xo2, _ = O.detect(P, imgs[0], P.num_features)
        O.lib().orc_set_orb(1)
        assert np.array_equal(x2, xo2) and x2[:, 0].min() < 19
        f2.destroy()
    finally:
        ctx.close()


def test_orb_descriptor_limits(binding, sw):
    """Positions closer than 19 px to the border of their level: SDVLB_ERR_ARG (the reference asserts), zero bytes."""
    cfg, poses, imgs = sw.sequence("C2", 0, 1)
    ctx = binding.Context(cfg["params"], cfg["cam"])
    try:
        f = ctx.frame(imgs[0], corners=False)
        with pytest.raises(Exception):
            f.orb_descriptors(np.array([[18, 100, 0]], np.int32))
        d, a = f.orb_descriptors(np.array([[19, 19, 0], [752 - 20, 480 - 20, 0], [30, 25, 2]], np.int32))
        assert d.any() and (a >= 0).all() and (a < 360).all()
        f.destroy()
    finally:
        ctx.close()
