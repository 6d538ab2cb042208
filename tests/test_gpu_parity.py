"""GPU parity tests: the sm_100a path, called through the C-ABI, against the CPU oracle on identical inputs.
Bars (BASELINE.json): pyramid bytes and FAST corner sets/scores/order bit-exact; H, b, x per Gauss-Newton iteration
within 1e-4 relative (teacher-forced); LK positions within 0.01 px; discrete outputs of SearchPoint identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-4     # north_star tolerance for H, b, x (relative to the infinity norm of the quantity)
LK_PX = 0.01   # north_star tolerance for refined positions


def _ctx(binding, cfg):
    return binding.Context(cfg["params"], cfg["cam"])


@pytest.mark.parametrize("name,seed", [("C1", 3), ("C2", 0), ("C5", 1)])
def test_pyramid_bit_exact(binding, sw, O, name, seed):
    cfg, poses, imgs = sw.sequence(name, seed, 2)
    ctx = _ctx(binding, cfg)
    try:
        for img in imgs:
            f = ctx.frame(img, corners=False)
            ref = O.pyramid(img, cfg["params"].pyramid_levels)
            for l, r in enumerate(ref):
                got = f.level(l)
                assert got.shape == r.shape
                assert np.array_equal(got, r), f"{name} level {l}: {(got != r).sum()} bytes differ"
            f.destroy()
    finally:
        ctx.close()


def test_pyramid_random_and_odd_sizes(binding, abi, O):
    rng = np.random.default_rng(5)
    # word-aligned rows (dp4a path, incl. widths that are not a multiple of 8), unaligned rows (generic path), sizes
    # whose tail levels run from shared memory and sizes too large for that
    for (w, h, levels) in [(101, 99, 3), (94, 60, 2), (47, 30, 2), (640, 480, 5), (333, 250, 4), (752, 480, 5),
                           (376, 240, 4), (188, 120, 3), (72, 40, 3), (20, 16, 2), (1920, 1080, 5), (1284, 724, 5),
                           (2048, 2048, 4),
                           # rolling separable kernel: rows aligned to 8 / 4 bytes only, with 8 and 4 output rows per thread,
                           # partial row blocks, partial column groups, destination rows that are not word-aligned
                           (1000, 450, 3), (1096, 808, 3), (1284, 402, 2), (100, 36, 2), (136, 100, 2), (24, 410, 2)]:
        p = abi.default_params()
        p.pyramid_levels = levels
        p.max_align_level = levels - 1
        p.min_align_level = min(2, levels - 1)
        p.max_fast_levels = min(3, levels)
        cam = abi.Camera(w, h, 300, 300, w / 2, h / 2)
        ctx = binding.Context(p, cam)
        try:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
            f = ctx.frame(img, corners=False)
            for l, r in enumerate(O.pyramid(img, levels)):
                assert np.array_equal(f.level(l), r), f"{w}x{h} level {l}"
            f.destroy()
        finally:
            ctx.close()


def _cmp_corners(f, img, P, O, nfeatures, tag):
    xyl, sc = f.corners()
    rx, rs = O.detect(P, img, nfeatures)
    assert len(xyl) == len(rx), f"{tag}: {len(xyl)} corners vs oracle {len(rx)}"
    same_order = np.array_equal(xyl, rx) and np.array_equal(sc, rs)
    if not same_order:
        a = sorted(map(tuple, np.c_[xyl, sc]))
        b = sorted(map(tuple, np.c_[rx, rs]))
        assert a == b, f"{tag}: corner SETS differ"
        raise AssertionError(f"{tag}: same set, different order")


@pytest.mark.parametrize("name,seed", [("C1", 3), ("C2", 0), ("C5", 1)])
def test_fast_bit_exact(binding, sw, O, name, seed):
    cfg, poses, imgs = sw.sequence(name, seed, 2)
    ctx = _ctx(binding, cfg)
    P = cfg["params"]
    try:
        for k, img in enumerate(imgs):
            f = ctx.frame(img, corners=True)
            _cmp_corners(f, img, P, O, P.num_features, f"{name}[{k}]")
            # Frame::CreateCorners with another budget (first frame uses 2*num_features, sdvl.cc:135)
            f.detect(2 * P.num_features)
            _cmp_corners(f, img, P, O, 2 * P.num_features, f"{name}[{k}] 2x")
            f.detect(150)
            _cmp_corners(f, img, P, O, 150, f"{name}[{k}] 150")
            f.destroy()
    finally:
        ctx.close()


def test_fast_noise_flat_and_ties(binding, abi, O):
    """Dense noise (every cell over quota, many score ties), a flat image (no corners) and a blocky image."""
    rng = np.random.default_rng(11)
    w, h = 752, 480
    p = abi.default_params()
    cam = abi.Camera(w, h, 458.654, 457.296, 367.215, 248.375)
    ctx = binding.Context(p, cam)
    try:
        noise = rng.integers(0, 256, (h, w), dtype=np.uint8)
        flat = np.full((h, w), 77, np.uint8)
        blocks = (np.kron(rng.integers(0, 2, (h // 16, w // 16)), np.ones((16, 16))) * 200 + 20).astype(np.uint8)
        lowc = (noise // 32 * 12).astype(np.uint8)   # few distinct scores -> ties at every retainBest boundary
        for tag, img in [("noise", noise), ("flat", flat), ("blocks", blocks), ("lowc", lowc)]:
            f = ctx.frame(img, corners=True)
            _cmp_corners(f, img, p, O, p.num_features, tag)
            f.destroy()
    finally:
        ctx.close()


def _align_case(binding, sw, scenes, O, name, seed, k_ref, k_cur, perturb):
    cfg, poses, imgs = sw.sequence(name, seed, max(k_ref, k_cur) + 1)
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = O.detect(P, imgs[k_ref], P.num_features)
    pts = scenes.seed_points(cfg, xyl, poses[k_ref], max_points=cfg["n_feat"])
    feats = scenes.align_feats(pts, poses[k_ref], invalid_every=17)
    T_ref = poses[k_ref]
    T_prior = poses[k_ref].copy() if perturb else poses[k_cur].copy()
    return cfg, imgs, pts, feats, T_ref, T_prior


@pytest.mark.parametrize("name,seed,kc,perturb", [("C2", 0, 1, True), ("C2", 7, 2, True), ("C1", 3, 1, True),
                                                  ("C3", 2, 1, True), ("C2", 0, 1, False)])
def test_image_align_teacher_forced(binding, sw, scenes, O, name, seed, kc, perturb):
    cfg, imgs, pts, feats, T_ref, T_prior = _align_case(binding, sw, scenes, O, name, seed, 0, kc, perturb)
    P, cam = cfg["params"], cfg["cam"]
    T_o, nt_o, err_o, tr_o = O.image_align(P, cam, imgs[0], imgs[kc], feats, pts["pos"], T_ref, T_prior)
    assert len(tr_o) >= 3
    ctx = _ctx(binding, cfg)
    try:
        ref = ctx.frame(imgs[0], corners=False)
        cur = ctx.frame(imgs[kc], corners=False)
        iters = np.zeros(8, np.int32)
        for r in tr_o:
            iters[r["level"]] += 1
        T_g, nt_g, err_g, tr_g = ctx.image_align(ref, cur, feats, T_ref, T_prior, forced=(tr_o["T_in"], iters))
        assert len(tr_g) == len(tr_o)
        worst = dict(H=0.0, b=0.0, x=0.0, chi2=0.0)
        for a, b in zip(tr_g, tr_o):
            assert a["level"] == b["level"] and a["iter"] == b["iter"]
            assert a["n_meas"] == b["n_meas"], f"n_meas differs at level {a['level']} iter {a['iter']}"
            for key in ("H", "b", "x"):
                den = np.abs(b[key]).max()
                if den > 0:
                    worst[key] = max(worst[key], np.abs(a[key] - b[key]).max() / den)
            worst["chi2"] = max(worst["chi2"], abs(a["chi2"] - b["chi2"]) / max(abs(b["chi2"]), 1e-30))
        print(f"{name} forced GN parity over {len(tr_o)} iterations: {worst}")
        assert worst["H"] <= REL and worst["b"] <= REL and worst["x"] <= REL and worst["chi2"] <= REL, worst

        # free-running: same pose within GN noise, same tracked count
        T_f, nt_f, err_f, tr_f = ctx.image_align(ref, cur, feats, T_ref, T_prior)
        dC = np.linalg.norm(sw.cam_center(T_f) - sw.cam_center(T_o))
        print(f"{name} free run: iters gpu {len(tr_f)} vs oracle {len(tr_o)}, |dC| = {dC:.3e} m, tracked {nt_f} vs {nt_o}")
        assert dC < 1e-4
        assert abs(nt_f - nt_o) <= 2
        ref.destroy(); cur.destroy()
    finally:
        ctx.close()


def test_image_align_identity_kat(binding, sw, scenes, O):
    """Known answer: aligning a frame against itself from the true pose gives b = 0, x = 0."""
    cfg, poses, imgs = sw.sequence("C2", 4, 1)
    P = cfg["params"]
    xyl, _ = O.detect(P, imgs[0], P.num_features)
    pts = scenes.seed_points(cfg, xyl, poses[0], max_points=200)
    feats = scenes.align_feats(pts, poses[0])
    ctx = _ctx(binding, cfg)
    try:
        a = ctx.frame(imgs[0], corners=False)
        b = ctx.frame(imgs[0], corners=False)
        T, nt, err, tr = ctx.image_align(a, b, feats, poses[0], poses[0])
        assert nt > 100
        assert np.abs(tr[0]["b"]).max() < 1e-6 * max(1.0, np.abs(tr[0]["H"]).max())
        assert np.abs(tr[0]["x"]).max() < 1e-9
        assert np.allclose(T, poses[0], atol=1e-9)
        # no features: returns 0 and leaves the pose (image_align.cc:55-58)
        T2, nt2, _, _ = ctx.image_align(a, b, feats[:0], poses[0], poses[0])
        assert nt2 == 0 and np.array_equal(T2, poses[0])
        a.destroy(); b.destroy()
    finally:
        ctx.close()


@pytest.mark.parametrize("name,seed,fixed", [("C2", 0, True), ("C2", 5, False), ("C1", 3, True), ("C3", 2, True)])
def test_search_points_parity(binding, sw, scenes, abi, O, name, seed, fixed):
    cfg, poses, imgs = sw.sequence(name, seed, 3)
    P, cam = cfg["params"], cfg["cam"]
    xyl, _ = O.detect(P, imgs[0], P.num_features)
    pts = scenes.seed_points(cfg, xyl, poses[0], one_per_cell=False, margin=0)   # includes points that fail the margins
    ctx = _ctx(binding, cfg)
    try:
        ref = ctx.frame(imgs[0], corners=False)
        cur = ctx.frame(imgs[2], corners=True)
        cg = scenes.candidates(pts, poses[0], ref.h, fixed=fixed, project=True, std_frac=0.05 if fixed else 0.02)
        co = cg.copy()
        co["ref_frame"] = 0
        got = ctx.search_points(cur, cg, poses[2])
        exp = O.search_points(P, cam, imgs[2], poses[2], [imgs[0]], co)
        assert np.array_equal(got["status"], exp["status"]), \
            f"status differs for {(got['status'] != exp['status']).sum()} of {len(exp)} candidates"
        assert np.array_equal(got["n_in_range"], exp["n_in_range"])
        assert np.array_equal(got["zmssd"], exp["zmssd"])
        f = exp["status"] == abi.MATCH_FOUND
        assert f.sum() > 50
        assert np.array_equal(got["level"][f], exp["level"][f])
        d = np.abs(got["px"][f] - exp["px"][f]).max()
        seen = exp["status"] != abi.MATCH_UNSEEN
        dp = np.abs(got["proj"][seen] - exp["proj"][seen]).max()
        print(f"{name} fixed={fixed}: {f.sum()} found of {len(exp)}, max |dpx| = {d:.2e}, max |dproj| = {dp:.2e}")
        assert d <= LK_PX
        assert dp <= 1e-9
        ref.destroy(); cur.destroy()
    finally:
        ctx.close()


@pytest.mark.parametrize("name,seed,n_locked", [("C2", 0, 0), ("C2", 3, 120), ("C1", 3, 40), ("C5", 1, 300)])
def test_filter_corners_bit_exact(binding, sw, O, name, seed, n_locked):
    """Frame::FilterCorners (Shi-Tomasi best corner per free cell): same index list as the oracle."""
    cfg, poses, imgs = sw.sequence(name, seed, 2)
    P = cfg["params"]
    rng = np.random.default_rng(seed)
    locked = np.stack([rng.uniform(0, cfg["w"] - 1, n_locked), rng.uniform(0, cfg["h"] - 1, n_locked)], axis=1)
    ctx = binding.Context(P, cfg["cam"])
    f = ctx.frame(imgs[1], corners=True)
    got = f.filter_corners(locked, 50)
    ref = O.filter_corners(P, imgs[1], P.num_features, locked, 50)
    n_cells = int(np.ceil(cfg["w"] / 32) * np.ceil(cfg["h"] / 32))
    print(f"{name}: {len(got)} filtered corners over {n_cells} cells ({n_locked} locked positions)")
    assert len(ref) > 0.3 * (n_cells - n_locked)
    assert np.array_equal(got, ref)
    f.destroy()
    ctx.close()


# ------------------------------------------------------------------ Camera::UndistortImage
EUROC_D = (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0)
TUM1_D = (0.2624, -0.9531, -0.0054, 0.0026, 1.1633)


@pytest.mark.parametrize("name,seed,D", [("C2", 0, EUROC_D), ("C1", 3, TUM1_D), ("C5", 1, (-0.2, 0.05, 0.001, -0.0005, 0.0))])
def test_undistort_bit_exact_and_fused_frame_build(binding, sw, O, name, seed, D):
    """Device Camera::UndistortImage == oracle (== cv2.undistort, tests/test_oracle_cpu.py), and a frame built from a
    distorted image by a context that knows the distortion equals a frame built from the undistorted image."""
    cfg, poses, imgs = sw.sequence(name, seed, 2)
    P, cam = cfg["params"], cfg["cam"]
    exp = [O.undistort(cam, D, im) for im in imgs]
    assert (exp[0] != imgs[0]).mean() > 0.1           # the distortion does something
    plain = _ctx(binding, cfg)
    ctx = _ctx(binding, cfg)
    try:
        # off: a copy, as the reference does (camera.cc:103-104)
        assert np.array_equal(ctx.undistort(imgs[0]), imgs[0])
        ctx.set_distortion(D)
        got = ctx.undistort(imgs[0])
        assert np.array_equal(got, exp[0]), f"{(got != exp[0]).sum()} pixels differ"
        # fused: upload -> undistort -> pyramid -> FAST
        f_d = ctx.frame(imgs[1], corners=True)
        f_u = plain.frame(exp[1], corners=True)
        for l in range(P.pyramid_levels):
            assert np.array_equal(f_d.level(l), f_u.level(l)), f"level {l}"
        xa, sa = f_d.corners()
        xb, sb = f_u.corners()
        assert np.array_equal(xa, xb) and np.array_equal(sa, sb)
        f_d.destroy(); f_u.destroy()
        ctx.set_distortion((0, 0, 0, 0, 0))
        assert np.array_equal(ctx.undistort(imgs[0]), imgs[0])
    finally:
        ctx.close(); plain.close()


@pytest.mark.parametrize("loc", [1, 2])
def test_undistort_in_async_frame_batches(binding, sw, O, loc):
    """sdvlb_frames_submit with distortion set: images in device memory or pinned host memory go through the raw
    scratch ring and the undistortion kernel before the pyramid (12 batches: the ring of 8 sets wraps)."""
    import ctypes as C
    import torch
    cfg, poses, imgs = sw.sequence("C2", 6, 3)
    ctx = _ctx(binding, cfg)
    L = binding.load()
    try:
        ctx.set_distortion(EUROC_D)
        exp = [O.undistort(cfg["cam"], EUROC_D, im) for im in imgs]
        n = len(imgs)
        t = torch.from_numpy(np.ascontiguousarray(imgs))
        t = t.cuda() if loc == 1 else t.pin_memory()
        stride = imgs.shape[1] * imgs.shape[2]
        ptrs = (C.c_void_p * n)(*[t.data_ptr() + i * stride for i in range(n)])
        for rep in range(12):
            out = (C.c_void_p * n)()
            assert L.sdvlb_frames_submit(C.c_void_p(ctx.h), ptrs, n, loc, 1, cfg["params"].num_features, out) == 0
            assert L.sdvlb_frames_wait(C.c_void_p(ctx.h), out, n) == 0
            for i in range(n):
                f = binding.Frame(ctx, out[i])
                assert np.array_equal(f.level(0), exp[i]), f"batch {rep} frame {i}"
                f.destroy()
    finally:
        ctx.close()


def test_host_camera_undistort_image(binding, sw, O):
    """sdvl::Camera::UndistortImage (C++ host mirror) == oracle == cv2.undistort; without distortion it clones."""
    cfg, poses, imgs = sw.sequence("C2", 2, 1)
    got = binding.host_camera_undistort(cfg["params"], cfg["cam"], EUROC_D, imgs[0])
    assert np.array_equal(got, O.undistort(cfg["cam"], EUROC_D, imgs[0]))
    assert np.array_equal(binding.host_camera_undistort(cfg["params"], cfg["cam"], (0, 0, 0, 0, 0), imgs[0]), imgs[0])
