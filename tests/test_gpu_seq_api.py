"""Resident sequences through the C-ABI itself (sdvlb_seq_*): capacities, the submission queue, the device-side
Map::NeedKeyframe hold and SDVL::CalcTrackingQuality, several contexts / devices in one process."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _seed_points(abi, sw, cfg, frame, T, max_points, first_id=0, occupied=None):
    """Fixed map points for the corners of `frame` (pose T, world->camera) on the synthetic plane z = 0: one per free
    32-px cell, like the bench harness (host/tracker.cc:SeedResident)."""
    cam = cfg["cam"]
    xyl, _ = frame.corners()
    R = sw.quat_R(T[:4])
    C = sw.cam_center(T)
    gw = int(np.ceil(cam.width / 32))
    occ = set() if occupied is None else set(occupied)
    pts = []
    n = len(xyl)
    for i in range(n):
        x, y, l = (int(v) for v in xyl[(i * 7919) % n])
        lw, lh = int(cam.width) >> l, int(cam.height) >> l
        if x < 6 or y < 6 or x >= lw - 6 or y >= lh - 6:
            continue
        px = np.array([x * (1 << l), y * (1 << l)], float)
        cell = int(px[1] // 32) * gw + int(px[0] // 32)
        if cell in occ:
            continue
        v = np.array([(px[0] - cam.u0) / cam.fx, (px[1] - cam.v0) / cam.fy, 1.0])
        d = R.T @ (v / np.linalg.norm(v))
        if abs(d[2]) < 1e-9:
            continue
        s = -C[2] / d[2]
        if s <= 0:
            continue
        p3d = C + d * s
        rho = 1.0 / np.linalg.norm(p3d - C)
        pts.append((p3d, px, px, rho, 0.05 * rho, first_id + len(pts), l, l, abi.CAND_FIXED, 0, 0, 0))
        occ.add(cell)
        if len(pts) >= max_points:
            break
    return np.array(pts, abi.SEQ_POINT_DT)


def _occupied(feats, cam):
    gw = int(np.ceil(cam.width / 32))
    return {int(f["px"][1] // 32) * gw + int(f["px"][0] // 32) for f in feats if f["flags"] & 1}


def _track(ctx, seqs, frames):
    ctx.seq_submit(seqs, frames)
    return ctx.seq_collect()


def test_mixed_capacities_in_one_submission(binding, abi, sw):
    """Sequences with different feature capacities in ONE submission (each carves its own scratch by its own capacity):
    same results as tracking them one by one."""
    cfg, poses, imgs = sw.sequence("C2", 4, 4)
    out = {}
    for together in (True, False):
        ctx = binding.Context(cfg["params"], cfg["cam"])
        seqs = [binding.Sequence(ctx, 256), binding.Sequence(ctx, 1024)]
        f0 = [ctx.frame(imgs[0]) for _ in seqs]
        for s, f in zip(seqs, f0):
            s.reset(f, poses[0])
            s.add_points(f, poses[0], _seed_points(abi, sw, cfg, f, poses[0], 200))
        res = []
        for k in range(1, 4):
            fr = [ctx.frame(imgs[k]) for _ in seqs]
            if together:
                r = _track(ctx, seqs, fr)
            else:
                r = [_track(ctx, [seqs[0]], [fr[0]])[0], _track(ctx, [seqs[1]], [fr[1]])[0]]
            res.append(r)
        out[together] = res
        ctx.close()
    for ra, rb in zip(out[True], out[False]):
        for a, b in zip(ra, rb):
            assert a["status"] == abi.SEQ_TRACKED and a["matches"] > 50
            assert np.array_equal(a["pose"], b["pose"]) and a["matches"] == b["matches"] and a["inliers"] == b["inliers"]
    # both capacities saw the same inputs: identical tracks
    assert np.array_equal(out[True][-1][0]["pose"], out[True][-1][1]["pose"])


def test_feature_capacity_overflow_is_reported(binding, abi, sw):
    """Appending more points than the sequence can hold: the device truncates and every later result of the sequence
    carries the error (SDVLB_ERR_OVERFLOW at collect) until the next reset."""
    cfg, poses, imgs = sw.sequence("C2", 4, 3)
    ctx = binding.Context(cfg["params"], cfg["cam"])
    s = binding.Sequence(ctx, 64)          # capacity 64 features
    f0 = ctx.frame(imgs[0])
    s.reset(f0, poses[0])
    pts = _seed_points(abi, sw, cfg, f0, poses[0], 60)
    s.add_points(f0, poses[0], pts[:40])
    s.add_points(f0, poses[0], pts[:40])     # 80 > 64: second batch is cut short on the device
    with pytest.raises(binding.SdvlbError, match="capacity"):
        _track(ctx, [s], [ctx.frame(imgs[1])])
    with pytest.raises(binding.SdvlbError):
        s.add_points(f0, poses[0], np.zeros(65, abi.SEQ_POINT_DT))   # more than the capacity in one call: refused at once
    s.reset(f0, poses[0])                    # a reset clears the condition
    s.add_points(f0, poses[0], pts[:40])
    r = _track(ctx, [s], [ctx.frame(imgs[1])])[0]
    assert r["status"] == abi.SEQ_TRACKED and r["matches"] > 10
    ctx.close()


def test_keyframe_slots_exhausted(binding, abi, sw):
    """SDVLB_SEQ_KF_CAP keyframes may be referenced by live points; one more is refused with a status, slots come back
    when a result reports that nothing references them any more."""
    cfg, poses, imgs = sw.sequence("C2", 4, 3)
    ctx = binding.Context(cfg["params"], cfg["cam"])
    s = binding.Sequence(ctx, 512)
    f0 = ctx.frame(imgs[0])
    s.reset(f0, poses[0])
    pts = _seed_points(abi, sw, cfg, f0, poses[0], 128)
    slots = [s.add_points(f0, poses[0], pts[2 * k:2 * k + 2]) for k in range(abi.SEQ_KF_CAP)]
    assert sorted(slots) == list(range(abi.SEQ_KF_CAP))
    with pytest.raises(binding.SdvlbError, match="keyframe slot"):
        s.add_points(f0, poses[0], pts[:1])
    r = _track(ctx, [s], [ctx.frame(imgs[1])])[0]
    assert r["status"] == abi.SEQ_TRACKED
    live = int((r["kf_live"] > 0).sum())
    assert 0 < live <= abi.SEQ_KF_CAP
    if live < abi.SEQ_KF_CAP:               # slots whose two points were both lost are free again
        s.add_points(f0, poses[0], pts[:1])
    ctx.close()


def test_submission_queue_depth_and_order(binding, abi, sw):
    """Up to SDVLB_SEQ_DEPTH submissions in flight; results come back oldest first and equal a one-at-a-time run."""
    cfg, poses, imgs = sw.sequence("C2", 6, 2 + abi.SEQ_DEPTH)
    runs = []
    for queued in (False, True):
        ctx = binding.Context(cfg["params"], cfg["cam"])
        s = binding.Sequence(ctx, 512)
        f0 = ctx.frame(imgs[0])
        s.reset(f0, poses[0])
        s.add_points(f0, poses[0], _seed_points(abi, sw, cfg, f0, poses[0], 200))
        frames = [ctx.frame(imgs[k]) for k in range(1, 1 + abi.SEQ_DEPTH)]
        res = []
        if queued:
            for f in frames:
                ctx.seq_submit([s], [f])
            assert ctx.seq_inflight() == abi.SEQ_DEPTH
            with pytest.raises(binding.SdvlbError, match="in flight"):
                ctx.seq_submit([s], [frames[0]])
            for _ in frames:
                res.append(ctx.seq_collect()[0])
            assert ctx.seq_inflight() == 0
        else:
            for f in frames:
                res.append(_track(ctx, [s], [f])[0])
        runs.append(res)
        ctx.close()
    for a, b in zip(*runs):
        assert a["status"] == b["status"] == abi.SEQ_TRACKED
        assert np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["feats"], b["feats"])


def test_keyframe_rule_holds_the_sequence(binding, abi, sw):
    """policy.keyframe_rule (Map::NeedKeyframe on the device): when it fires the result says so, steps queued behind it
    come back HELD without consuming their frame, and the sequence continues once the caller answers (new points, or a
    plain release)."""
    cfg, poses, imgs = sw.sequence("C2", 6, 6)
    ctx = binding.Context(cfg["params"], cfg["cam"])
    s = binding.Sequence(ctx, 512)
    # lost_ratio 2.0: "npoints < 2 * last_matches" is true as soon as last_matches > 0, i.e. from the second frame on
    s.set_policy(keyframe_rule=1, min_keyframe_its=1000, lost_ratio=2.0)
    f0 = ctx.frame(imgs[0])
    s.reset(f0, poses[0])
    s.add_points(f0, poses[0], _seed_points(abi, sw, cfg, f0, poses[0], 200))
    fr = [ctx.frame(imgs[k]) for k in range(6)]
    r1 = _track(ctx, [s], [fr[1]])[0]
    assert r1["status"] == abi.SEQ_TRACKED and r1["need_keyframe"] == 0   # last_matches was 0: nothing lost yet
    ctx.seq_submit([s], [fr[2]])
    ctx.seq_submit([s], [fr[3]])          # queued behind frame 2, which raises the hold
    r2, r3 = ctx.seq_collect()[0], ctx.seq_collect()[0]
    assert r2["status"] == abi.SEQ_TRACKED and r2["need_keyframe"] == 1
    assert r3["status"] == abi.SEQ_HELD
    r3b = _track(ctx, [s], [fr[3]])[0]    # still held: nobody answered
    assert r3b["status"] == abi.SEQ_HELD
    # answer with new points seeded on frame 2 (what the mapping thread would do), then frame 3 again
    new = _seed_points(abi, sw, cfg, fr[2], r2["pose"], 200 - int((r2["feats"]["flags"] & 1).sum()), first_id=1000,
                       occupied=_occupied(r2["feats"], cfg["cam"]))
    s.add_points(fr[2], r2["pose"], new)
    r3c = _track(ctx, [s], [fr[3]])[0]
    assert r3c["status"] == abi.SEQ_TRACKED and r3c["matches"] > r2["inliers"] * 0.8
    assert r3c["need_keyframe"] == 1      # the rule keeps firing with this policy
    s.release()                           # a plain release answers it too
    r4 = _track(ctx, [s], [fr[4]])[0]
    assert r4["status"] == abi.SEQ_TRACKED
    d = np.linalg.norm(sw.cam_center(r4["pose"]) - sw.cam_center(poses[4]))
    assert d < 5e-3, f"trajectory lost after the hold: {d}"
    ctx.close()


def test_tracking_quality_keeps_the_reference_frame(binding, abi, sw):
    """policy.tracking_quality (SDVL::CalcTrackingQuality, sdvl.cc:240-264 and :99,119): a frame that tracks badly is
    not adopted as the next alignment reference, lost_frames counts up and the third loss holds the sequence for
    relocalisation.  Without the policy the blank frame becomes the reference and the track is gone."""
    cfg, poses, imgs = sw.sequence("C2", 7, 6)
    blank = np.full_like(imgs[0], 128)
    outcomes = {}
    for quality in (1, 0):
        ctx = binding.Context(cfg["params"], cfg["cam"])
        s = binding.Sequence(ctx, 512)
        s.set_policy(tracking_quality=quality)
        f0 = ctx.frame(imgs[0])
        s.reset(f0, poses[0])
        s.add_points(f0, poses[0], _seed_points(abi, sw, cfg, f0, poses[0], 200))
        r1 = _track(ctx, [s], [ctx.frame(imgs[1])])[0]
        rb = _track(ctx, [s], [ctx.frame(blank)])[0]          # nothing to match in a flat image
        r2 = _track(ctx, [s], [ctx.frame(imgs[2])])[0]
        outcomes[quality] = (r1, rb, r2)
        if quality:
            assert r1["quality"] == abi.TRACKING_GOOD and r1["lost_frames"] == 0
            assert rb["matches"] < cfg["params"].min_matches and rb["quality"] == abi.TRACKING_BAD and rb["lost_frames"] == 1
            assert r2["quality"] == abi.TRACKING_GOOD and r2["lost_frames"] == 0
            assert r2["matches"] > 0.5 * r1["matches"], "the reference frame was not kept"
            d = np.linalg.norm(sw.cam_center(r2["pose"]) - sw.cam_center(poses[2]))
            assert d < 5e-3
            # three bad frames in a row: held for relocalisation
            for k in range(3):
                rk = _track(ctx, [s], [ctx.frame(blank)])[0]
                assert rk["status"] == abi.SEQ_TRACKED and rk["lost_frames"] == k + 1
            assert _track(ctx, [s], [ctx.frame(imgs[3])])[0]["status"] == abi.SEQ_HELD
        else:
            assert rb["quality"] == abi.TRACKING_GOOD          # not evaluated
            assert r2["matches"] == 0, "without the policy the blank frame is the new reference: nothing left to track"
        ctx.close()


def _short_track(binding, abi, sw, cfg, poses, imgs, device):
    import ctypes as C
    ctx = binding.Context(cfg["params"], cfg["cam"], device=device)
    s = binding.Sequence(ctx, 512)
    f0 = ctx.frame(imgs[0])
    s.reset(f0, poses[0])
    s.add_points(f0, poses[0], _seed_points(abi, sw, cfg, f0, poses[0], 200))
    r = [_track(ctx, [s], [ctx.frame(imgs[k])])[0] for k in (1, 2)]
    # the standalone pose-refinement kernel too (its dynamic shared memory depends on the observation count)
    rng = abi.Rand()
    binding.load().sdvlb_rand_seed(C.byref(rng), 3)
    obs = np.zeros(300, abi.POSE_OBS_DT)
    obs["v"] = [0.0, 0.0, 1.0]
    obs["pos"] = np.random.default_rng(0).normal(size=(300, 3)) + [0, 0, 5]
    ctx.select_inliers(obs, poses[0], rng)
    ctx.close()
    return r


def test_second_context_in_the_same_process(binding, abi, sw):
    """Kernel attributes (dynamic shared-memory opt-in, carve-out) are recorded per (device, kernel): a context created
    after another one was destroyed runs every kernel of the path and gives the same results."""
    cfg, poses, imgs = sw.sequence("C2", 4, 3)
    a = _short_track(binding, abi, sw, cfg, poses, imgs, 0)
    b = _short_track(binding, abi, sw, cfg, poses, imgs, 0)
    for x, y in zip(a, b):
        assert x["status"] == abi.SEQ_TRACKED and x["matches"] > 50
        assert np.array_equal(x["pose"], y["pose"]) and x["matches"] == y["matches"]


def test_context_on_a_second_device(binding, abi, sw):
    """cudaFuncSetAttribute applies to the current device only: the opt-ins must be repeated for device 1."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg, poses, imgs = sw.sequence("C2", 4, 3)
    a = _short_track(binding, abi, sw, cfg, poses, imgs, 0)
    b = _short_track(binding, abi, sw, cfg, poses, imgs, 1)
    for x, y in zip(a, b):
        assert np.array_equal(x["pose"], y["pose"]) and x["matches"] == y["matches"]


def test_corner_list_longer_than_the_pinned_mirror(binding, sw, O):
    """The pinned corner mirror holds max(2048, 2 * num_features) records; a budget that yields more is fetched from the
    device list in full (C5-sized frame, budget 6000)."""
    cfg = sw.config("C5")
    params = copy.copy(cfg["params"])
    params.num_features = 1000          # mirror: 2048 records
    poses = sw.trajectory(cfg, 1, 1)
    img = sw.render(cfg, poses)[0]
    ctx = binding.Context(params, cfg["cam"])
    f = ctx.frame(img, corners=True, nfeatures=6000)
    xyl, sc = f.corners()
    rx, rs = O.detect(params, img, 6000)
    assert len(rx) > 2048, f"the case needs more corners than the mirror holds ({len(rx)})"
    assert np.array_equal(xyl, rx) and np.array_equal(sc, rs)
    ctx.close()
