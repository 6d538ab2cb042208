"""GPU end-to-end parity: the host mirror (classic class-API path and batched path) against the CPU oracle on
synthetic sequences.  Bar (BASELINE.json): trajectory within 1 mm ATE of the reference CPU path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATE_MM = 1.0


def _run_oracle(O, sw, cfg, poses, imgs, kf_every=20):
    tr = O.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], kf_every)
    est, stats, _ = tr.run(imgs, poses)
    tr.close()
    return est, stats


def _run_host(binding, sw, cfg, seqs, classic, n_groups=1, kf_every=20):
    n_seq = len(seqs)
    n = seqs[0][1].shape[0]
    t = binding.HostTracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], kf_every, n_seq, n_groups)
    est = np.zeros((n_seq, n, 7))
    stats = np.zeros((n_seq, n, 8), np.int32)
    for k in range(n):
        imgs = np.stack([s[1][k] for s in seqs])
        gt = np.stack([s[0][k] for s in seqs])
        e, st = t.step(imgs, gt, classic=classic)
        est[:, k] = e
        stats[:, k] = st
    t.close()
    return est, stats


@pytest.mark.parametrize("name,n_frames", [("C2", 40), ("C1", 30), ("C3", 30)])
def test_trajectory_vs_oracle(binding, sw, O, name, n_frames):
    cfg = sw.config(name)
    seqs = []
    for seed in (0, 1, 2):
        poses = sw.trajectory(cfg, seed, n_frames)
        seqs.append((poses, sw.render(cfg, poses)))
    est_b, st_b = _run_host(binding, sw, cfg, seqs, classic=False, n_groups=2)
    est_c, st_c = _run_host(binding, sw, cfg, seqs[:1], classic=True)
    # the batched submission and the per-call class API are the same computation
    assert np.array_equal(est_b[0], est_c[0]), "batched and classic paths diverge"
    assert np.array_equal(st_b[0][:, [1, 2, 3, 4, 5, 7]], st_c[0][:, [1, 2, 3, 4, 5, 7]])
    for i, (poses, imgs) in enumerate(seqs):
        est_o, st_o = _run_oracle(O, sw, cfg, poses, imgs)
        d = np.array([np.linalg.norm(sw.cam_center(a) - sw.cam_center(b)) for a, b in zip(est_b[i], est_o)])
        ate_vs_oracle = float(np.sqrt((d ** 2).mean())) * 1e3
        ate_gt_gpu = sw.ate(est_b[i], poses) * 1e3
        ate_gt_cpu = sw.ate(est_o, poses) * 1e3
        same_stats = float((st_b[i][:, 1:6] == st_o[:, 1:6]).all(axis=1).mean())
        print(f"{name} seq {i}: ATE gpu-vs-oracle {ate_vs_oracle:.4f} mm (max {d.max()*1e3:.4f}), vs GT gpu {ate_gt_gpu:.3f} mm / "
              f"cpu {ate_gt_cpu:.3f} mm; frames with identical match stats {same_stats:.2%}; "
              f"matches/frame {st_b[i][1:,1].mean():.0f}, keyframes {st_b[i][:,7].sum()}")
        assert ate_vs_oracle <= ATE_MM
        assert ate_gt_gpu <= 5.0
        assert st_b[i][1:, 1].mean() > 0.4 * cfg["n_feat"]


@pytest.mark.parametrize("n_groups,n_threads", [(1, 1), (3, 2), (5, 0)])
def test_pipelined_run_equals_lockstep(binding, sw, n_groups, n_threads):
    """sdvlh_tracker_run (frame batches built one step ahead, asynchronous tracking, groups interleaved on fewer host
    threads) is the same computation as stepping all sequences in lock-step."""
    cfg = sw.config("C2")
    n_seq, n = 5, 12
    seqs = []
    for seed in range(n_seq):
        poses = sw.trajectory(cfg, 10 + seed, n)
        seqs.append((poses, sw.render(cfg, poses)))
    est_s, st_s = _run_host(binding, sw, cfg, seqs, classic=False, n_groups=2)
    imgs = np.stack([s[1] for s in seqs])
    gt = np.stack([s[0] for s in seqs])
    t = binding.HostTracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20, n_seq, n_groups, n_threads=n_threads)
    # two consecutive runs on the same tracker continue the sequences
    e1, s1 = t.run(imgs[:, :5], gt[:, :5])
    e2, s2 = t.run(imgs[:, 5:], gt[:, 5:])
    t.close()
    est_r = np.concatenate([e1, e2], axis=1)
    st_r = np.concatenate([s1, s2], axis=1)
    assert np.array_equal(est_r, est_s)
    assert np.array_equal(st_r, st_s)


@pytest.mark.parametrize("loc", [0, 1, 2])
def test_async_frame_batches_match_sync_frames(binding, sw, O, loc):
    """sdvlb_frames_submit (build stream; images in pageable host, device or pinned host memory) produces the same
    pyramid bytes and corner list as sdvlb_frame_create."""
    import ctypes as C
    import torch
    cfg, poses, imgs = sw.sequence("C2", 4, 3)
    ctx = binding.Context(cfg["params"], cfg["cam"])
    L = binding.load()
    n = len(imgs)
    t = torch.from_numpy(np.ascontiguousarray(imgs))
    if loc == 1:
        t = t.cuda()
    elif loc == 2:
        t = t.pin_memory()
    stride = imgs.shape[1] * imgs.shape[2]
    ptrs = (C.c_void_p * n)(*[t.data_ptr() + i * stride for i in range(n)])
    out = (C.c_void_p * n)()
    rc = L.sdvlb_frames_submit(C.c_void_p(ctx.h), ptrs, n, loc, 1, cfg["params"].num_features, out)
    assert rc == 0, L.sdvlb_last_error()
    assert L.sdvlb_frames_wait(C.c_void_p(ctx.h), out, n) == 0
    for i in range(n):
        fa = binding.Frame(ctx, out[i])
        fs = ctx.frame(imgs[i], corners=True)
        xa, sa = fa.corners()
        xs, ss = fs.corners()
        assert len(xa) > 500 and np.array_equal(xa, xs) and np.array_equal(sa, ss)
        for l in range(cfg["params"].pyramid_levels):
            assert np.array_equal(fa.level(l), fs.level(l))
        fa.destroy()
        fs.destroy()
    ctx.close()
