#!/usr/bin/env python
"""bench.py — tracked frames/s of the SDVL tracking front-end (pyramid + FAST + ImageAlign + FeatureAlign) on B200.

Workload (BASELINE.json configs[1], "C2"): EuRoC-shaped synthetic 752x480 mono sequences, 5-level pyramid, 200
features, full front-end per frame.  One GPU runs `--seqs` independent sequences in lock-step; a *step* is one new
frame for every sequence of this GPU.  Sequences are sharded over ranks with no data-path collective (weak scaling).

  value : frames/s with the frames already resident in HBM when the timed region starts
  e2e   : frames/s through the host-facing API with frames in pinned host memory (H2D of every frame and D2H of
          poses / matches / corner lists inside the timed region)
  roofline    : dominant kernel, algorithmic bytes per launch / CUDA-event duration, vs measured HBM copy peak
  cpu_baseline: the reference's own CPU code (oracle/_ref, kind "reference": SDVL's sources compiled unmodified
          against stand-in Eigen / OpenCV headers) on one host core, bounded sample; the oracle port if _ref is absent
  --impl reference : the same on all host cores (one sequence per core, one process each), no GPU
"""
import argparse
import importlib
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: see slam-sdvl_b200/binding.py
IMG_HOST, IMG_DEVICE, IMG_PINNED = 0, 1, 2   # SDVLB_IMG_* (include/sdvl_b200.h)


def load_pkg():
    name = "slam_sdvl_b200"
    if name not in sys.modules:
        d = os.path.join(ROOT, "slam-sdvl_b200")
        spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"),
                                                      submodule_search_locations=[d])
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
    return sys.modules[name]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


WORKLOADS = {
    "C1": "C1: 640x480 TUM-shaped synthetic textured-plane sequences, 4-level pyramid, ~100 features",
    "C2": "C2: 752x480 EuRoC-shaped synthetic sequences, 5-level pyramid, 200 features",
    "C3": "C3: 640x480 TUM-shaped synthetic sequences with fast camera motion, 5-level pyramid, 200 features",
    "C5": "C5: 1920x1080 synthetic sequences, 5-level pyramid, 2000 features per frame",
}


def workload(name):
    return WORKLOADS[name] + ", pyramid+FAST+ImageAlign+FeatureAlign"


def ncu_summary_path():
    """The newest committed `ncu --set full` summary (profiles/rNN_*_ncu_full.txt)."""
    d = os.path.join(ROOT, "profiles")
    c = sorted(f for f in os.listdir(d) if f.endswith("_ncu_full.txt") and "seed" not in f) if os.path.isdir(d) else []
    return os.path.join(d, c[-1]) if c else os.path.join(d, "missing")


NCU_SUMMARY = ncu_summary_path()
NCU_KERNELS = {"pyr_down_kernel": "pyramid", "pyr_down_roll_kernel": "pyramid", "pyr_tail_kernel": "pyramid", "fast_cells_kernel": "fast",
               "fast_select_kernel": "select", "seq_align_kernel": "align", "image_align_kernel": "align",
               "search_seq_kernel": "search", "seq_post_kernel": "pose", "orb_frames_kernel": "orb"}
NCU_SEQS_PER_LAUNCH = 64


def ncu_traffic_per_launch(kernel, seqs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` (bench key) from the ncu capture under
    profiles/ (64 sequences per launch; scaled to `seqs`).  A key made of several launches per step (pyramid) is the
    sum of their per-launch averages.  None when the summary is not there."""
    if not os.path.exists(NCU_SUMMARY):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    cur = None
    with open(NCU_SUMMARY) as f:
        for line in f:
            if line.startswith("## "):
                name, grid = line[3:line.index("grid")].strip(), line[line.index("grid"):].strip()
                name = name.removeprefix("void ").split("<")[0]          # template kernels: "void name<args>"
                cur = (name, grid) if NCU_KERNELS.get(name) == kernel else None
                if cur:
                    per.setdefault(cur, []).append(0.0)
            elif cur and line.strip().startswith(("dram read", "dram write")):
                parts = line.split()
                per[cur][-1] += float(parts[2]) * unit.get(parts[3], 1.0)
    if not per:
        return None
    return sum(sum(v) / len(v) for v in per.values()) * seqs / NCU_SEQS_PER_LAUNCH


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        if self.index is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU boxes have one PCIe root per socket: a rank whose pinned frame buffers sit on the other socket pays the
    socket interconnect on every upload.  Restrict the rank to the cores of its GPU's NUMA node before anything is
    allocated (first touch then places the pinned memory there).  Returns a short description for the JSON line."""
    if os.environ.get("SDVLB_NUMA_BIND", "1") == "0":
        return {"bound": False, "why": "SDVLB_NUMA_BIND=0"}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"bound": False, "why": "no NUMA node reported", "pci": bus}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "node has no usable cpus", "pci": bus, "node": node}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "pci": bus, "node": node, "cpus": len(cpus)}
    except Exception as e:   # sysfs layout differs / attribute missing: run unbound
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


def algorithmic_bytes(cfg, kernel, stats_prev, stats_cur, n_corners=1000, kbar=3.0):
    """SURVEY.md §8(d) per-frame formulas, summed over the frames of one launch (one step of one group)."""
    w, h, L = cfg["w"], cfg["h"], cfg["params"].pyramid_levels
    dims = [(w >> l, h >> l) for l in range(L)]
    if kernel == "pyramid":
        per = sum(a * b for a, b in dims)
        return per * len(stats_cur)
    if kernel == "fast":
        per = sum(a * b for a, b in dims[:cfg["params"].max_fast_levels]) + 16 * n_corners
        return per * len(stats_cur)
    if kernel == "select":
        return 16 * n_corners * len(stats_cur)
    levels = cfg["params"].max_align_level - cfg["params"].min_align_level + 1
    n_feat = stats_prev[:, 5].astype(np.int64)          # features of the alignment reference = candidates searched
    if kernel == "align":
        iters = stats_cur[:, 6].astype(np.int64)
        return int((levels * n_feat * 49 + iters * n_feat * 25 + 64 * n_feat).sum())
    if kernel == "search":
        return int((n_feat * (121 + 64 * kbar + 810)).sum())
    if kernel == "orb":      # per corner: the radius-15 circular patch (709 bytes) + 32 descriptor bytes
        return (709 + 32) * n_corners * len(stats_cur)
    if kernel == "pose":     # read matches (40 B) + features, write the new list, its host mirror and the obs arrays
        found = stats_cur[:, 1].astype(np.int64)
        return int((n_feat * (40 + 160)).sum() + (found * (160 + 32 + 7 * 8)).sum())
    raise ValueError(kernel)


def cpu_impl():
    """The CPU arm's implementation: the reference's own sources (oracle/_ref/libsdvlref.so, compiled unmodified from
    /root/reference against oracle/ref_shim, kind "reference") when that library was built, else the oracle port."""
    from oracle import oracle_py, ref_py
    if os.path.exists(ref_py.path()):
        try:
            ref_py.lib()
            return ref_py, "reference", ("SDVL's own image_align.cc / matcher.cc / feature_align.cc / frame.cc / fast_detector.cc "
                                         "(oracle/_ref: compiled unmodified, -O3; cv::pyrDown / cv::FAST served by the "
                                         "cv2-pinned restatements)")
        except Exception as e:   # a library built for another machine: fall back to the port, and say so
            print(f"bench.py: oracle/_ref unusable ({e}); timing the oracle port", file=sys.stderr)
    oracle_py.lib()
    return oracle_py, "port", "CPU oracle (restatement of the reference path)"


def opencv_sanity(cfg, sw):
    """the two third-party image operations of the CPU arm are the oracle's restatements of
    cv::pyrDown / cv::FAST (OpenCV's C++ library is not installed); this times them next to the real OpenCV (python
    cv2, one thread) on one frame of the workload, so the reader can see how much of the CPU arm's time per frame could
    be owed to the restatement rather than to OpenCV.  cv2's FAST is timed on whole levels (one call per level): a
    lower bound of the reference's 480 per-cell calls."""
    try:
        import cv2
        from oracle import oracle_py as O
        cv2.setNumThreads(1)
        P = cfg["params"]
        img = sw.render(cfg, sw.trajectory(cfg, 4242, 1), threads=1)[0]

        def ms(fn, reps=5):
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0) / reps * 1e3

        def cv_pyr():
            pyr = [img]
            for _ in range(1, P.pyramid_levels):
                pyr.append(cv2.pyrDown(pyr[-1], dstsize=(pyr[-1].shape[1] // 2, pyr[-1].shape[0] // 2)))
            return pyr

        fd = cv2.FastFeatureDetector_create(threshold=P.fast_threshold, nonmaxSuppression=True)
        levels = cv_pyr()[:P.max_fast_levels]
        t_pyr_o = ms(lambda: O.pyramid(img, P.pyramid_levels))
        # the reference's own call pattern: one cv::FAST per 32 x 32 cell ROI (fast_detector.cc:81-95); the same number
        # of calls on 7 x 7 ROIs (nothing to test) measures the per-call overhead, most of which is the python binding
        m = 1 + P.patch_size // 2
        rois, tiny = [], []
        for im in levels:
            H, W = im.shape
            for i in range((H + P.cell_size - 1) // P.cell_size):
                y0, y1 = max(m, P.cell_size * i), min(H - m, P.cell_size * (i + 1))
                for j in range((W + P.cell_size - 1) // P.cell_size):
                    x0, x1 = max(m, P.cell_size * j), min(W - m, P.cell_size * (j + 1))
                    if y1 > y0 and x1 > x0:
                        rois.append(im[y0:y1, x0:x1])
                        tiny.append(im[y0:y0 + 7, x0:x0 + 7])
        return {"cv2_version": cv2.__version__, "cv2_pyrdown_ms_per_frame": ms(cv_pyr),
                "cv2_fast_whole_levels_ms_per_frame": ms(lambda: [fd.detect(x) for x in levels]),
                "cv2_fast_per_cell_ms_per_frame": ms(lambda: [fd.detect(r) for r in rois]),
                "cv2_fast_per_cell_call_overhead_ms_per_frame": ms(lambda: [fd.detect(r) for r in tiny]),
                "cells": len(rois),
                "restated_pyrdown_ms_per_frame": t_pyr_o,
                "restated_fast_and_selection_ms_per_frame": ms(lambda: O.detect(P, img, P.num_features)) - t_pyr_o}
    except Exception as e:   # noqa: BLE001 -- informational only
        return {"unavailable": repr(e)}


def opencv_sanity_subprocess(config):
    """opencv_sanity in a fresh interpreter: cv2 is never imported into the process that holds the CUDA context."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--opencv-sanity", "--config", config],
                           capture_output=True, text=True, timeout=120)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001 -- informational only
        return {"unavailable": repr(e)}


def _reference_worker(i, cfg_name, F, W, kf_every, start, q):
    """One process per host core: the reference keeps process-wide state (rand(), Frame::counter_, Config singleton),
    so sequences run in separate processes, as separate SDVL instances would."""
    try:
        load_pkg()
        sw = importlib.import_module("slam_sdvl_b200.synthworld")
        impl, _, _ = cpu_impl()
        cfg = sw.config(cfg_name)
        poses = sw.trajectory(cfg, 1000 + i, F)
        imgs = sw.render(cfg, poses, threads=1)
        tr = impl.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], kf_every)
        tr.run(imgs[:1 + W], poses[:1 + W])
        start.wait()
        est, st, sec = tr.run(imgs[1 + W:], poses[1 + W:])
        t_end = time.perf_counter()
        tr.close()
        q.put((i, t_end, sec, float(sw.ate(est, poses[1 + W:])), int(np.clip(st[:, 6], 0, None).sum()), None))
    except Exception as e:   # noqa: BLE001 -- reported by the parent
        try:
            start.abort()
        except Exception:
            pass
        q.put((i, 0.0, 0.0, 0.0, 0, repr(e)))


def run_reference(args, cfg, sw, rank, world):
    """The reference path's CPU implementation on all host cores, one sequence per core (one process each): the
    reference's own code when oracle/_ref was built (see cpu_impl), else the oracle port.  No GPU, no CUDA."""
    if rank != 0:
        return
    import multiprocessing as mp
    impl, kind, what = cpu_impl()
    T = os.cpu_count() or 1
    W, K = args.warmup, args.steps
    F = 1 + W + K
    ctx = mp.get_context("fork")
    start = ctx.Barrier(T + 1)
    q = ctx.Queue()
    procs = [ctx.Process(target=_reference_worker, args=(i, args.config, F, W, args.kf_every, start, q)) for i in range(T)]
    for p in procs:
        p.start()
    try:
        start.wait()                  # every worker has rendered its sequence and run its warm-up frames
    except Exception:                 # a worker failed before the barrier (it aborts it): its message is in the queue
        pass
    t0 = time.perf_counter()
    res = [q.get() for _ in range(T)]
    for p in procs:
        p.join()
    bad = [r for r in res if r[5]]
    if bad:
        raise SystemExit(f"bench.py --impl reference: worker failed: {bad[0][5]}")
    dt = max(r[1] for r in res) - t0  # CLOCK_MONOTONIC is system-wide: the slowest worker ends the step
    gn = sum(r[4] for r in res)
    value = T * K / dt
    ate = max(r[3] for r in res) * 1e3
    out = {
        "impl": "reference", "metric": "tracked_frames_per_sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
        "config": {"workload": workload(args.config), "sequences": T, "processes": T,
                   "step": "one frame for each of the sequences (one per host core)"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": T, "kind": kind,
                         "sample": f"{T} sequences x {K} frames, one sequence per core; {what}",
                         "frames_per_s_per_core": K * T / sum(r[2] for r in res) if sum(r[2] for r in res) > 0 else None,
                         "opencv_sanity": opencv_sanity(cfg, sw)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "us_per_gn_iter_cpu": None, "gn_iters": gn if kind == "port" else None, "max_ate_mm_vs_gt": ate,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seqs", type=int, default=64, help="sequences per GPU (weak scaling) / in total (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --seqs sequences per GPU; strong: --seqs sequences in total, split over the ranks "
                         "(BASELINE.json configs[3] as written: 64 sequences over 1/2/4/8 GPUs)")
    ap.add_argument("--repeats", type=int, default=5, help="timed blocks per leg; the median is reported")
    ap.add_argument("--groups", type=int, default=0, help="contexts (stream pairs) per GPU (0 = 2 per host thread)")
    ap.add_argument("--threads", type=int, default=0, help="host threads per GPU (0 = cores / ranks)")
    ap.add_argument("--kf-every", type=int, default=20)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--opencv-sanity", action="store_true", help="print the cv2-vs-restatement timings as JSON and exit")
    ap.add_argument("--sweep", default="", help="comma list of GROUPSxTHREADS to time (e2e only), e.g. 8x4,16x8")
    ap.add_argument("--prefetch", type=int, default=2, help="frame batches are built this many steps ahead")
    ap.add_argument("--depth", type=int, default=1, help="tracking submissions in flight per group (resident sequences)")
    ap.add_argument("--sweep-cycles", action="store_true", help="print the in-kernel latency breakdown per sweep entry")
    ap.add_argument("--sweep-device", action="store_true", help="sweep with the frames resident in HBM")
    ap.add_argument("--e2e-upload", default="kernel", choices=["dma", "kernel"],
                    help="how level 0 crosses PCIe in the e2e run: copy engine (cudaMemcpyAsync per frame) or the "
                         "one-kernel upload reading pinned host memory")
    ap.add_argument("--host-replay", action="store_true",
                    help="previous design: FeatureAlign bookkeeping + pose refinement on the host (for comparison)")
    ap.add_argument("--no-extras", action="store_true", help="skip the pageable / single-sequence / CPU legs")
    ap.add_argument("--orb", action="store_true",
                    help="Config::UseORB() (the mode every shipped cfg of the reference sets): FAST with the ORB margin, "
                         "corner descriptors at frame construction, SearchPoint scored by descriptor distance; GPU arm only")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    load_pkg()
    sw = importlib.import_module("slam_sdvl_b200.synthworld")
    sharding = importlib.import_module("slam_sdvl_b200.sharding")
    cfg = sw.config(args.config)

    if args.opencv_sanity:
        print(json.dumps(opencv_sanity(cfg, sw)), flush=True)
        return
    if args.impl == "reference":
        run_reference(args, cfg, sw, rank, world)
        return

    import torch
    import torch.distributed as dist
    binding = importlib.import_module("slam_sdvl_b200.binding")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "why": "single GPU"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    binding.load()
    binding.load_host()

    # sequences of this rank: a fixed number per GPU (weak) or an equal share of a fixed total (strong)
    if args.scaling == "strong":
        S = (args.seqs + world - 1 - rank) // world
        S_total = args.seqs
        first_seq = sum((args.seqs + world - 1 - r) // world for r in range(rank))
    else:
        S, S_total, first_seq = args.seqs, args.seqs * world, rank * args.seqs
    if S <= 0:
        raise SystemExit("bench.py: fewer sequences than ranks")
    W, K, R = args.warmup, args.steps, max(1, args.repeats)
    F = 1 + W + K                       # frame 0 initialises every sequence (ground-truth pose + map seeding)
    w, h = cfg["w"], cfg["h"]
    ncpu = os.cpu_count() or 1
    cores = max(1, ncpu // max(1, world))
    if world > 1 and cores <= 4:
        cores = max(1, cores - 1)    # leave a core per rank to the Python thread, the driver's threads and the sampler
    if args.host_replay:     # the host replays FeatureAlign: every core is needed
        threads = args.threads or max(1, min(S, cores))
        groups = args.groups or max(1, min(S, 2 * threads))
    else:                    # resident sequences: the host only submits and plays the mapping thread
        threads = args.threads or max(1, min(6, cores))
        groups = args.groups or max(1, min(S, 6))

    # ---- synthetic frames, rendered once into pinned host memory
    host = torch.empty((S, F, h, w), dtype=torch.uint8).pin_memory()
    host_np = host.numpy()
    gt = np.zeros((S, F, 7))
    for s in range(S):
        gt[s] = sw.trajectory(cfg, first_seq + s, F)
        sw.render(cfg, gt[s], threads=max(1, ncpu // max(1, world)), out=host_np[s])
    frame_bytes = w * h

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    extras = {}

    def pcie_h2d_gbs():
        """Host->device copy bandwidth of this box with every rank copying at once (copy engine, pinned source): what
        the e2e leg could reach if the front-end cost nothing, i.e. its roofline."""
        n = 256 << 20
        src = torch.empty(n, dtype=torch.uint8).pin_memory()
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        dst.copy_(src, non_blocking=True)
        barrier()
        best = 0.0
        for _ in range(5):   # the best of five passes: single passes of this measurement spread from 41 to 55 GB/s
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            ev1.record()
            torch.cuda.synchronize()
            sec = sharding.max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
            best = max(best, 4 * n / sec / 1e9)
        return best

    pcie_gbs = pcie_h2d_gbs()

    def timed_run(base_ptr, on_device, n_groups, timing=False, pipelined=True, seqs=None, n_threads=None):
        """Fresh tracker; init + warm-up untimed; K timed steps. Returns seconds (max over ranks), est, stats, extras.
        pipelined: sdvlh_tracker_run (frame batches ahead of the tracking they feed, groups free-running); otherwise
        every step is one synchronous lock-step submission (used for the per-kernel timing pass, where launches must
        not overlap).  seqs: only the first `seqs` sequences (single-sequence latency leg)."""
        n = S if seqs is None else seqs
        n_groups = min(n_groups, n)
        trk = binding.HostTracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], args.kf_every, n, n_groups,
                                  device=local_rank, timing=timing, n_threads=min(n_groups, n_threads or threads),
                                  resident=not args.host_replay, use_orb=args.orb)
        trk.set_prefetch(args.prefetch)
        if not args.host_replay:
            trk.set_depth(args.depth)
        est = np.zeros((n, F, 7))
        stats = np.zeros((F, n, 8), np.int32)

        ptr_tab = (base_ptr + (np.arange(n, dtype=np.uint64)[:, None] * F + np.arange(F, dtype=np.uint64)[None, :])
                   * np.uint64(frame_bytes))
        def go(k0, k1):
            if pipelined:
                e, st = trk.run_ptrs(ptr_tab[:, k0:k1], gt[:n, k0:k1], on_device=on_device)
                est[:, k0:k1] = e
                stats[k0:k1] = st.transpose(1, 0, 2)
            else:
                for k in range(k0, k1):
                    e, st = trk.step_ptrs([int(p) for p in ptr_tab[:, k]], gt[:n, k], on_device=on_device)
                    est[:, k] = e
                    stats[k] = st

        # init frame + warm-up steps (untimed), then K timed steps
        go(0, 1 + W)
        trk.counters(reset=True)
        trk.phases(reset=True)
        if timing:
            trk.timing_read(reset=True)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        go(1 + W, F)
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        sec = ev0.elapsed_time(ev1) * 1e-3
        sec = sharding.max_over_ranks(sec)
        counters = trk.counters(reset=True) + (trk.phases(reset=True),)
        ktimes = trk.timing_read(reset=True) if timing else None
        if not args.host_replay:
            extras["post_cycles_timing" if timing else "post_cycles_run"] = trk.post_cycles(reset=True)
            extras["slowest_timing" if timing else "slowest_run"] = trk.slowest_cycles(reset=True)
            if timing:
                extras["post_cycles"] = extras["post_cycles_timing"]
        ngroups = trk.groups()
        trk.close()
        return sec, wall, est, stats, counters, ktimes, ngroups

    def repeated(base_ptr, on_device, n_groups, **kw):
        """R timed blocks, each a fresh tracker over the same frames (init + warm-up untimed).  Returns the block with
        the median time, plus every block's seconds."""
        runs = [timed_run(base_ptr, on_device, n_groups, **kw) for _ in range(R)]
        for r in runs[1:]:
            assert np.array_equal(r[2], runs[0][2]), "repeated blocks must be the same computation"
        order = sorted(range(R), key=lambda i: runs[i][0])
        return runs[order[R // 2]], [r[0] for r in runs]

    e2e_loc = IMG_HOST if args.e2e_upload == "dma" else IMG_PINNED
    if args.sweep:   # host-side configuration sweep (groups x threads), e2e placement; prints a table and exits
        for spec in args.sweep.split(","):
            g_, t_ = (int(v) for v in spec.split("x"))
            threads = t_
            if args.sweep_device:
                if "dev" not in extras:
                    extras["dev"] = host.cuda(non_blocking=False)
                sec, wall, *_ = timed_run(extras["dev"].data_ptr(), IMG_DEVICE, g_)
            else:
                sec, wall, *_ = timed_run(host.data_ptr(), e2e_loc, g_)
            print(f"sweep groups={g_:3d} threads={t_:3d}: {S_total * K / sec:10.0f} frames/s  ({sec / K * 1e3:.3f} ms/step)",
                  file=sys.stderr, flush=True)
            if args.sweep_cycles:
                print("   in-kernel us/frame:", {k: round(v / 1965.0, 1) for k, v in extras["post_cycles_run"].items()},
                      file=sys.stderr, flush=True)
        return

    clocks = ClockSampler(local_rank if rank == 0 else None)   # rank 0 samples its GPU; 8 samplers only add noise
    clocks.start()
    # ---- e2e: frames in pinned host memory
    (e2e_sec, e2e_wall, est_e, stats_e, cnt_e, _, ngroups), e2e_all = repeated(host.data_ptr(), e2e_loc, groups)
    # ---- value: frames resident in HBM
    dev = host.cuda(non_blocking=False)
    (val_sec, val_wall, est_v, stats_v, cnt_v, _, _), val_all = repeated(dev.data_ptr(), IMG_DEVICE, groups)
    clock_info = clocks.stop()
    # ---- kernel pass: one context so launches do not overlap, per-kernel CUDA events on the launching stream
    # (a submission takes at most 64 sequences: more than that are split over contexts that run one after the other)
    k_sec, _, est_k, stats_k, _, ktimes, _ = timed_run(dev.data_ptr(), IMG_DEVICE, (S + 63) // 64, timing=True, pipelined=False,
                                                       n_threads=1)
    assert np.array_equal(est_k, est_v), "pipelined and lock-step runs must be the same computation"
    assert np.array_equal(est_e, est_v), "host-resident and HBM-resident runs must be the same computation"
    total_frames = S_total * K
    value = total_frames / val_sec
    e2e = total_frames / e2e_sec
    ate_mm = max(sw.ate(est_v[s, 1 + W:], gt[s, 1 + W:]) for s in range(S)) * 1e3
    gn_iters = int(stats_k[1 + W:, :, 6].sum())

    # ---- further legs (N = 1 and weak scaling only: they describe one GPU)
    e2e_pageable = None
    latency = None
    if not args.no_extras and world == 1:
        # the reference hands cv::Mat (pageable memory) to HandleFrame: the same e2e leg from pageable host memory,
        # one cudaMemcpyAsync per frame (SDVLB_IMG_HOST)
        pageable = np.array(host_np, copy=True)
        p_sec, _, est_p, _, cnt_p, _, _ = timed_run(pageable.ctypes.data, IMG_HOST, groups)
        assert np.array_equal(est_p, est_v), "pageable and pinned inputs must be the same computation"
        e2e_pageable = {"value": total_frames / p_sec, "unit": "frames/s", "ms_per_step": p_sec / K * 1e3,
                        "h2d_bytes_per_step": cnt_p[1] / K,
                        "source": "pageable host memory (numpy): sdvlb_frames_submit copies it into the context's pinned staging sets (calling threads, non-temporal stores) and uploads from there"}
        del pageable
        # the reference's real deployment is ONE camera (main.cc:126-159): latency of a single sequence through the same
        # API (frames in pinned host memory, one group, results read back), and with the frames already in HBM
        l_sec, *_ = timed_run(host.data_ptr(), e2e_loc, 1, seqs=1, n_threads=1)
        ld_sec, *_ = timed_run(dev.data_ptr(), IMG_DEVICE, 1, seqs=1, n_threads=1)
        latency = {"sequences": 1, "ms_per_frame_e2e": l_sec / K * 1e3, "ms_per_frame_hbm": ld_sec / K * 1e3,
                   "frames_per_s_e2e": K / l_sec, "note": "one camera, frames pipelined one step ahead of the tracking"}

    # ---- roofline: every kernel of the kernel pass against the measured HBM peak, the dominant one first
    peak, peak_kind = measured_peaks()
    kshare = {k: v[0] for k, v in ktimes.items() if v[1] > 0}
    tot_ms = sum(kshare.values())
    dom = max(kshare, key=kshare.get)
    per_kernel = {}
    for name in kshare:
        alg = 0
        for k in range(1 + W, F):
            alg += algorithmic_bytes(cfg, name, stats_k[k - 1], stats_k[k])
        n_launch = max(1, ktimes[name][1])
        avg_ms = ktimes[name][0] / n_launch
        per_kernel[name] = {"us_per_launch": avg_ms * 1e3, "algorithmic_bytes_per_launch": alg / n_launch,
                            "gbs": (alg / n_launch) / (avg_ms * 1e-3) / 1e9,
                            "frac": (alg / n_launch) / (avg_ms * 1e-3) / 1e9 / peak,
                            "launches_per_step": ktimes[name][1] / K,
                            "traffic": ncu_traffic_per_launch(name, S)}
    alg_all = sum(v["algorithmic_bytes_per_launch"] * v["launches_per_step"] for v in per_kernel.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": per_kernel[dom]["traffic"],
                "traffic_source": os.path.relpath(NCU_SUMMARY, ROOT) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                "peak_source": peak_kind,
                "avg_launch_us": per_kernel[dom]["us_per_launch"],
                "algorithmic_bytes_per_launch": per_kernel[dom]["algorithmic_bytes_per_launch"],
                "kernel_ms_share": {k: (v / tot_ms if tot_ms else 0.0) for k, v in kshare.items()},
                "kernel_us_per_step": {k: v[0] * 1e3 / K for k, v in ktimes.items() if v[1] > 0},
                "per_kernel": per_kernel,
                "whole_step": {"algorithmic_bytes": alg_all, "gbs_at_value": alg_all / (val_sec / K) / 1e9,
                               "frac_at_value": alg_all / (val_sec / K) / 1e9 / peak},
                "pcie": {"h2d_gbs_per_gpu_achieved": cnt_e[1] / e2e_sec / 1e9, "h2d_gbs_per_gpu_peak": pcie_gbs,
                         "frac": cnt_e[1] / e2e_sec / 1e9 / pcie_gbs,
                         "peak_source": "pinned cudaMemcpyAsync 4 x 256 MB, all ranks at once, slowest rank, best of 5 passes",
                         "frames_per_s_at_peak": pcie_gbs * 1e9 / frame_bytes * world}}
    us_per_gn_iter = ktimes["align"][0] * 1e3 / max(1, gn_iters) * S   # one CTA per sequence runs its own GN loop

    # ---- CPU baseline on one host core (rank 0, bounded sample)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_extras:   # N = 1 only: at N > 1 the host cores belong to the ranks
        O, cpu_kind, cpu_what = cpu_impl()
        ns = min(S, 32 if w * h < 1000000 else 4)    # bounded sample, about 5-20 s of CPU work
        t_cpu, frames_cpu, gn_cpu = 0.0, 0, 0
        for s in range(ns):
            tr = O.Tracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], args.kf_every)
            tr.run(host_np[s, :1], gt[s, :1])
            _, st, sec = tr.run(host_np[s, 1:], gt[s, 1:])
            t_cpu += sec
            frames_cpu += F - 1
            gn_cpu += int(np.clip(st[:, 6], 0, None).sum())
            tr.close()
        cpu_baseline = {"value": frames_cpu / t_cpu, "unit": "frames/s", "cores": 1, "kind": cpu_kind,
                        "sample": f"{ns} sequences x {F - 1} frames of the same workload, 1 thread; {cpu_what}",
                        "ms_per_frame": t_cpu / frames_cpu * 1e3,
                        "host_cores_available": ncpu, "opencv_sanity": opencv_sanity_subprocess(args.config)}
        if latency is not None:
            latency["ms_per_frame_cpu_reference_1_thread"] = t_cpu / frames_cpu * 1e3

    def spread(all_sec):
        v = sorted(total_frames / x for x in all_sec)
        return {"median": v[len(v) // 2], "min": v[0], "max": v[-1], "runs": len(v)}

    if rank == 0:
        out = {
            "metric": "tracked_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": val_sec / K * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
            "value_runs": spread(val_all),
            "config": {"workload": workload(args.config) + (" [ORB descriptor mode, use_orb: 1]" if args.orb else ""),
                       "sequences_per_gpu": S, "sequences_total": S_total,
                       "step": "one new frame for every sequence of the GPU", "host_groups_per_gpu": ngroups,
                       "host_threads_per_gpu": min(ngroups, threads), "submissions_in_flight_per_group": args.depth,
                       "timed_blocks": f"{R} blocks of {K} steps per leg, each a fresh tracker (init + {W} warm-up "
                                       "steps untimed); value / e2e are the median block",
                       "cache": "every step consumes frames never seen before (inputs "
                                f"{frame_bytes / 1e6:.2f} MB x sequences per step, "
                                f"{S * F * frame_bytes / 1e6:.0f} MB per block, larger than L2); no L2 flush needed",
                       "kf_every": args.kf_every, "numa": numa,
                       "sequence_state": "host replay" if args.host_replay else "resident in HBM (sdvlb_seq_*)"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": cnt_e[1] / K, "d2h_bytes_per_step": cnt_e[2] / K,
                    "ms_per_step": e2e_sec / K * 1e3, "runs": spread(e2e_all),
                    "source": "pinned host memory, read over PCIe by the upload kernel",
                    "pcie": roofline["pcie"]},
            "e2e_pageable": e2e_pageable,
            "latency_single_sequence": latency,
            "gpu_launches": cnt_v[0],
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clock_info,
            "us_per_gn_iter": us_per_gn_iter,
            "align_and_feature_align_kernel_us_per_frame": {k: v / (clock_info.get("sm_mhz") or 1965.0) for k, v in
                                                  extras.get("post_cycles", {}).items()},
            "slowest_sequence_of_a_group_step_us": {k: v / (clock_info.get("sm_mhz") or 1965.0) for k, v in
                                                    extras.get("slowest_run", {}).items()},
            "gn_iters_per_frame": gn_iters / (S * K),
            "max_ate_mm_vs_gt": ate_mm,
            "matches_per_frame": float(stats_v[1 + W:, :, 1].mean()),
            "keyframes_per_frame": float(stats_v[1 + W:, :, 7].mean()),
            "wall_check_s": {"value": val_wall, "e2e": e2e_wall},
            "host_phase_thread_seconds": {"value": cnt_v[3], "e2e": cnt_e[3]},
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
