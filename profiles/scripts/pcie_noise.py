"""Does PCIe traffic slow the compute kernels down?  Tracks HBM-resident frames (no upload in the tracker) while a
background thread keeps the host->device direction of PCIe busy, either with the copy engine (cudaMemcpyAsync of pinned
memory) or with an SM kernel reading pinned memory (torch index copy through a mapped tensor is not available, so the
kernel variant uses the library's own upload through sdvlb_frames_submit without corners).
  python profiles/scripts/pcie_noise.py [none|ce|ce_small]"""
import importlib
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.load_pkg()
sw = importlib.import_module("slam_sdvl_b200.synthworld")
binding = importlib.import_module("slam_sdvl_b200.binding")
mode = sys.argv[1] if len(sys.argv) > 1 else "none"
G, T, S = 8, 4, 64
cfg = sw.config("C2")
w, h = cfg["w"], cfg["h"]
K = 60
F = 1 + 5 + K
host = torch.empty((S, F, h, w), dtype=torch.uint8).pin_memory()
gt = np.zeros((S, F, 7))
for s in range(S):
    gt[s] = sw.trajectory(cfg, s, F)
    sw.render(cfg, gt[s], threads=16, out=host.numpy()[s])
dev = host.cuda()
trk = binding.HostTracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20, S, G, n_threads=T, resident=True)
ptr_tab = (dev.data_ptr() + (np.arange(S, dtype=np.uint64)[:, None] * F + np.arange(F, dtype=np.uint64)[None, :]) * np.uint64(w * h))
trk.run_ptrs(ptr_tab[:, :6], gt[:, :6], on_device=1)
torch.cuda.synchronize()

stop = False
moved = [0]


def noise():
    st = torch.cuda.Stream()
    if mode == "ce":
        src = torch.empty(32 << 20, dtype=torch.uint8).pin_memory()
        dst = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(st):
            while not stop:
                for _ in range(4):
                    dst.copy_(src, non_blocking=True)
                    moved[0] += src.numel()
                st.synchronize()
    elif mode == "ce_small":   # frame-sized copies
        src = torch.empty((64, w * h), dtype=torch.uint8).pin_memory()
        dst = torch.empty((64, w * h), dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(st):
            while not stop:
                for i in range(64):
                    dst[i].copy_(src[i], non_blocking=True)
                    moved[0] += w * h
                st.synchronize()


th = threading.Thread(target=noise)
if mode != "none":
    th.start()
    time.sleep(0.05)
m0 = moved[0]
t0 = time.perf_counter()
trk.run_ptrs(ptr_tab[:, 6:], gt[:, 6:], on_device=1)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
m1 = moved[0]
stop = True
if mode != "none":
    th.join()
trk.close()
print(f"noise={mode}: {S * K / dt:.0f} frames/s; background H2D {(m1 - m0) / dt / 1e9:.1f} GB/s")
