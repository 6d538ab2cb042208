"""Kernel timeline of a short tracker run through torch.profiler (CUPTI activity records of every kernel in the process,
also the ones this library launches): which kernels of the build stream and the tracking stream overlap?
  python profiles/scripts/timeline.py GROUPS THREADS [SEQS]  -> gpurun_out/timeline_<groups>.json (+ summary on stdout)"""
import importlib
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.load_pkg()
sw = importlib.import_module("slam_sdvl_b200.synthworld")
binding = importlib.import_module("slam_sdvl_b200.binding")
G, T = int(sys.argv[1]), int(sys.argv[2])
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
LOC = int(sys.argv[4]) if len(sys.argv) > 4 else 1   # 1 frames in HBM, 2 pinned host memory (upload kernel), 0 DMA
cfg = sw.config("C2")
w, h = cfg["w"], cfg["h"]
NSTEPS = int(sys.argv[5]) if len(sys.argv) > 5 else 8
F = 1 + 5 + NSTEPS
host = torch.empty((S, F, h, w), dtype=torch.uint8).pin_memory()
gt = np.zeros((S, F, 7))
for s in range(S):
    gt[s] = sw.trajectory(cfg, s, F)
    sw.render(cfg, gt[s], threads=16, out=host.numpy()[s])
dev = host.cuda()
trk = binding.HostTracker(cfg["params"], cfg["cam"], sw.PLANE, cfg["n_feat"], 20, S, G, n_threads=T, resident=True)
ptr_tab = ((dev.data_ptr() if LOC == 1 else host.data_ptr()) + (np.arange(S, dtype=np.uint64)[:, None] * F + np.arange(F, dtype=np.uint64)[None, :]) * np.uint64(w * h))
trk.run_ptrs(ptr_tab[:, :6], gt[:, :6], on_device=LOC)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trk.run_ptrs(ptr_tab[:, 6:], gt[:, 6:], on_device=LOC)
    torch.cuda.synchronize()
trk.close()
out = os.path.join(ROOT, "gpurun_out", f"timeline_{G}_{LOC}.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
busy = 0.0
end = t0
for e in ev:
    s_, e_ = e["ts"], e["ts"] + e["dur"]
    if e_ > end:
        busy += e_ - max(s_, end)
        end = e_
span = end - t0
print(f"groups={G}: {len(ev)} kernels over {span:.0f} us, GPU has >=1 kernel running {100 * busy / span:.0f}% of the time, "
      f"sum of kernel durations {sum(e['dur'] for e in ev):.0f} us (= {sum(e['dur'] for e in ev) / span:.2f} kernels in flight on average)")
for e in ev[:60]:
    print(f"{e['ts'] - t0:9.1f} +{e['dur']:7.1f}  stream {e['args'].get('stream')}  {e['name'][:50]}")
