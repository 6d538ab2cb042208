"""Phase timing of fast_select_kernel (debug build of the library with -DSDVLB_SELECT_DEBUG: the kernel prints, for
frame 0 of the batch, the %globaltimer span of every phase of every CTA).  64 C2 frames per batch, second batch shown."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(sys.path[0], "tests"))
import importlib, conftest
conftest.load_pkg()
binding = importlib.import_module("slam_sdvl_b200.binding")
sw = importlib.import_module("slam_sdvl_b200.synthworld")
binding.LIB_PATH = os.path.join(os.path.dirname(binding.LIB_PATH), "libsdvl_b200_dbg.so")
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg, poses, imgs = sw.sequence(name, 0, 2)
ctx = binding.Context(cfg["params"], cfg["cam"])
L = binding.load()
n = 64
arr = (C.c_void_p * n)(*[imgs[k % 2].ctypes.data for k in range(n)])
out = (C.c_void_p * n)()
for rep in range(2):
    print(f"---- batch {rep}", flush=True)
    assert L.sdvlb_frames_submit(C.c_void_p(ctx.h), arr, n, 0, 1, cfg["params"].num_features, out) == 0
    assert L.sdvlb_frames_wait(C.c_void_p(ctx.h), out, n) == 0
    for k in range(n):
        L.sdvlb_frame_destroy(C.c_void_p(ctx.h), C.c_void_p(out[k]))
ctx.close()
