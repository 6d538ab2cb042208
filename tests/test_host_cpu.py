"""CPU suite, part 2: the product's host-side pieces that do not need a GPU — the corner-selection ordering logic
(shared header compiled for the host), the C-ABI surface, and the sequence sharding used by the multi-GPU bench."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def select_shim(tmp_path_factory):
    so = tmp_path_factory.mktemp("shim") / "libselect_shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared",
                           os.path.join(ROOT, "tests", "host_shims", "select_shim.cc"), "-o", str(so)])
    return C.CDLL(str(so))


def test_device_retain_best_order_equals_libstdcxx(select_shim, O):
    """slam-sdvl_b200/csrc/select_impl.h (the code the select kernel runs) vs std::nth_element/std::partition."""
    rng = np.random.default_rng(1)
    for t in range(1500):
        n = int(rng.integers(1, 900)) if t % 3 else int(rng.integers(1, 40))
        hi = int(rng.choice([3, 10, 40, 245]))
        score = rng.integers(10, 10 + hi, n).astype(np.uint32)
        if t % 7 == 0:
            score = np.sort(score)
        if t % 11 == 0:
            score = np.sort(score)[::-1].copy()
        pay = np.arange(n, dtype=np.uint32)
        keep = int(rng.integers(0, n + 3))
        a = ((score << 22) | pay).astype(np.uint32)
        m = select_shim.shim_retain_best22(a.ctypes.data_as(C.c_void_p), n, keep)
        ref = O.retain_best(np.stack([pay.astype(np.float32), np.zeros(n, np.float32), score.astype(np.float32)], 1), keep)
        assert m == len(ref) and np.array_equal(a[:m] & 0x3FFFFF, ref[:, 0].astype(np.uint32)), (n, keep)


def test_device_retain_best_killer_sequences(select_shim, O):
    """Inputs that drive introselect into its heap-select fallback (median-of-3 killer, organ pipes, all equal)."""
    def killer(n):   # Musser's median-of-3 killer permutation
        k = n // 2
        a = [0] * n
        for i in range(1, k + 1):
            a[2 * i - 2] = i if i % 2 else k + i - 1
            a[2 * i - 1] = k + i
        return np.array(a, np.uint32) + 1

    for n in (64, 256, 512, 850):
        for arr in (killer(n), np.r_[np.arange(n // 2), np.arange(n // 2)[::-1]].astype(np.uint32) + 1,
                    np.full(n, 5, np.uint32)):
            arr = (arr % 250).astype(np.uint32) + 1
            for keep in (1, n // 3, n // 2, n - 1):
                pay = np.arange(len(arr), dtype=np.uint32)
                a = ((arr << 22) | pay).astype(np.uint32)
                m = select_shim.shim_retain_best22(a.ctypes.data_as(C.c_void_p), len(arr), keep)
                ref = O.retain_best(np.stack([pay.astype(np.float32), np.zeros(len(arr), np.float32), arr.astype(np.float32)], 1), keep)
                assert m == len(ref) and np.array_equal(a[:m] & 0x3FFFFF, ref[:, 0].astype(np.uint32))


def test_c_abi_exports_every_declared_symbol(binding):
    """Every function include/sdvl_b200.h declares is exported by libsdvl_b200.so (load only, no compute)."""
    hdr = open(os.path.join(ROOT, "include", "sdvl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sdvlb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = binding.load()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/sdvl_b200.h but not exported"
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)
    hlib = binding.load_host()
    for sym in binding.HOST_EXPORTS:
        assert hasattr(hlib, sym)


def test_params_default_matches_reference_config(binding, abi):
    p = abi.Params()
    binding.load().sdvlb_params_default(C.byref(p))
    d = abi.default_params()
    for name, _ in abi.Params._fields_:
        assert getattr(p, name) == getattr(d, name), name
    assert (p.pyramid_levels, p.max_matches, p.max_align_level, p.min_align_level, p.max_img_align_its) == (5, 150, 4, 2, 30)


def test_no_cuda_is_a_loud_failure(binding, abi):
    """Without a device the product path refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    p = abi.default_params()
    cam = abi.Camera(640, 480, 300, 300, 320, 240)
    with pytest.raises(binding.SdvlbError):
        binding.Context(p, cam)
    with pytest.raises(binding.SdvlbError):
        binding.HostTracker(p, cam, np.array([0, 0, 1.0, 0]), 100, 20, 1)


def test_product_never_touches_the_oracle():
    """Nothing under slam-sdvl_b200/ may import, link or call oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "slam-sdvl_b200")):
        if "_build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle_py|liboracle|\borc_[a-z]|#include\s+\"[^\"]*oracle", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(ROOT, "slam-sdvl_b200", "libsdvl_b200_host.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import load_pkg
load_pkg()
import importlib
sh = importlib.import_module("slam_sdvl_b200.sharding")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
seeds = sh.shard_seeds(rank, world, 3)
sec = sh.max_over_ranks(1.0 + rank, device="cpu")
tot = sh.sum_over_ranks(len(seeds), device="cpu")
with open(os.path.join({out!r}, "rank%d.json" % rank), "w") as f:   # one file per rank: stdout lines can interleave
    json.dump(dict(rank=rank, seeds=seeds, sec=sec, tot=tot), f)
dist.barrier()
dist.destroy_process_group()
"""


def test_sequence_sharding_two_ranks_gloo(tmp_path):
    """world_size-2 gloo run of the sharding helpers the multi-GPU bench uses: disjoint sequences, max-over-ranks time."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    rows = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    seeds = sorted(sum((r["seeds"] for r in rows), []))
    assert seeds == list(range(6))
    assert all(r["sec"] == 2.0 for r in rows) and all(r["tot"] == 6 for r in rows)


def test_product_rand_matches_glibc_and_libstdcxx(binding, abi):
    """sdvlb_rand_* (the rand() stream every FeatureAlign / resident sequence owns) == glibc rand() after srand(seed),
    and sdvlb_rand_shuffle == libstdc++ std::random_shuffle driven by it (same loop as the oracle's, which
    tests/test_oracle_cpu.py pins against the real libstdc++)."""
    libc = C.CDLL("libc.so.6")
    L = binding.load()
    for seed in (1, 7, 12345):
        libc.srand(seed)
        ref = [libc.rand() for _ in range(2000)]
        r = abi.Rand()
        L.sdvlb_rand_seed(C.byref(r), seed)
        assert [L.sdvlb_rand_next(C.byref(r)) for _ in range(2000)] == ref
    for n in (1, 2, 17, 360):
        libc.srand(1)
        v = np.arange(n, dtype=np.int32)
        expect = v.copy()
        for i in range(1, n):
            j = libc.rand() % (i + 1)
            expect[i], expect[j] = expect[j], expect[i]
        r = abi.Rand()
        L.sdvlb_rand_seed(C.byref(r), 1)
        L.sdvlb_rand_shuffle(C.byref(r), abi.ptr(v), n)
        assert np.array_equal(v, expect)


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` (no GPU): one JSON line with the keys the driver reads, `impl: reference`, a
    `cpu_baseline` describing the run and the e2e object with zero transfer bytes; non-zero ranks print nothing."""
    import json
    import subprocess
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "tracked_frames_per_sec" and d["unit"] == "frames/s"
    assert d["steps"] == 4 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 runs the arm
    env["RANK"], env["WORLD_SIZE"] = "1", "2"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_rolling_pyramid_model_matches_the_oracle(O):
    """The separable rolling-window form of pyr_down_roll_kernel (its numpy model, profiles/scripts/pyr_roll_model.py:
    horizontal sums as 16-bit pairs, borders folded into the dot-product coefficients, 4 / 8 output rows per thread)
    against the oracle's cv::pyrDown restatement, for every row alignment the kernel is instantiated for."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pyr_roll_model", os.path.join(ROOT, "profiles", "scripts", "pyr_roll_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(11)
    for (w, h) in [(64, 24), (40, 26), (44, 18), (24, 50), (20, 10)]:   # rows aligned to 16 / 8 / 4 / 8 / 4 bytes
        for img in (rng.integers(0, 256, (h, w), dtype=np.uint8), np.full((h, w), 255, np.uint8)):
            ref = O.pyramid(img, 2)[1]
            for rows_per_thread in (4, 8):
                assert np.array_equal(m.rolling(img, rows_per_thread), ref), f"{w}x{h}, {rows_per_thread} rows per thread"


def test_bench_reads_dram_traffic_of_every_kernel_from_the_ncu_summary():
    """roofline.traffic comes from the committed `ncu --set full` summary: every bench key must resolve, including the
    kernels ncu prints as `void name<args>` (templates) and keys made of several launches (pyramid)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    assert os.path.exists(bench.NCU_SUMMARY), "no ncu summary under profiles/"
    for key in ("pyramid", "fast", "select", "align", "search", "pose"):
        t = bench.ncu_traffic_per_launch(key, 64)
        assert t is not None and t > 0, key
    assert bench.ncu_traffic_per_launch("pyramid", 32) == pytest.approx(bench.ncu_traffic_per_launch("pyramid", 64) / 2)
