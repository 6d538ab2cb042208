"""Model of the warp-parallel retainBest of slam-sdvl_b200/csrc/select_warp.cuh, checked against libstdc++.

The kernel replaces the serial replay of std::nth_element / std::partition by "two-sided passes": the stops of the left
and of the right cursor are the elements of the ORIGINAL range that satisfy the respective stop condition (ascending /
descending), the swaps are the pairs (L[k], R[k]) with L[k] < R[k], the cut is min(L[K], R[K-1]).  This script states
that formulation in numpy and compares the resulting array, ORDER included, with std::nth_element + std::partition
(compiled here with g++, as cv::KeyPointsFilter::retainBest calls them) on random inputs with many ties.

  python profiles/scripts/select_model.py        (needs g++; no GPU)"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

SRC = r"""
#include <algorithm>
#include <cstdint>
#include <vector>
extern "C" int retain_best_ref(uint32_t* a, int n, int n_points, int shift) {
  std::vector<uint32_t> v(a, a + n);
  if (n_points >= 0 && n > n_points) {
    if (n_points == 0) return 0;
    auto gt = [shift](uint32_t x, uint32_t y) { return (x >> shift) > (y >> shift); };
    std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), gt);
    const uint32_t amb = v[n_points - 1] >> shift;
    auto it = std::partition(v.begin() + n_points, v.end(), [=](uint32_t x) { return (x >> shift) >= amb; });
    const int m = int(it - v.begin());
    std::copy(v.begin(), v.end(), a);
    return m;
  }
  return n;
}
"""


def build_ref():
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "ref.cc"), "w") as f:
        f.write(SRC)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", os.path.join(d, "ref.cc"), "-o", os.path.join(d, "libref.so")])
    return ctypes.CDLL(os.path.join(d, "libref.so"))


def ref(lib, a, npts, shift):
    b = np.ascontiguousarray(a, np.uint32).copy()
    m = lib.retain_best_ref(b.ctypes.data_as(ctypes.c_void_p), len(b), npts, shift)
    return b[:m]


def two_sided(a, lo, hi, stop_l, stop_r):
    """One pass: returns (K, L, R); swaps applied in place."""
    r = a[lo:hi]
    idx = np.arange(lo, hi)
    L = idx[stop_l(r)]
    R = idx[stop_r(r)][::-1]
    m = min(len(L), len(R))
    K = int(np.sum(L[:m] < R[:m]))
    for k in range(K):
        a[L[k]], a[R[k]] = a[R[k]], a[L[k]]
    return K, L, R


def nth_element(a, nth, shift):
    n = len(a)
    if n == 0 or nth == n:
        return
    first, last = 0, n
    depth = 2 * (n.bit_length() - 1)
    gt = lambda x, y: (int(x) >> shift) > (int(y) >> shift)
    while last - first > 3:
        if depth == 0:
            raise RuntimeError("heap_select fallback")
        depth -= 1
        mid = first + (last - first) // 2
        x, y, z = first + 1, mid, last - 1
        if gt(a[x], a[y]):
            w = y if gt(a[y], a[z]) else (z if gt(a[x], a[z]) else x)
        else:
            w = x if gt(a[x], a[z]) else (z if gt(a[y], a[z]) else y)
        a[first], a[w] = a[w], a[first]
        p = int(a[first]) >> shift
        K, L, R = two_sided(a, first + 1, last, lambda r: (r >> shift) <= p, lambda r: (r >> shift) >= p)
        cut = min(L[K] if K < len(L) else 1 << 30, R[K - 1] if K > 0 else 1 << 30)
        if cut <= nth:
            first = cut
        else:
            last = cut
    for i in range(first + 1, last):   # __insertion_sort
        val = a[i]
        if gt(val, a[first]):
            a[first + 1:i + 1] = a[first:i].copy()
            a[first] = val
        else:
            l, nx = i, i - 1
            while gt(val, a[nx]):
                a[l] = a[nx]
                l, nx = nx, nx - 1
            a[l] = val


def retain_best(a, npts, shift):
    a = np.array(a, np.uint32)
    n = len(a)
    if not (npts >= 0 and n > npts):
        return a
    if npts == 0:
        return a[:0]
    nth_element(a, npts - 1, shift)
    amb = int(a[npts - 1]) >> shift
    _, _, B = two_sided(a, npts, n, lambda r: (r >> shift) < amb, lambda r: (r >> shift) >= amb)
    return a[:npts + len(B)]


def main(trials=6000):
    lib = build_ref()
    rng = np.random.default_rng(0)
    bad = tot = 0
    for t in range(trials):
        n = int(rng.integers(1, 400))
        kind = t % 4
        if kind == 0:
            resp = rng.integers(0, 256, n)
        elif kind == 1:
            resp = rng.integers(0, 8, n)
        elif kind == 2:
            resp = rng.integers(10, 40, n)
        else:
            resp = np.sort(rng.integers(0, 256, n))[::-1] if t % 8 < 4 else np.full(n, 7)
        shift = 10
        a = (resp.astype(np.uint32) << shift) | rng.permutation(n).astype(np.uint32)
        npts = int(rng.integers(0, n + 2))
        try:
            got = retain_best(a, npts, shift)
        except RuntimeError:
            continue
        exp = ref(lib, a, npts, shift)
        tot += 1
        bad += int(len(got) != len(exp) or not np.array_equal(got, exp))
    print(f"{tot} cases, {bad} differ from libstdc++")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
