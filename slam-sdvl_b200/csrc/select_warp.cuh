// select_warp.cuh — cv::KeyPointsFilter::retainBest by one WARP, leaving the survivors in exactly the order libstdc++'s
// std::nth_element + std::partition leave them (that order becomes Frame::corners_ order, which breaks ZMSSD ties in
// Matcher::SearchFeatures, matcher.cc:278 -- it is part of the contract; select_impl.h is the serial restatement).
//
// The serial algorithms are sequences of two-sided scans:
//   __unguarded_partition(first, last, pivot):  f runs up to the next element with !(a[f] > pivot), l runs down to the
//       next element with !(pivot > a[l]); if they have not crossed the two are swapped; repeat.  Returns f.
//   std::partition(first, last, pred):          the same with "!pred" from the left and "pred" from the right.
// f only moves right and l only moves left, and a swapped element lands behind the cursor that found it, so neither
// cursor ever meets an element that was moved: the k-th stop of f is the k-th element (ascending) that satisfies the
// left stop condition in the ORIGINAL array, the k-th stop of l is the k-th element (descending) that satisfies the
// right one, and the swaps are exactly the pairs (L[k], R[k]) for k < K = #{k : L[k] < R[k]} (L ascending, R descending:
// the predicate is monotone).  The returned cut is min(L[K], R[K-1]) -- the cursor f stops at the next original left
// stop or, if that lies beyond it, at the element the last swap put at R[K-1].  Both stop lists come from ballots and
// population counts; the K swaps are independent.  What stays serial (one lane) is tiny: the median of three, the final
// insertion sort of at most three elements, and the bookkeeping of nth_element's range.
// (tests: profiles/scripts/select_model.py replays this formulation against libstdc++ on random inputs; the GPU parity
// tests compare whole corner lists, order included, against the oracle.)
#pragma once
#include <stdint.h>

#include "select_impl.h"

namespace sdvlb_sel {

// One two-sided pass over a[lo, hi) by a warp.  stopL / stopR: the stop conditions of the left and right cursor as
// functions of the response.  pos: scratch for 2 * (hi - lo) positions (type P holds any index < hi).
// Performs the swaps; returns K and the totals through nl / nr; L[k] = pos[k], R[k] = pos[m + nr - 1 - k].
template <int SHIFT, typename P, typename FL, typename FR>
__device__ __forceinline__ int warp_two_sided(uint32_t* __restrict__ a, int lo, int hi, FL stopL, FR stopR,
                                              P* __restrict__ pos, int& nl, int& nr) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int m = hi - lo;
  P* const Lp = pos;
  P* const Ra = pos + m;   // right stops in ASCENDING order
  nl = 0; nr = 0;
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    const bool valid = i < hi;
    const uint32_t r = valid ? (a[i] >> SHIFT) : 0u;
    const bool isL = valid && stopL(r), isR = valid && stopR(r);
    const unsigned bl = __ballot_sync(0xffffffffu, isL), br = __ballot_sync(0xffffffffu, isR);
    if (isL) Lp[nl + __popc(bl & lt)] = P(i);
    if (isR) Ra[nr + __popc(br & lt)] = P(i);
    nl += __popc(bl);
    nr += __popc(br);
  }
  __syncwarp();
  const int m2 = min(nl, nr);
  int K = 0;
  for (int k0 = 0; k0 < m2; k0 += 32) {
    const int k = k0 + lane;
    const bool pred = k < m2 && int(Lp[k]) < int(Ra[nr - 1 - k]);
    const unsigned b = __ballot_sync(0xffffffffu, pred);
    K += __popc(b);
    if (b != 0xffffffffu) break;   // monotone: the first chunk that is not full ends it
  }
  for (int k = lane; k < K; k += 32) {
    const int i = Lp[k], j = Ra[nr - 1 - k];
    const uint32_t t = a[i]; a[i] = a[j]; a[j] = t;
  }
  __syncwarp();
  return K;
}

// std::nth_element(a, a + nth, a + n, response greater) by a warp (restates select_impl.h::nth_element_desc).
template <int SHIFT, typename P>
__device__ __forceinline__ void warp_nth_element_desc(uint32_t* __restrict__ a, int nth, int n, P* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  if (n == 0 || nth == n) return;
  int first = 0, last = n;
  int depth_limit = lg2(n) * 2;
  while (last - first > 3) {
    if (depth_limit == 0) {   // introselect's fallback (never seen on FAST scores; kept for exactness)
      if (lane == 0) {
        heap_select<SHIFT>(a + first, nth + 1 - first, last - first);
        swap_u32(a[first], a[nth]);
      }
      __syncwarp();
      return;
    }
    --depth_limit;
    const int mid = first + (last - first) / 2;
    {   // __move_median_to_first(first, first + 1, mid, last - 1): every lane reads, lane 0 swaps
      const int x = first + 1, y = mid, z = last - 1;
      const uint32_t ax = a[x], ay = a[y], az = a[z];
      int with;
      if (greater<SHIFT>(ax, ay)) {
        if (greater<SHIFT>(ay, az)) with = y;
        else if (greater<SHIFT>(ax, az)) with = z;
        else with = x;
      } else if (greater<SHIFT>(ax, az)) with = x;
      else if (greater<SHIFT>(ay, az)) with = z;
      else with = y;
      __syncwarp();
      if (lane == 0) swap_u32(a[first], a[with]);
      __syncwarp();
    }
    const uint32_t p = a[first] >> SHIFT;
    int nl, nr;
    const int lo = first + 1, m = last - lo;
    const int K = warp_two_sided<SHIFT, P>(
        a, lo, last, [p](uint32_t r) { return !(r > p); }, [p](uint32_t r) { return !(p > r); }, pos, nl, nr);
    int cut = 0x7fffffff;
    if (K < nl) cut = min(cut, int(pos[K]));
    if (K > 0) cut = min(cut, int(pos[m + nr - K]));   // R[K - 1]
    __syncwarp();   // pos is rewritten by the next pass
    if (cut <= nth) first = cut;
    else last = cut;
  }
  // __insertion_sort(first, last) on at most three elements
  if (lane == 0) {
    for (int i = first + 1; i < last; ++i) {
      const uint32_t val = a[i];
      if (greater<SHIFT>(val, a[first])) {
        for (int k = i; k > first; --k) a[k] = a[k - 1];
        a[first] = val;
      } else {
        int l = i, next = i - 1;
        while (greater<SHIFT>(val, a[next])) { a[l] = a[next]; l = next; --next; }
        a[l] = val;
      }
    }
  }
  __syncwarp();
}

// cv::KeyPointsFilter::retainBest by a warp: returns the new size (uniform); survivors occupy a[0..ret) in the order the
// serial algorithm leaves them.  a: shared (or global) memory, pos: scratch for 2 * size positions.
template <int SHIFT, typename P>
__device__ __forceinline__ int warp_retain_best(uint32_t* __restrict__ a, int size, int n_points, P* __restrict__ pos) {
  if (!(n_points >= 0 && size > n_points)) return size;
  if (n_points == 0) return 0;
  warp_nth_element_desc<SHIFT, P>(a, n_points - 1, size, pos);
  const uint32_t amb = a[n_points - 1] >> SHIFT;
  // std::partition on [n_points, size) with pred: response >= amb (the ties of the last kept response survive)
  int nl, nr;
  warp_two_sided<SHIFT, P>(
      a, n_points, size, [amb](uint32_t r) { return !(r >= amb); }, [amb](uint32_t r) { return r >= amb; }, pos, nl, nr);
  return n_points + nr;
}

}  // namespace sdvlb_sel
