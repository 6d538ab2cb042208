// select_impl.h — cv::KeyPointsFilter::retainBest on packed keypoints, usable from host and device.
//
// FastDetector::SelectPixels (extra/fast_detector.cc:138-148) calls retainBest once per cell and once per level.
// retainBest (OpenCV 4.x) = std::nth_element(begin, begin+n-1, end, response desc) + std::partition(begin+n, end,
// response >= kps[n-1].response) + resize.  The *order* std leaves the survivors in becomes the order of
// Frame::GetCorners(), which Matcher::SearchFeatures (matcher.cc:278) uses to break ZMSSD ties, so the libstdc++
// (GCC 13) algorithms are restated here step by step: __introselect / __unguarded_partition_pivot /
// __move_median_to_first / __unguarded_partition / __insertion_sort / __heap_select (bits/stl_algo.h, stl_heap.h).
// Keys are u32 with the response in the top bits; `SDVLB_KEY_SHIFT` low bits carry the payload (position).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SDVLB_HD __host__ __device__ __forceinline__
#else
#define SDVLB_HD inline
#endif

namespace sdvlb_sel {

// comp(a, b) of KeypointResponseGreater: a.response > b.response
template <int SHIFT>
SDVLB_HD bool greater(uint32_t a, uint32_t b) { return (a >> SHIFT) > (b >> SHIFT); }

SDVLB_HD void swap_u32(uint32_t& a, uint32_t& b) { const uint32_t t = a; a = b; b = t; }

SDVLB_HD int lg2(int n) {  // std::__lg
  int r = 0;
  while (n > 1) { n >>= 1; r++; }
  return r;
}

template <int SHIFT>
SDVLB_HD void adjust_heap(uint32_t* first, int holeIndex, int len, uint32_t value) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (greater<SHIFT>(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    first[holeIndex] = first[secondChild - 1];
    holeIndex = secondChild - 1;
  }
  // __push_heap
  int parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && greater<SHIFT>(first[parent], value)) {
    first[holeIndex] = first[parent];
    holeIndex = parent;
    parent = (holeIndex - 1) / 2;
  }
  first[holeIndex] = value;
}

template <int SHIFT>
SDVLB_HD void heap_select(uint32_t* first, int middle, int last) {
  // __make_heap(first, middle)
  const int len = middle;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    while (true) {
      const uint32_t value = first[parent];
      adjust_heap<SHIFT>(first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  for (int i = middle; i < last; ++i)
    if (greater<SHIFT>(first[i], first[0])) {
      // __pop_heap(first, middle, i)
      const uint32_t value = first[i];
      first[i] = first[0];
      adjust_heap<SHIFT>(first, 0, middle, value);
    }
}

template <int SHIFT>
SDVLB_HD void nth_element_desc(uint32_t* a, int nth, int n) {
  if (n == 0 || nth == n) return;
  int first = 0, last = n;
  int depth_limit = lg2(n) * 2;
  while (last - first > 3) {
    if (depth_limit == 0) {
      heap_select<SHIFT>(a + first, nth + 1 - first, last - first);
      swap_u32(a[first], a[nth]);
      return;
    }
    --depth_limit;
    // __unguarded_partition_pivot
    const int mid = first + (last - first) / 2;
    {
      const int r = first, x = first + 1, y = mid, z = last - 1;  // __move_median_to_first(result, a, b, c)
      if (greater<SHIFT>(a[x], a[y])) {
        if (greater<SHIFT>(a[y], a[z])) swap_u32(a[r], a[y]);
        else if (greater<SHIFT>(a[x], a[z])) swap_u32(a[r], a[z]);
        else swap_u32(a[r], a[x]);
      } else if (greater<SHIFT>(a[x], a[z])) swap_u32(a[r], a[x]);
      else if (greater<SHIFT>(a[y], a[z])) swap_u32(a[r], a[z]);
      else swap_u32(a[r], a[y]);
    }
    int f = first + 1, l = last;
    const int pivot = first;
    while (true) {  // __unguarded_partition
      while (greater<SHIFT>(a[f], a[pivot])) ++f;
      --l;
      while (greater<SHIFT>(a[pivot], a[l])) --l;
      if (!(f < l)) break;
      swap_u32(a[f], a[l]);
      ++f;
    }
    const int cut = f;
    if (cut <= nth) first = cut;
    else last = cut;
  }
  // __insertion_sort(first, last)
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
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

// Returns the new size; survivors occupy a[0..ret) in the order std:: leaves them.
template <int SHIFT>
SDVLB_HD int retain_best(uint32_t* a, int size, int n_points) {
  if (n_points >= 0 && size > n_points) {
    if (n_points == 0) return 0;
    nth_element_desc<SHIFT>(a, n_points - 1, size);
    const uint32_t amb = a[n_points - 1] >> SHIFT;
    // std::partition (bidirectional) on [n_points, size) with pred: response >= amb
    int first = n_points, last = size;
    while (true) {
      while (true) {
        if (first == last) return first;
        else if ((a[first] >> SHIFT) >= amb) ++first;
        else break;
      }
      --last;
      while (true) {
        if (first == last) return first;
        else if (!((a[last] >> SHIFT) >= amb)) --last;
        else break;
      }
      swap_u32(a[first], a[last]);
      ++first;
    }
  }
  return size;
}

}  // namespace sdvlb_sel
