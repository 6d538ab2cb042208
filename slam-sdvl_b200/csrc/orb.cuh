// orb.cuh -- device functions of the ORB descriptor mode, shared by orb.cu (descriptors of
// positions / of every corner of a frame / of the init features of new map points) and search.cu (the depth-filter
// candidates, whose init-feature descriptor is recomputed from the keyframe on the fly).
//
// One warp per descriptor.  Orientation: the intensity-centroid moments m_10, m_01 over the circular patch of radius
// 15 are integer sums (lane r owns row r - 15), so any summation order is exact; cv::fastAtan2 is OpenCV's degree-7
// polynomial, evaluated here with explicitly rounded float operations (no FMA contraction) so that it returns
// OpenCV's bits.  Descriptor: lane i computes byte i, i.e. the 8 learned tests 8i..8i+7 of csrc/orb_pattern.h
// rotated by the orientation, sample coordinates rounded half-to-even as cvRound does.  cos / sin of the float angle
// are taken in double and rounded to float (the correctly rounded value; glibc's cosf / sinf, which the reference
// calls, return the same float except in rare last-place cases).
#pragma once
#include <cfloat>

#include "common.cuh"
#include "orb_pattern.h"

namespace sdvlb_orb {

constexpr int kOrbHalf = 15;                 // Config::ORBSize() / 2, orb_size = 31
constexpr int kOrbLimit = kOrbHalf + 4;      // ORBDetector::IsInsideLimits (extra/orb_detector.cc:439-446)
// umax_ of ORBDetector::InitParameters (extra/orb_detector.cc:326-348) for a half patch of 15
static __constant__ int8_t c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {   // cv::fastAtan2 (atan_f32)
  const float scale = float(180.0 / 3.1415926535897932384626433832795);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, float(DBL_EPSILON)));
    const float c2 = __fmul_rn(c, c);
    float t = __fadd_rn(__fmul_rn(p7, c2), p5);
    t = __fadd_rn(__fmul_rn(t, c2), p3);
    t = __fadd_rn(__fmul_rn(t, c2), p1);
    a = __fmul_rn(t, c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, float(DBL_EPSILON)));
    const float c2 = __fmul_rn(c, c);
    float t = __fadd_rn(__fmul_rn(p7, c2), p5);
    t = __fadd_rn(__fmul_rn(t, c2), p3);
    t = __fadd_rn(__fmul_rn(t, c2), p1);
    a = __fsub_rn(90.f, __fmul_rn(t, c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// img: one pyramid level (stride = W); (x, y) inside the limits.  Every lane returns the orientation in degrees; lane
// i returns descriptor byte i in *byte_out.
__device__ __forceinline__ float orb_describe(const uint8_t* __restrict__ img, int W, int x, int y, uint32_t* byte_out) {
  const int lane = threadIdx.x & 31;
  const uint8_t* __restrict__ center = img + size_t(y) * W + x;
  int m10 = 0, m01 = 0;
  if (lane < 2 * kOrbHalf + 1) {
    const int v = lane - kOrbHalf;
    const int d = c_umax[v < 0 ? -v : v];
    const uint8_t* __restrict__ row = center + v * W;
    int s = 0;
    for (int u = -d; u <= d; ++u) {
      const int val = __ldg(row + u);
      m10 += u * val;
      s += val;
    }
    m01 = v * s;
  }
  m10 = int(__reduce_add_sync(0xffffffffu, unsigned(m10)));
  m01 = int(__reduce_add_sync(0xffffffffu, unsigned(m01)));
  const float deg = fast_atan2_deg(float(m01), float(m10));
  const float factorPI = float(3.1415926535897932384626433832795 / 180.f);
  const float angle = float(double(deg) * double(factorPI));
  const float a = float(cos(double(angle))), b = float(sin(double(angle)));
  const signed char* pat = kOrbPattern31 + lane * 32;
  uint32_t val = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    int t[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float px = float(pat[4 * k + 2 * j]), py = float(pat[4 * k + 2 * j + 1]);
      const float r = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
      const float c = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
      t[j] = __ldg(center + __float2int_rn(r) * W + __float2int_rn(c));
    }
    val |= uint32_t(t[0] < t[1]) << k;
  }
  *byte_out = val;
  return deg;
}


// Descriptor bytes (one per lane) -> words: every lane returns word (lane & 7) of the 32-byte descriptor, in the
// little-endian layout the byte array has in memory.
__device__ __forceinline__ uint32_t orb_word_of_bytes(uint32_t byte) {
  const int w = (threadIdx.x & 7) * 4;
  const uint32_t b0 = __shfl_sync(0xffffffffu, byte, w), b1 = __shfl_sync(0xffffffffu, byte, w + 1);
  const uint32_t b2 = __shfl_sync(0xffffffffu, byte, w + 2), b3 = __shfl_sync(0xffffffffu, byte, w + 3);
  return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// Feature::descriptor_ of a point's init feature (whole warp): the reference computes it where the feature is created,
// ORBDetector::GetDescriptor(pyramid[level], corner) with corner = feature position / 2^level truncated to integers
// (frame.cc:148-161 + map.cc:319-323; homography_init.cc:141-149), i.e. a function of the keyframe and the feature
// alone.  A feature outside ORBDetector::IsInsideLimits never gets one: its descriptor_ stays the 32 zero bytes of the
// Feature constructor (feature.cc:34-35).  Every lane returns word (lane & 7).
__device__ __forceinline__ uint32_t orb_feature_word(const uint8_t* __restrict__ pyr, const PyrGeom& G, int level, double px0,
                                                     double px1) {
  const int x = int(px0 / double(1 << level)), y = int(px1 / double(1 << level));
  uint32_t byte = 0;
  if (level >= 0 && level < G.levels && x >= kOrbLimit && x < G.w[level] - kOrbLimit && y >= kOrbLimit &&
      y < G.h[level] - kOrbLimit)
    orb_describe(pyr + G.off[level], G.w[level], x, y, &byte);
  return orb_word_of_bytes(byte);
}

}  // namespace sdvlb_orb
