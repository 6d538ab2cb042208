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
// umax_ as 16 nibbles (entry v in bits 4v..4v+3): a table in constant memory would be read at a different address by every
// lane, which the constant cache serialises
constexpr unsigned long long kUmaxNibbles = 0x3689ABCDDEEEFFFFull;
static_assert(((kUmaxNibbles >> 0) & 15) == 15 && ((kUmaxNibbles >> 16) & 15) == 14 && ((kUmaxNibbles >> 28) & 15) == 13 &&
              ((kUmaxNibbles >> 36) & 15) == 12 && ((kUmaxNibbles >> 40) & 15) == 11 && ((kUmaxNibbles >> 44) & 15) == 10 &&
              ((kUmaxNibbles >> 48) & 15) == 9 && ((kUmaxNibbles >> 52) & 15) == 8 && ((kUmaxNibbles >> 56) & 15) == 6 &&
              ((kUmaxNibbles >> 60) & 15) == 3, "umax_ = {15,15,15,15,14,14,14,13,13,12,11,10,9,8,6,3}");

// The 8 learned tests of descriptor byte `lane` (32 bytes of the pattern table) as 8 words, one test per word:
// (x0, y0, x1, y1) signed bytes.  Two 16-byte loads per lane, a contiguous kilobyte per warp.  (The table used to live
// in constant memory: 32 different addresses per load, serialised -- 486 us per 64 frames for orb_frames_kernel.)
__device__ __forceinline__ void orb_load_pattern(uint32_t pw[8]) {
  const uint4* __restrict__ pp = reinterpret_cast<const uint4*>(kOrbPattern31) + (threadIdx.x & 31) * 2;
  const uint4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
  pw[0] = p0.x; pw[1] = p0.y; pw[2] = p0.z; pw[3] = p0.w; pw[4] = p1.x; pw[5] = p1.y; pw[6] = p1.z; pw[7] = p1.w;
}

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
// The three steps of a descriptor.  (1) intensity-centroid moments of the circular patch, by a warp; every lane
// returns both sums.
__device__ __forceinline__ void orb_moments(const uint8_t* __restrict__ center, int W, int& m10_out, int& m01_out) {
  // lane u owns COLUMN u - 15: a warp-wide load reads one row of the patch (31 consecutive bytes, one or two sectors).
  // With a lane per row every load touched 31 different rows -- 31 L1 wavefronts per instruction, ~960 per descriptor,
  // which is what bounded the kernel.  The sums are integer, so any order gives the reference's value.
  const int lane = threadIdx.x & 31;
  const int u = lane - kOrbHalf;
  const int au = u < 0 ? -u : u;
  // All 31 row loads are issued before the first one is used (they are independent; issued four at a time behind
  // their own consumers they cost eight L2 round trips per descriptor).  Every address is inside the image: the corner
  // is at least kOrbLimit = 19 pixels from the border, the loads reach 16.
  int vals[2 * kOrbHalf + 1];
#pragma unroll
  for (int v = -kOrbHalf; v <= kOrbHalf; ++v) vals[v + kOrbHalf] = __ldg(center + v * W + u);
  int col = 0, colv = 0;   // sum of the column's pixels inside the circle, and of v * pixel
#pragma unroll
  for (int v = -kOrbHalf; v <= kOrbHalf; ++v) {
    const int d = int((kUmaxNibbles >> (4 * (v < 0 ? -v : v))) & 15ull);
    const int val = au <= d ? vals[v + kOrbHalf] : 0;
    col += val;
    colv += v * val;
  }
  m10_out = int(__reduce_add_sync(0xffffffffu, unsigned(u * col)));
  m01_out = int(__reduce_add_sync(0xffffffffu, unsigned(colv)));
}
// (2) orientation in degrees (cv::fastAtan2) and the rotation (a, b) = (cos, sin) of it, by one thread: the double-
// precision cos / sin are a few hundred fp64 instructions, which a warp should spend on 32 corners, not on one.
__device__ __forceinline__ float orb_angle(int m10, int m01, float& a, float& b) {
  const float deg = fast_atan2_deg(float(m01), float(m10));
  const float factorPI = float(3.1415926535897932384626433832795 / 180.f);
  const float angle = float(double(deg) * double(factorPI));
  double sn, cs;
  sincos(double(angle), &sn, &cs);
  a = float(cs); b = float(sn);
  return deg;
}
// (3) the 8 rotated tests of descriptor byte `lane`, by a warp.
__device__ __forceinline__ uint32_t orb_sample(const uint8_t* __restrict__ center, int W, float a, float b, const uint32_t pw[8]) {
  int t[16];
#pragma unroll
  for (int k = 0; k < 8; k++) {
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float px = float(int(int8_t(pw[k] >> (16 * j)))), py = float(int(int8_t(pw[k] >> (16 * j + 8))));
      const float r = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
      const float c = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
      t[2 * k + j] = __ldg(center + __float2int_rn(r) * W + __float2int_rn(c));
    }
  }
  uint32_t val = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) val |= uint32_t(t[2 * k] < t[2 * k + 1]) << k;
  return val;
}

// img: one pyramid level (stride = W); (x, y) inside the limits.  Every lane returns the orientation in degrees; lane
// i returns descriptor byte i in *byte_out.  (All three steps by one warp: positions lists, single features.)
__device__ __forceinline__ float orb_describe(const uint8_t* __restrict__ img, int W, int x, int y, uint32_t* byte_out,
                                              const uint32_t pw[8]) {
  const uint8_t* __restrict__ center = img + size_t(y) * W + x;
  int m10, m01;
  orb_moments(center, W, m10, m01);
  float a, b;
  const float deg = orb_angle(m10, m01, a, b);
  *byte_out = orb_sample(center, W, a, b, pw);
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
      y < G.h[level] - kOrbLimit) {
    uint32_t pw[8];
    orb_load_pattern(pw);
    orb_describe(pyr + G.off[level], G.w[level], x, y, &byte, pw);
  }
  return orb_word_of_bytes(byte);
}

}  // namespace sdvlb_orb
