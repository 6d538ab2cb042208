// orb.cu — ORB descriptor mode on the device: ORBDetector::GetOrientation and
// GetDescriptor (extra/orb_detector.cc:350-437) for a list of positions or for every corner of a frame.
//
// One warp per position.  Orientation: the intensity-centroid moments m_10, m_01 over the circular patch of radius
// 15 are integer sums (lane r owns row r - 15), so any summation order is exact; cv::fastAtan2 is OpenCV's degree-7
// polynomial, evaluated here with explicitly rounded float operations (no FMA contraction) so that it returns
// OpenCV's bits.  Descriptor: lane i computes byte i, i.e. the 8 learned tests 8i..8i+7 of csrc/orb_pattern.h
// rotated by the orientation, sample coordinates rounded half-to-even as cvRound does.  cos / sin of the float angle
// are taken in double and rounded to float (the correctly rounded value; glibc's cosf / sinf, which the reference
// calls, return the same float except in rare last-place cases).
#include <cfloat>

#include "common.cuh"
#include "orb_pattern.h"

namespace {

constexpr int kOrbHalf = 15;                 // Config::ORBSize() / 2, orb_size = 31
constexpr int kOrbLimit = kOrbHalf + 4;      // ORBDetector::IsInsideLimits (extra/orb_detector.cc:439-446)
// umax_ of ORBDetector::InitParameters (extra/orb_detector.cc:326-348) for a half patch of 15
__constant__ int8_t c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

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

constexpr int ORB_WARPS = 8;

// positions: n x (x, y, level) in level coordinates
__global__ void __launch_bounds__(ORB_WARPS * 32) orb_positions_kernel(const uint8_t* __restrict__ pyr,
                                                                        const __grid_constant__ PyrGeom G, int levels,
                                                                        const int32_t* __restrict__ xyl, int n,
                                                                        uint8_t* __restrict__ desc, float* __restrict__ angle) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * ORB_WARPS + (threadIdx.x >> 5);
  if (i >= n) return;
  const int x = __ldg(xyl + 3 * i), y = __ldg(xyl + 3 * i + 1), l = __ldg(xyl + 3 * i + 2);
  uint32_t byte = 0;
  float deg = -1.f;
  if (l >= 0 && l < levels && x >= kOrbLimit && x < G.w[l] - kOrbLimit && y >= kOrbLimit && y < G.h[l] - kOrbLimit)
    deg = orb_describe(pyr + G.off[l], G.w[l], x, y, &byte);
  desc[size_t(i) * 32 + lane] = uint8_t(byte);
  if (angle && lane == 0) angle[i] = deg;
}

// every corner of a frame (the lazily filled Frame::descriptors_ of matcher.cc:265-269, all at once)
__global__ void __launch_bounds__(ORB_WARPS * 32) orb_corners_kernel(const FrameDev f, const __grid_constant__ PyrGeom G,
                                                                      int levels, int cap, uint8_t* __restrict__ desc) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * ORB_WARPS + (threadIdx.x >> 5);
  const int n = min(*f.n_corners, cap);
  if (i >= n) return;
  const int4 c = __ldg(f.corners + i);
  uint32_t byte = 0;
  if (c.z >= 0 && c.z < levels && c.x >= kOrbLimit && c.x < G.w[c.z] - kOrbLimit && c.y >= kOrbLimit &&
      c.y < G.h[c.z] - kOrbLimit)
    orb_describe(f.pyr + G.off[c.z], G.w[c.z], c.x, c.y, &byte);
  desc[size_t(i) * 32 + lane] = uint8_t(byte);
}

}  // namespace

cudaError_t sdvlb_launch_orb_positions(const uint8_t* pyr, const PyrGeom& g, int levels, const int32_t* d_xyl, int n,
                                       uint8_t* d_desc, float* d_angle, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  orb_positions_kernel<<<(n + ORB_WARPS - 1) / ORB_WARPS, ORB_WARPS * 32, 0, stream>>>(pyr, g, levels, d_xyl, n, d_desc, d_angle);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_orb_corners(const FrameDev& f, const PyrGeom& g, int levels, int cap, uint8_t* d_desc,
                                     cudaStream_t stream) {
  if (cap <= 0) return cudaSuccess;
  orb_corners_kernel<<<(cap + ORB_WARPS - 1) / ORB_WARPS, ORB_WARPS * 32, 0, stream>>>(f, g, levels, cap, d_desc);
  return cudaGetLastError();
}
