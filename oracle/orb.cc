// orb.cc — restatement of the ORB descriptor mode (oracle; TEST INFRASTRUCTURE ONLY).
// ORBDetector (extra/orb_detector.cc:326-448): intensity-centroid orientation over a circular patch of radius 15,
// rotated 256-test BRIEF (the learned pattern of csrc/orb_pattern.h), Hamming distance; cv::fastAtan2 as OpenCV
// computes it (call site extra/orb_detector.cc:436).  Config::UseORB() also widens the margins of FAST
// (extra/fast_detector.cc:63-66), FilterCorners (:183-186) and GetCornersInRange (matcher.cc:131-134,199-202) to
// 4 + orb_size/2 and replaces the ZMSSD comparison of SearchFeatures by the descriptor distance with threshold 100
// (matcher.cc:243-277, matcher.h:37).
//
// Every float product below is rounded on its own (volatile temporaries): the reference's build flags let the
// compiler contract a*b+c into an FMA where it likes; this file states the uncontracted arithmetic, which is what the
// reference's code gives when compiled with -ffp-contract=off (oracle/_ref strict build) and what the device computes.
#include <cfloat>

#include "oracle.h"
#include "../slam-sdvl_b200/csrc/orb_pattern.h"

namespace oracle {

static int g_use_orb = 0;
void SetUseOrb(int on) { g_use_orb = on; }
bool UseOrb() { return g_use_orb != 0; }
int BorderMargin(const sdvlb_params& P) { return g_use_orb ? 4 + kOrbSize / 2 : 1 + P.patch_size / 2; }

// cv::fastAtan2 (OpenCV 4.x modules/core/src/mathfuncs_core.simd.hpp, atan_f32): degree-7 odd polynomial on the
// octant-reduced ratio, degrees in [0, 360).  Pinned against cv2.fastAtan2 (tests/test_oracle_cpu.py).
float FastAtan2(float y, float x) {
  const float scale = float(180.0 / 3.1415926535897932384626433832795);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float ax = std::fabs(x), ay = std::fabs(y);
  volatile float a, c, c2, t;
  if (ax >= ay) {
    c = ay / (ax + float(DBL_EPSILON));
    c2 = c * c;
    t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
    a = t * c;
  } else {
    c = ax / (ay + float(DBL_EPSILON));
    c2 = c * c;
    t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
    t = t * c;
    a = 90.f - t;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

static inline int CvRound(double v) { return int(std::nearbyint(v)); }   // cvRound: round half to even

OrbDetector::OrbDetector() {   // ORBDetector::InitParameters (extra/orb_detector.cc:326-348)
  const int half = kOrbSize / 2;
  const int vmax = int(std::floor(half * std::sqrt(2.f) / 2 + 1));
  const int vmin = int(std::ceil(half * std::sqrt(2.f) / 2));
  const double hp2 = half * half;
  umax_.assign(half + 1, 0);
  for (int v = 0; v <= vmax; ++v) umax_[v] = CvRound(std::sqrt(hp2 - v * v));
  for (int v = half, v0 = 0; v >= vmin; --v) {
    while (umax_[v0] == umax_[v0 + 1]) ++v0;
    umax_[v] = v0;
    ++v0;
  }
}

bool OrbDetector::IsInsideLimits(const Mat8& src, int x, int y) const {   // :439-446
  const int m = kOrbSize / 2 + 4;
  return x >= m && x < src.cols - m && y >= m && y < src.rows - m;
}

double OrbDetector::GetOrientation(const Mat8& src, int x, int y) const {   // :412-437
  int m_01 = 0, m_10 = 0;
  const int half = kOrbSize / 2;
  const int step = src.cols;
  const uint8_t* center = src.ptr(y) + x;
  for (int u = -half; u <= half; ++u) m_10 += u * center[u];
  for (int v = 1; v <= half; ++v) {
    int v_sum = 0;
    const int d = umax_[v];
    for (int u = -d; u <= d; ++u) {
      const int val_plus = center[u + v * step], val_minus = center[u - v * step];
      v_sum += (val_plus - val_minus);
      m_10 += u * (val_plus + val_minus);
    }
    m_01 += v * v_sum;
  }
  return FastAtan2(float(m_01), float(m_10));
}

void OrbDetector::GetDescriptor(const Mat8& src, int x, int y, uint8_t desc[32]) const {   // :350-396
  const float factorPI = float(3.1415926535897932384626433832795 / 180.f);
  const int step = src.cols;
  const uint8_t* center = src.ptr(y) + x;
  const float angle = float(GetOrientation(src, x, y) * factorPI);
  const float a = std::cos(angle), b = std::sin(angle);   // float overloads (cosf / sinf)
  const signed char* pat = kOrbPattern31;
  auto sample = [&](int k) {
    const int px = pat[2 * k], py = pat[2 * k + 1];
    volatile float xb = px * b, ya = py * a, xa = px * a, yb = py * b;
    volatile float r = xb + ya, c = xa - yb;
    return int(center[CvRound(r) * step + CvRound(c)]);
  };
  for (int i = 0; i < 32; ++i, pat += 32) {
    int val = 0;
    for (int k = 0; k < 8; k++) val |= (sample(2 * k) < sample(2 * k + 1)) << k;
    desc[i] = uint8_t(val);
  }
}

int OrbDetector::Distance(const uint8_t* a, const uint8_t* b) {   // :399-410
  int dist = 0;
  for (int i = 0; i < 32; i++) dist += __builtin_popcount(unsigned(a[i] ^ b[i]));
  return dist;
}

const std::vector<uint8_t>& CornerDescriptor(Frame* f, int index) {   // lazy fill, matcher.cc:265-269
  if (f->descriptors.size() != f->corners.size()) f->descriptors.resize(f->corners.size());
  std::vector<uint8_t>& d = f->descriptors[size_t(index)];
  if (d.empty()) {
    static const OrbDetector det;
    d.resize(32);
    const Corner& c = f->corners[size_t(index)];
    det.GetDescriptor(f->pyramid[size_t(c.level)], c.x, c.y, d.data());
  }
  return d;
}

}  // namespace oracle

using namespace oracle;

extern "C" {

void orc_set_orb(int on) { SetUseOrb(on); }
float orc_fast_atan2(float y, float x) { return FastAtan2(y, x); }

// ORB descriptors at n positions (x, y, level) of the pyramid of img; also the orientation in degrees.
int orc_orb_descriptors(const sdvlb_params* P, const uint8_t* img, int w, int h, const int32_t* xyl, int n, uint8_t* desc,
                        float* angle) {
  Camera cam{double(w), double(h), 1, 1, 0, 0};
  auto f = MakeFrame(*P, &cam, img, w, h, false, 0);
  const OrbDetector det;
  for (int i = 0; i < n; i++) {
    const int x = xyl[3 * i], y = xyl[3 * i + 1], l = xyl[3 * i + 2];
    if (l < 0 || l >= int(f->pyramid.size()) || !det.IsInsideLimits(f->pyramid[size_t(l)], x, y)) return -2;
    det.GetDescriptor(f->pyramid[size_t(l)], x, y, desc + 32 * size_t(i));
    if (angle) angle[i] = float(det.GetOrientation(f->pyramid[size_t(l)], x, y));
  }
  return 0;
}

}  // extern "C"
