// compat.h — the few Eigen / OpenCV types that appear in the SDVL class interfaces of the hot path.
//
// When the real headers are installed (as in an SDVL build) they are used.  This image has neither, so a minimal
// stand-in with the same spelling is provided: only what the interfaces of Frame / ImageAlign / FeatureAlign /
// Matcher mention (frame.h:45-123, image_align.h:41, feature_align.h:46-57, matcher.h:45-46).  The implementation
// files only use operator()(int), x()/y()/z(), component constructors, cv::Mat{rows, cols, data, step} — all valid
// on the real types too.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>

#if defined(SDVL_HAVE_EIGEN_OPENCV)
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>
#else

namespace Eigen {
template <typename T, int N>
struct CompatVec {
  T d[N];
  CompatVec() { for (int i = 0; i < N; i++) d[i] = T(0); }
  CompatVec(T a, T b) { static_assert(N == 2, "size"); d[0] = a; d[1] = b; }
  CompatVec(T a, T b, T c) { static_assert(N == 3, "size"); d[0] = a; d[1] = b; d[2] = c; }
  T& operator()(int i) { return d[i]; }
  const T& operator()(int i) const { return d[i]; }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
  T x() const { return d[0]; }
  T y() const { return d[1]; }
  T z() const { static_assert(N >= 3, "size"); return d[2]; }
  CompatVec operator+(const CompatVec& o) const { CompatVec r; for (int i = 0; i < N; i++) r.d[i] = d[i] + o.d[i]; return r; }
  CompatVec operator-(const CompatVec& o) const { CompatVec r; for (int i = 0; i < N; i++) r.d[i] = d[i] - o.d[i]; return r; }
  CompatVec operator*(T s) const { CompatVec r; for (int i = 0; i < N; i++) r.d[i] = d[i] * s; return r; }
  CompatVec operator/(T s) const { CompatVec r; for (int i = 0; i < N; i++) r.d[i] = d[i] / s; return r; }
  T dot(const CompatVec& o) const { T s = 0; for (int i = 0; i < N; i++) s += d[i] * o.d[i]; return s; }
  T squaredNorm() const { return dot(*this); }
  T norm() const;
};
typedef CompatVec<double, 2> Vector2d;
typedef CompatVec<double, 3> Vector3d;
typedef CompatVec<int, 2> Vector2i;
typedef CompatVec<int, 3> Vector3i;
}  // namespace Eigen

#include <cmath>
namespace Eigen {
template <typename T, int N>
T CompatVec<T, N>::norm() const { return T(std::sqrt(double(squaredNorm()))); }
}  // namespace Eigen

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

#ifndef CV_8UC1
#define CV_8UC1 0
#define CV_8U 0
#endif
namespace cv {
// Continuous or strided single-channel u8 image header; owns its buffer only when created with (rows, cols, type).
class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  size_t step = 0;
  Mat() {}
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), step(size_t(c)) {
    owner_.reset(new uint8_t[size_t(r) * c], std::default_delete<uint8_t[]>());
    data = owner_.get();
  }
  Mat(int r, int c, int /*type*/, void* ptr, size_t stp = 0) : rows(r), cols(c), data(static_cast<uint8_t*>(ptr)), step(stp ? stp : size_t(c)) {}
  bool isContinuous() const { return step == size_t(cols); }
  bool empty() const { return data == nullptr; }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; y++) std::memcpy(m.data + size_t(y) * cols, data + size_t(y) * step, size_t(cols));
    return m;
  }
 private:
  std::shared_ptr<uint8_t> owner_;
};
}  // namespace cv
#endif  // SDVL_HAVE_EIGEN_OPENCV
