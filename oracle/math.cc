// math.cc — SE3, camera, small utils, 6x6 LDLT, libc randomness (oracle; test infrastructure only).
// Follows extra/se3.cc, extra/utils.cc, camera.cc of the reference; Eigen pieces (quaternion
// algebra, LDLT) are restated from Eigen 3's documented algorithms.
#include <algorithm>
#include <limits>

#include "oracle.h"

namespace oracle {

static const double SMALL_EPS = 1e-10;  // extra/se3.h:30

// Eigen::Quaterniond::toRotationMatrix()
M3 SE3::Rotation() const {
  M3 R;
  const double tx = 2.0 * q1, ty = 2.0 * q2, tz = 2.0 * q3;
  const double twx = tx * q0, twy = ty * q0, twz = tz * q0;
  const double txx = tx * q1, txy = ty * q1, txz = tz * q1;
  const double tyy = ty * q2, tyz = tz * q2, tzz = tz * q3;
  R.m[0][0] = 1.0 - (tyy + tzz); R.m[0][1] = txy - twz;         R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;         R.m[1][1] = 1.0 - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;         R.m[2][1] = tyz + twx;         R.m[2][2] = 1.0 - (txx + tyy);
  return R;
}

static V3 MulM3(const M3& R, const V3& p) {
  return V3(R.m[0][0] * p.x + R.m[0][1] * p.y + R.m[0][2] * p.z,
            R.m[1][0] * p.x + R.m[1][1] * p.y + R.m[1][2] * p.z,
            R.m[2][0] * p.x + R.m[2][1] * p.y + R.m[2][2] * p.z);
}

// se3.cc:59-70 — Quaterniond::inverse() = conjugate / squaredNorm
SE3 SE3::Inverse() const {
  SE3 r;
  const double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
  if (n2 > 0.0) {
    r.q0 = q0 / n2; r.q1 = -q1 / n2; r.q2 = -q2 / n2; r.q3 = -q3 / n2;
  } else {
    r.q0 = r.q1 = r.q2 = r.q3 = 0.0;
  }
  V3 rt = MulM3(r.Rotation(), t);
  r.t = V3(-rt.x, -rt.y, -rt.z);
  return r;
}

// se3.cc:166-177 — quaternion product then normalize()
SE3 SE3::operator*(const SE3& o) const {
  SE3 r;
  double w = q0 * o.q0 - q1 * o.q1 - q2 * o.q2 - q3 * o.q3;
  double x = q0 * o.q1 + q1 * o.q0 + q2 * o.q3 - q3 * o.q2;
  double y = q0 * o.q2 + q2 * o.q0 + q3 * o.q1 - q1 * o.q3;
  double z = q0 * o.q3 + q3 * o.q0 + q1 * o.q2 - q2 * o.q1;
  const double n = std::sqrt(w * w + x * x + y * y + z * z);
  r.q0 = w / n; r.q1 = x / n; r.q2 = y / n; r.q3 = z / n;
  r.t = t + MulM3(Rotation(), o.t);
  return r;
}

V3 SE3::operator*(const V3& p) const { return MulM3(Rotation(), p) + t; }  // se3.h:68

// se3.cc:72-94 with RotationExp se3.cc:114-130
SE3 SE3::Exp(const Vec6 u) {
  const V3 upsilon(u[0], u[1], u[2]);
  const V3 omega(u[3], u[4], u[5]);
  const double theta = omega.norm();
  const double half_theta = 0.5 * theta;
  double imag_factor;
  const double real_factor = std::cos(half_theta);
  if (theta < SMALL_EPS) {
    const double theta_sq = theta * theta;
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
  } else {
    imag_factor = std::sin(half_theta) / theta;
  }
  SE3 r;
  r.q0 = real_factor; r.q1 = imag_factor * omega.x; r.q2 = imag_factor * omega.y; r.q3 = imag_factor * omega.z;

  M3 Om = {{{0, -omega.z, omega.y}, {omega.z, 0, -omega.x}, {-omega.y, omega.x, 0}}};
  M3 V;
  if (theta < SMALL_EPS) {
    V = r.Rotation();
  } else {
    M3 Om2;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Om.m[i][k] * Om.m[k][j];
        Om2.m[i][j] = s;
      }
    const double theta_sq = theta * theta;
    const double a = (1 - std::cos(theta)) / theta_sq;
    const double b = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) V.m[i][j] = (i == j ? 1.0 : 0.0) + a * Om.m[i][j] + b * Om2.m[i][j];
  }
  r.t = MulM3(V, upsilon);
  return r;
}

// se3.cc:96-112 with RotationLog se3.cc:140-164
void SE3::Log(const SE3& s, Vec6 out) {
  const double n = std::sqrt(s.q1 * s.q1 + s.q2 * s.q2 + s.q3 * s.q3);
  const double w = s.q0;
  const double squared_w = w * w;
  double two_atan_nbyw_by_n;
  if (n < SMALL_EPS) {
    two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * squared_w);
  } else {
    // (the |w|<eps branch of the reference is overwritten by the next statement, se3.cc:152-160)
    two_atan_nbyw_by_n = 2 * std::atan(n / w) / n;
  }
  const double theta = two_atan_nbyw_by_n * n;
  const V3 om(two_atan_nbyw_by_n * s.q1, two_atan_nbyw_by_n * s.q2, two_atan_nbyw_by_n * s.q3);
  M3 Om = {{{0, -om.z, om.y}, {om.z, 0, -om.x}, {-om.y, om.x, 0}}};
  M3 Om2;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0;
      for (int k = 0; k < 3; k++) acc += Om.m[i][k] * Om.m[k][j];
      Om2.m[i][j] = acc;
    }
  const double c = (theta < SMALL_EPS) ? (1. / 12.) : (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
  M3 Vinv;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Vinv.m[i][j] = (i == j ? 1.0 : 0.0) - 0.5 * Om.m[i][j] + c * Om2.m[i][j];
  const V3 ups = MulM3(Vinv, s.t);
  out[0] = ups.x; out[1] = ups.y; out[2] = ups.z; out[3] = om.x; out[4] = om.y; out[5] = om.z;
}

// camera.cc:74-79 (Vector3d::normalize)
V3 Camera::Unproject(const V2& p) const {
  V3 r((p.x - u0) / fx, (p.y - v0) / fy, 1.0);
  const double n = r.norm();
  return V3(r.x / n, r.y / n, r.z / n);
}

double AbsMax6(const Vec6 v) {  // utils.cc:28-42
  double max = -1;
  for (int i = 0; i < 6; i++) {
    const double a = std::fabs(v[i]);
    if (a > max) max = a;
  }
  return max;
}

float Interpolate8U(const Mat8& mat, float u, float v) {  // utils.cc:44-59
  const int x = int(std::floor(u));
  const int y = int(std::floor(v));
  const float subpix_x = u - x;
  const float subpix_y = v - y;
  const float w00 = (1.0f - subpix_x) * (1.0f - subpix_y);
  const float w01 = (1.0f - subpix_x) * subpix_y;
  const float w10 = subpix_x * (1.0f - subpix_y);
  const float w11 = 1.0f - w00 - w01 - w10;
  const int stride = mat.cols;
  const uint8_t* ptr = mat.data.data() + y * stride + x;
  return w00 * ptr[0] + w01 * ptr[stride] + w10 * ptr[1] + w11 * ptr[stride + 1];
}

void Jacobian3DToPlane(const V3& p, double J[2][6]) {  // utils.cc:99-118
  const double x = p.x, y = p.y;
  const double z_inv = 1. / p.z;
  const double z_inv_2 = z_inv * z_inv;
  J[0][0] = -z_inv;
  J[0][1] = 0.0;
  J[0][2] = x * z_inv_2;
  J[0][3] = y * J[0][2];
  J[0][4] = -(1.0 + x * J[0][2]);
  J[0][5] = y * z_inv;
  J[1][0] = 0.0;
  J[1][1] = -z_inv;
  J[1][2] = y * z_inv_2;
  J[1][3] = 1.0 + y * J[1][2];
  J[1][4] = -J[0][3];
  J[1][5] = -x * z_inv;
}

double GetMedianVector(std::vector<double>* v) {  // utils.cc:215-220
  auto it = v->begin() + (v->size() / 2);
  std::nth_element(v->begin(), it, v->end());
  return *it;
}

// Eigen LDLT<Matrix6d>: in-place lower-triangular factorisation with diagonal pivoting
// (ldlt_inplace<Lower>::unblocked) and _solve_impl with pseudo-inverse of D.
void LdltSolve6(const Mat6 Hin, const Vec6 b, Vec6 x) {
  const int n = 6;
  double m[6][6];
  int tr[6];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) m[i][j] = Hin[i][j];
  bool zero_matrix = false;
  for (int k = 0; k < n; ++k) {
    int big = k;
    double best = std::fabs(m[k][k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(m[i][i]) > best) { best = std::fabs(m[i][i]); big = i; }
    tr[k] = big;
    if (k != big) {
      for (int j = 0; j < k; ++j) std::swap(m[k][j], m[big][j]);
      for (int i = big + 1; i < n; ++i) std::swap(m[i][k], m[i][big]);
      std::swap(m[k][k], m[big][big]);
      for (int i = k + 1; i < big; ++i) std::swap(m[i][k], m[big][i]);
    }
    const int rs = n - k - 1;
    if (k > 0) {
      double temp[6];
      for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
      double s = 0;
      for (int j = 0; j < k; ++j) s += m[k][j] * temp[j];
      m[k][k] -= s;
      for (int i = 0; i < rs; ++i) {
        double a = 0;
        for (int j = 0; j < k; ++j) a += m[k + 1 + i][j] * temp[j];
        m[k + 1 + i][k] -= a;
      }
    }
    const double akk = m[k][k];
    const bool pivot_is_valid = std::fabs(akk) > 0.0;
    if (k == 0 && !pivot_is_valid) {
      for (int j = 0; j < n; ++j) tr[j] = j;
      zero_matrix = true;
      break;
    }
    if (rs > 0 && pivot_is_valid)
      for (int i = 0; i < rs; ++i) m[k + 1 + i][k] /= akk;
  }
  double d[6];
  for (int i = 0; i < n; i++) d[i] = b[i];
  for (int k = 0; k < n; ++k)
    if (tr[k] != k) std::swap(d[k], d[tr[k]]);
  if (!zero_matrix) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) d[i] -= m[i][j] * d[j];
  }
  const double tol = std::numeric_limits<double>::min();
  for (int i = 0; i < n; ++i) {
    if (std::fabs(m[i][i]) > tol) d[i] /= m[i][i];
    else d[i] = 0.0;
  }
  if (!zero_matrix) {
    for (int i = n - 1; i >= 0; --i)
      for (int j = i + 1; j < n; ++j) d[i] -= m[j][i] * d[j];
  }
  for (int k = n - 1; k >= 0; --k)
    if (tr[k] != k) std::swap(d[k], d[tr[k]]);
  for (int i = 0; i < n; i++) x[i] = d[i];
}

// glibc random_r TYPE_3 (degree 31, separation 3) as used by rand() with the default seed.
void GlibcRand::Seed(unsigned s) {
  if (s == 0) s = 1;
  int32_t init[34];
  init[0] = int32_t(s);
  for (int i = 1; i < 31; i++) {
    const long hi = init[i - 1] / 127773;
    const long lo = init[i - 1] % 127773;
    long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    init[i] = int32_t(word);
  }
  for (int i = 31; i < 34; i++) init[i] = init[i - 31];
  for (int i = 0; i < 34; i++) r[i] = uint32_t(init[i]);
  n = 34;
  for (int i = 34; i < 344; i++) {  // discard the first 310 outputs
    r[n % 34] = r[(n - 31) % 34] + r[(n - 3) % 34];
    n++;
  }
}

int GlibcRand::Next() {
  const uint32_t v = r[(n - 31) % 34] + r[(n - 3) % 34];
  r[n % 34] = v;
  n++;
  return int(v >> 1);
}

// libstdc++ std::random_shuffle(first, last) (bits/stl_algo.h), used at feature_align.cc:53,103
void RandomShuffle(std::vector<int>* v, GlibcRand* rng) {
  const int size = int(v->size());
  for (int i = 1; i < size; ++i) {
    const int j = rng->Next() % (i + 1);
    if (i != j) std::swap((*v)[i], (*v)[j]);
  }
}

}  // namespace oracle
