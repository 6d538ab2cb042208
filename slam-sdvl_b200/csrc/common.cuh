// common.cuh — device-side data model and SE3/camera math shared by the sm_100a kernels.
// Math follows extra/se3.cc, camera.cc, extra/utils.cc of the reference (file:line cited per function).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdvl_b200.h"

#if defined(__CUDA_ARCH__)
#define SDVLB_UNROLL _Pragma("unroll")
#elif defined(__CUDACC__)
#define SDVLB_UNROLL
#else
#define SDVLB_UNROLL _Pragma("GCC unroll 8")
#endif

#define SDVLB_MAX_LEVELS 8
#define SDVLB_CELL 32
#define SDVLB_CELL_CAP 176          // > 13*13: upper bound of NMS survivors in a 26x26 tested area
#define SDVLB_MAX_DIM 2048          // x,y packed in 11 bits each by the corner selector

// Geometry of one pyramid (all frames of a context share it).
struct PyrGeom {
  int levels;
  int w[SDVLB_MAX_LEVELS], h[SDVLB_MAX_LEVELS];
  int off[SDVLB_MAX_LEVELS];        // byte offset of each level inside a frame's pyramid allocation (256-B aligned)
  int total;                        // bytes per frame
  // FAST cell grid per level
  int wcells[SDVLB_MAX_LEVELS], hcells[SDVLB_MAX_LEVELS], cell_off[SDVLB_MAX_LEVELS];
  int total_cells;                  // over levels 0..max_fast_levels-1
};

// Device view of one frame.
struct FrameDev {
  uint8_t* pyr;        // levels at PyrGeom::off
  int4* corners;       // corners in the reference's order: (x, y, level, FAST score), level coordinates
  int32_t* n_corners;  // device counter; lives in the 16-byte header right before `corners` (one D2H mirrors both)
  double* pose;        // 7 doubles, world->camera, written by the ImageAlign kernel / uploaded by the host
  int32_t* host_mirror;  // pinned, device-visible host copy of header + corners the selector writes directly
                         // (16-byte header then int4 records), or nullptr
  int32_t mirror_cap;    // corners the host mirror can hold
  int32_t pad_;
  // Spatial index of the corners for Matcher::GetCornersInRange: 32-px cells of the level-0 image (the level-0 FAST
  // grid), corners binned by their level-0 position.  Layout: start[cells + 1], cursor[cells], item[corner_cap].
  int32_t* grid;
  // Config::UseORB(): Frame::descriptors_, 8 words per corner (same order as `corners`), written when the frame is
  // built; nullptr for frames of a context that was not in ORB mode when their slot was allocated.
  uint32_t* desc;
};

// Frames of one build submission, passed to the build kernels BY VALUE (kernel parameter space), so a frame batch
// needs no descriptor upload.
#define SDVLB_BATCH_MAX 64
struct FrameBatch {
  int n;
  int scratch_base;    // index of frame 0 of this batch in the FAST scratch arrays
  FrameDev f[SDVLB_BATCH_MAX];
};
struct ImageBatch {    // level-0 sources of a frame batch (device-visible pinned host memory or device memory)
  const uint8_t* src[SDVLB_BATCH_MAX];
};

struct UndistortArgs {   // cv::undistort with K = (fx, fy, u0, v0), D = (k1, k2, p1, p2, k3)
  int w, h;
  double fx, fy, u0, v0, k1, k2, p1, p2, k3;
};

struct DevParams {
  sdvlb_params p;
  sdvlb_camera cam;
};

// ImageAlign scratch in global memory per feature, used by the features that do not fit the kernel's shared-memory
// cache: 20 doubles (xyz[3], j0[6], j1[6], px[2], gradient moments[3]), 52 floats (patch[16], dx[16], dy[16] + 4 pad)
// and one flag word.
#define SDVLB_ALIGN_SC_DOUBLES 20
#define SDVLB_ALIGN_SC_FLOATS 52
#define SDVLB_ALIGN_SC_BYTES(n) (size_t(n) * (SDVLB_ALIGN_SC_DOUBLES * 8 + SDVLB_ALIGN_SC_FLOATS * 4 + 4))
// One ImageAlign::ComputePose call (device descriptor).
struct AlignJobDev {
  FrameDev ref, cur;
  const sdvlb_align_feat* feats;
  int n;
  int fast;
  double T_ref[7];
  double T_cur[7];
  // outputs
  double* out_pose;        // 7 doubles (also written to cur.pose)
  int32_t* out_info;       // [0] = n_meas of the last ComputeResiduals, [1] = iterations run
  double* out_error;       // GetError()
  int32_t* out_cycles;     // optional: [4] SM cycles in PrecomputePatches / residuals / reduction / solve+update
  sdvlb_gn_iter* trace;    // optional
  int trace_cap;
  int forced_n;
  // teacher forcing (parity tests)
  const double* forced_T;
  const int32_t* forced_iters;
  // scratch
  float* sc_f;             // patch[n*16], dx[n*16], dy[n*16]
  double* sc_d;            // xyz[n*3], j0[n*6], j1[n*6], H[21][n]
  int32_t* sc_flags;       // bit0 visible (sticky), bit1 has a Jacobian at this level
};

// One Matcher::SearchPoint call (device descriptor; frame handles resolved to device pointers).
struct SearchCandDev {
  const uint8_t* ref_pyr;
  double ref_T[7];
  double ref_px[2];
  double ref_v[3];
  double idepth, idepth_std;
  double px[2];
  double pos[3];
  int32_t ref_level;
  int32_t flags;
  int32_t cur_index;       // index into the FrameDev array passed to the kernel
  int32_t pad_;
};

// FAST detection + selection launch parameters (fast.cu).
struct FastArgs {
  PyrGeom g;
  int n_fast_levels;
  int margin;            // 1 + patch_size/2 (fast_detector.cc:66, use_orb == 0)
  int threshold;
  int nfeat[SDVLB_MAX_LEVELS];   // per-level budgets (fast_detector.cc:160-173)
  int corner_cap;
  int level_cap[SDVLB_MAX_LEVELS];   // capacity of the per-level scratch list
  int level_kp_off[SDVLB_MAX_LEVELS];
  int level_kp_total;
  int32_t* overflow_flag;   // pinned, device-visible: set when a fixed capacity was exceeded
  int part_off[SDVLB_MAX_LEVELS + 1];   // selection kernel: first CTA (part) of each level, [n_fast_levels] = total
};
#define SDVLB_TICKET_STRIDE 16   // ints of selection tickets per frame: [0] levels done, [1 + l] parts of level l done,
#define SDVLB_TICKET_OVERFLOW 12 //   [12] a capacity of the frame's selection was exceeded
struct FastPlan {
  FastArgs args;
  int max_cells_level;
  int nfeatures;
};

// 1/d on the serial chains of the Gauss-Newton steps (LDL^T pivots, quaternion normalisation, the Rodrigues factors):
// on the device a MUFU seed (about 20 good bits) and two Newton steps, accurate to an ulp for normal operands, a fifth
// of the fp64 instructions of the IEEE division sequence.  fp64 issue is what bounds those single-thread phases.
__host__ __device__ inline double pivot_rcp(double d) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  return r;
#else
  return 1.0 / d;
#endif
}

// ----------------------------------------------------------------------------- fp64 SE3
struct DSE3 {
  double q0, q1, q2, q3, tx, ty, tz;
};

__host__ __device__ inline DSE3 se3_load(const double* a) {
  DSE3 s;
  s.q0 = a[0]; s.q1 = a[1]; s.q2 = a[2]; s.q3 = a[3]; s.tx = a[4]; s.ty = a[5]; s.tz = a[6];
  return s;
}
__host__ __device__ inline void se3_store(const DSE3& s, double* a) {
  a[0] = s.q0; a[1] = s.q1; a[2] = s.q2; a[3] = s.q3; a[4] = s.tx; a[5] = s.ty; a[6] = s.tz;
}

// Eigen::Quaterniond::toRotationMatrix (used by SE3::GetRotation, extra/se3.h:41)
__host__ __device__ inline void se3_rot(const DSE3& s, double R[9]) {
  const double tx = 2.0 * s.q1, ty = 2.0 * s.q2, tz = 2.0 * s.q3;
  const double twx = tx * s.q0, twy = ty * s.q0, twz = tz * s.q0;
  const double txx = tx * s.q1, txy = ty * s.q1, txz = tz * s.q1;
  const double tyy = ty * s.q2, tyz = tz * s.q2, tzz = tz * s.q3;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

__host__ __device__ inline void mat3_mul_vec(const double R[9], double x, double y, double z, double& ox, double& oy,
                                             double& oz) {
  ox = R[0] * x + R[1] * y + R[2] * z;
  oy = R[3] * x + R[4] * y + R[5] * z;
  oz = R[6] * x + R[7] * y + R[8] * z;
}

// SE3::operator*(Vector3d) (extra/se3.h:68)
__host__ __device__ inline void se3_apply(const DSE3& s, double x, double y, double z, double& ox, double& oy,
                                          double& oz) {
  double R[9];
  se3_rot(s, R);
  mat3_mul_vec(R, x, y, z, ox, oy, oz);
  ox += s.tx; oy += s.ty; oz += s.tz;
}

// SE3::Inverse (extra/se3.cc:59-70)
__host__ __device__ inline DSE3 se3_inverse(const DSE3& s) {
  DSE3 r;
  const double n2 = s.q0 * s.q0 + s.q1 * s.q1 + s.q2 * s.q2 + s.q3 * s.q3;
  if (n2 > 0.0) {
    const double inv = pivot_rcp(n2);
    r.q0 = s.q0 * inv; r.q1 = -s.q1 * inv; r.q2 = -s.q2 * inv; r.q3 = -s.q3 * inv;
  } else {
    r.q0 = r.q1 = r.q2 = r.q3 = 0.0;
  }
  double R[9], x, y, z;
  se3_rot(r, R);
  mat3_mul_vec(R, s.tx, s.ty, s.tz, x, y, z);
  r.tx = -x; r.ty = -y; r.tz = -z;
  return r;
}

// SE3::operator*(SE3) (extra/se3.cc:166-177): quaternion product, normalise, t = t_a + R_a t_b
__host__ __device__ inline DSE3 se3_mul(const DSE3& a, const DSE3& b) {
  DSE3 r;
  const double w = a.q0 * b.q0 - a.q1 * b.q1 - a.q2 * b.q2 - a.q3 * b.q3;
  const double x = a.q0 * b.q1 + a.q1 * b.q0 + a.q2 * b.q3 - a.q3 * b.q2;
  const double y = a.q0 * b.q2 + a.q2 * b.q0 + a.q3 * b.q1 - a.q1 * b.q3;
  const double z = a.q0 * b.q3 + a.q3 * b.q0 + a.q1 * b.q2 - a.q2 * b.q1;
  const double inv = pivot_rcp(sqrt(w * w + x * x + y * y + z * z));
  r.q0 = w * inv; r.q1 = x * inv; r.q2 = y * inv; r.q3 = z * inv;
  double R[9], ox, oy, oz;
  se3_rot(a, R);
  mat3_mul_vec(R, b.tx, b.ty, b.tz, ox, oy, oz);
  r.tx = a.tx + ox; r.ty = a.ty + oy; r.tz = a.tz + oz;
  return r;
}

// SE3::Exp (extra/se3.cc:72-94,114-130); u = [upsilon; omega].  Loop-free: V = I + a*Om + b*Om*Om with the products of
// the skew matrix written out (the same sums the reference's 3x3 product forms).
__host__ __device__ inline DSE3 se3_exp(const double u[6]) {
  const double SMALL_EPS = 1e-10;
  const double ox = u[3], oy = u[4], oz = u[5];
  const double theta = sqrt(ox * ox + oy * oy + oz * oz);
  const double half_theta = 0.5 * theta;
  double imag_factor;
  double sin_half, real_factor;
  sincos(half_theta, &sin_half, &real_factor);
  if (theta < SMALL_EPS) {
    const double theta_sq = theta * theta;
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
  } else {
    imag_factor = sin_half * pivot_rcp(theta);
  }
  DSE3 r;
  r.q0 = real_factor; r.q1 = imag_factor * ox; r.q2 = imag_factor * oy; r.q3 = imag_factor * oz;
  double V[9];
  if (theta < SMALL_EPS) {
    se3_rot(r, V);
  } else {
    double sin_theta, cos_theta;
    sincos(theta, &sin_theta, &cos_theta);
    const double inv_t = pivot_rcp(theta), inv_t2 = inv_t * inv_t;
    const double a = (1 - cos_theta) * inv_t2;
    const double b = (theta - sin_theta) * (inv_t2 * inv_t);
    // Om = [0 -oz oy; oz 0 -ox; -oy ox 0], Om2 = Om * Om
    const double m00 = -(oz * oz) - oy * oy, m01 = oy * ox, m02 = oz * ox;
    const double m11 = -(oz * oz) - ox * ox, m12 = oz * oy;
    const double m22 = -(oy * oy) - ox * ox;
    V[0] = 1.0 + b * m00;          V[1] = a * -oz + b * m01;      V[2] = a * oy + b * m02;
    V[3] = a * oz + b * m01;       V[4] = 1.0 + b * m11;          V[5] = a * -ox + b * m12;
    V[6] = a * -oy + b * m02;      V[7] = a * ox + b * m12;       V[8] = 1.0 + b * m22;
  }
  mat3_mul_vec(V, u[0], u[1], u[2], r.tx, r.ty, r.tz);
  return r;
}

// SE3::Log (extra/se3.cc:96-112,140-164); out = [upsilon; omega]
__host__ __device__ inline void se3_log(const DSE3& s, double out[6]) {
  const double SMALL_EPS = 1e-10;
  const double n = sqrt(s.q1 * s.q1 + s.q2 * s.q2 + s.q3 * s.q3);
  const double w = s.q0;
  double two_atan_nbyw_by_n;
  if (n < SMALL_EPS) two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * w * w);
  else two_atan_nbyw_by_n = 2 * atan(n / w) / n;   // the |w|<eps branch is overwritten in the reference (se3.cc:152-160)
  const double theta = two_atan_nbyw_by_n * n;
  const double ox = two_atan_nbyw_by_n * s.q1, oy = two_atan_nbyw_by_n * s.q2, oz = two_atan_nbyw_by_n * s.q3;
  const double Om[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double a = 0;
      for (int k = 0; k < 3; k++) a += Om[i * 3 + k] * Om[k * 3 + j];
      Om2[i * 3 + j] = a;
    }
  const double c = (theta < SMALL_EPS) ? (1. / 12.) : (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
  double V[9];
  for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * Om[i] + c * Om2[i];
  mat3_mul_vec(V, s.tx, s.ty, s.tz, out[0], out[1], out[2]);
  out[3] = ox; out[4] = oy; out[5] = oz;
}

// Camera::Unproject (camera.cc:74-79): unit bearing of a pixel
__host__ __device__ inline void cam_unproject_unit(const sdvlb_camera& c, double u, double v, double out[3]) {
  const double x = (u - c.u0) / c.fx, y = (v - c.v0) / c.fy, z = 1.0;
  const double n = sqrt(x * x + y * y + z * z);
  out[0] = x / n; out[1] = y / n; out[2] = z / n;
}

// glibc rand() (random_r TYPE_3, 31-word additive feedback) as an explicit state: the reference draws from the
// process-wide rand() (feature_align.cc:53,103,180); every FeatureAlign here owns one stream seeded like srand(1).
__host__ __device__ inline void rand_seed(sdvlb_rand* s, unsigned seed) {   // srandom_r
  if (seed == 0) seed = 1;
  int32_t init[34];
  init[0] = int32_t(seed);
  for (int i = 1; i < 31; i++) {
    const long long hi = init[i - 1] / 127773, lo = init[i - 1] % 127773;
    long long word = 16807 * lo - 2836 * hi;
    if (word < 0) word += 2147483647;
    init[i] = int32_t(word);
  }
  for (int i = 31; i < 34; i++) init[i] = init[i - 31];
  for (int i = 0; i < 34; i++) s->r[i] = uint32_t(init[i]);
  s->n = 34;
  for (int i = 34; i < 344; i++) { s->r[s->n % 34] = s->r[(s->n - 31) % 34] + s->r[(s->n - 3) % 34]; s->n++; }
}
__host__ __device__ inline int rand_next(sdvlb_rand* s) {
  const uint32_t v = s->r[(s->n - 31) % 34] + s->r[(s->n - 3) % 34];
  s->r[s->n % 34] = v;
  s->n = s->n >= 34 * 1000000 ? s->n + 1 - 34 * 999999 : s->n + 1;   // keep the counter bounded (same value mod 34)
  return int(v >> 1);
}

// Unpivoted LDL^T solve of a symmetric positive definite 6x6 system held entirely in registers (every loop has a
// compile-time trip count).  Returns false -- leaving x untouched -- when a pivot is not safely positive; the caller
// then uses ldlt_solve6 below, which reproduces Eigen's pivoted LDLT including its handling of singular systems.  For a
// well-conditioned SPD matrix the two differ only in rounding (~1e-15 relative).
__host__ __device__ inline bool ldlt_solve6_spd(const double A[6][6], const double b[6], double x[6]) {
  double L[6][6], Dinv[6], W[6][6];
  double dmax = 0.0;
SDVLB_UNROLL
  for (int i = 0; i < 6; i++) dmax = fmax(dmax, A[i][i]);
  const double tiny = dmax * 1e-13;
  bool ok = dmax > 0.0;
SDVLB_UNROLL
  for (int k = 0; k < 6; k++) {
    double d = A[k][k];
SDVLB_UNROLL
    for (int j = 0; j < k; j++) d -= L[k][j] * W[k][j];
    ok = ok && (d > tiny);
    const double inv = pivot_rcp(d);
    Dinv[k] = inv;
SDVLB_UNROLL
    for (int i = k + 1; i < 6; i++) {
      double s = A[i][k];
SDVLB_UNROLL
      for (int j = 0; j < k; j++) s -= L[i][j] * W[k][j];
      W[i][k] = s;           // L[i][k] * D[k]
      L[i][k] = s * inv;
    }
  }
  if (!ok) return false;
  double y[6];
SDVLB_UNROLL
  for (int i = 0; i < 6; i++) {
    double s = b[i];
SDVLB_UNROLL
    for (int j = 0; j < i; j++) s -= L[i][j] * y[j];
    y[i] = s;
  }
SDVLB_UNROLL
  for (int i = 0; i < 6; i++) y[i] = y[i] * Dinv[i];
SDVLB_UNROLL
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
SDVLB_UNROLL
    for (int j = i + 1; j < 6; j++) s -= L[j][i] * x[j];
    x[i] = s;
  }
  return true;
}

// Eigen LDLT<Matrix6d>::solve (image_align.cc:102): diagonal-pivoted LDL^T, pseudo-inverse of D.
__host__ __device__ inline void ldlt_solve6(const double Hin[36], const double b[6], double x[6]) {
  const int n = 6;
  double m[36];
  int tr[6];
  for (int i = 0; i < 36; i++) m[i] = Hin[i];
  bool zero_matrix = false;
  for (int k = 0; k < n; ++k) {
    int big = k;
    double best = fabs(m[k * 6 + k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(m[i * 6 + i]) > best) { best = fabs(m[i * 6 + i]); big = i; }
    tr[k] = big;
    if (k != big) {
      for (int j = 0; j < k; ++j) { double t = m[k * 6 + j]; m[k * 6 + j] = m[big * 6 + j]; m[big * 6 + j] = t; }
      for (int i = big + 1; i < n; ++i) { double t = m[i * 6 + k]; m[i * 6 + k] = m[i * 6 + big]; m[i * 6 + big] = t; }
      { double t = m[k * 6 + k]; m[k * 6 + k] = m[big * 6 + big]; m[big * 6 + big] = t; }
      for (int i = k + 1; i < big; ++i) { double t = m[i * 6 + k]; m[i * 6 + k] = m[big * 6 + i]; m[big * 6 + i] = t; }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      double temp[6];
      for (int j = 0; j < k; ++j) temp[j] = m[j * 6 + j] * m[k * 6 + j];
      double s = 0;
      for (int j = 0; j < k; ++j) s += m[k * 6 + j] * temp[j];
      m[k * 6 + k] -= s;
      for (int i = 0; i < rs; ++i) {
        double a = 0;
        for (int j = 0; j < k; ++j) a += m[(k + 1 + i) * 6 + j] * temp[j];
        m[(k + 1 + i) * 6 + k] -= a;
      }
    }
    const double akk = m[k * 6 + k];
    const bool pivot_is_valid = fabs(akk) > 0.0;
    if (k == 0 && !pivot_is_valid) {
      for (int j = 0; j < n; ++j) tr[j] = j;
      zero_matrix = true;
      break;
    }
    if (rs > 0 && pivot_is_valid)
      for (int i = 0; i < rs; ++i) m[(k + 1 + i) * 6 + k] /= akk;
  }
  double d[6];
  for (int i = 0; i < n; i++) d[i] = b[i];
  for (int k = 0; k < n; ++k)
    if (tr[k] != k) { double t = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t; }
  if (!zero_matrix)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) d[i] -= m[i * 6 + j] * d[j];
  const double tol = 2.2250738585072014e-308;  // std::numeric_limits<double>::min()
  for (int i = 0; i < n; ++i) {
    if (fabs(m[i * 6 + i]) > tol) d[i] /= m[i * 6 + i];
    else d[i] = 0.0;
  }
  if (!zero_matrix)
    for (int i = n - 1; i >= 0; --i)
      for (int j = i + 1; j < n; ++j) d[i] -= m[j * 6 + i] * d[j];
  for (int k = n - 1; k >= 0; --k)
    if (tr[k] != k) { double t = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t; }
  for (int i = 0; i < n; i++) x[i] = d[i];
}

// Jacobian3DToPlane, 2x6 (extra/utils.cc:99-118)
__host__ __device__ inline void jacobian3d_to_plane(double x, double y, double z, double J0[6], double J1[6]) {
  const double z_inv = 1. / z;
  const double z_inv_2 = z_inv * z_inv;
  J0[0] = -z_inv;
  J0[1] = 0.0;
  J0[2] = x * z_inv_2;
  J0[3] = y * J0[2];
  J0[4] = -(1.0 + x * J0[2]);
  J0[5] = y * z_inv;
  J1[0] = 0.0;
  J1[1] = -z_inv;
  J1[2] = y * z_inv_2;
  J1[3] = 1.0 + y * J1[2];
  J1[4] = -J0[3];
  J1[5] = -x * z_inv;
}

// One-time per (device, kernel) launch preparation, safe to call from any host thread before every launch: the common
// shared-memory carve-out and the opt-in to `dyn_smem_bytes` of dynamic shared memory (raised when a later call asks for
// more).  cudaFuncSetAttribute applies to the CURRENT device only, so the record is kept per device.
//
// Every kernel of the library asks for the SAME carve-out: an SM runs CTAs of different kernels side by side only if
// they agree on the L1 / shared-memory split; with per-kernel defaults the big-smem kernels (ImageAlign, FeatureAlign)
// and the small-smem ones (FAST, pyramid, SearchPoint) of different streams could not share SMs and the build stream
// serialised against the tracking stream.
#if defined(__CUDACC__)
cudaError_t sdvlb_kernel_prepare_ptr(const void* kernel, int dyn_smem_bytes);
template <typename K>
inline cudaError_t sdvlb_kernel_prepare(K kernel, size_t dyn_smem_bytes = 0) {
  return sdvlb_kernel_prepare_ptr(reinterpret_cast<const void*>(kernel), int(dyn_smem_bytes));
}
// Programmatic dependent launch: a kernel launched with sdvlb_launch_dependent may start while the previous kernel of
// its stream is still running, once every CTA of that kernel has called sdvlb_launch_dependents() (or exited); it must
// call sdvlb_grid_dependency_wait() before it touches anything the previous kernel writes (the wait returns when that
// grid has completed and its memory operations are visible).  The tracking chain of a sequence is three short
// kernels per frame, so the launch latency between them is a visible part of a frame's latency.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void sdvlb_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sdvlb_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
inline void sdvlb_grid_dependency_wait() {}
inline void sdvlb_launch_dependents() {}
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t sdvlb_launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t dyn_smem,
                                          cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = dyn_smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// inside a launcher that returns cudaError_t
#define SDVLB_PREPARE(kernel, dyn)                                             \
  do {                                                                         \
    const cudaError_t pe_ = sdvlb_kernel_prepare(kernel, dyn);                 \
    if (pe_ != cudaSuccess) return pe_;                                        \
  } while (0)
#endif

// error handling shared by the host side
#define SDVLB_CUDA_TRY(expr)                                                   \
  do {                                                                         \
    cudaError_t e_ = (expr);                                                   \
    if (e_ != cudaSuccess) return sdvlb_set_cuda_error(e_, #expr, __FILE__, __LINE__); \
  } while (0)
int sdvlb_set_cuda_error(cudaError_t e, const char* expr, const char* file, int line);
int sdvlb_set_error(int code, const char* msg);
