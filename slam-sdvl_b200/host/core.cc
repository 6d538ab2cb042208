// core.cc — SE3, Camera, Config, Point, Feature, Map stand-in, Device and host randomness for the host mirror.
// The SE3 / LDLT arithmetic is the same __host__ __device__ code the kernels use (csrc/common.cuh), which follows
// extra/se3.cc and Eigen's documented algorithms.
#include <atomic>
#include <cmath>
#include <stdexcept>
#include <string>

#include "../csrc/common.cuh"
#include "sdvl_host.h"

namespace sdvl {

// ---------------------------------------------------------------- SE3
static DSE3 ToD(const SE3& s) { double a[7]; s.ToArray(a); return se3_load(a); }
static SE3 FromD(const DSE3& d) { double a[7]; se3_store(d, a); return SE3(a); }

void SE3::GetRotation(double R[9]) const { se3_rot(ToD(*this), R); }
SE3 SE3::Inverse() const { return FromD(se3_inverse(ToD(*this))); }
SE3 SE3::Exp(const double update[6]) { return FromD(se3_exp(update)); }
SE3 SE3::operator*(const SE3& o) const { return FromD(se3_mul(ToD(*this), ToD(o))); }
Eigen::Vector3d SE3::operator*(const Eigen::Vector3d& p) const {
  double x, y, z;
  se3_apply(ToD(*this), p(0), p(1), p(2), x, y, z);
  return Eigen::Vector3d(x, y, z);
}

void SE3::Log(const SE3& se3, double out[6]) {   // se3.cc:96-112,140-164
  const double SMALL_EPS = 1e-10;
  double a[7];
  se3.ToArray(a);
  const double n = std::sqrt(a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
  const double w = a[0];
  double two_atan_nbyw_by_n;
  if (n < SMALL_EPS) two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * w * w);
  else two_atan_nbyw_by_n = 2 * std::atan(n / w) / n;   // the |w|<eps branch is overwritten in the reference (se3.cc:152-160)
  const double theta = two_atan_nbyw_by_n * n;
  const double ox = two_atan_nbyw_by_n * a[1], oy = two_atan_nbyw_by_n * a[2], oz = two_atan_nbyw_by_n * a[3];
  const double Om[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Om[i * 3 + k] * Om[k * 3 + j];
      Om2[i * 3 + j] = s;
    }
  const double c = (theta < SMALL_EPS) ? (1. / 12.) : (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
  double V[9];
  for (int i = 0; i < 9; i++) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * Om[i] + c * Om2[i];
  mat3_mul_vec(V, a[4], a[5], a[6], out[0], out[1], out[2]);
  out[3] = ox; out[4] = oy; out[5] = oz;
}

// ---------------------------------------------------------------- Config / Camera
sdvlb_params& Config::params_() {
  static sdvlb_params p = [] { sdvlb_params q; sdvlb_params_default(&q); return q; }();
  return p;
}
sdvlb_camera& Config::camera_() {
  static sdvlb_camera c = {640, 480, 300.0, 300.0, 320.0, 240.0};   // config.cc:35-40
  return c;
}

Camera::Camera() {
  const sdvlb_camera& c = Config::CameraParams();
  width_ = c.width; height_ = c.height; fx_ = c.fx; fy_ = c.fy; u0_ = c.u0; v0_ = c.v0;
}
void Camera::Project(const Eigen::Vector3d& p, Eigen::Vector2d* o) const {
  (*o)(0) = u0_ + fx_ * p(0) / p(2);
  (*o)(1) = v0_ + fy_ * p(1) / p(2);
}
void Camera::Unproject(const Eigen::Vector2d& p, Eigen::Vector3d* o) const {
  double x = (p(0) - u0_) / fx_, y = (p(1) - v0_) / fy_, z = 1.0;
  const double n = std::sqrt(x * x + y * y + z * z);
  (*o)(0) = x / n; (*o)(1) = y / n; (*o)(2) = z / n;
}

// ---------------------------------------------------------------- Point / Feature / Map
static std::atomic<int> g_point_counter{0};
Point::Point() { id_ = g_point_counter++; }
double Point::GetStd() { return std::sqrt(sigma2_); }
Eigen::Vector3d Point::GetPosition() const {
  if (fixed_) return p3d_;
  SE3 se3 = feature_->GetFrame()->GetWorldPose();
  return se3 * (feature_->GetVector() * (1.0 / rho_));
}
void Point::InitFixed(const std::shared_ptr<Feature>& f, const Eigen::Vector3d& p3d, double rho, double sigma2) {
  feature_ = f;
  p3d_ = p3d;
  rho_ = rho;
  sigma2_ = sigma2;
  fixed_ = true;
}

Feature::Feature(const std::shared_ptr<Frame>& f, const Eigen::Vector2d& p, int l) {
  frame_ = f;
  point_ = nullptr;
  p2d_ = p;
  v_ = frame_->GetCamera()->Unproject(p2d_);
  level_ = l;
}

void Map::DeletePoint(const std::shared_ptr<Point>& point) {
  std::unique_lock<std::mutex> lock(mutex_map_);
  points_trash_.push_back(point);
}
void Map::EmptyTrash() {
  std::vector<std::shared_ptr<Point>> cp;
  {
    std::unique_lock<std::mutex> lock(mutex_map_);
    cp.swap(points_trash_);
  }
  for (auto& p : cp) p->SetDelete();
}

// ---------------------------------------------------------------- Device
static thread_local sdvlb_ctx* t_ctx = nullptr;
static sdvlb_ctx* g_default_ctx = nullptr;
static std::mutex g_ctx_mutex;

sdvlb_ctx* Device::Current() {
  if (t_ctx) return t_ctx;
  std::unique_lock<std::mutex> lock(g_ctx_mutex);
  if (!g_default_ctx) {
    const int rc = sdvlb_ctx_create(0, &Config::Params(), &Config::CameraParams(), &g_default_ctx);
    if (rc) throw std::runtime_error(std::string("sdvl-b200: cannot create CUDA context (no CPU fallback): ") + sdvlb_last_error());
  }
  return g_default_ctx;
}
void Device::SetCurrent(sdvlb_ctx* ctx) { t_ctx = ctx; }

// ---------------------------------------------------------------- randomness (feature_align.cc:53,103,180)
void RandomShuffle(std::vector<int>* v, HostRand* rng) {
  const int size = int(v->size());
  for (int i = 1; i < size; ++i) {
    const int j = rng->Next() % (i + 1);
    if (i != j) std::swap((*v)[i], (*v)[j]);
  }
}

}  // namespace sdvl
