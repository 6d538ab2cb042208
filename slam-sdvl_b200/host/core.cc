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

void Point::InitCandidate(const std::shared_ptr<Feature>& p, double depth) {   // point.cc:48-61
  feature_ = p;
  a_ = 10;
  b_ = 10;
  rho_ = 1.0 / depth;
  sigma2_ = 1.0;
  z_range_ = std::sqrt(sigma2_ * 36);
  cos_alpha_ = 1.0;
  last_distance_ = 1.0 / rho_;
}
void Point::ToSeed(sdvlb_seed* s) const {
  const std::shared_ptr<Frame> rf = feature_->GetFrame();
  std::memset(s, 0, sizeof(*s));
  s->ref_frame = rf->Handle();
  rf->GetPose().ToArray(s->ref_T);
  s->ref_px[0] = feature_->GetPosition()(0); s->ref_px[1] = feature_->GetPosition()(1);
  for (int i = 0; i < 3; i++) s->ref_v[i] = feature_->GetVector()(i);
  s->rho = rho_; s->sigma2 = sigma2_; s->a = a_; s->b = b_; s->z_range = z_range_;
  s->cos_alpha = cos_alpha_; s->last_distance = last_distance_;
  s->ref_level = feature_->GetLevel();
  s->n_failed = n_failed_;
  s->last_kf_id = last_kf_id_;
}
void Point::FromSeed(const sdvlb_seed& s) {
  rho_ = s.rho; sigma2_ = s.sigma2; a_ = s.a; b_ = s.b;
  cos_alpha_ = s.cos_alpha; last_distance_ = s.last_distance;
  n_failed_ = s.n_failed;
  if (s.status == SDVLB_SEED_CONVERGED) {   // Point::HasConverged (point.cc:168-174)
    p3d_ = Eigen::Vector3d(s.p3d[0], s.p3d[1], s.p3d[2]);
    fixed_ = true;
  }
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
// map.cc:397-498.  The reference walks candidates_ and erases as it goes; every candidate's outcome depends only on
// its own state and the frame, so all SearchPoint / triangulation / Point::Update calls go to the device in one batch
// and the walk is replayed over the results (same erase / DeletePoint decisions, same order).  The reference's
// candidates_updating_halt_ early return (map.cc:414-420) belongs to its two-thread protocol and is the caller's.
void Map::UpdateCandidates(const std::shared_ptr<Frame>& frame, double depth_mean, int min_kf_id) {
  std::vector<std::shared_ptr<Point>> kept;
  std::vector<int> slot(candidates_.size(), -1);
  seeds_.clear();
  for (size_t i = 0; i < candidates_.size(); i++) {
    const std::shared_ptr<Point>& p = candidates_[i];
    if (p->ToDelete()) continue;   // map.cc:426-430: DeletePoint + erase below
    slot[i] = int(seeds_.size());
    seeds_.emplace_back();
    p->ToSeed(&seeds_.back());
  }
  sdvlb_seed_params sp;
  sp.depth_mean = depth_mean;
  sp.map_scale = 1.0;          // Config::MapScale()      (config.cc:70)
  sp.scale_min_dist = 0.25;    // Config::ScaleMinDist()  (config.cc:75)
  sp.min_kf_id = min_kf_id;
  sp.pad_ = 0;
  double T[7];
  frame->GetPose().ToArray(T);
  const int rc = sdvlb_update_candidates(frame->Context(), frame->Handle(), T, seeds_.data(), int(seeds_.size()), &sp);
  if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_update_candidates failed: ") + sdvlb_last_error());
  for (size_t i = 0; i < candidates_.size(); i++) {
    const std::shared_ptr<Point>& p = candidates_[i];
    if (slot[i] < 0) { DeletePoint(p); continue; }
    const sdvlb_seed& s = seeds_[slot[i]];
    p->FromSeed(s);
    switch (s.status) {
      case SDVLB_SEED_DELETE_OLD: DeletePoint(p); break;                 // map.cc:433-436: erased
      case SDVLB_SEED_DELETE_FAILED: DeletePoint(p); kept.push_back(p); break;   // map.cc:449-452: deleted, `it++`
      case SDVLB_SEED_CONVERGED: break;                                  // map.cc:486-489: erased, now a fixed point
      default: kept.push_back(p);
    }
  }
  candidates_.swap(kept);
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
