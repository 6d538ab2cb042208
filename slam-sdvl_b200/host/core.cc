// core.cc — SE3, Camera, Config, Point, Feature, Map stand-in, Device and host randomness for the host mirror.
// The SE3 / LDLT arithmetic is the same __host__ __device__ code the kernels use (csrc/common.cuh), which follows
// extra/se3.cc and Eigen's documented algorithms.
#include <atomic>
#include <cstring>
#include <set>
#include <cassert>
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
bool& Config::use_orb_() {
  static bool on = false;   // kUseORB_ (config.cc:79)
  return on;
}

Camera::Camera() {
  const sdvlb_camera& c = Config::CameraParams();
  width_ = c.width; height_ = c.height; fx_ = c.fx; fy_ = c.fy; u0_ = c.u0; v0_ = c.v0;
}
void Camera::SetDistortions(double d0, double d1, double d2, double d3, double d4) {   // camera.cc:38-67
  d_[0] = d0; d_[1] = d1; d_[2] = d2; d_[3] = d3; d_[4] = d4;
  // the reference tests d0 five times (camera.cc:45); the intent -- and cv::undistort's behaviour -- is "all zero"
  has_distortion_ = !(d0 == 0.0 && d1 == 0.0 && d2 == 0.0 && d3 == 0.0 && d4 == 0.0);
}

void Camera::UndistortImage(const cv::Mat& in, cv::Mat* out) const {   // camera.cc:100-105
  if (!has_distortion_) { *out = in.clone(); return; }
  if (in.cols != int(width_) || in.rows != int(height_) || !in.isContinuous())
    throw std::runtime_error("sdvl-b200: UndistortImage needs a continuous image of the camera's size");
  sdvlb_ctx* ctx = Device::Current();
  if (sdvlb_ctx_set_distortion(ctx, d_))
    throw std::runtime_error(std::string("sdvl-b200: sdvlb_ctx_set_distortion failed: ") + sdvlb_last_error());
  cv::Mat dst(in.rows, in.cols, CV_8UC1);
  const int rc = sdvlb_undistort(ctx, in.data, dst.data);
  const double off[5] = {0, 0, 0, 0, 0};
  sdvlb_ctx_set_distortion(ctx, off);   // the caller's Frame() calls get the undistorted image, as in main.cc:133-137
  if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_undistort failed: ") + sdvlb_last_error());
  *out = dst;
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
  if (s.status == SDVLB_SEED_CONVERGED && !fixed_) {   // Point::HasConverged (point.cc:162-174); a fixed point keeps p3d_
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
// map.cc:397-498.  The reference walks candidates_ and erases as it goes; a candidate's outcome depends only on its
// own state and the frame, so the SearchPoint / triangulation / Point::Update calls of a pass go to the device as a
// batch and the walk is replayed over the results (same erase / DeletePoint decisions).  InitCandidates lists every
// non-fixed candidate twice (map.cc:386,393), so the reference applies each observation twice, the second time to the
// state the first left behind: the k-th occurrences of the points form the k-th batch.  The reference's
// candidates_updating_halt_ early return (map.cc:414-420) belongs to its two-thread protocol and is the caller's.
void Map::UpdateCandidates(const std::shared_ptr<Frame>& frame, double depth_mean, int min_kf_id) {
  const size_t n = candidates_.size();
  std::vector<int> rank(n, 0);
  int rounds = 0;
  for (size_t i = 0; i < n; i++) {
    for (size_t j = 0; j < i; j++)
      if (candidates_[j] == candidates_[i]) rank[i]++;
    rounds = std::max(rounds, rank[i] + 1);
  }
  sdvlb_seed_params sp;
  sp.depth_mean = depth_mean;
  sp.map_scale = 1.0;          // Config::MapScale()      (config.cc:70)
  sp.scale_min_dist = 0.25;    // Config::ScaleMinDist()  (config.cc:75)
  sp.min_kf_id = min_kf_id;
  sp.mode = SDVLB_SEEDS_UPDATE;
  double T[7];
  frame->GetPose().ToArray(T);
  seeds_.assign(n, sdvlb_seed());
  for (auto& s : seeds_) s.status = -1;
  std::vector<char> erase(n, 0);
  std::vector<sdvlb_seed> batch;
  std::vector<size_t> who;
  for (int r = 0; r < rounds; r++) {
    batch.clear();
    who.clear();
    for (size_t i = 0; i < n; i++) {
      if (rank[i] != r) continue;
      const std::shared_ptr<Point>& p = candidates_[i];
      if (p->ToDelete()) { DeletePoint(p); erase[i] = 1; continue; }   // map.cc:426-430
      batch.emplace_back();
      p->ToSeed(&batch.back());
      who.push_back(i);
    }
    const int rc = sdvlb_update_candidates(frame->Context(), frame->Handle(), T, batch.data(), int(batch.size()), &sp);
    if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_update_candidates failed: ") + sdvlb_last_error());
    for (size_t k = 0; k < who.size(); k++) {
      const size_t i = who[k];
      const std::shared_ptr<Point>& p = candidates_[i];
      const sdvlb_seed& s = batch[k];
      seeds_[i] = s;
      const bool was_fixed = p->IsFixed();
      p->FromSeed(s);
      switch (s.status) {
        case SDVLB_SEED_DELETE_OLD: DeletePoint(p); erase[i] = 1; break;   // map.cc:433-436
        case SDVLB_SEED_DELETE_FAILED: DeletePoint(p); break;              // map.cc:449-452: deleted, stays listed
        case SDVLB_SEED_CONVERGED: erase[i] = 1; break;                    // map.cc:486-489: now a fixed point
        case SDVLB_SEED_UPDATED: if (was_fixed) erase[i] = 1; break;       // HasConverged() of a fixed point
        default: break;
      }
    }
  }
  std::vector<std::shared_ptr<Point>> kept;
  for (size_t i = 0; i < n; i++)
    if (!erase[i]) kept.push_back(candidates_[i]);
  candidates_.swap(kept);
}

static double Distance2D(const Eigen::Vector2d& a, const Eigen::Vector2d& b) {   // extra/utils.cc:222-226
  const double d1 = a(0) - b(0), d2 = a(1) - b(1);
  return std::sqrt(d1 * d1 + d2 * d2);
}

int Map::InitCandidates(const std::shared_ptr<Frame>& frame, const std::vector<std::shared_ptr<Frame>>& best_kfs,
                        double depth_mean) {
  if (best_kfs.empty()) return 0;
  frame->FilterCorners();
  std::vector<Eigen::Vector3i>& corners = frame->GetCorners();
  std::vector<int>& fcorners = frame->GetFilteredCorners();
  std::vector<char> imatches(fcorners.size(), 0);
  sdvlb_seed_params sp;
  sp.depth_mean = depth_mean;
  sp.map_scale = 1.0;
  sp.scale_min_dist = 0.25;
  sp.min_kf_id = 0;
  sp.mode = SDVLB_SEEDS_INIT;
  double T_frame[7];
  frame->GetPose().ToArray(T_frame);
  int created = 0;
  std::vector<sdvlb_seed> batch;
  std::vector<size_t> who;
  std::vector<std::shared_ptr<Feature>> feats;
  for (const std::shared_ptr<Frame>& cframe : best_kfs) {
    // pure rotation leads to wrong triangulation (map.cc:301-304)
    const double distance = (frame->GetWorldPosition() - cframe->GetWorldPosition()).norm();
    if (distance / depth_mean < 0.01) continue;
    batch.clear(); who.clear(); feats.clear();
    for (size_t count = 0; count < fcorners.size(); count++) {
      if (imatches[count]) continue;
      const Eigen::Vector3i corner = corners[fcorners[count]];
      const int scale = 1 << corner(2);
      auto feature = std::make_shared<Feature>(frame, Eigen::Vector2d(corner(0) * scale, corner(1) * scale), corner(2));
      if (Config::UseORB()) {   // map.cc:319-323: save the descriptor
        std::vector<std::vector<unsigned char>>& descriptors = frame->GetDescriptors();
        assert(!descriptors[size_t(fcorners[count])].empty());
        feature->SetDescriptor(descriptors[size_t(fcorners[count])]);
      }
      sdvlb_seed s;
      std::memset(&s, 0, sizeof(s));
      s.ref_frame = frame->Handle();
      std::memcpy(s.ref_T, T_frame, sizeof(T_frame));
      s.ref_px[0] = feature->GetPosition()(0); s.ref_px[1] = feature->GetPosition()(1);
      for (int i = 0; i < 3; i++) s.ref_v[i] = feature->GetVector()(i);
      s.rho = 1.0 / depth_mean;   // SearchPoint(cframe, feature, 1.0/depth_mean, 1.0, false, ...) (map.cc:320)
      s.sigma2 = 1.0;
      s.a = s.b = 10; s.z_range = 6; s.cos_alpha = 1.0; s.last_distance = depth_mean;
      s.ref_level = corner(2);
      batch.push_back(s);
      who.push_back(count);
      feats.push_back(feature);
    }
    double T_c[7];
    cframe->GetPose().ToArray(T_c);
    const int rc = sdvlb_update_candidates(cframe->Context(), cframe->Handle(), T_c, batch.data(), int(batch.size()), &sp);
    if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_update_candidates failed: ") + sdvlb_last_error());
    for (size_t k = 0; k < batch.size(); k++) {
      const sdvlb_seed& s = batch[k];
      if (s.status < SDVLB_SEED_NO_DEPTH) continue;   // SearchPoint failed
      const Eigen::Vector2d imgpos(s.px[0], s.px[1]);
      const std::shared_ptr<Feature>& feature = feats[k];
      // compare to the 3D points seen from the selected keyframe (map.cc:324-345)
      bool mfound = false;
      for (auto& f2 : cframe->GetFeatures()) {
        if (mfound) break;
        if (!f2) continue;
        std::shared_ptr<Point> point = f2->GetPoint();
        if (!point || point->ToDelete()) continue;
        if (Distance2D(imgpos, f2->GetPosition()) < 1.0) {
          std::unique_lock<std::mutex> lock(mutex_map_);
          feature->SetPoint(point);
          frame->AddFeature(feature);
          mfound = true;
        }
      }
      if (mfound) continue;
      if (s.status != SDVLB_SEED_UPDATED) continue;   // triangulation, parallax, minimum distance (map.cc:348-366)
      auto candidate = std::make_shared<Point>();
      auto feature2 = std::make_shared<Feature>(cframe, imgpos, s.level);
      {
        std::unique_lock<std::mutex> lock(mutex_map_);
        candidate->InitCandidate(feature, s.depth);
        frame->AddFeature(feature);
        feature->SetPoint(candidate);
        cframe->AddFeature(feature2);
        feature2->SetPoint(candidate);
      }
      imatches[who[k]] = 1;
      candidates_.push_back(candidate);   // map.cc:386
      candidates_.push_back(candidate);   // map.cc:393: `fixed` is false, the reference lists the candidate again
      created++;
    }
  }
  return created;
}

int Map::AddConnectionsPoints(const std::shared_ptr<Frame>& frame, const std::vector<std::shared_ptr<Frame>>& best_kfs) {
  if (best_kfs.empty()) return 0;
  // points seen from the connected keyframes and not from this frame, in std::set<shared_ptr<Point>> order (map.cc:570-588)
  std::set<std::shared_ptr<Point>> points;
  for (const auto& kf : best_kfs)
    for (auto& f : kf->GetFeatures()) {
      if (!f) continue;
      std::shared_ptr<Point> point = f->GetPoint();
      if (!point || point->ToDelete()) continue;
      bool seen = false;   // Point::SeenFrom(frame) (point.cc:178-184) through the frame's feature list
      for (auto& ff : frame->GetFeatures())
        if (ff && ff->GetPoint() == point) { seen = true; break; }
      if (!seen) points.insert(point);
    }
  std::vector<std::shared_ptr<Point>> list;
  std::vector<sdvlb_candidate> cands;
  for (const auto& pt : points) {
    std::shared_ptr<Feature> feature = pt->GetInitFeature();
    if (!feature) continue;
    sdvlb_candidate c;
    std::memset(&c, 0, sizeof(c));
    const std::shared_ptr<Frame> rf = feature->GetFrame();
    c.ref_frame = rf->Handle();
    rf->GetPose().ToArray(c.ref_T);
    c.ref_px[0] = feature->GetPosition()(0); c.ref_px[1] = feature->GetPosition()(1);
    for (int i = 0; i < 3; i++) c.ref_v[i] = feature->GetVector()(i);
    c.idepth = pt->GetInverseDepth();
    c.idepth_std = pt->GetStd();
    const Eigen::Vector3d pos = pt->GetPosition();
    for (int i = 0; i < 3; i++) c.pos[i] = pos(i);
    c.ref_level = feature->GetLevel();
    // Project + IsInsideImage(pos, PatchSize) (map.cc:597-602) run on the device (SDVLB_CAND_PROJECT)
    c.flags = SDVLB_CAND_PROJECT | (pt->IsFixed() ? SDVLB_CAND_FIXED : 0);
    cands.push_back(c);
    list.push_back(pt);
  }
  std::vector<sdvlb_match> matches(cands.size());
  double T[7];
  frame->GetPose().ToArray(T);
  std::vector<unsigned char> descs;
  if (Config::UseORB())
    for (const auto& pt : list) {
      const std::vector<unsigned char>& d = pt->GetInitFeature()->GetDescriptor();
      descs.insert(descs.end(), d.begin(), d.end());
    }
  const int rc = Config::UseORB()
      ? sdvlb_search_points_orb(frame->Context(), frame->Handle(), cands.data(), int(cands.size()), T, descs.data(), matches.data())
      : sdvlb_search_points(frame->Context(), frame->Handle(), cands.data(), int(cands.size()), T, matches.data());
  if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_search_points failed: ") + sdvlb_last_error());
  int linked = 0;
  for (size_t i = 0; i < list.size(); i++) {
    if (matches[i].status != SDVLB_MATCH_FOUND) continue;
    std::unique_lock<std::mutex> lock(mutex_map_);
    auto feature = std::make_shared<Feature>(frame, Eigen::Vector2d(matches[i].px[0], matches[i].px[1]), matches[i].level);
    feature->SetPoint(list[i]);
    frame->AddFeature(feature);
    linked++;
  }
  return linked;
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
    if (Config::UseORB() && sdvlb_ctx_set_orb(g_default_ctx, 1))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_ctx_set_orb failed: ") + sdvlb_last_error());
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
