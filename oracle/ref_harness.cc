// ref_harness.cc — flat C entry points over the REFERENCE's own classes (TEST INFRASTRUCTURE ONLY).
//
// oracle/Makefile (target `ref`) compiles the reference's hot-path sources UNMODIFIED, where they lie under
// /root/reference (image_align.cc, matcher.cc, feature_align.cc, frame.cc, camera.cc, feature.cc, point.cc, map.cc,
// config.cc, extra/{se3,utils,fast_detector,orb_detector}.cc), against the stand-in headers of oracle/ref_shim
// (Eigen and OpenCV are not installed), and links them with this file into oracle/_ref/libsdvlref.so.  The entry
// points mirror oracle/capi.cc one for one (ref_* vs orc_*), so tests/test_oracle_vs_ref.py can push the same seeded
// inputs through the reference's code and through the oracle's restatement and compare.  Nothing here restates
// reference arithmetic: this file only builds Frames / Features / Points, calls the reference's methods and copies
// results out.  `#define private public` is used to read private members (H_, Jres_, inliers_, Config fields).
//
// What is NOT the reference's own code in libsdvlref.so: cv::pyrDown / cv::FAST / retainBest / cv::undistort
// (oracle/ref_cv.cc -> the oracle's restatements, pinned bit-exactly against OpenCV 4.13) and the Eigen primitives
// of oracle/ref_shim/Eigen/Dense (LDLT, small inverses, quaternion conversions: Eigen 3's published formulas).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <Eigen/Dense>
#include <opencv2/core/core.hpp>

#define private public
#define protected public
#include "config.h"
#include "camera.h"
#include "frame.h"
#include "feature.h"
#include "point.h"
#include "image_align.h"
#include "matcher.h"
#include "feature_align.h"
#include "map.h"
#include "extra/bundle.h"
#include "extra/fast_detector.h"
#include "extra/se3.h"
#include "extra/utils.h"
#undef private
#undef protected

#include "../include/sdvl_b200.h"

using sdvl::Config;
using std::shared_ptr;
typedef Eigen::Matrix<double, 6, 1> Vec6;

// extra/bundle.cc needs g2o (not built, out of scope): Map::BundleAdjustment is never reached by the harness.
namespace sdvl {
int Bundle::counter_ = 0;
Bundle::Bundle(Map* map) : map_(map) { std::abort(); }
void Bundle::Local(const std::vector<std::shared_ptr<Frame>>&) { std::abort(); }
}  // namespace sdvl

namespace {

// The reference prints [DEBUG] lines on every frame (feature_align.cc:71-72, sdvl.cc:182,197): silence std::cout.
struct Quiet {
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(nullptr)) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

bool g_use_orb = false;   // Config::UseORB() for the calls that follow (ref_set_orb)

void Configure(const sdvlb_params* P, const sdvlb_camera* cam) {
  Config& c = Config::GetInstance();
  c.kPyramidLevels_ = P->pyramid_levels; c.kCellSize_ = P->cell_size; c.kMaxMatches_ = P->max_matches;
  c.kMaxAlignLevel_ = P->max_align_level; c.kMinAlignLevel_ = P->min_align_level;
  c.kMaxImgAlignIts_ = P->max_img_align_its; c.kAlignPatchSize_ = P->align_patch_size; c.kPatchSize_ = P->patch_size;
  c.kMaxAlignIts_ = P->max_align_its; c.kSearchSize_ = P->search_size; c.kMaxFastLevels_ = P->max_fast_levels;
  c.kFastThreshold_ = P->fast_threshold; c.kNumFeatures_ = P->num_features; c.kMaxFailed_ = P->max_failed;
  c.kMaxOptimPoseIts_ = P->max_optim_pose_its; c.kMaxRansacPoints_ = P->max_ransac_points;
  c.kMaxRansacIts_ = P->max_ransac_its; c.kMinMatches_ = P->min_matches;
  c.kInlierErrorThreshold_ = P->inlier_error_threshold;
  c.kUseORB_ = g_use_orb;
  c.kORBSize_ = 31;
  if (cam) {
    sdvl::CameraParameters& k = c.camera_params_;
    k.width = int(cam->width); k.height = int(cam->height);
    k.fx = cam->fx; k.fy = cam->fy; k.u0 = cam->u0; k.v0 = cam->v0;
    k.d1 = k.d2 = k.d3 = k.d4 = k.d5 = 0.0;
  }
}

sdvl::SE3 ToSE3(const double a[7]) {
  sdvl::SE3 s;
  s.q0_ = a[0]; s.q1_ = a[1]; s.q2_ = a[2]; s.q3_ = a[3];
  s.t_ = Eigen::Vector3d(a[4], a[5], a[6]);
  return s;
}
void FromSE3(const sdvl::SE3& s, double a[7]) {
  a[0] = s.q0_; a[1] = s.q1_; a[2] = s.q2_; a[3] = s.q3_;
  a[4] = s.t_(0); a[5] = s.t_(1); a[6] = s.t_(2);
}

cv::Mat ImageMat(const uint8_t* img, int w, int h) {   // the clone SDVL::HandleFrame hands to Frame (sdvl.cc:59)
  return cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(img)).clone();
}

shared_ptr<sdvl::Frame> MakeFrame(sdvl::Camera* cam, sdvl::ORBDetector* orb, const uint8_t* img, int w, int h,
                                  bool corners, int id) {
  sdvl::Frame::counter_ = id;
  return std::make_shared<sdvl::Frame>(cam, orb, ImageMat(img, w, h), corners);
}

shared_ptr<sdvl::Point> FixedPoint(const shared_ptr<sdvl::Feature>& init, const Eigen::Vector3d& pos) {
  auto pt = std::make_shared<sdvl::Point>();
  pt->a_ = 10; pt->b_ = 10; pt->rho_ = 1.0; pt->sigma2_ = 1.0;
  pt->fixed_ = true;
  pt->p3d_ = pos;
  pt->feature_ = init;
  return pt;
}

}  // namespace

extern "C" {

// Config::ReadParameters (config.cc:88-163) on a cfg file, then the fields sdvlb_params carries are read back:
// checks the reference's own defaults and its config parser against orc_params_default.
int ref_read_config(const char* filename, sdvlb_params* p, sdvlb_camera* cam) {
  Config& c = Config::GetInstance();
  if (filename && filename[0] && !c.ReadParameters(filename)) return -1;
  p->pyramid_levels = Config::PyramidLevels(); p->cell_size = Config::CellSize(); p->max_matches = Config::MaxMatches();
  p->max_align_level = Config::MaxAlignLevel(); p->min_align_level = Config::MinAlignLevel();
  p->max_img_align_its = Config::MaxImgAlignIts(); p->align_patch_size = Config::AlignPatchSize();
  p->patch_size = Config::PatchSize(); p->max_align_its = Config::MaxAlignIts(); p->search_size = Config::SearchSize();
  p->max_fast_levels = Config::MaxFastLevels(); p->fast_threshold = Config::FastThreshold();
  p->num_features = Config::NumFeatures(); p->max_failed = Config::MaxFailed();
  p->max_optim_pose_its = Config::MaxOptimPoseIts(); p->max_ransac_points = Config::MaxRansacPoints();
  p->max_ransac_its = Config::MaxRansacIts(); p->min_matches = Config::MinMatches();
  p->inlier_error_threshold = Config::InlierErrorThreshold();
  if (cam) {
    const sdvl::CameraParameters& k = Config::GetCameraParameters();
    cam->width = k.width; cam->height = k.height; cam->fx = k.fx; cam->fy = k.fy; cam->u0 = k.u0; cam->v0 = k.v0;
  }
  return 0;
}

// Frame::Frame + CreatePyramid (frame.cc:34-56,114-120): levels concatenated, returns total bytes.
int64_t ref_pyramid(const uint8_t* img, int w, int h, int levels, uint8_t* out) {
  sdvlb_params P;
  std::memset(&P, 0, sizeof(P));
  Config& c = Config::GetInstance();
  const int keep = c.kPyramidLevels_;
  c.kPyramidLevels_ = levels;
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto f = MakeFrame(&cam, &orb, img, w, h, false, 0);
  c.kPyramidLevels_ = keep;
  int64_t off = 0;
  for (auto& m : f->GetPyramid()) {
    if (!m.isContinuous()) return -1;
    if (out) std::memcpy(out + off, m.data, size_t(m.cols) * m.rows);
    off += int64_t(m.cols) * m.rows;
  }
  return off;
}

// Frame(img, corners=true) -> GetCorners() (frame.cc:122-131, extra/fast_detector.cc:58-175).
int ref_detect(const sdvlb_params* P, const uint8_t* img, int w, int h, int nfeatures, int32_t* xyl, int cap) {
  sdvlb_params p = *P;
  p.num_features = nfeatures;
  sdvlb_camera k{double(w), double(h), 1, 1, 0, 0};
  Configure(&p, &k);
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto f = MakeFrame(&cam, &orb, img, w, h, true, 0);
  const auto& cs = f->GetCorners();
  const int n = std::min<int>(cap, int(cs.size()));
  for (int i = 0; i < n; i++) { xyl[3 * i] = cs[i](0); xyl[3 * i + 1] = cs[i](1); xyl[3 * i + 2] = cs[i](2); }
  return int(cs.size());
}

// Frame::FilterCorners (frame.cc:133-146) with the frame's features at `locked`.
int ref_filter_corners(const sdvlb_params* P, const uint8_t* img, int w, int h, int nfeatures, const double* locked,
                       int n_locked, int min_feature_score, int32_t* indices, int cap) {
  sdvlb_params p = *P;
  p.num_features = nfeatures;
  sdvlb_camera k{double(w), double(h), 1, 1, 0, 0};
  Configure(&p, &k);
  Config::GetInstance().kMinFeatureScore_ = min_feature_score;
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto f = MakeFrame(&cam, &orb, img, w, h, true, 0);
  for (int i = 0; i < n_locked; i++)
    f->AddFeature(std::make_shared<sdvl::Feature>(f, Eigen::Vector2d(locked[2 * i], locked[2 * i + 1]), 0));
  f->FilterCorners();
  const auto& idx = f->GetFilteredCorners();
  for (int i = 0; i < int(idx.size()) && i < cap; i++) indices[i] = idx[i];
  const int n = int(idx.size());
  f->RemoveFeatures();
  return n;
}
double ref_shi_tomasi(const uint8_t* img, int w, int h, int px, int py) {
  return sdvl::FindShiTomasiScoreAtPoint(cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(img)), px, py);
}

// Camera::SetDistortions + UndistortImage (camera.cc:38-67,100-105).
int ref_undistort(const sdvlb_camera* cam_, const double d[5], const uint8_t* img, int w, int h, uint8_t* out) {
  sdvlb_params P;
  ref_read_config(nullptr, &P, nullptr);
  Configure(&P, cam_);
  sdvl::Camera cam;
  cam.SetDistortions(d[0], d[1], d[2], d[3], d[4]);
  cv::Mat o;
  cam.UndistortImage(cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(img)), &o);
  for (int y = 0; y < h; y++) std::memcpy(out + size_t(y) * w, o.ptr<uchar>(y), size_t(w));
  return 0;
}

// ImageAlign::ComputePose (image_align.cc:46-84) on two images; same contract as orc_image_align.  The trace is
// produced by a second ImageAlign driven through its private ComputeResiduals / H_ / Jres_ with the control flow of
// ComputePose/Optimize replayed here (that replay is harness code; its final pose is checked against the pose the
// reference's own ComputePose returned, and -3 is returned if they differ in any bit).
int ref_image_align(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* ref_img, const uint8_t* cur_img,
                    int w, int h, const sdvlb_align_feat* feats, const double* pos3, int n, const double T_ref[7],
                    double T_cur[7], int fast, int* n_tracked, double* error, sdvlb_gn_iter* trace, int trace_cap,
                    int* trace_n) {
  Quiet q;
  Configure(P, cam_);
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto f1 = MakeFrame(&cam, &orb, ref_img, w, h, false, 0);
  auto f2 = MakeFrame(&cam, &orb, cur_img, w, h, false, 1);
  f1->SetPose(ToSE3(T_ref));
  const sdvl::SE3 prior = ToSE3(T_cur);
  f2->SetPose(prior);
  for (int i = 0; i < n; i++) {
    auto ft = std::make_shared<sdvl::Feature>(f1, nullptr, Eigen::Vector2d(feats[i].px[0], feats[i].px[1]),
                                              Eigen::Vector3d(feats[i].v[0], feats[i].v[1], feats[i].v[2]), 0);
    if (feats[i].valid) ft->SetPoint(FixedPoint(ft, Eigen::Vector3d(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2])));
    f1->AddFeature(ft);
  }
  sdvl::ImageAlign ia;
  const int r = ia.ComputePose(f1, f2, fast != 0);
  FromSE3(f2->GetPose(), T_cur);
  if (n_tracked) *n_tracked = r;
  if (error) *error = ia.GetError();

  int tn = 0, rc = 0;
  if (trace_n && n > 0) {
    sdvl::ImageAlign ib;
    f2->SetPose(prior);
    ib.frame1_ = f1; ib.frame2_ = f2;
    const int area = Config::AlignPatchSize() * Config::AlignPatchSize();
    ib.patch_cache_ = cv::Mat(n, area, CV_32F);
    ib.jacobian_cache_.resize(Eigen::NoChange, n * area);
    ib.visible_fts_.resize(n, false);
    sdvl::SE3 se3 = f2->GetPose() * f1->GetPose().Inverse();
    for (int level = Config::MaxAlignLevel(); level >= Config::MinAlignLevel(); level--) {
      ib.jacobian_cache_.setZero();
      sdvl::SE3 bk = se3;
      for (int i = 0; i < Config::MaxImgAlignIts(); i++) {
        ib.H_.setZero(); ib.Jres_.setZero(); ib.n_meas_ = 0;
        sdvlb_gn_iter rec;
        std::memset(&rec, 0, sizeof(rec));
        rec.level = level; rec.iter = i;
        FromSE3(se3, rec.T_in);
        const double chi2 = ib.ComputeResiduals(se3, level, true, i == 0);
        if (ib.n_meas_ == 0) ib.stop_ = true;
        const Vec6 x = ib.H_.ldlt().solve(ib.Jres_);
        const bool nan = std::isnan(x[0]);
        if (nan) ib.stop_ = true;
        rec.n_meas = int(ib.n_meas_);
        for (int a = 0; a < 6; a++) { rec.b[a] = ib.Jres_(a); rec.x[a] = x(a); for (int b = 0; b < 6; b++) rec.H[a * 6 + b] = ib.H_(a, b); }
        rec.chi2 = chi2;
        rec.flags = (nan ? 2 : 0) | (ib.n_meas_ == 0 ? 4 : 0);
        const bool reject = (i > 0 && chi2 > ib.chi2_) || ib.stop_;
        if (reject) rec.flags |= 1;
        if (trace && tn < trace_cap) trace[tn] = rec;
        tn++;
        if (reject) { se3 = bk; break; }
        bk = se3;
        se3 = se3 * sdvl::SE3::Exp(-x);
        ib.chi2_ = chi2;
        ib.error_ = sdvl::AbsMax(x);
        if (ib.error_ <= 1e-10) break;
      }
      if (fast && ib.error_ > 0.01) break;
    }
    double Tb[7];
    FromSE3(se3 * f1->GetPose(), Tb);
    if (std::memcmp(Tb, T_cur, sizeof(Tb)) != 0) rc = -3;
    f2->SetPose(ToSE3(T_cur));
  }
  if (trace_n) *trace_n = tn;
  f1->RemoveFeatures();
  return rc;
}

// Matcher::SearchPoint (matcher.cc:45-121), + FeatureAlign::ProjectPoint's visibility test (feature_align.cc:323-333)
// when SDVLB_CAND_PROJECT; same contract as orc_search_points (zmssd / n_in_range are oracle-only debug: left -1 / 0).
int ref_search_points(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* cur_img, int w, int h,
                      const double T_cur[7], const uint8_t* const* ref_imgs, int n_refs, const sdvlb_candidate* cands,
                      int n, sdvlb_match* out) {
  Quiet q;
  Configure(P, cam_);
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto cur = MakeFrame(&cam, &orb, cur_img, w, h, true, 1000);
  cur->SetPose(ToSE3(T_cur));
  std::vector<shared_ptr<sdvl::Frame>> refs(n_refs);
  for (int i = 0; i < n_refs; i++) refs[i] = MakeFrame(&cam, &orb, ref_imgs[i], w, h, false, i);
  sdvl::Matcher m(Config::PatchSize());
  for (int i = 0; i < n; i++) {
    const sdvlb_candidate& c = cands[i];
    const int ri = int(reinterpret_cast<intptr_t>(c.ref_frame));
    if (ri < 0 || ri >= n_refs) return -2;
    auto rf = refs[ri];
    rf->SetPose(ToSE3(c.ref_T));
    auto ft = std::make_shared<sdvl::Feature>(rf, nullptr, Eigen::Vector2d(c.ref_px[0], c.ref_px[1]),
                                              Eigen::Vector3d(c.ref_v[0], c.ref_v[1], c.ref_v[2]), c.ref_level);
    if (g_use_orb) {   // the descriptor the feature got where it was created (frame.cc:148-161, map.cc:319-323)
      std::vector<uchar> d(32);
      const Eigen::Vector2i lp = ft->GetLevelPosition().cast<int>();
      if (!orb.IsInsideLimits(rf->GetPyramid()[c.ref_level], lp)) return -3;
      orb.GetDescriptor(rf->GetPyramid()[c.ref_level], lp, &d);
      ft->SetDescriptor(d);
    }
    sdvlb_match& o = out[i];
    std::memset(&o, 0, sizeof(o));
    o.zmssd = -1;
    Eigen::Vector2d px(c.px[0], c.px[1]);
    if (c.flags & SDVLB_CAND_PROJECT) {
      Eigen::Vector2d p;
      if (!cur->Project(Eigen::Vector3d(c.pos[0], c.pos[1], c.pos[2]), &p) ||
          !cam.IsInsideImage(p.cast<int>(), Config::PatchSize())) {
        o.status = SDVLB_MATCH_UNSEEN;
        continue;
      }
      px = p;
    }
    o.proj[0] = px(0); o.proj[1] = px(1);
    int level = 0;
    const bool found = m.SearchPoint(cur, ft, c.idepth, c.idepth_std, (c.flags & SDVLB_CAND_FIXED) != 0, &px, &level);
    o.status = found ? SDVLB_MATCH_FOUND : SDVLB_MATCH_NOT_FOUND;
    if (found) { o.px[0] = px(0); o.px[1] = px(1); o.level = level; }
  }
  return 0;
}

// Matcher::AlignPatch (matcher.cc:359-445) alone; same contract as orc_align_patch.
int ref_align_patch(const sdvlb_params* P, const uint8_t* img, int w, int h, const uint8_t* border_patch, double px[2]) {
  Configure(P, nullptr);
  sdvl::Matcher m(Config::PatchSize());
  uint8_t bp[100], patch[64];
  std::memcpy(bp, border_patch, 100);
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) patch[y * 8 + x] = bp[(y + 1) * 10 + x + 1];
  Eigen::Vector2d p(px[0], px[1]);
  const bool ok = m.AlignPatch(cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(img)), bp, patch, &p);
  px[0] = p(0); px[1] = p(1);
  return ok ? 1 : 0;
}

// FeatureAlign::SelectInliers (mode 0, rand() stream = srand(seed) right before the call) and
// FeatureAlign::OptimizePose(frame) (mode 1) (feature_align.cc:74-83,152-283,341-431); contract of orc_pose_refine.
int ref_pose_refine(const sdvlb_params* P, const sdvlb_camera* cam_, sdvlb_pose_obs* obs, int n, double T[7],
                    unsigned seed, int mode) {
  Quiet q;
  Configure(P, cam_);
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  sdvl::Map map;
  sdvl::FeatureAlign fa(&map, &cam, Config::MaxMatches());
  std::vector<uint8_t> blank(size_t(cam_->width) * size_t(cam_->height), 0);
  Config& c = Config::GetInstance();
  const int keep = c.kPyramidLevels_;
  c.kPyramidLevels_ = 1;
  auto frame = MakeFrame(&cam, &orb, blank.data(), int(cam_->width), int(cam_->height), false, 0);
  c.kPyramidLevels_ = keep;
  frame->SetPose(ToSE3(T));
  std::vector<shared_ptr<sdvl::Feature>> fs;
  std::vector<shared_ptr<sdvl::Point>> points;
  for (int i = 0; i < n; i++) {
    auto ft = std::make_shared<sdvl::Feature>(frame, nullptr, Eigen::Vector2d(0, 0),
                                              Eigen::Vector3d(obs[i].v[0], obs[i].v[1], obs[i].v[2]), obs[i].level);
    points.push_back(FixedPoint(ft, Eigen::Vector3d(obs[i].pos[0], obs[i].pos[1], obs[i].pos[2])));
    ft->SetPoint(points.back());
    fs.push_back(ft);
  }
  int n_in = 0;
  if (mode == 0) {
    srand(seed);
    fa.SelectInliers(frame, fs, &fa.inliers_, &fa.outliers_);
  } else {
    for (int i = 0; i < n; i++) {
      if (obs[i].flags == SDVLB_OBS_INLIER) fa.inliers_.push_back(fs[i]);
      else if (obs[i].flags == SDVLB_OBS_OUTLIER) fa.outliers_.push_back(fs[i]);
    }
    fa.OptimizePose(frame);
    FromSE3(frame->GetPose(), T);
  }
  for (int i = 0; i < n; i++) obs[i].flags = SDVLB_OBS_OUTLIER;
  for (auto& f : fa.inliers_)
    for (int i = 0; i < n; i++) if (fs[i] == f) obs[i].flags = SDVLB_OBS_INLIER;
  n_in = int(fa.inliers_.size());
  for (auto& p : points) p->feature_ = nullptr;   // break the Feature <-> Point cycles
  for (auto& f : fs) f->SetPoint(nullptr);
  frame->RemoveFeatures();
  return n_in;
}

// Map::UpdateCandidates (map.cc:402-498; private) on a Map whose candidates_ are built from `seeds`: the reference's
// own list walk, SearchPoint, triangulation and Point::Update / HasConverged / Unpromote (point.cc:63-116,162-176).
// Written back per seed: the depth-filter state (rho, sigma2, a, b, n_failed, cos_alpha, last_distance, p3d when it
// converged) and `status` = SDVLB_SEED_CONVERGED (fixed and dropped from candidates_), SDVLB_SEED_DELETE_OLD (handed to
// DeletePoint, for whichever reason) or -1 (still a candidate).  depth_mean is Frame::GetSceneDepth() in the
// reference: the current frame is given one fixed point at that depth on its optical axis.
int ref_update_candidates(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* cur_img, int w, int h,
                          const double T_cur[7], const uint8_t* const* ref_imgs, int n_refs, sdvlb_seed* seeds, int n,
                          const sdvlb_seed_params* sp) {
  Quiet q;
  if (sp->mode != SDVLB_SEEDS_UPDATE) return -5;
  Configure(P, cam_);
  Config& c = Config::GetInstance();
  c.kMapScale_ = sp->map_scale;
  c.kScaleMinDist_ = sp->scale_min_dist;
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto cur = MakeFrame(&cam, &orb, cur_img, w, h, true, 1000);
  cur->SetPose(ToSE3(T_cur));
  {   // GetSceneDepth() == depth_mean
    auto anchor = std::make_shared<sdvl::Feature>(cur, Eigen::Vector2d(cam_->u0, cam_->v0), 0);
    anchor->SetPoint(FixedPoint(anchor, cur->GetWorldPose() * Eigen::Vector3d(0, 0, sp->depth_mean)));
    cur->AddFeature(anchor);
  }
  std::vector<shared_ptr<sdvl::Frame>> refs(n_refs);
  std::vector<int> ref_set(n_refs, 0);
  for (int i = 0; i < n_refs; i++) refs[i] = MakeFrame(&cam, &orb, ref_imgs[i], w, h, false, i);
  // tiny frames that only carry a keyframe id (point->GetLastFeature()->GetFrame()->GetKeyframeID(), map.cc:431)
  const int keep = c.kPyramidLevels_;
  c.kPyramidLevels_ = 1;
  std::vector<uint8_t> blank(64, 0);
  std::map<int, shared_ptr<sdvl::Frame>> kf_by_id;
  auto kf_frame = [&](int id) {
    auto it = kf_by_id.find(id);
    if (it != kf_by_id.end()) return it->second;
    auto f = MakeFrame(&cam, &orb, blank.data(), 8, 8, false, 2000 + int(kf_by_id.size()));
    f->SetKeyframeID(id);
    kf_by_id[id] = f;
    return f;
  };
  sdvl::Map map;
  map.last_kf_ = kf_frame(sp->min_kf_id + 2 * Config::MaxSearchKeyframes());
  c.kPyramidLevels_ = keep;
  std::vector<shared_ptr<sdvl::Point>> pts(n);
  int rc = 0;
  for (int i = 0; i < n; i++) {
    const sdvlb_seed& S = seeds[i];
    const int ri = int(reinterpret_cast<intptr_t>(S.ref_frame));
    if (ri < 0 || ri >= n_refs) return -2;
    double Tr[7];
    if (ref_set[ri]) { FromSE3(refs[ri]->GetPose(), Tr); if (std::memcmp(Tr, S.ref_T, sizeof(Tr)) != 0) rc = -4; }
    refs[ri]->SetPose(ToSE3(S.ref_T));
    ref_set[ri] = 1;
    auto ft = std::make_shared<sdvl::Feature>(refs[ri], nullptr, Eigen::Vector2d(S.ref_px[0], S.ref_px[1]),
                                              Eigen::Vector3d(S.ref_v[0], S.ref_v[1], S.ref_v[2]), S.ref_level);
    if (g_use_orb) {   // the descriptor the init feature got at its keyframe (frame.cc:148-161, map.cc:319-323)
      const Eigen::Vector2i lp(int(S.ref_px[0] / (1 << S.ref_level)), int(S.ref_px[1] / (1 << S.ref_level)));
      if (orb.IsInsideLimits(refs[ri]->GetPyramid()[S.ref_level], lp)) {
        std::vector<uchar> d(32);
        orb.GetDescriptor(refs[ri]->GetPyramid()[S.ref_level], lp, &d);
        ft->SetDescriptor(d);
      }
    }
    auto pt = std::make_shared<sdvl::Point>();
    pt->feature_ = ft;
    pt->a_ = S.a; pt->b_ = S.b; pt->rho_ = S.rho; pt->sigma2_ = S.sigma2; pt->z_range_ = S.z_range;
    pt->cos_alpha_ = S.cos_alpha; pt->last_distance_ = S.last_distance; pt->n_failed_ = S.n_failed;
    c.kPyramidLevels_ = 1;
    pt->AddFeature(std::make_shared<sdvl::Feature>(kf_frame(S.last_kf_id), pt, Eigen::Vector2d(0, 0), Eigen::Vector3d(0, 0, 1), 0));
    c.kPyramidLevels_ = keep;
    ft->SetPoint(pt);
    pts[i] = pt;
    map.candidates_.push_back(pt);
  }
  if (rc) return rc;
  map.UpdateCandidates(cur);
  for (int i = 0; i < n; i++) {
    sdvlb_seed& S = seeds[i];
    const auto& pt = pts[i];
    S.rho = pt->rho_; S.sigma2 = pt->sigma2_; S.a = pt->a_; S.b = pt->b_; S.n_failed = pt->n_failed_;
    S.cos_alpha = pt->cos_alpha_; S.last_distance = pt->last_distance_;
    const bool trashed = std::find(map.points_trash_.begin(), map.points_trash_.end(), pt) != map.points_trash_.end();
    const bool listed = std::find(map.candidates_.begin(), map.candidates_.end(), pt) != map.candidates_.end();
    S.status = -1;
    if (pt->fixed_) {
      S.status = SDVLB_SEED_CONVERGED;
      S.p3d[0] = pt->p3d_(0); S.p3d[1] = pt->p3d_(1); S.p3d[2] = pt->p3d_(2);
      if (listed) rc = -6;
    } else if (trashed) {
      S.status = SDVLB_SEED_DELETE_OLD;
    } else if (!listed) {
      rc = -7;
    }
  }
  // break the shared_ptr cycles
  for (auto& pt : pts) { pt->feature_->SetPoint(nullptr); pt->features_.clear(); pt->feature_ = nullptr; }
  cur->RemoveFeatures();
  map.candidates_.clear();
  map.points_trash_.clear();
  map.last_kf_ = nullptr;
  return rc;
}

// Map::InitCandidates (map.cc:262-400; private) on a new keyframe `new_img` connected to one older keyframe `old_img`:
// the reference's own FilterCorners -> per-corner SearchPoint in the connected keyframe -> triangulation -> parallax /
// minimum-distance screens -> Point::InitCandidate, and its candidates_ list (where every candidate is pushed twice,
// map.cc:384,392).  The new keyframe carries one fixed point on its optical axis at depth_mean so that
// Frame::GetSceneDepth() returns it (its cell is locked for FilterCorners, as any feature's is); the old keyframe has
// no features, so the 1-px link test (map.cc:325-343) never fires.  Out, per unique candidate in list order: init
// feature position / level, InitCandidate depth (1/rho), matched position / level in the old keyframe.
int ref_init_candidates(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* new_img, const double T_new[7],
                        const uint8_t* old_img, const double T_old[7], int w, int h, double depth_mean, double map_scale,
                        double scale_min_dist, int min_feature_score, double* ref_px, int32_t* ref_level, double* depth,
                        double* px2, int32_t* level2, int cap, int* n_listed) {
  Quiet q;
  Configure(P, cam_);
  Config& c = Config::GetInstance();
  c.kMapScale_ = map_scale;
  c.kScaleMinDist_ = scale_min_dist;
  c.kMinFeatureScore_ = min_feature_score;
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto frame = MakeFrame(&cam, &orb, new_img, w, h, true, 1);
  auto cframe = MakeFrame(&cam, &orb, old_img, w, h, true, 0);
  frame->SetPose(ToSE3(T_new));
  cframe->SetPose(ToSE3(T_old));
  frame->SetKeyframe();
  cframe->SetKeyframe();
  auto anchor = std::make_shared<sdvl::Feature>(frame, Eigen::Vector2d(cam_->u0, cam_->v0), 0);
  auto anchor_pt = FixedPoint(anchor, frame->GetWorldPose() * Eigen::Vector3d(0, 0, depth_mean));
  anchor->SetPoint(anchor_pt);
  frame->AddFeature(anchor);
  frame->AddConnection(std::make_pair(cframe, 1));
  sdvl::Map map;
  map.InitCandidates(frame);
  if (n_listed) *n_listed = int(map.candidates_.size());
  std::vector<shared_ptr<sdvl::Point>> uniq;
  for (auto& p : map.candidates_)
    if (std::find(uniq.begin(), uniq.end(), p) == uniq.end()) uniq.push_back(p);
  int rc = int(uniq.size());
  for (int i = 0; i < int(uniq.size()) && i < cap; i++) {
    const auto& p = uniq[i];
    const auto& f1 = p->GetInitFeature();
    const auto& f2 = p->GetFeatures().front();   // feature2 was pushed to the front last (map.cc:381)
    if (f1->GetFrame() != frame || f2->GetFrame() != cframe || p->GetFeatures().size() != 2) rc = -8;
    ref_px[2 * i] = f1->GetPosition()(0); ref_px[2 * i + 1] = f1->GetPosition()(1);
    ref_level[i] = f1->GetLevel();
    depth[i] = 1.0 / p->GetInverseDepth();
    px2[2 * i] = f2->GetPosition()(0); px2[2 * i + 1] = f2->GetPosition()(1);
    level2[i] = f2->GetLevel();
  }
  // break the shared_ptr cycles
  for (auto& p : uniq) { for (auto& f : p->GetFeatures()) f->SetPoint(nullptr); p->GetFeatures().clear(); p->feature_ = nullptr; }
  anchor->SetPoint(nullptr);
  anchor_pt->feature_ = nullptr;
  map.candidates_.clear();
  frame->connections_.clear();
  frame->RemoveFeatures();
  cframe->RemoveFeatures();
  return rc;
}

// SDVL::Relocalize's body for one keyframe (sdvl.cc:209-237) with the reference's own ImageAlign (fast = true) and
// FeatureAlign::Reproject (reloc = true); contract of orc_relocalize.
int ref_relocalize(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* kf_img, const uint8_t* cur_img, int w,
                   int h, const sdvlb_align_feat* feats, const double* pos3, const int32_t* levels, int n,
                   const double T_kf[7], double T_out[7], double* error, int32_t out[4]) {
  Quiet q;
  Configure(P, cam_);
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto kf = MakeFrame(&cam, &orb, kf_img, w, h, false, 0);
  auto cur = MakeFrame(&cam, &orb, cur_img, w, h, true, 1);
  kf->SetPose(ToSE3(T_kf));
  cur->SetPose(kf->GetPose());                                   // sdvl.cc:211
  const Eigen::Vector3d C = kf->GetWorldPosition();
  std::vector<shared_ptr<sdvl::Point>> pts;
  for (int i = 0; i < n; i++) {
    auto ft = std::make_shared<sdvl::Feature>(kf, nullptr, Eigen::Vector2d(feats[i].px[0], feats[i].px[1]),
                                              Eigen::Vector3d(feats[i].v[0], feats[i].v[1], feats[i].v[2]), levels[i]);
    auto pt = FixedPoint(ft, Eigen::Vector3d(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]));
    pt->id_ = i;
    pt->rho_ = 1.0 / (pt->p3d_ - C).norm();
    pt->sigma2_ = (0.05 * pt->rho_) * (0.05 * pt->rho_);
    ft->SetPoint(pt);
    kf->AddFeature(ft);
    pts.push_back(pt);
  }
  sdvl::ImageAlign ia;
  ia.ComputePose(kf, cur, true);                                 // sdvl.cc:217
  *error = ia.GetError();
  FromSE3(cur->GetPose(), T_out);
  out[0] = -1; out[1] = 0; out[2] = 0; out[3] = 0;
  if (!(ia.GetError() >= 0.001)) {                               // sdvl.cc:221-222
    srand(1);
    sdvl::Map map;
    sdvl::FeatureAlign fa(&map, &cam, Config::MaxMatches());
    fa.Reproject(cur, kf, kf, true);                             // sdvl.cc:225
    out[0] = fa.GetMatches();
    out[1] = fa.GetAttempts();
  }
  out[2] = cur->GetNumFeatures();
  for (auto& p : pts) { out[3] += p->Score(); p->feature_ = nullptr; }
  for (auto& f : kf->GetFeatures()) f->SetPoint(nullptr);
  kf->RemoveFeatures();
  cur->RemoveFeatures();
  return 0;
}

// Map::AddConnectionsPoints (map.cc:560-617; private): the points seen from a connected keyframe are projected into the
// new keyframe `cur_img`, searched with Matcher::SearchPoint and linked when found.  cands describe the points (init
// feature in the connected keyframe, inverse depth or fixed position); out[i].status = SDVLB_MATCH_FOUND with the linked
// feature's position / level when point i gained a feature in the new keyframe, SDVLB_MATCH_NOT_FOUND otherwise.
int ref_add_connections_points(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* cur_img, const double T_cur[7],
                               const uint8_t* kf_img, const double T_kf[7], int w, int h, const sdvlb_candidate* cands,
                               int n, sdvlb_match* out) {
  Quiet q;
  Configure(P, cam_);
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto kf = MakeFrame(&cam, &orb, kf_img, w, h, false, 0);
  auto frame = MakeFrame(&cam, &orb, cur_img, w, h, true, 1);
  kf->SetPose(ToSE3(T_kf));
  frame->SetPose(ToSE3(T_cur));
  kf->SetKeyframe();
  frame->SetKeyframe();
  std::vector<shared_ptr<sdvl::Point>> pts;
  for (int i = 0; i < n; i++) {
    const sdvlb_candidate& c = cands[i];
    auto ft = std::make_shared<sdvl::Feature>(kf, nullptr, Eigen::Vector2d(c.ref_px[0], c.ref_px[1]),
                                              Eigen::Vector3d(c.ref_v[0], c.ref_v[1], c.ref_v[2]), c.ref_level);
    auto pt = std::make_shared<sdvl::Point>();
    pt->id_ = i;
    pt->feature_ = ft;
    pt->a_ = 10; pt->b_ = 10; pt->z_range_ = 6; pt->cos_alpha_ = 1; pt->last_distance_ = 1.0 / c.idepth;
    pt->rho_ = c.idepth;
    pt->sigma2_ = c.idepth_std * c.idepth_std;
    if (c.flags & SDVLB_CAND_FIXED) { pt->fixed_ = true; pt->p3d_ = Eigen::Vector3d(c.pos[0], c.pos[1], c.pos[2]); }
    pt->AddFeature(ft);
    ft->SetPoint(pt);
    kf->AddFeature(ft);
    pts.push_back(pt);
    std::memset(&out[i], 0, sizeof(out[i]));
    out[i].status = SDVLB_MATCH_NOT_FOUND;
    out[i].zmssd = -1;
  }
  frame->AddConnection(std::make_pair(kf, n));
  sdvl::Map map;
  map.AddConnectionsPoints(frame);
  int linked = 0;
  for (auto& f : frame->GetFeatures()) {
    const auto pt = f->GetPoint();
    if (!pt || pt->GetID() < 0 || pt->GetID() >= n || pts[pt->GetID()] != pt) return -9;
    sdvlb_match& o = out[pt->GetID()];
    if (o.status == SDVLB_MATCH_FOUND) return -10;            // a point linked twice
    o.status = SDVLB_MATCH_FOUND;
    o.px[0] = f->GetPosition()(0); o.px[1] = f->GetPosition()(1);
    o.level = f->GetLevel();
    if (pt->GetFeatures().front() != f) return -11;           // Point::AddFeature pushes to the front (point.h:100-102)
    linked++;
  }
  for (auto& p : pts) { for (auto& f : p->GetFeatures()) f->SetPoint(nullptr); p->GetFeatures().clear(); p->feature_ = nullptr; }
  frame->connections_.clear();
  frame->RemoveFeatures();
  kf->RemoveFeatures();
  return linked;
}

// ---- ORB descriptor mode ------------------------------------------------------------------------------------------
void ref_set_orb(int on) { g_use_orb = on != 0; }
// ORBDetector::GetDescriptor / GetOrientation (extra/orb_detector.cc:350-437) at n positions (x, y, level).
int ref_orb_descriptors(const sdvlb_params* P, const uint8_t* img, int w, int h, const int32_t* xyl, int n, uint8_t* desc,
                        float* angle) {
  sdvlb_camera k{double(w), double(h), 1, 1, 0, 0};
  Configure(P, &k);
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  auto f = MakeFrame(&cam, &orb, img, w, h, false, 0);
  std::vector<uchar> d(32);
  for (int i = 0; i < n; i++) {
    const Eigen::Vector2i p(xyl[3 * i], xyl[3 * i + 1]);
    const int l = xyl[3 * i + 2];
    if (l < 0 || l >= int(f->GetPyramid().size()) || !orb.IsInsideLimits(f->GetPyramid()[l], p)) return -2;
    orb.GetDescriptor(f->GetPyramid()[l], p, &d);
    std::memcpy(desc + 32 * size_t(i), d.data(), 32);
    if (angle) angle[i] = float(orb.GetOrientation(f->GetPyramid()[l], p));
  }
  return 0;
}
int ref_orb_distance(const uint8_t* a, const uint8_t* b) {   // ORBDetector::Distance (:399-410)
  sdvl::ORBDetector orb;
  return orb.Distance(std::vector<uchar>(a, a + 32), std::vector<uchar>(b, b + 32));
}

// ---- primitives -------------------------------------------------------------------------------------------------
void ref_se3_exp(const double u[6], double T[7]) {
  Vec6 v;
  for (int i = 0; i < 6; i++) v(i) = u[i];
  FromSE3(sdvl::SE3::Exp(v), T);
}
void ref_se3_log(const double T[7], double u[6]) {
  const Vec6 v = sdvl::SE3::Log(ToSE3(T));
  for (int i = 0; i < 6; i++) u[i] = v(i);
}
void ref_se3_mul(const double A[7], const double B[7], double C[7]) { FromSE3(ToSE3(A) * ToSE3(B), C); }
void ref_se3_inv(const double A[7], double C[7]) { FromSE3(ToSE3(A).Inverse(), C); }
void ref_se3_apply(const double A[7], const double p[3], double q[3]) {
  const Eigen::Vector3d r = ToSE3(A) * Eigen::Vector3d(p[0], p[1], p[2]);
  q[0] = r(0); q[1] = r(1); q[2] = r(2);
}
void ref_jacobian_3d_to_plane(const double p[3], double J[12]) {   // extra/utils.cc:99-118, row-major 2x6
  Eigen::Matrix<double, 2, 6> M;
  sdvl::Jacobian3DToPlane(Eigen::Vector3d(p[0], p[1], p[2]), &M);
  for (int r = 0; r < 2; r++) for (int c = 0; c < 6; c++) J[r * 6 + c] = M(r, c);
}
float ref_interpolate8u(const uint8_t* img, int w, int h, float u, float v) {   // extra/utils.cc:44-59
  return sdvl::Interpolate8U(cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(img)), u, v);
}

// ---- sequence driver ----------------------------------------------------------------------------------------------
// The RUNNING branch of SDVL::HandleFrame (sdvl.cc:90-95,127) with the reference's own ProcessFrame body
// (sdvl.cc:185-200), motion model (sdvl.cc:266-281), Map::NeedKeyframe / DeletePoint / EmptyTrash (map.cc:165-253).
// HomographyInit and the mapping thread are replaced by the same ground-truth seeding as oracle/tracker.cc, so the two
// drivers see identical maps.  The reference draws from the process-wide rand(): one tracker at a time.
struct RefTracker {
  sdvl::Camera cam;
  sdvl::ORBDetector orb;
  sdvl::Map map;
  std::unique_ptr<sdvl::FeatureAlign> fa;
  shared_ptr<sdvl::Frame> last_frame, last_kf;
  Vec6 vel;
  double plane[4];
  int max_points, kf_every, frame_counter = 0, point_counter = 0;
  double sec_total = 0;
  sdvlb_params P;
  sdvlb_camera K;
  std::vector<std::weak_ptr<sdvl::Point>> all_points;   // to break the Point <-> init Feature cycles at destruction
};

static void SeedKeyframe(RefTracker* t, const shared_ptr<sdvl::Frame>& f, const sdvl::SE3& gt_pose) {
  f->SetKeyframe();
  const int cell = Config::CellSize();
  const int gw = int(std::ceil(t->cam.GetWidth() / cell)), gh = int(std::ceil(t->cam.GetHeight() / cell));
  std::vector<char> occupied(size_t(gw) * gh, 0);
  int n_points = 0;
  for (auto& ft : f->GetFeatures()) {
    if (!ft->GetPoint() || ft->GetPoint()->ToDelete()) continue;
    n_points++;
    const int cx = int(ft->GetPosition()(0) / cell), cy = int(ft->GetPosition()(1) / cell);
    if (cx >= 0 && cx < gw && cy >= 0 && cy < gh) occupied[size_t(cy) * gw + cx] = 1;
  }
  const sdvl::SE3 gt_wc = gt_pose.Inverse();
  const Eigen::Matrix3d Rwc = gt_wc.GetRotation();
  const Eigen::Vector3d C = gt_wc.GetTranslation();
  const Eigen::Vector3d est_C = f->GetWorldPosition();
  const auto& corners = f->GetCorners();
  const int n = int(corners.size());
  const int margin = Config::PatchSize() / 2 + 2;
  const Eigen::Vector3d nrm(t->plane[0], t->plane[1], t->plane[2]);
  for (int i = 0; i < n && n_points < t->max_points; i++) {
    const Eigen::Vector3i c = corners[size_t((long long)i * 7919 % n)];
    const cv::Mat& lvl = f->GetPyramid()[c(2)];
    if (c(0) < margin || c(1) < margin || c(0) >= lvl.cols - margin || c(1) >= lvl.rows - margin) continue;
    const Eigen::Vector2d px(double(c(0) * (1 << c(2))), double(c(1) * (1 << c(2))));
    const int cx = int(px(0) / cell), cy = int(px(1) / cell);
    if (occupied[size_t(cy) * gw + cx]) continue;
    auto ft = std::make_shared<sdvl::Feature>(f, px, c(2));
    if (g_use_orb) {   // the descriptor an init feature gets at its keyframe (frame.cc:148-161, map.cc:319-323)
      const Eigen::Vector2i lp(c(0), c(1));
      if (!t->orb.IsInsideLimits(lvl, lp)) continue;
      std::vector<uchar> d(32);
      t->orb.GetDescriptor(lvl, lp, &d);
      ft->SetDescriptor(d);
    }
    const Eigen::Vector3d dir = Rwc * ft->GetVector();
    const double denom = nrm.dot(dir);
    if (std::fabs(denom) < 1e-9) continue;
    const double s = (t->plane[3] - nrm.dot(C)) / denom;
    if (s <= 0) continue;
    auto pt = FixedPoint(ft, C + dir * s);
    pt->id_ = t->point_counter++;
    const double depth = (pt->p3d_ - est_C).norm();
    pt->rho_ = 1.0 / depth;
    pt->sigma2_ = (0.05 * pt->rho_) * (0.05 * pt->rho_);
    ft->SetPoint(pt);
    t->all_points.push_back(pt);
    f->AddFeature(ft);
    occupied[size_t(cy) * gw + cx] = 1;
    n_points++;
  }
  t->last_kf = f;
  t->map.last_kf_ = f;
}

void* ref_tracker_create(const sdvlb_params* P, const sdvlb_camera* cam, const double plane[4], int max_points,
                         int kf_every) {
  Configure(P, cam);
  Config::GetInstance().kMinKeyframeIts_ = kf_every;
  Config::GetInstance().kLostRatio_ = 0.7;
  srand(1);   // glibc's state before the first rand() of a process; FeatureAlign's constructor shuffles first
  RefTracker* t = new RefTracker;
  t->P = *P; t->K = *cam;
  t->fa.reset(new sdvl::FeatureAlign(&t->map, &t->cam, Config::MaxMatches()));   // sdvl.cc:38
  t->vel.setZero();
  for (int i = 0; i < 4; i++) t->plane[i] = plane[i];
  t->max_points = max_points;
  t->kf_every = kf_every;
  return t;
}
void ref_tracker_destroy(void* h) {
  RefTracker* t = static_cast<RefTracker*>(h);
  // break the Point <-> init Feature -> keyframe shared_ptr cycles of every point this tracker created
  for (auto& w : t->all_points)
    if (auto p = w.lock()) p->feature_ = nullptr;
  if (t->last_frame) t->last_frame->RemoveFeatures();
  t->last_frame = nullptr;
  t->last_kf = nullptr;
  t->map.last_kf_ = nullptr;
  delete t;
}
int ref_tracker_step(void* h, const uint8_t* img, int w, int h_, const double gt_pose[7], double est_pose[7],
                     int32_t stats[8]) {
  RefTracker* t = static_cast<RefTracker*>(h);
  Quiet q;
  Configure(&t->P, &t->K);
  Config::GetInstance().kMinKeyframeIts_ = t->kf_every;
  const auto t0 = std::chrono::steady_clock::now();
  int st[8] = {0, 0, 0, 0, 0, 0, -1, 0};
  auto frame = MakeFrame(&t->cam, &t->orb, img, w, h_, true, t->frame_counter);   // sdvl.cc:59
  const sdvl::SE3 gt = ToSE3(gt_pose);
  if (!t->last_frame) {
    frame->SetPose(gt);
    SeedKeyframe(t, frame, gt);
    st[7] = 1;
  } else {
    frame->SetPose(sdvl::SE3::Exp(t->vel) * t->last_frame->GetPose());   // SetMotionModel, sdvl.cc:278-281
    {
      sdvl::ImageAlign image_align;                                       // ProcessFrame, sdvl.cc:185-200
      st[0] = image_align.ComputePose(t->last_frame, frame);
    }
    t->fa->Reproject(frame, t->last_frame, t->last_kf);
    st[1] = t->fa->GetMatches();
    st[2] = t->fa->GetAttempts();
    t->fa->OptimizePose(frame);
    st[3] = int(t->fa->inliers_.size());
    st[4] = int(t->fa->outliers_.size());
    const sdvl::SE3 mov = frame->GetPose() * t->last_frame->GetPose().Inverse();   // GetMotionModel, sdvl.cc:266-276
    const Vec6 vel = sdvl::SE3::Log(mov);
    const Vec6 old_vel = t->vel;
    t->vel = 0.9 * (0.5 * vel + 0.5 * old_vel);
    if (t->map.NeedKeyframe(frame, st[1])) {   // map.cc:170-188
      SeedKeyframe(t, frame, gt);
      st[7] = 1;
    }
  }
  int nf = 0;
  for (auto& ft : frame->GetFeatures())
    if (ft->GetPoint() && !ft->GetPoint()->ToDelete()) nf++;
  st[5] = nf;
  if (t->last_frame) t->last_frame->RemoveFeatures();
  t->last_frame = frame;
  t->frame_counter++;
  t->map.EmptyTrash();   // sdvl.cc:127
  FromSE3(frame->GetPose(), est_pose);
  t->sec_total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) for (int i = 0; i < 8; i++) stats[i] = st[i];
  return 0;
}
double ref_tracker_run(void* h, const uint8_t* imgs, int n, int w, int h_, const double* gt_poses, double* est_poses,
                       int32_t* stats) {
  RefTracker* t = static_cast<RefTracker*>(h);
  const double before = t->sec_total;
  for (int i = 0; i < n; i++)
    ref_tracker_step(h, imgs + size_t(i) * w * h_, w, h_, gt_poses + 7 * i, est_poses + 7 * i, stats ? stats + 8 * i : nullptr);
  return t->sec_total - before;
}
int ref_tracker_last_features(void* h, double* px, int32_t* level, int32_t* pid, int cap) {
  RefTracker* t = static_cast<RefTracker*>(h);
  if (!t->last_frame) return 0;
  int n = 0;
  for (auto& ft : t->last_frame->GetFeatures()) {
    if (!ft->GetPoint() || ft->GetPoint()->ToDelete()) continue;
    if (n < cap) { px[2 * n] = ft->GetPosition()(0); px[2 * n + 1] = ft->GetPosition()(1); level[n] = ft->GetLevel(); pid[n] = ft->GetPoint()->GetID(); }
    n++;
  }
  return n;
}

}  // extern "C"
