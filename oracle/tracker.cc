// tracker.cc — sequence driver for the oracle (test infrastructure only).
// Reproduces the RUNNING branch of SDVL::HandleFrame (sdvl.cc:55-130): motion-model prior (sdvl.cc:278-281),
// ProcessFrame = ImageAlign::ComputePose + FeatureAlign::Reproject + OptimizePose (sdvl.cc:179-203), motion-model
// update (sdvl.cc:266-276), EmptyTrash (map.cc:207-253).  HomographyInit and the mapping thread are out of scope
// (SURVEY.md §2 rows 11-12): instead every `kf_every`-th frame becomes a keyframe whose FAST corners are turned into
// fixed map points placed on a known world plane (ground-truth depth), at most one per 32-px cell, up to max_points.
#include <algorithm>
#include <chrono>

#include "oracle.h"

namespace oracle {

Tracker::Tracker(const sdvlb_params& P, const Camera& cam, const SeedPlane& plane, int max_points, int kf_every)
    : P_(P), cam_(cam), plane_(plane), max_points_(max_points), kf_every_(kf_every) {
  feature_align_.reset(new FeatureAlign(P_, &cam_, P_.max_matches, &rng_, &trash_));  // sdvl.cc:38
  for (int i = 0; i < 6; i++) vel_[i] = 0;
}

Tracker::~Tracker() {
  for (auto& w : all_points_)
    if (auto p = w.lock()) p->feature = nullptr;
  if (last_frame_) last_frame_->features.clear();
}

void Tracker::SeedKeyframe(const std::shared_ptr<Frame>& f, const SE3& gt_pose) {
  f->is_keyframe = true;
  const int gw = int(std::ceil(cam_.width / P_.cell_size));
  const int gh = int(std::ceil(cam_.height / P_.cell_size));
  std::vector<char> occupied(size_t(gw) * gh, 0);
  int n_points = 0;
  for (auto& ft : f->features) {
    if (!ft->point || ft->point->del) continue;
    n_points++;
    const int cx = int(ft->p2d.x / P_.cell_size), cy = int(ft->p2d.y / P_.cell_size);
    if (cx >= 0 && cx < gw && cy >= 0 && cy < gh) occupied[size_t(cy) * gw + cx] = 1;
  }
  const SE3 gt_wc = gt_pose.Inverse();   // camera -> world (ground truth)
  const M3 Rwc = gt_wc.Rotation();
  const V3 C = gt_wc.t;
  const V3 est_C = f->GetWorldPosition();
  const int n = int(f->corners.size());
  const int margin = P_.patch_size / 2 + 2;
  for (int i = 0; i < n && n_points < max_points_; i++) {
    const Corner& c = f->corners[size_t((long long)i * 7919 % n)];   // fixed pseudo-random visiting order
    const Mat8& lvl = f->pyramid[c.level];
    if (c.x < margin || c.y < margin || c.x >= lvl.cols - margin || c.y >= lvl.rows - margin) continue;
    V2 px; px.x = double(c.x * (1 << c.level)); px.y = double(c.y * (1 << c.level));
    const int cx = int(px.x / P_.cell_size), cy = int(px.y / P_.cell_size);
    if (occupied[size_t(cy) * gw + cx]) continue;
    std::shared_ptr<Feature> ft = MakeFeature(f, px, c.level);
    if (UseOrb()) {   // the descriptor an init feature gets at its keyframe (frame.cc:148-161, map.cc:319-323)
      static const OrbDetector det;
      if (!det.IsInsideLimits(lvl, c.x, c.y)) continue;
      ft->descriptor.resize(32);
      det.GetDescriptor(lvl, c.x, c.y, ft->descriptor.data());
    }
    const V3 dir(Rwc.m[0][0] * ft->v.x + Rwc.m[0][1] * ft->v.y + Rwc.m[0][2] * ft->v.z,
                 Rwc.m[1][0] * ft->v.x + Rwc.m[1][1] * ft->v.y + Rwc.m[1][2] * ft->v.z,
                 Rwc.m[2][0] * ft->v.x + Rwc.m[2][1] * ft->v.y + Rwc.m[2][2] * ft->v.z);
    const V3 nrm(plane_.n[0], plane_.n[1], plane_.n[2]);
    const double denom = nrm.dot(dir);
    if (std::fabs(denom) < 1e-9) continue;
    const double s = (plane_.d - nrm.dot(C)) / denom;
    if (s <= 0) continue;
    auto pt = std::make_shared<Point>();
    pt->id = point_counter_++;
    pt->fixed = true;
    pt->p3d = C + dir * s;
    const double depth = (pt->p3d - est_C).norm();
    pt->rho = 1.0 / depth;
    pt->sigma2 = (0.05 * pt->rho) * (0.05 * pt->rho);
    pt->feature = ft;
    ft->point = pt;
    all_points_.push_back(pt);
    f->features.push_back(ft);
    occupied[size_t(cy) * gw + cx] = 1;
    n_points++;
  }
  last_kf_ = f;
}

void Tracker::EmptyTrash() {  // map.cc:207-253 (points only)
  for (auto& p : trash_) p->del = true;
  trash_.clear();
}

void Tracker::HandleFrame(const uint8_t* img, int w, int h, const SE3& gt_pose, SE3* est_pose, TrackStats* st) {
  TrackStats s;
  std::shared_ptr<Frame> frame = MakeFrame(P_, &cam_, img, w, h, true, frame_counter_);  // sdvl.cc:59
  if (!last_frame_) {
    frame->pose = gt_pose;
    SeedKeyframe(frame, gt_pose);
    s.keyframe = 1;
  } else {
    frame->pose = SE3::Exp(vel_) * last_frame_->pose;  // SetMotionModel, sdvl.cc:278-281
    ImageAlign image_align(P_);
    s.n_tracked = image_align.ComputePose(last_frame_, frame);  // sdvl.cc:185-190
    s.gn_iters = int(image_align.trace.size());
    feature_align_->Reproject(frame, last_frame_);             // sdvl.cc:193
    s.matches = feature_align_->GetMatches();
    s.attempts = feature_align_->GetAttempts();
    feature_align_->OptimizePose(frame);                        // sdvl.cc:200
    s.inliers = feature_align_->n_inliers();
    s.outliers = feature_align_->n_outliers();
    // GetMotionModel, sdvl.cc:266-276
    const SE3 mov = frame->pose * last_frame_->pose.Inverse();
    Vec6 vel;
    SE3::Log(mov, vel);
    for (int i = 0; i < 6; i++) vel_[i] = 0.9 * (0.5 * vel[i] + 0.5 * vel_[i]);
    // Map::NeedKeyframe (map.cc:170-188) with MinKeyframeIts = kf_every and LostRatio = 0.7 (config.cc:62,77)
    int npoints = 0;
    for (auto& ft : frame->features)
      if (ft->point) npoints++;   // Frame::GetNumPoints, frame.cc:165-180
    const bool enough_its = (frame->id - last_kf_->id) >= kf_every_;
    const bool lost_many = npoints < last_matches_ * 0.7;
    const bool lost_some = npoints < last_matches_ * 0.9;
    last_matches_ = std::max(last_matches_, npoints);
    if ((enough_its && lost_some) || lost_many) {
      last_matches_ = npoints;
      SeedKeyframe(frame, gt_pose);
      s.keyframe = 1;
    }
  }
  int nf = 0;
  for (auto& ft : frame->features)
    if (ft->point && !ft->point->del) nf++;
  s.n_feats = nf;
  // Frame <-> Feature shared_ptr cycle: the reference breaks it in Map::EmptyTrash (RemoveFeatures); here a frame's
  // feature list is dropped as soon as it stops being the alignment reference (keyframes stay alive through the init
  // features their live points hold).
  if (last_frame_) last_frame_->features.clear();
  last_frame_ = frame;
  frame_counter_++;
  EmptyTrash();  // sdvl.cc:127
  *est_pose = frame->pose;
  if (st) *st = s;
}

}  // namespace oracle
