// feature_align.cc — FeatureAlign of the host mirror (reference: feature_align.cc:33-431).
//
// The reference walks a shuffled 32-px grid and calls Matcher::SearchPoint point by point until each visited cell has a
// match.  SearchPoint has no side effects, so here the GPU evaluates ProjectPoint + SearchPoint for EVERY point that
// ProjectPoints would visit, in one launch (sdvlb_search_points with SDVLB_CAND_PROJECT), and ApplyMatches then
// replays ProjectPoint's bookkeeping and the SelectPoints loop over the results: same visiting order, same early
// exits, same Promote/Unpromote/DeletePoint side effects, same rand() consumption.  RANSAC inlier selection and the
// Tukey-weighted pose refinement stay on the host (tiny 6x6 fp64 problems), as in the reference.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <stdexcept>
#include <string>

#include "../csrc/common.cuh"   // ldlt_solve6, jacobian3d_to_plane (same code as the device)
#include "sdvl_host.h"

using std::shared_ptr;
using std::vector;

namespace sdvl {

FeatureAlign::FeatureAlign(Map* map, Camera* camera, int max_matches) {   // :33-54
  map_ = map;
  cell_size_ = Config::CellSize();
  max_matches_ = max_matches;
  matches_ = 0;
  num_attempts_ = 0;
  relocalizing_ = false;
  grid_width_ = int(std::ceil(double(camera->GetWidth()) / cell_size_));
  grid_height_ = int(std::ceil(double(camera->GetHeight()) / cell_size_));
  const int size = grid_width_ * grid_height_;
  grid_.resize(size);
  for (int i = 0; i < size; ++i) cell_order_.push_back(i);
  RandomShuffle(&cell_order_, &rng_);
}

FeatureAlign::~FeatureAlign() {}

void FeatureAlign::ResetGrid() {   // :285-294
  matches_ = 0;
  num_attempts_ = 0;
  for (auto& c : grid_) c.clear();
}

// ProjectPoints' iteration and filters (:296-321); the projection itself runs on the device.
void FeatureAlign::CollectCandidates(int frame_id, const shared_ptr<Frame>& last_frame, bool reloc,
                                     vector<sdvlb_candidate>* cands, vector<shared_ptr<Point>>* points) {
  cands->clear();
  points->clear();
  relocalizing_ = reloc;
  vector<shared_ptr<Feature>>& features = last_frame->GetFeatures();
  for (auto it_fts = features.begin(); it_fts != features.end(); it_fts++) {
    if (*it_fts == nullptr) continue;
    shared_ptr<Point> point = (*it_fts)->GetPoint();
    if (!point || point->ToDelete()) continue;
    if (frame_id == point->GetLastFrame()) continue;
    shared_ptr<Feature> feature = point->GetInitFeature();
    sdvlb_candidate c;
    if (feature) {
      Matcher::FillCandidate(feature, point->GetInverseDepth(), point->GetStd(), point->IsFixed(), &c);
    } else {   // never searched (feature_align.cc:120-122) but still binned by ProjectPoint: project only
      Matcher::FillCandidate((*it_fts), point->GetInverseDepth(), point->GetStd(), point->IsFixed(), &c);
    }
    const Eigen::Vector3d pos = point->GetPosition();
    c.pos[0] = pos(0); c.pos[1] = pos(1); c.pos[2] = pos(2);
    c.flags |= SDVLB_CAND_PROJECT;
    cands->push_back(c);
    points->push_back(point);
    if (!reloc) point->SetLastFrame(frame_id);
  }
}

void FeatureAlign::ApplyMatches(const shared_ptr<Frame>& frame, const vector<shared_ptr<Point>>& points,
                                const sdvlb_match* matches) {
  vector<shared_ptr<Feature>> fs_found;
  inliers_.clear();
  outliers_.clear();

  // ---- ProjectPoint bookkeeping (:323-339), in ProjectPoints order
  ResetGrid();
  for (size_t i = 0; i < points.size(); i++) {
    const sdvlb_match& m = matches[i];
    const shared_ptr<Point>& point = points[i];
    if (m.status == SDVLB_MATCH_UNSEEN) {
      point->SetStatus(Point::P_UNSEEN);
      continue;
    }
    const int k = int(m.proj[1] / cell_size_) * grid_width_ + int(m.proj[0] / cell_size_);
    grid_.at(k).push_back(std::make_pair(point, Eigen::Vector2d(m.proj[0], m.proj[1])));
    point->SetStatus(Point::P_SEEN);
  }
  // result lookup by point (a point is visited at most once per frame thanks to GetLastFrame)
  std::vector<std::pair<Point*, const sdvlb_match*>> lookup;
  lookup.reserve(points.size());
  for (size_t i = 0; i < points.size(); i++) lookup.push_back(std::make_pair(points[i].get(), &matches[i]));
  std::sort(lookup.begin(), lookup.end());
  auto find_match = [&lookup](Point* p) -> const sdvlb_match* {
    auto it = std::lower_bound(lookup.begin(), lookup.end(), std::make_pair(p, (const sdvlb_match*)nullptr));
    return (it != lookup.end() && it->first == p) ? it->second : nullptr;
  };

  // ---- SelectPoints loop (:98-149)
  matches_ = 0;
  num_attempts_ = 0;
  RandomShuffle(&cell_order_, &rng_);
  const int size = int(grid_.size());
  for (int i = 0; i < size && matches_ < max_matches_; i++) {
    bool found = false;
    GridCell& cell = grid_.at(cell_order_[i]);
    cell.sort([](const PointInfo& a, const PointInfo& b) { return a.first->Score() > b.first->Score(); });
    for (auto it = cell.begin(); it != cell.end() && !found; it++) {
      shared_ptr<Point> point = it->first;
      if (point->ToDelete()) continue;
      shared_ptr<Feature> feature = point->GetInitFeature();
      if (!feature) continue;
      num_attempts_++;
      const sdvlb_match* m = find_match(point.get());
      found = m && m->status == SDVLB_MATCH_FOUND;
      if (found) {
        if (!relocalizing_) {
          point->Promote();
          shared_ptr<Feature> nf = std::make_shared<Feature>(frame, Eigen::Vector2d(m->px[0], m->px[1]), m->level);
          nf->SetPoint(point);
          frame->AddFeature(nf);
          point->SetStatus(Point::P_FOUND);
          fs_found.push_back(nf);
        }
        matches_++;
      } else {
        if (!relocalizing_) {
          if (point->Unpromote()) map_->DeletePoint(point);
          point->SetStatus(Point::P_NOT_FOUND);
        }
      }
    }
  }
  const auto t0 = std::chrono::steady_clock::now();
  SelectInliers(frame, fs_found, &inliers_, &outliers_);   // :68
  ransac_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void FeatureAlign::Reproject(const shared_ptr<Frame>& frame, const shared_ptr<Frame>& last_frame,
                             const shared_ptr<Frame>& /*last_kf*/, bool reloc) {   // :59-71
  relocalizing_ = reloc;
  vector<sdvlb_candidate> cands;
  vector<shared_ptr<Point>> points;
  CollectCandidates(frame->GetID(), last_frame, reloc, &cands, &points);
  vector<sdvlb_match> res(cands.size());
  if (!cands.empty()) {
    double T_cur[7];
    frame->GetPose().ToArray(T_cur);
    const int rc = sdvlb_search_points(frame->Context(), frame->Handle(), cands.data(), int(cands.size()), T_cur, res.data());
    if (rc) throw std::runtime_error(std::string("sdvl-b200: FeatureAlign::Reproject failed: ") + sdvlb_last_error());
  }
  ApplyMatches(frame, points, res.data());
}

bool FeatureAlign::device_pose_refinement_ = false;

// fs -> C-ABI observations (FeatureAlign lists as arrays); flag = which list each feature is in
static void FillObs(const vector<shared_ptr<Feature>>& fs, int flag, vector<sdvlb_pose_obs>* obs) {
  for (const auto& f : fs) {
    shared_ptr<Point> p = f->GetPoint();
    if (!p) continue;
    sdvlb_pose_obs o;
    const Eigen::Vector3d pos = p->GetPosition();
    o.v[0] = f->GetVector()(0); o.v[1] = f->GetVector()(1); o.v[2] = f->GetVector()(2);
    o.pos[0] = pos(0); o.pos[1] = pos(1); o.pos[2] = pos(2);
    o.level = f->GetLevel();
    o.flags = flag;
    obs->push_back(o);
  }
}

bool FeatureAlign::OptimizePose(const shared_ptr<Frame>& frame) {   // :73-82
  if (device_pose_refinement_) {   // the same three steps as one device call (sdvlb_optimize_pose)
    vector<shared_ptr<Feature>> all;
    vector<sdvlb_pose_obs> obs;
    for (const auto& f : inliers_) if (f->GetPoint()) all.push_back(f);
    FillObs(inliers_, SDVLB_OBS_INLIER, &obs);
    for (const auto& f : outliers_) if (f->GetPoint()) all.push_back(f);
    FillObs(outliers_, SDVLB_OBS_OUTLIER, &obs);
    double T[7];
    frame->GetPose().ToArray(T);
    if (sdvlb_optimize_pose(frame->Context(), obs.data(), int(obs.size()), T))
      throw std::runtime_error(std::string("sdvl-b200: FeatureAlign::OptimizePose failed: ") + sdvlb_last_error());
    frame->SetPose(SE3(T));
    inliers_.clear();
    outliers_.clear();
    for (size_t i = 0; i < all.size(); i++)
      (obs[i].flags == SDVLB_OBS_INLIER ? inliers_ : outliers_).push_back(all[i]);
    RemoveOutliers(frame, &outliers_);
    return true;
  }
  OptimizePose(frame, &inliers_, &outliers_);
  if (RescueOutliers(frame, &inliers_, &outliers_)) OptimizePose(frame, &inliers_, &outliers_);
  RemoveOutliers(frame, &outliers_);
  return true;
}

void FeatureAlign::SelectInliers(const shared_ptr<Frame>& frame, vector<shared_ptr<Feature>>& fs_found,
                                 vector<shared_ptr<Feature>>* inliers, vector<shared_ptr<Feature>>* outliers) {   // :152-216
  vector<shared_ptr<Feature>> selected, best_fs;
  int supporters, best_supporters;
  SE3 se3, best_se3;
  inliers->clear();
  outliers->clear();
  if (fs_found.empty()) return;
  if (device_pose_refinement_) {   // RANSAC on the device; rng_ advances by the reference's number of rand() calls
    vector<sdvlb_pose_obs> obs;
    FillObs(fs_found, 0, &obs);
    double T[7];
    frame->GetPose().ToArray(T);
    if (sdvlb_select_inliers(frame->Context(), obs.data(), int(obs.size()), T, rng_.State()))
      throw std::runtime_error(std::string("sdvl-b200: FeatureAlign::SelectInliers failed: ") + sdvlb_last_error());
    size_t k = 0;
    for (const auto& f : fs_found) {
      if (!f->GetPoint()) continue;
      (obs[k++].flags == SDVLB_OBS_INLIER ? inliers : outliers)->push_back(f);
    }
    return;
  }
  const int size = int(fs_found.size());
  const int npoints = std::min(Config::MaxRansacPoints(), size);
  vector<int> indexes(npoints);
  const double sprob = 0.99;
  int nits = Config::MaxRansacIts();
  best_supporters = 0;
  int it = 0;
  const double thr = Config::InlierErrorThreshold() / frame->GetCamera()->GetFx();
  while (it < nits) {
    selected.clear();
    const int index = rng_.Next() % size;
    for (int i = 0; i < npoints; i++) {
      indexes[i] = (index + i) % size;
      selected.push_back(fs_found.at(indexes[i]));
    }
    if (!ConvergePose(frame, selected, &se3)) { it++; continue; }
    supporters = CheckReprojectionError(fs_found, se3, thr);
    if (supporters > best_supporters) {
      best_fs = selected;
      best_supporters = supporters;
      best_se3 = se3;
      const double epsilon = 1.0 - (double(supporters) / double(size));
      double tmp = 1.0 - epsilon;
      for (int k = 1; k < npoints; k++) tmp *= tmp;
      if (tmp < 1e-5) nits = Config::MaxRansacIts();
      else nits = std::min(Config::MaxRansacIts(), int(std::log(1.0 - sprob) / std::log(1.0 - tmp)));
    }
    it++;
  }
  CheckReprojectionError(fs_found, best_se3, thr, inliers, outliers);
}

void FeatureAlign::OptimizePose(const shared_ptr<Frame>& frame, vector<shared_ptr<Feature>>* features,
                                vector<shared_ptr<Feature>>* outliers) {   // :218-230
  SE3 se3 = frame->GetPose();
  if (!ConvergePose(frame, *features, &se3)) return;
  frame->SetPose(se3);
  vector<shared_ptr<Feature>> cfeatures = *features;
  features->clear();
  CheckReprojectionError(cfeatures, frame->GetPose(), Config::InlierErrorThreshold() / frame->GetCamera()->GetFx(), features, outliers);
}

bool FeatureAlign::RescueOutliers(const shared_ptr<Frame>& frame, vector<shared_ptr<Feature>>* inliers,
                                  vector<shared_ptr<Feature>>* outliers) {   // :232-243
  const int init_inliers = int(inliers->size());
  vector<shared_ptr<Feature>> cfeatures = *outliers;
  outliers->clear();
  CheckReprojectionError(cfeatures, frame->GetPose(), 2 * Config::InlierErrorThreshold() / frame->GetCamera()->GetFx(), inliers, outliers);
  return int(inliers->size()) > init_inliers;
}

void FeatureAlign::RemoveOutliers(const shared_ptr<Frame>& frame, vector<shared_ptr<Feature>>* outliers) {   // :245-256
  for (auto it = outliers->begin(); it != outliers->end(); it++) {
    shared_ptr<Point> p = (*it)->GetPoint();
    if (!p) continue;
    (*it)->SetPoint(nullptr);
    p->SetStatus(Point::P_NOT_FOUND);
    frame->AddOutlier((*it)->GetPosition());
  }
}

int FeatureAlign::CheckReprojectionError(const vector<shared_ptr<Feature>>& features, const SE3& se3, double threshold,
                                         vector<shared_ptr<Feature>>* inliers, vector<shared_ptr<Feature>>* outliers) {   // :258-283
  int valids = 0;
  for (auto it = features.begin(); it != features.end(); it++) {
    shared_ptr<Point> point = (*it)->GetPoint();
    if (!point) continue;
    const Eigen::Vector3d pos = se3 * point->GetPosition();
    const Eigen::Vector2d a = Camera::SimpleProject((*it)->GetVector()), b = Camera::SimpleProject(pos);
    const double sqrt_inv_cov = 1.0 / (1 << (*it)->GetLevel());
    const double ex = (a(0) - b(0)) * sqrt_inv_cov, ey = (a(1) - b(1)) * sqrt_inv_cov;
    if (std::sqrt(ex * ex + ey * ey) <= threshold) {
      valids++;
      if (inliers != NULL) inliers->push_back(*it);
    } else {
      if (outliers != NULL) outliers->push_back(*it);
    }
  }
  return valids;
}

bool FeatureAlign::ConvergePose(const shared_ptr<Frame>& frame, const vector<shared_ptr<Feature>>& features, SE3* se3) {   // :341-421
  double A[36], b[6], J0[6], J1[6];
  SE3 last_se3 = frame->GetPose();
  Camera* camera = frame->GetCamera();
  *se3 = last_se3;
  double chi2 = 0.0;

  vector<double> errors;
  for (auto it = features.begin(); it != features.end(); it++) {
    shared_ptr<Point> point = (*it)->GetPoint();
    if (!point) continue;
    const Eigen::Vector3d pos = (*se3) * point->GetPosition();
    const Eigen::Vector2d a = Camera::SimpleProject((*it)->GetVector()), c = Camera::SimpleProject(pos);
    const double s = 1.0 / (1 << (*it)->GetLevel());
    const double ex = (a(0) - c(0)) * s, ey = (a(1) - c(1)) * s;
    errors.push_back(std::sqrt(ex * ex + ey * ey));
  }
  if (errors.empty()) return false;
  // GetMedianVector (utils.cc:215-220)
  auto mid = errors.begin() + (errors.size() / 2);
  std::nth_element(errors.begin(), mid, errors.end());
  double scale = KMADNorm * (*mid);

  for (int i = 0; i < Config::MaxOptimPoseIts(); i++) {
    for (int r = 0; r < 36; r++) A[r] = 0;
    for (int r = 0; r < 6; r++) b[r] = 0;
    double new_chi2 = 0.0;
    if (i == 5) scale = 0.85 / camera->GetFx();
    for (auto it = features.begin(); it != features.end(); it++) {
      shared_ptr<Point> point = (*it)->GetPoint();
      if (!point) continue;
      const Eigen::Vector3d pos = (*se3) * point->GetPosition();
      jacobian3d_to_plane(pos(0), pos(1), pos(2), J0, J1);
      const Eigen::Vector2d a = Camera::SimpleProject((*it)->GetVector()), c = Camera::SimpleProject(pos);
      const double sqrt_inv_cov = 1.0 / (1 << (*it)->GetLevel());
      const double ex = (a(0) - c(0)) * sqrt_inv_cov, ey = (a(1) - c(1)) * sqrt_inv_cov;
      for (int r = 0; r < 6; r++) { J0[r] *= sqrt_inv_cov; J1[r] *= sqrt_inv_cov; }
      const double weight = GetTukeyValue(std::sqrt(ex * ex + ey * ey) / scale);
      for (int r = 0; r < 6; r++) {
        for (int q = 0; q < 6; q++) A[r * 6 + q] += (J0[r] * J0[q] + J1[r] * J1[q]) * weight;
        b[r] -= (J0[r] * ex + J1[r] * ey) * weight;
      }
      new_chi2 += (ex * ex + ey * ey) * weight;
    }
    double dT[6];
    ldlt_solve6(A, b, dT);
    if ((i > 0 && new_chi2 > chi2) || std::isnan(dT[0])) {
      *se3 = last_se3;
      break;
    }
    const SE3 T_new = SE3::Exp(dT) * (*se3);
    last_se3 = *se3;
    *se3 = T_new;
    chi2 = new_chi2;
    double amax = -1;
    for (int r = 0; r < 6; r++) amax = std::max(amax, std::fabs(dT[r]));
    if (amax <= 1e-10) break;
  }
  return true;
}

double FeatureAlign::GetTukeyValue(double x) {   // :423-431
  const double x_square = x * x;
  if (x_square <= KTukeyC) {
    const double tmp = 1.0 - x_square / KTukeyC;
    return tmp * tmp;
  }
  return 0.0;
}

}  // namespace sdvl
