// feature_align.cc — restatement of FeatureAlign (feature_align.cc:33-431) (oracle; test infrastructure only).
#include <algorithm>
#include <cassert>

#include "oracle.h"

namespace oracle {

static const double KMADNorm = 1.4826;             // feature_align.h:114
static const double KTukeyC = 4.6851 * 4.6851;     // feature_align.h:115

static inline V2 SimpleProject(const V3& p) { V2 r; r.x = p.x / p.z; r.y = p.y / p.z; return r; }  // camera.h:110-112

FeatureAlign::FeatureAlign(const sdvlb_params& P, const Camera* cam, int max_matches, GlibcRand* rng,
                           std::vector<std::shared_ptr<Point>>* trash)
    : P_(P), cam_(cam), rng_(rng), trash_(trash) {  // :33-54
  cell_size_ = P.cell_size;
  max_matches_ = max_matches;
  grid_width_ = int(std::ceil(double(cam->width) / cell_size_));
  grid_height_ = int(std::ceil(double(cam->height) / cell_size_));
  const int size = grid_width_ * grid_height_;
  grid_.resize(size);
  for (int i = 0; i < size; ++i) cell_order_.push_back(i);
  RandomShuffle(&cell_order_, rng_);
}

void FeatureAlign::Reproject(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Frame>& last_frame, bool reloc) {  // :59-71
  std::vector<std::shared_ptr<Feature>> selected_fs;
  inliers_.clear();
  outliers_.clear();
  relocalizing_ = reloc;
  SelectPoints(frame, last_frame, &selected_fs);
  SelectInliers(frame, selected_fs, &inliers_, &outliers_);
  selected_ = selected_fs;
}

bool FeatureAlign::OptimizePose(const std::shared_ptr<Frame>& frame) {  // :73-82
  OptimizePose(frame, &inliers_, &outliers_);
  if (RescueOutliers(frame, &inliers_, &outliers_)) OptimizePose(frame, &inliers_, &outliers_);
  RemoveOutliers(frame, &outliers_);
  return true;
}

void FeatureAlign::SelectPoints(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Frame>& last_frame,
                                std::vector<std::shared_ptr<Feature>>* fs_found) {  // :88-150
  Matcher matcher(P_, P_.patch_size);
  V2 pos;
  bool found;
  int level = 0;

  ProjectPoints(frame, last_frame);
  matches_ = 0;
  num_attempts_ = 0;

  RandomShuffle(&cell_order_, rng_);
  const int size = int(grid_.size());
  for (int i = 0; i < size && matches_ < max_matches_; i++) {
    found = false;
    GridCell& cell = grid_.at(cell_order_[i]);
    cell.sort([](const PointInfo& a, const PointInfo& b) { return a.first->n_successful > b.first->n_successful; });
    for (auto it = cell.begin(); it != cell.end() && !found; ++it) {
      std::shared_ptr<Point> point = it->first;
      if (point->del) continue;
      std::shared_ptr<Feature> feature = point->feature;
      if (!feature) continue;
      num_attempts_++;
      pos = it->second;
      found = matcher.SearchPoint(frame, feature, point->rho, point->GetStd(), point->fixed, &pos, &level);
      if (found) {
        if (!relocalizing_) {
          point->Promote();
          std::shared_ptr<Feature> nf = MakeFeature(frame, pos, level);
          nf->point = point;
          frame->features.push_back(nf);
          point->status = 0;  // P_FOUND
          fs_found->push_back(nf);
        }
        matches_++;
      } else {
        if (!relocalizing_) {
          if (point->Unpromote(P_.max_failed)) trash_->push_back(point);  // map_->DeletePoint
          point->status = 1;  // P_NOT_FOUND
        }
      }
    }
  }
}

void FeatureAlign::SelectInliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>& fs_found,
                                 std::vector<std::shared_ptr<Feature>>* inliers,
                                 std::vector<std::shared_ptr<Feature>>* outliers) {  // :152-216
  std::vector<std::shared_ptr<Feature>> selected, best_fs;
  int supporters, best_supporters;
  SE3 se3, best_se3;
  inliers->clear();
  outliers->clear();
  if (fs_found.empty()) return;

  const int size = int(fs_found.size());
  const int npoints = std::min(P_.max_ransac_points, size);
  std::vector<int> indexes(npoints);
  const double sprob = 0.99;
  int nits = P_.max_ransac_its;
  best_supporters = 0;
  int it = 0;
  const double thr = P_.inlier_error_threshold / frame->cam->fx;
  while (it < nits) {
    selected.clear();
    const int index = rng_->Next() % size;
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
      if (tmp < 1e-5)
        nits = P_.max_ransac_its;
      else
        nits = std::min(P_.max_ransac_its, int(std::log(1.0 - sprob) / std::log(1.0 - tmp)));
    }
    it++;
  }
  CheckReprojectionError(fs_found, best_se3, thr, inliers, outliers);
}

void FeatureAlign::OptimizePose(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>* features,
                                std::vector<std::shared_ptr<Feature>>* outliers) {  // :218-230
  SE3 se3 = frame->pose;
  if (!ConvergePose(frame, *features, &se3)) return;
  frame->pose = se3;
  std::vector<std::shared_ptr<Feature>> cfeatures = *features;
  features->clear();
  CheckReprojectionError(cfeatures, frame->pose, P_.inlier_error_threshold / frame->cam->fx, features, outliers);
}

bool FeatureAlign::RescueOutliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>* inliers,
                                  std::vector<std::shared_ptr<Feature>>* outliers) {  // :232-243
  const int init_inliers = int(inliers->size());
  std::vector<std::shared_ptr<Feature>> cfeatures = *outliers;
  outliers->clear();
  CheckReprojectionError(cfeatures, frame->pose, 2 * P_.inlier_error_threshold / frame->cam->fx, inliers, outliers);
  return int(inliers->size()) > init_inliers;
}

void FeatureAlign::RemoveOutliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>* outliers) {  // :245-256
  for (auto it = outliers->begin(); it != outliers->end(); ++it) {
    std::shared_ptr<Point> p = (*it)->point;
    if (!p) continue;
    (*it)->point = nullptr;
    p->status = 1;  // P_NOT_FOUND
    frame->outliers.push_back((*it)->p2d);
  }
}

int FeatureAlign::CheckReprojectionError(const std::vector<std::shared_ptr<Feature>>& features, const SE3& se3,
                                         double threshold, std::vector<std::shared_ptr<Feature>>* inliers,
                                         std::vector<std::shared_ptr<Feature>>* outliers) {  // :258-283
  int valids = 0;
  for (auto it = features.begin(); it != features.end(); ++it) {
    std::shared_ptr<Point> point = (*it)->point;
    if (!point) continue;
    const V3 pos = se3 * point->GetPosition();
    const V2 a = SimpleProject((*it)->v), b = SimpleProject(pos);
    const double sqrt_inv_cov = 1.0 / (1 << (*it)->level);
    const double ex = (a.x - b.x) * sqrt_inv_cov, ey = (a.y - b.y) * sqrt_inv_cov;
    if (std::sqrt(ex * ex + ey * ey) <= threshold) {
      valids++;
      if (inliers) inliers->push_back(*it);
    } else {
      if (outliers) outliers->push_back(*it);
    }
  }
  return valids;
}

void FeatureAlign::ResetGrid() {  // :285-294
  matches_ = 0;
  num_attempts_ = 0;
  for (auto& c : grid_) c.clear();
}

void FeatureAlign::ProjectPoints(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Frame>& last_frame) {  // :296-321
  ResetGrid();
  for (auto it = last_frame->features.begin(); it != last_frame->features.end(); ++it) {
    if (*it == nullptr) continue;
    std::shared_ptr<Point> point = (*it)->point;
    if (!point || point->del) continue;
    if (frame->id == point->last_frame) continue;
    ProjectPoint(frame, point);
    if (!relocalizing_) point->last_frame = frame->id;
  }
}

bool FeatureAlign::ProjectPoint(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Point>& point) {  // :323-339
  V2 p;
  if (!frame->Project(point->GetPosition(), &p)) { point->status = 3; return false; }          // P_UNSEEN
  if (!frame->cam->IsInsideImage(int(p.x), int(p.y), P_.patch_size)) { point->status = 3; return false; }
  const int k = int(p.y / cell_size_) * grid_width_ + int(p.x / cell_size_);
  grid_.at(k).push_back(std::make_pair(point, p));
  point->status = 2;  // P_SEEN
  return true;
}

bool FeatureAlign::ConvergePose(const std::shared_ptr<Frame>& frame, const std::vector<std::shared_ptr<Feature>>& features,
                                SE3* se3) {  // :341-421
  Mat6 A;
  Vec6 b;
  double J[2][6];
  SE3 last_se3 = frame->pose;
  *se3 = last_se3;
  double chi2 = 0.0;

  std::vector<double> errors;
  for (auto it = features.begin(); it != features.end(); ++it) {
    std::shared_ptr<Point> point = (*it)->point;
    if (!point) continue;
    const V3 pos = (*se3) * point->GetPosition();
    const V2 a = SimpleProject((*it)->v), c = SimpleProject(pos);
    const double s = 1.0 / (1 << (*it)->level);
    const double ex = (a.x - c.x) * s, ey = (a.y - c.y) * s;
    errors.push_back(std::sqrt(ex * ex + ey * ey));
  }
  if (errors.empty()) return false;
  double scale = KMADNorm * GetMedianVector(&errors);

  for (int i = 0; i < P_.max_optim_pose_its; i++) {
    for (int r = 0; r < 6; r++) {
      b[r] = 0;
      for (int c = 0; c < 6; c++) A[r][c] = 0;
    }
    double new_chi2 = 0.0;
    if (i == 5) scale = 0.85 / frame->cam->fx;

    for (auto it = features.begin(); it != features.end(); ++it) {
      std::shared_ptr<Point> point = (*it)->point;
      if (!point) continue;
      const V3 pos = (*se3) * point->GetPosition();
      Jacobian3DToPlane(pos, J);
      const V2 a = SimpleProject((*it)->v), c = SimpleProject(pos);
      const double sqrt_inv_cov = 1.0 / (1 << (*it)->level);
      const double ex = (a.x - c.x) * sqrt_inv_cov, ey = (a.y - c.y) * sqrt_inv_cov;
      for (int r = 0; r < 6; r++) { J[0][r] *= sqrt_inv_cov; J[1][r] *= sqrt_inv_cov; }
      const double weight = GetTukeyValue(std::sqrt(ex * ex + ey * ey) / scale);
      for (int r = 0; r < 6; r++) {
        for (int cc = 0; cc < 6; cc++) A[r][cc] += (J[0][r] * J[0][cc] + J[1][r] * J[1][cc]) * weight;
        b[r] -= (J[0][r] * ex + J[1][r] * ey) * weight;
      }
      new_chi2 += (ex * ex + ey * ey) * weight;
    }

    Vec6 dT;
    LdltSolve6(A, b, dT);
    if ((i > 0 && new_chi2 > chi2) || std::isnan(dT[0])) {
      *se3 = last_se3;
      break;
    }
    const SE3 T_new = SE3::Exp(dT) * (*se3);
    last_se3 = *se3;
    *se3 = T_new;
    chi2 = new_chi2;
    if (AbsMax6(dT) <= 1e-10) break;
  }
  return true;
}

double FeatureAlign::GetTukeyValue(double x) {  // :423-431
  const double x_square = x * x;
  if (x_square <= KTukeyC) {
    const double tmp = 1.0 - x_square / KTukeyC;
    return tmp * tmp;
  }
  return 0.0;
}

}  // namespace oracle
