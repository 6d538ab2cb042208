// feature_align.cc — FeatureAlign of the host mirror (reference: feature_align.cc:33-431).
//
// The reference walks a shuffled 32-px grid and calls Matcher::SearchPoint point by point until each visited cell has a
// match.  SearchPoint has no side effects, so here the GPU evaluates ProjectPoint + SearchPoint for EVERY point that
// ProjectPoints would visit, in one launch (sdvlb_search_points with SDVLB_CAND_PROJECT), and ApplyMatches then
// replays ProjectPoint's bookkeeping and the SelectPoints loop over the results: same visiting order, same early
// exits, same Promote/Unpromote/DeletePoint side effects, same rand() consumption.  RANSAC inlier selection and the
// Tukey-weighted pose refinement are device calls as well (sdvlb_select_inliers / sdvlb_optimize_pose).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <stdexcept>
#include <string>

#include "sdvl_host.h"

using std::shared_ptr;
using std::vector;

namespace sdvl {

FeatureAlign::FeatureAlign(Map* map, Camera* camera, int max_matches)   // :33-54
    : map_(map), cell_size_(Config::CellSize()), max_matches_(max_matches), matches_(0), num_attempts_(0), relocalizing_(false) {
  grid_width_ = int(std::ceil(double(camera->GetWidth()) / cell_size_));
  grid_height_ = int(std::ceil(double(camera->GetHeight()) / cell_size_));
  grid_.resize(size_t(grid_width_) * grid_height_);
  cell_order_.resize(grid_.size());
  for (size_t i = 0; i < cell_order_.size(); ++i) cell_order_[i] = int(i);
  RandomShuffle(&cell_order_, &rng_);
}

FeatureAlign::~FeatureAlign() {}

void FeatureAlign::ResetGrid() {   // :285-294
  matches_ = num_attempts_ = 0;
  for (auto& c : grid_) c.clear();
}

// ProjectPoints' iteration and filters (:296-321); the projection itself runs on the device.
void FeatureAlign::CollectCandidates(int frame_id, const shared_ptr<Frame>& last_frame, bool reloc,
                                     vector<sdvlb_candidate>* cands, vector<shared_ptr<Point>>* points,
                                     vector<unsigned char>* descs) {
  cands->clear();
  points->clear();
  if (descs) descs->clear();
  relocalizing_ = reloc;
  for (const shared_ptr<Feature>& seen : last_frame->GetFeatures()) {
    if (!seen) continue;
    const shared_ptr<Point> point = seen->GetPoint();
    if (!point || point->ToDelete() || frame_id == point->GetLastFrame()) continue;
    const shared_ptr<Feature> feature = point->GetInitFeature();
    sdvlb_candidate c;
    // without an init feature the point is never searched (feature_align.cc:120-122) but still binned by ProjectPoint:
    // it goes to the device for the projection only, described by the feature that sees it
    Matcher::FillCandidate(feature ? feature : seen, point->GetInverseDepth(), point->GetStd(), point->IsFixed(), &c);
    const Eigen::Vector3d pos = point->GetPosition();
    c.pos[0] = pos(0); c.pos[1] = pos(1); c.pos[2] = pos(2);
    c.flags |= SDVLB_CAND_PROJECT;
    cands->push_back(c);
    points->push_back(point);
    if (descs) {
      const vector<unsigned char>& d = (feature ? feature : seen)->GetDescriptor();
      descs->insert(descs->end(), d.begin(), d.end());
    }
    if (!reloc) point->SetLastFrame(frame_id);
  }
}

void FeatureAlign::ApplyMatches(const shared_ptr<Frame>& frame, const vector<shared_ptr<Point>>& points,
                                const sdvlb_match* matches) {
  vector<shared_ptr<Feature>> fs_found;
  inliers_.clear(); outliers_.clear();

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
  matches_ = num_attempts_ = 0;
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
  vector<unsigned char> descs;
  CollectCandidates(frame->GetID(), last_frame, reloc, &cands, &points, Config::UseORB() ? &descs : nullptr);
  vector<sdvlb_match> res(cands.size());
  if (!cands.empty()) {
    double T_cur[7];
    frame->GetPose().ToArray(T_cur);
    const int rc = Config::UseORB()
        ? sdvlb_search_points_orb(frame->Context(), frame->Handle(), cands.data(), int(cands.size()), T_cur, descs.data(), res.data())
        : sdvlb_search_points(frame->Context(), frame->Handle(), cands.data(), int(cands.size()), T_cur, res.data());
    if (rc) throw std::runtime_error(std::string("sdvl-b200: FeatureAlign::Reproject failed: ") + sdvlb_last_error());
  }
  ApplyMatches(frame, points, res.data());
}

// ---- pose refinement (:73-82, :152-283, :341-431): RANSAC inlier selection and the Tukey-weighted Gauss-Newton run on
// the device (pose_call_kernel, csrc/seq.cu) behind sdvlb_select_inliers / sdvlb_optimize_pose; the host keeps the
// reference's lists in step with the flags the device returns.  There is no host implementation of either.
namespace {

// The features of `fs` that still observe a point, as C-ABI observations tagged with `flag` (the list they are in).
// `kept` receives the same features, index for index.
void AppendObservations(const vector<shared_ptr<Feature>>& fs, int flag, vector<sdvlb_pose_obs>* obs,
                        vector<shared_ptr<Feature>>* kept) {
  for (const auto& f : fs) {
    const shared_ptr<Point> p = f->GetPoint();
    if (!p) continue;
    sdvlb_pose_obs o;
    const Eigen::Vector3d pos = p->GetPosition();
    const Eigen::Vector3d v = f->GetVector();
    for (int k = 0; k < 3; k++) { o.v[k] = v(k); o.pos[k] = pos(k); }
    o.level = f->GetLevel();
    o.flags = flag;
    obs->push_back(o);
    kept->push_back(f);
  }
}

// Splits `kept` by the flags the device wrote back.
void SplitByFlag(const vector<shared_ptr<Feature>>& kept, const vector<sdvlb_pose_obs>& obs,
                 vector<shared_ptr<Feature>>* inliers, vector<shared_ptr<Feature>>* outliers) {
  inliers->clear(); outliers->clear();
  for (size_t i = 0; i < kept.size(); i++)
    (obs[i].flags == SDVLB_OBS_INLIER ? inliers : outliers)->push_back(kept[i]);
}

}  // namespace

bool FeatureAlign::OptimizePose(const shared_ptr<Frame>& frame) {   // :73-82, one device call for the three steps
  vector<shared_ptr<Feature>> kept;
  vector<sdvlb_pose_obs> obs;
  AppendObservations(inliers_, SDVLB_OBS_INLIER, &obs, &kept);
  AppendObservations(outliers_, SDVLB_OBS_OUTLIER, &obs, &kept);
  double T[7];
  frame->GetPose().ToArray(T);
  if (sdvlb_optimize_pose(frame->Context(), obs.data(), int(obs.size()), T))
    throw std::runtime_error(std::string("sdvl-b200: FeatureAlign::OptimizePose failed: ") + sdvlb_last_error());
  frame->SetPose(SE3(T));
  SplitByFlag(kept, obs, &inliers_, &outliers_);
  RemoveOutliers(frame, &outliers_);
  return true;
}

void FeatureAlign::SelectInliers(const shared_ptr<Frame>& frame, vector<shared_ptr<Feature>>& fs_found,
                                 vector<shared_ptr<Feature>>* inliers, vector<shared_ptr<Feature>>* outliers) {   // :152-216
  inliers->clear(); outliers->clear();
  if (fs_found.empty()) return;
  vector<shared_ptr<Feature>> kept;
  vector<sdvlb_pose_obs> obs;
  AppendObservations(fs_found, 0, &obs, &kept);
  double T[7];
  frame->GetPose().ToArray(T);
  // rng_ advances by exactly the number of rand() calls the reference's loop makes
  if (sdvlb_select_inliers(frame->Context(), obs.data(), int(obs.size()), T, rng_.State()))
    throw std::runtime_error(std::string("sdvl-b200: FeatureAlign::SelectInliers failed: ") + sdvlb_last_error());
  SplitByFlag(kept, obs, inliers, outliers);
}

void FeatureAlign::RemoveOutliers(const shared_ptr<Frame>& frame, vector<shared_ptr<Feature>>* outliers) {   // :245-256
  for (const auto& f : *outliers) {
    const shared_ptr<Point> p = f->GetPoint();
    if (!p) continue;
    f->SetPoint(nullptr);
    p->SetStatus(Point::P_NOT_FOUND);
    frame->AddOutlier(f->GetPosition());
  }
}

}  // namespace sdvl
