// frame_align_matcher.cc — Frame, ImageAlign and Matcher of the host mirror: thin marshalling over the C-ABI.
// Reference behaviour: frame.cc:34-131, image_align.cc:46-84, matcher.cc:45-121.
#include <atomic>
#include <cassert>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <string>

#include "sdvl_host.h"

namespace sdvl {

static void Check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string("sdvl-b200: ") + what + " failed: " + sdvlb_last_error());
}

// ---------------------------------------------------------------- Frame
int Frame::counter_ = 0;

Frame::Frame(Camera* camera, ORBDetector* /*detector*/, const cv::Mat& img, bool corners) {
  id_ = counter_;
  camera_ = camera;
  width_ = img.cols;
  height_ = img.rows;
  ctx_ = Device::Current();
  handle_ = nullptr;
  // CreatePyramid + (optionally) CreateCorners(MaxFastLevels, NumFeatures) on the device (frame.cc:46-53)
  Check(sdvlb_frame_create(ctx_, img.data, img.cols, img.rows, int(img.step), corners ? 1 : 0, Config::NumFeatures(),
                           &handle_),
        "Frame");
  counter_ += 1;
}

Frame::Frame(Camera* camera, sdvlb_ctx* ctx, sdvlb_frame* adopted, int id) {
  id_ = id;
  camera_ = camera;
  ctx_ = ctx;
  handle_ = adopted;
  width_ = int(camera->GetWidth());
  height_ = int(camera->GetHeight());
}

Frame::~Frame() {
  RemoveFeatures();
  if (handle_) sdvlb_frame_destroy(ctx_, handle_);
}

std::vector<cv::Mat>& Frame::GetPyramid() {
  if (!pyramid_fetched_) {
    const int levels = Config::PyramidLevels();
    pyramid_.clear();
    for (int l = 0; l < levels; l++) {
      const uint8_t* data = nullptr;
      int w = 0, h = 0;
      Check(sdvlb_frame_level(handle_, l, &data, &w, &h), "GetPyramid");
      pyramid_.push_back(cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(data)));   // header over the pinned mirror
    }
    pyramid_fetched_ = true;
  }
  return pyramid_;
}

std::vector<Eigen::Vector3i>& Frame::GetCorners() {
  if (!corners_fetched_) {
    const int32_t* xyls = nullptr;
    int n = 0;
    const int rc = sdvlb_frame_corners(handle_, &xyls, &n);
    corners_.clear();
    corner_scores_.clear();
    if (rc == 0) {
      corners_.reserve(n);
      for (int i = 0; i < n; i++) {
        corners_.push_back(Eigen::Vector3i(xyls[4 * i], xyls[4 * i + 1], xyls[4 * i + 2]));
        corner_scores_.push_back(xyls[4 * i + 3]);
      }
    } else if (rc != SDVLB_ERR_STATE) {   // a frame built with corners=false simply has none (frame.cc:52-53)
      Check(rc, "GetCorners");
    }
    corners_fetched_ = true;
  }
  return corners_;
}

const std::vector<int>& Frame::GetCornerScores() { GetCorners(); return corner_scores_; }

void Frame::CreateCorners(int /*levels*/, int nfeatures) {   // `levels` is ignored by the reference too (frame.cc:122-126)
  Check(sdvlb_frame_detect(ctx_, handle_, nfeatures), "CreateCorners");
  corners_fetched_ = false;
  descriptors_fetched_ = false;
}

std::vector<std::vector<unsigned char>>& Frame::GetDescriptors() {
  if (!descriptors_fetched_) {
    descriptors_.clear();
    if (Config::UseORB()) {
      const int n = int(GetCorners().size());
      std::vector<unsigned char> flat(size_t(n) * 32);
      int got = 0;
      if (n > 0) Check(sdvlb_frame_descriptors(ctx_, handle_, flat.data(), n, &got), "GetDescriptors");
      descriptors_.resize(size_t(n));
      for (int i = 0; i < n; i++) descriptors_[size_t(i)].assign(flat.begin() + size_t(i) * 32, flat.begin() + size_t(i + 1) * 32);
    }
    descriptors_fetched_ = true;
  }
  return descriptors_;
}

int Frame::GetNumPoints() const {
  int count = 0;
  for (auto it = features_.begin(); it != features_.end(); it++) {
    if (!(*it)) continue;
    if (!(*it)->GetPoint()) continue;
    count++;
  }
  return count;
}

bool Frame::Project(const Eigen::Vector3d& p3D, Eigen::Vector2d* p2D) {
  const Eigen::Vector3d rel = pose_ * p3D;
  if (rel(2) < 0.0) return false;
  camera_->Project(rel, p2D);
  return true;
}

void Frame::RemoveFeatures() {
  for (auto it = features_.begin(); it != features_.end(); it++) *it = nullptr;
  features_.clear();
}

void Frame::FilterCorners() {   // frame.cc:133-146: lock the cells that already hold features, best corner per free cell
  assert(filtered_corners_.empty());
  std::vector<double> locked;
  locked.reserve(2 * features_.size());
  for (auto it = features_.begin(); it != features_.end(); it++) {
    assert(*it != nullptr);
    locked.push_back((*it)->GetPosition()(0));
    locked.push_back((*it)->GetPosition()(1));
  }
  const int cap = int(std::ceil(width_ / double(Config::CellSize())) * std::ceil(height_ / double(Config::CellSize())));
  filtered_corners_.resize(size_t(cap));
  int n = 0;
  Check(sdvlb_frame_filter_corners(ctx_, handle_, locked.data(), int(locked.size() / 2), Config::MinFeatureScore(),
                                   filtered_corners_.data(), cap, &n),
        "FilterCorners");
  filtered_corners_.resize(size_t(n));
}

// ---------------------------------------------------------------- ImageAlign
ImageAlign::ImageAlign() { error_ = 1e10; }
ImageAlign::~ImageAlign() {}

void ImageAlign::CollectFeatures(const std::shared_ptr<Frame>& frame1, std::vector<sdvlb_align_feat>* out) {
  const Eigen::Vector3d first_pos = frame1->GetWorldPosition();
  std::vector<std::shared_ptr<Feature>>& features = frame1->GetFeatures();
  out->resize(features.size());
  size_t i = 0;
  for (auto it = features.begin(); it != features.end(); it++, i++) {
    const std::shared_ptr<Feature>& feature = *it;
    sdvlb_align_feat& f = (*out)[i];
    f.px[0] = feature->GetPosition()(0);
    f.px[1] = feature->GetPosition()(1);
    f.v[0] = feature->GetVector()(0); f.v[1] = feature->GetVector()(1); f.v[2] = feature->GetVector()(2);
    f.pad_ = 0;
    std::shared_ptr<Point> point = feature->GetPoint();
    if (point && !point->ToDelete()) {   // image_align.cc:154,229
      f.valid = 1;
      f.depth = (point->GetPosition() - first_pos).norm();   // image_align.cc:159,234
    } else {
      f.valid = 0;
      f.depth = 1.0;
    }
  }
}

int ImageAlign::ComputePose(const std::shared_ptr<Frame>& frame1, const std::shared_ptr<Frame>& frame2, bool fast) {
  assert(Config::PyramidLevels() >= Config::MaxAlignLevel());   // image_align.cc:52
  const int size = int(frame1->GetFeatures().size());
  if (size == 0) {
    std::cerr << "[ERROR] No points to track!" << std::endl;    // image_align.cc:55-58
    return 0;
  }
  std::vector<sdvlb_align_feat> feats;
  CollectFeatures(frame1, &feats);
  double T_ref[7], T_cur[7];
  frame1->GetPose().ToArray(T_ref);
  frame2->GetPose().ToArray(T_cur);
  int n_tracked = 0, trace_n = 0;
  Check(sdvlb_image_align(frame1->Context(), frame1->Handle(), frame2->Handle(), feats.data(), size, T_ref, T_cur,
                          fast ? 1 : 0, &n_tracked, &error_, nullptr, 0, &trace_n, nullptr),
        "ImageAlign::ComputePose");
  iterations_ = trace_n;
  frame2->SetPose(SE3(T_cur));   // image_align.cc:79
  return n_tracked;
}

// ---------------------------------------------------------------- Matcher
Matcher::Matcher(int size) { patch_size_ = size; }
Matcher::~Matcher() {}

void Matcher::FillCandidate(const std::shared_ptr<Feature>& feature, double idepth, double idepth_std, bool fixed,
                            sdvlb_candidate* c) {
  std::shared_ptr<Frame> ref_frame = feature->GetFrame();
  assert(ref_frame);
  c->ref_frame = ref_frame->Handle();
  ref_frame->GetPose().ToArray(c->ref_T);
  c->ref_px[0] = feature->GetPosition()(0); c->ref_px[1] = feature->GetPosition()(1);
  c->ref_v[0] = feature->GetVector()(0); c->ref_v[1] = feature->GetVector()(1); c->ref_v[2] = feature->GetVector()(2);
  c->idepth = idepth;
  c->idepth_std = idepth_std;
  c->px[0] = c->px[1] = 0.0;
  c->pos[0] = c->pos[1] = c->pos[2] = 0.0;
  c->ref_level = feature->GetLevel();
  c->flags = fixed ? SDVLB_CAND_FIXED : 0;
}

bool Matcher::SearchPoint(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Feature>& feature, double idepth,
                          double idepth_std, bool fixed, Eigen::Vector2d* px, int* flevel) {
  sdvlb_candidate c;
  FillCandidate(feature, idepth, idepth_std, fixed, &c);
  c.px[0] = (*px)(0); c.px[1] = (*px)(1);
  double T_cur[7];
  frame->GetPose().ToArray(T_cur);
  sdvlb_match m;
  if (Config::UseORB()) {   // matcher.cc:79-80,109: scored against feature->GetDescriptor()
    assert(feature->HasDescriptor());
    Check(sdvlb_search_points_orb(frame->Context(), frame->Handle(), &c, 1, T_cur, feature->GetDescriptor().data(), &m),
          "Matcher::SearchPoint");
  } else {
    Check(sdvlb_search_points(frame->Context(), frame->Handle(), &c, 1, T_cur, &m), "Matcher::SearchPoint");
  }
  if (m.status != SDVLB_MATCH_FOUND) return false;
  (*px)(0) = m.px[0]; (*px)(1) = m.px[1];
  *flevel = m.level;
  return true;
}

}  // namespace sdvl
