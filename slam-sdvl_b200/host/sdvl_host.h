// sdvl_host.h — host-side mirror of the SDVL class interfaces on the tracking hot path, implemented over the C-ABI
// (include/sdvl_b200.h).  Same names, argument meaning and return/error behaviour as the reference:
//   Frame          frame.h:43-172      (ctor, GetPyramid, GetCorners, CreateCorners, pose / feature accessors)
//   ImageAlign     image_align.h:33-66 (ComputePose, GetError)
//   Matcher        matcher.h:39-84     (SearchPoint)
//   FeatureAlign   feature_align.h:42-116 (Reproject, OptimizePose, GetMatches, GetAttempts)
// SE3 / Camera / Feature / Point / Config / Map are the reference's own (unchanged, out of scope) classes in an SDVL
// build; the reduced versions below carry exactly the members the hot path touches so that this library and its
// tests are self-contained (no OpenCV / Eigen in this image).
#pragma once
#include <algorithm>
#include <list>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/sdvl_b200.h"
#include "compat.h"

namespace sdvl {

class ORBDetector;   // the descriptors are computed on the device (csrc/orb.cu); pointer kept for signature compatibility
class Frame;
class Feature;
class Point;

// ------------------------------------------------------------------------------------------------ SE3 (extra/se3.h)
class SE3 {
 public:
  SE3() { q_[0] = 1; q_[1] = q_[2] = q_[3] = 0; t_[0] = t_[1] = t_[2] = 0; }
  explicit SE3(const double a[7]) { for (int i = 0; i < 4; i++) q_[i] = a[i]; for (int i = 0; i < 3; i++) t_[i] = a[4 + i]; }
  Eigen::Vector3d GetTranslation() const { return Eigen::Vector3d(t_[0], t_[1], t_[2]); }
  void GetRotation(double R[9]) const;
  SE3 Inverse() const;                                  // se3.cc:59-70
  static SE3 Exp(const double update[6]);               // se3.cc:72-94
  static void Log(const SE3& se3, double out[6]);       // se3.cc:96-112
  Eigen::Vector3d operator*(const Eigen::Vector3d& pos) const;   // se3.h:68
  SE3 operator*(const SE3& se3) const;                  // se3.cc:166-177
  void ToArray(double a[7]) const { for (int i = 0; i < 4; i++) a[i] = q_[i]; for (int i = 0; i < 3; i++) a[4 + i] = t_[i]; }
 private:
  double q_[4];   // w x y z
  double t_[3];
};

// ------------------------------------------------------------------------------------------------ Config (config.h)
// Static getters with the reference's names (config.h:64-104), backed by one sdvlb_params + camera.
class Config {
 public:
  static void Set(const sdvlb_params& p, const sdvlb_camera& cam) { params_() = p; camera_() = cam; }
  static const sdvlb_params& Params() { return params_(); }
  static const sdvlb_camera& CameraParams() { return camera_(); }
  static int PyramidLevels() { return params_().pyramid_levels; }
  static int CellSize() { return params_().cell_size; }
  static int MaxMatches() { return params_().max_matches; }
  static int MaxAlignLevel() { return params_().max_align_level; }
  static int MinAlignLevel() { return params_().min_align_level; }
  static int MaxImgAlignIts() { return params_().max_img_align_its; }
  static int AlignPatchSize() { return params_().align_patch_size; }
  static int PatchSize() { return params_().patch_size; }
  static int MaxAlignIts() { return params_().max_align_its; }
  static int SearchSize() { return params_().search_size; }
  static int MaxFastLevels() { return params_().max_fast_levels; }
  static int FastThreshold() { return params_().fast_threshold; }
  static int NumFeatures() { return params_().num_features; }
  static int MaxFailed() { return params_().max_failed; }
  static int MaxOptimPoseIts() { return params_().max_optim_pose_its; }
  static int MaxRansacPoints() { return params_().max_ransac_points; }
  static int MaxRansacIts() { return params_().max_ransac_its; }
  static int MinMatches() { return params_().min_matches; }
  static double InlierErrorThreshold() { return params_().inlier_error_threshold; }
  static int MinFeatureScore() { return 50; }   // kMinFeatureScore_ (config.cc:84)
  // Config::UseORB() / ORBSize() (config.h:139-140).  Set before the first Frame / tracker exists: contexts switch to
  // the ORB mode (sdvlb_ctx_set_orb) when they are created.
  static bool UseORB() { return use_orb_(); }
  static void SetUseORB(bool on) { use_orb_() = on; }
  static int ORBSize() { return 31; }
 private:
  static sdvlb_params& params_();
  static sdvlb_camera& camera_();
  static bool& use_orb_();
};

// ------------------------------------------------------------------------------------------------ Camera (camera.h)
class Camera {
 public:
  Camera();   // from Config (camera.cc:27-37)
  double GetWidth() const { return width_; }
  double GetHeight() const { return height_; }
  double GetFx() const { return fx_; }
  double GetFy() const { return fy_; }
  double GetU0() const { return u0_; }
  double GetV0() const { return v0_; }
  void Project(const Eigen::Vector3d& p3D, Eigen::Vector2d* p2D) const;     // camera.cc:69-72
  void Unproject(const Eigen::Vector2d& p2D, Eigen::Vector3d* p3D) const;   // camera.cc:74-79
  Eigen::Vector3d Unproject(const Eigen::Vector2d& p2D) const { Eigen::Vector3d r; Unproject(p2D, &r); return r; }
  bool IsInsideImage(const Eigen::Vector2i& p, int m = 0) const {           // camera.h:93-95
    return p(0) >= m && p(0) < width_ - m && p(1) >= m && p(1) < height_ - m;
  }
  static Eigen::Vector2d SimpleProject(const Eigen::Vector3d& p) { return Eigen::Vector2d(p(0) / p(2), p(1) / p(2)); }
  // camera.cc:38-67: (k1, k2, p1, p2, k3) of cv::undistort; all zero = no distortion
  void SetDistortions(double d0, double d1, double d2, double d3, double d4);
  bool HasDistortion() const { return has_distortion_; }
  const double* GetDistortions() const { return d_; }
  // camera.cc:100-105: cv::undistort on the device (sdvlb_undistort), a copy without distortion.  `in` must be
  // continuous.  A tracker that wants the undistortion fused into Frame construction instead calls
  // sdvlb_ctx_set_distortion(ctx, camera->GetDistortions()) once and hands the distorted image to Frame().
  void UndistortImage(const cv::Mat& in, cv::Mat* out) const;
  double GetPixelErrorAngle() const { return std::atan(1.0 / (2.0 * fx_)) * 2.0; }   // camera.h:104-107
 private:
  double width_, height_, fx_, fy_, u0_, v0_;
  double d_[5] = {0, 0, 0, 0, 0};
  bool has_distortion_ = false;
};

// ------------------------------------------------------------------------------------------------ Point (point.h)
class Point {
 public:
  enum PointStatus { P_FOUND, P_NOT_FOUND, P_SEEN, P_UNSEEN, P_OUTLIER };
  Point();
  double GetInverseDepth() { return rho_; }
  double GetStd();
  std::shared_ptr<Feature> GetInitFeature() { return feature_; }
  void SetInitFeature(const std::shared_ptr<Feature>& f) { feature_ = f; }
  int GetID() const { return id_; }
  Eigen::Vector3d GetPosition() const;                 // point.cc:128-142
  int Score() const { return n_successful_; }
  int GetLastFrame() const { return last_frame_; }
  void SetLastFrame(int id) { last_frame_ = id; }
  PointStatus GetStatus() const { return status_; }
  void SetStatus(PointStatus s) { status_ = s; }
  bool ToDelete() const { return delete_; }
  void SetDelete() { delete_ = true; }
  void SetFixed() { fixed_ = true; }
  bool IsFixed() { return fixed_; }
  bool Promote() { n_successful_++; n_failed_ = 0; return true; }                    // point.cc:102-106
  bool Unpromote() { n_failed_++; b_++; return n_failed_ > Config::MaxFailed(); }    // point.cc:108-115
  // map seeding (stand-in for InitCandidate + depth-filter convergence, point.cc:46-60,162-176)
  void InitFixed(const std::shared_ptr<Feature>& f, const Eigen::Vector3d& p3d, double rho, double sigma2);
  void InitCandidate(const std::shared_ptr<Feature>& p, double depth);                // point.cc:48-61
  // The depth-filter state as the device sees it (Map::UpdateCandidates): Point::Update and HasConverged
  // (point.cc:63-100,162-176) run inside sdvlb_update_candidates; these copy the state in and out.
  void ToSeed(sdvlb_seed* s) const;
  void FromSeed(const sdvlb_seed& s);
  int GetLastKeyframeID() const { return last_kf_id_; }        // GetLastFeature()->GetFrame()->GetKeyframeID()
  void SetLastKeyframeID(int id) { last_kf_id_ = id; }
 private:
  int id_;
  PointStatus status_ = P_FOUND;
  bool delete_ = false, fixed_ = false;
  int last_frame_ = -1, n_successful_ = 0, n_failed_ = 0, last_kf_id_ = 0;
  double a_ = 10, b_ = 10, rho_ = 1.0, sigma2_ = 1.0, z_range_ = 6.0, cos_alpha_ = 1.0, last_distance_ = 1.0;
  std::shared_ptr<Feature> feature_;
  Eigen::Vector3d p3d_;
};

// ------------------------------------------------------------------------------------------------ Feature (feature.h)
class Feature {
 public:
  Feature(const std::shared_ptr<Frame>& f, const Eigen::Vector2d& p, int l);          // feature.cc:28-36
  std::shared_ptr<Frame> GetFrame() { return frame_; }
  std::shared_ptr<Point> GetPoint() const { return point_; }
  void SetPoint(const std::shared_ptr<Point>& p) { point_ = p; }
  const Eigen::Vector2d& GetPosition() const { return p2d_; }
  const Eigen::Vector3d& GetVector() const { return v_; }
  int GetLevel() const { return level_; }
  Eigen::Vector2d GetLevelPosition() { return p2d_ / double(1 << level_); }           // feature.h:93-95
  // feature.h:78-89: the ORB descriptor of the feature (32 zero bytes until one is set)
  const std::vector<unsigned char>& GetDescriptor() const { return descriptor_; }
  void SetDescriptor(const std::vector<unsigned char>& d) { std::copy(d.begin(), d.end(), descriptor_.begin()); has_descriptor_ = true; }
  bool HasDescriptor() const { return has_descriptor_; }
 private:
  std::vector<unsigned char> descriptor_ = std::vector<unsigned char>(32, 0);
  bool has_descriptor_ = false;
  std::shared_ptr<Frame> frame_;
  std::shared_ptr<Point> point_;
  Eigen::Vector2d p2d_;
  Eigen::Vector3d v_;
  int level_;
};

// ------------------------------------------------------------------------------------------------ Map stand-in (map.h)
// What FeatureAlign needs -- DeletePoint (map.cc:165-168), EmptyTrash (map.cc:207-253, points part) -- and the
// mapping thread's candidate pass, UpdateCandidates (map.cc:397-498), which evaluates all candidates of a frame in one
// device call (sdvlb_update_candidates) and then applies the list surgery in the reference's order.
class Map {
 public:
  void DeletePoint(const std::shared_ptr<Point>& point);
  void EmptyTrash();
  std::mutex& GetMutex() { return mutex_map_; }
  void AddCandidate(const std::shared_ptr<Point>& p) { candidates_.push_back(p); }
  std::vector<std::shared_ptr<Point>>& GetCandidates() { return candidates_; }
  // depth_mean = frame->GetSceneDepth(); min_kf_id = last_kf_->GetKeyframeID() - 2 * Config::MaxSearchKeyframes()
  void UpdateCandidates(const std::shared_ptr<Frame>& frame, double depth_mean, int min_kf_id);
  // per list entry of the last pass (status -1: the entry was not evaluated)
  const std::vector<sdvlb_seed>& LastSeeds() const { return seeds_; }
  // map.cc:262-395.  best_kfs = frame->GetBestConnections(Config::MaxSearchKeyframes()) (the keyframe graph is the
  // caller's); depth_mean = frame->GetSceneDepth().  Returns the number of candidates created.
  int InitCandidates(const std::shared_ptr<Frame>& frame, const std::vector<std::shared_ptr<Frame>>& best_kfs,
                     double depth_mean);
  // map.cc:560-617: points seen from the connected keyframes but not from `frame` are searched in it (one
  // sdvlb_search_points batch) and linked when found.  Returns the number of links made.
  int AddConnectionsPoints(const std::shared_ptr<Frame>& frame, const std::vector<std::shared_ptr<Frame>>& best_kfs);
 private:
  std::mutex mutex_map_;
  std::vector<std::shared_ptr<Point>> points_trash_;
  std::vector<std::shared_ptr<Point>> candidates_;
  std::vector<sdvlb_seed> seeds_;
};

// ------------------------------------------------------------------------------------------------ device context
// Every Frame lives on one sdvlb_ctx. A thread picks its context with Device::SetCurrent; by default one context on
// device 0 is created from Config the first time it is needed. Fails loudly (std::runtime_error) without CUDA.
class Device {
 public:
  static sdvlb_ctx* Current();
  static void SetCurrent(sdvlb_ctx* ctx);
};

// ------------------------------------------------------------------------------------------------ Frame (frame.h)
class Frame {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  Frame(Camera* camera, ORBDetector* detector, const cv::Mat& img, bool corners);   // frame.cc:34-56
  Frame(Camera* camera, sdvlb_ctx* ctx, sdvlb_frame* adopted, int id);              // batched construction
  ~Frame();
  bool IsKeyframe() { return is_keyframe_; }
  void SetKeyframe() { is_keyframe_ = true; }
  SE3& GetPose() { return pose_; }
  const SE3& GetPose() const { return pose_; }
  void SetPose(const SE3& se3) { pose_ = se3; }
  std::vector<cv::Mat>& GetPyramid();                          // host mirror, fetched on first use
  std::vector<std::shared_ptr<Feature>>& GetFeatures() { return features_; }
  std::vector<Eigen::Vector3i>& GetCorners();                  // host mirror, fetched on first use
  std::vector<Eigen::Vector2d>& GetOutliers() { return outliers_; }
  Camera* GetCamera() const { return camera_; }
  int GetWidth() const { return width_; }
  int GetHeight() const { return height_; }
  int GetID() const { return id_; }
  void SetID(int id) { id_ = id; }   // extension: per-sequence ids when several sequences share a process
  SE3 GetWorldPose() const { return pose_.Inverse(); }
  Eigen::Vector3d GetWorldPosition() const { return pose_.Inverse().GetTranslation(); }
  void AddFeature(const std::shared_ptr<Feature>& f) { features_.push_back(f); }
  void AddOutlier(const Eigen::Vector2d& p) { outliers_.push_back(p); }
  int GetNumFeatures() const { return int(features_.size()); }
  int GetNumPoints() const;                                    // frame.cc:165-180
  bool Project(const Eigen::Vector3d& p3D, Eigen::Vector2d* p2D);   // frame.cc:93-102
  void CreateCorners(int levels, int nfeatures);               // frame.cc:122-131
  void RemoveFeatures();                                       // frame.cc:214-219
  void FilterCorners();                                        // frame.cc:133-163 (use_orb == 0)
  std::vector<int>& GetFilteredCorners() { return filtered_corners_; }
  // frame.h: per-corner ORB descriptors.  The reference fills them lazily (frame.cc:148-161, matcher.cc:265-269); the
  // device computes all of them when the frame is built, this fetches the host copy on first use (ORB mode only).
  std::vector<std::vector<unsigned char>>& GetDescriptors();
  // device side
  sdvlb_frame* Handle() const { return handle_; }
  sdvlb_ctx* Context() const { return ctx_; }
  const std::vector<int>& GetCornerScores();                   // cv::KeyPoint::response, parity checks only
 private:
  int id_;
  Camera* camera_;
  sdvlb_ctx* ctx_;
  sdvlb_frame* handle_;
  bool is_keyframe_ = false;
  bool pyramid_fetched_ = false, corners_fetched_ = false;
  std::vector<cv::Mat> pyramid_;
  int width_, height_;
  SE3 pose_;
  std::vector<std::shared_ptr<Feature>> features_;
  std::vector<Eigen::Vector3i> corners_;
  std::vector<int> corner_scores_;
  std::vector<Eigen::Vector2d> outliers_;
  std::vector<int> filtered_corners_;
  std::vector<std::vector<unsigned char>> descriptors_;
  bool descriptors_fetched_ = false;
  static int counter_;
};

// ------------------------------------------------------------------------------------------------ ImageAlign
class ImageAlign {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  ImageAlign();
  ~ImageAlign();
  // Compute Pose between frames (image_align.cc:46-84). Returns n_meas/patch_area.
  int ComputePose(const std::shared_ptr<Frame>& frame1, const std::shared_ptr<Frame>& frame2, bool fast = false);
  double GetError() { return error_; }
  // Marshals frame1's features into the C-ABI layout (shared with the batched tracker).
  static void CollectFeatures(const std::shared_ptr<Frame>& frame1, std::vector<sdvlb_align_feat>* out);
  int GetIterations() const { return iterations_; }
 private:
  double error_;
  int iterations_ = 0;
};

// ------------------------------------------------------------------------------------------------ Matcher
class Matcher {
 public:
  explicit Matcher(int size);
  ~Matcher();
  // Search a point in current frame close to a epipolar line (matcher.cc:45-121)
  bool SearchPoint(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Feature>& feature, double idepth,
                   double idepth_std, bool fixed, Eigen::Vector2d* px, int* flevel);
  // Fills one C-ABI candidate from the same arguments.
  static void FillCandidate(const std::shared_ptr<Feature>& feature, double idepth, double idepth_std, bool fixed,
                            sdvlb_candidate* c);
 private:
  int patch_size_;
};

// ------------------------------------------------------------------------------------------------ FeatureAlign
typedef std::pair<std::shared_ptr<Point>, Eigen::Vector2d> PointInfo;
typedef std::list<PointInfo> GridCell;

class HostRand {   // glibc rand() stream (TYPE_3, seed 1), one per FeatureAlign so sequences do not interleave
 public:
  HostRand() { Seed(1); }
  void Seed(unsigned s) { sdvlb_rand_seed(&s_, s); }
  int Next() { return sdvlb_rand_next(&s_); }
  sdvlb_rand* State() { return &s_; }
 private:
  sdvlb_rand s_;
};

class FeatureAlign {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  FeatureAlign(Map* map, Camera* camera, int max_matches);     // feature_align.cc:33-54
  ~FeatureAlign();
  // Reproject all points in current frame found in other frames (feature_align.cc:59-71)
  void Reproject(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Frame>& last_frame,
                 const std::shared_ptr<Frame>& last_kf, bool reloc = false);
  // Minimize reprojection error of a single frame (feature_align.cc:73-82)
  bool OptimizePose(const std::shared_ptr<Frame>& frame);
  int GetMatches() { return matches_; }
  int GetAttempts() { return num_attempts_; }

  // Batched form used by the multi-sequence tracker: the GPU evaluates ProjectPoint + SearchPoint for every point
  // ProjectPoints would visit (CollectCandidates), then ApplyMatches replays ProjectPoint's bookkeeping and the
  // SelectPoints / SelectInliers logic on the results, with identical side effects.
  // descs (ORB mode): 32 bytes per candidate, feature->GetDescriptor() of its init feature (matcher.cc:109)
  void CollectCandidates(int frame_id, const std::shared_ptr<Frame>& last_frame, bool reloc,
                         std::vector<sdvlb_candidate>* cands, std::vector<std::shared_ptr<Point>>* points,
                         std::vector<unsigned char>* descs = nullptr);
  void ApplyMatches(const std::shared_ptr<Frame>& frame, const std::vector<std::shared_ptr<Point>>& points,
                    const sdvlb_match* matches);
  int GetInliers() const { return int(inliers_.size()); }
  int GetOutliers() const { return int(outliers_.size()); }
  double ransac_seconds = 0;   // time spent in SelectInliers (host phase accounting)
 private:
  // RANSAC (feature_align.cc:152-216) and OptimizePose's Gauss-Newton rounds are device calls
  // (sdvlb_select_inliers / sdvlb_optimize_pose); the host only keeps the inlier / outlier lists.
  void SelectInliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>& fs_found,
                     std::vector<std::shared_ptr<Feature>>* inliers, std::vector<std::shared_ptr<Feature>>* outliers);
  void RemoveOutliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>* outliers);
  void ResetGrid();

  Map* map_;
  int cell_size_, max_matches_, grid_width_, grid_height_;
  std::vector<GridCell> grid_;
  std::vector<int> cell_order_;
  int matches_, num_attempts_;
  bool relocalizing_;
  std::vector<std::shared_ptr<Feature>> inliers_, outliers_;
  HostRand rng_;
};

void RandomShuffle(std::vector<int>* v, HostRand* rng);   // libstdc++ std::random_shuffle with rand()

}  // namespace sdvl
