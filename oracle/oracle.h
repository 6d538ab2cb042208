// oracle.h — CPU restatement of the SDVL tracking front-end (TEST INFRASTRUCTURE ONLY).
//
// This directory is the parity oracle for the B200 path.  It restates, function by
// function, the reference algorithm (JdeRobot/slam-SDVL; citations are file:line in
// the reference tree).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it; nothing under slam-sdvl_b200/
// links against or calls it.
//
// PARITY PIN STATUS: PINNED against the reference's own code.
//   * oracle/_ref/libsdvlref.so = the reference's hot-path sources compiled UNMODIFIED from /root/reference
//     against the stand-in Eigen / OpenCV headers of oracle/ref_shim (oracle/Makefile target `ref`).
//     tests/test_oracle_vs_ref.py runs the same seeded inputs through both: with -ffp-contract=off on both sides
//     every output (GN traces, matched pixels, refined poses, 45-frame trajectories) is bit-identical; with the
//     default flags they agree to rounding.  tests/golden/ref_golden.npz carries reference outputs for machines
//     without the reference tree.
//   * pyramid (cv::pyrDown), FAST (cv::FAST) and cv::undistort are pinned bit-exactly against OpenCV 4.13 via
//     python cv2 (tests/test_oracle_cpu.py, tests/golden/cv2_golden.npz); retainBest against libstdc++.
//   * rand()/random_shuffle are pinned against this machine's glibc/libstdc++.
//   * not pinned by reference code: the Eigen primitives (LDLT, small inverses, quaternion conversions), restated
//     from Eigen 3's published algorithms both here and in the stand-in header (Eigen is not installed).
#ifndef SDVL_ORACLE_H_
#define SDVL_ORACLE_H_

#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <utility>
#include <vector>

#include "../include/sdvl_b200.h"  // POD structs only (params, camera, trace records)

namespace oracle {

// ---------------------------------------------------------------- small linear algebra
struct V2 { double x = 0, y = 0; };
struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
  V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
  V3 operator*(double s) const { return V3(x * s, y * s, z * s); }
  double dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
  double norm() const { return std::sqrt(dot(*this)); }
};
struct M3 { double m[3][3]; };
typedef double Vec6[6];
typedef double Mat6[6][6];

// Eigen::Matrix<double,6,6>::ldlt().solve (image_align.cc:102, feature_align.cc:402):
// diagonal-pivoted LDL^T with pseudo-inverse of D.
void LdltSolve6(const Mat6 H, const Vec6 b, Vec6 x);

// ---------------------------------------------------------------- SE3 (extra/se3.cc)
struct SE3 {
  double q0 = 1, q1 = 0, q2 = 0, q3 = 0;  // w x y z
  V3 t;
  M3 Rotation() const;                 // se3.h:41 (Quaterniond::toRotationMatrix)
  SE3 Inverse() const;                 // se3.cc:59-70
  SE3 operator*(const SE3& o) const;   // se3.cc:166-177
  V3 operator*(const V3& p) const;     // se3.h:68
  static SE3 Exp(const Vec6 u);        // se3.cc:72-94
  static void Log(const SE3& s, Vec6 out);  // se3.cc:96-112
  void ToArray(double a[7]) const { a[0] = q0; a[1] = q1; a[2] = q2; a[3] = q3; a[4] = t.x; a[5] = t.y; a[6] = t.z; }
  static SE3 FromArray(const double a[7]) {
    SE3 s; s.q0 = a[0]; s.q1 = a[1]; s.q2 = a[2]; s.q3 = a[3]; s.t = V3(a[4], a[5], a[6]); return s;
  }
};

// ---------------------------------------------------------------- Camera (camera.cc/.h)
struct Camera {
  double width, height, fx, fy, u0, v0;
  void Project(const V3& p, V2* out) const {  // camera.cc:69-72
    out->x = u0 + fx * p.x / p.z;
    out->y = v0 + fy * p.y / p.z;
  }
  V3 Unproject(const V2& p) const;            // camera.cc:74-79
  bool IsInsideImage(int x, int y, int m = 0) const {  // camera.h:93-95
    return x >= m && x < width - m && y >= m && y < height - m;
  }
  bool IsInsideImage(int x, int y, int m, int l) const {  // camera.h:96-98
    return x >= m && x < width / (1 << l) - m && y >= m && y < height / (1 << l) - m;
  }
};

// ---------------------------------------------------------------- images
struct Mat8 {  // continuous CV_8UC1
  int cols = 0, rows = 0;
  std::vector<uint8_t> data;
  Mat8() {}
  Mat8(int c, int r) : cols(c), rows(r), data(size_t(c) * r) {}
  const uint8_t* ptr(int y) const { return data.data() + size_t(y) * cols; }
};

// cv::pyrDown(src, dst, Size(cols/2, rows/2)) for CV_8UC1 (call site frame.cc:119).
void PyrDown(const Mat8& src, Mat8* dst);

struct KeyPoint { float x, y, response; };
// cv::FAST(roi, kps, threshold, nonmax=true), TYPE_9_16 (call site fast_detector.cc:95).
void FastRoi(const uint8_t* roi, int stride, int cols, int rows, int threshold, std::vector<KeyPoint>* out);
// cv::KeyPointsFilter::retainBest (call sites fast_detector.cc:140,148).
void RetainBest(std::vector<KeyPoint>* kps, int n);

struct Corner { int x, y, level; };

// FastDetector::SelectPixels / DetectPyramid (fast_detector.cc:58-175).
void SelectPixels(const sdvlb_params& P, const Mat8& src, int level, int nfeatures,
                  std::vector<Corner>* corners, std::vector<int>* scores);
void DetectPyramid(const sdvlb_params& P, const std::vector<Mat8>& pyr, int nfeatures,
                   std::vector<Corner>* corners, std::vector<int>* scores);

// FindShiTomasiScoreAtPoint (extra/utils.cc:61-97) and FastDetector::FilterCorners (fast_detector.cc:177-218) as
// Frame::FilterCorners drives it (frame.cc:133-146): locked = level-0 positions of the frame's features.
double FindShiTomasiScoreAtPoint(const Mat8& img, int px, int py);
void FilterCorners(const sdvlb_params& P, const std::vector<Mat8>& pyr, const std::vector<Corner>& corners,
                   const std::vector<V2>& locked, int min_feature_score, std::vector<int>* indices);

// Camera::UndistortImage (camera.cc:100-105) = cv::undistort(in, out, K, D) with D = (k1, k2, p1, p2, k3)
// (camera.cc:38-67), i.e. cv::initUndistortRectifyMap(K, D, I, K, size, CV_16SC2) + cv::remap(INTER_LINEAR,
// BORDER_CONSTANT 0).  Pinned bit-exactly against OpenCV 4.13 (tests/golden, tests/test_oracle_cpu.py).
void UndistortImage(const Camera& cam, const double d[5], const Mat8& in, Mat8* out);

// ---------------------------------------------------------------- ORB descriptor mode (extra/orb_detector.cc)
const int kOrbSize = 31;               // Config::ORBSize(); the only size the learned pattern exists for
void SetUseOrb(int on);                // Config::UseORB() for this process (the oracle is single-threaded test code)
bool UseOrb();
int BorderMargin(const sdvlb_params& P);   // 4 + orb_size/2 with ORB, 1 + patch_size/2 without
float FastAtan2(float y, float x);     // cv::fastAtan2
class OrbDetector {
 public:
  OrbDetector();
  bool IsInsideLimits(const Mat8& src, int x, int y) const;
  double GetOrientation(const Mat8& src, int x, int y) const;
  void GetDescriptor(const Mat8& src, int x, int y, uint8_t desc[32]) const;
  static int Distance(const uint8_t* a, const uint8_t* b);
 private:
  std::vector<int> umax_;
};
struct Frame;
const std::vector<uint8_t>& CornerDescriptor(Frame* f, int index);

// ---------------------------------------------------------------- utils (extra/utils.cc)
double AbsMax6(const Vec6 v);                              // utils.cc:28-42
float Interpolate8U(const Mat8& m, float u, float v);      // utils.cc:44-59
void Jacobian3DToPlane(const V3& p, double J[2][6]);       // utils.cc:99-118
double GetMedianVector(std::vector<double>* v);            // utils.cc:215-220

// glibc rand() (TYPE_3, default seed 1) and libstdc++ std::random_shuffle, per sequence.
struct GlibcRand {
  uint32_t r[34];   // sliding window of the additive-feedback sequence o[i] = o[i-31] + o[i-3]
  int n;            // number of values produced so far (window position)
  GlibcRand() { Seed(1); }
  void Seed(unsigned s);
  int Next();
};
void RandomShuffle(std::vector<int>* v, GlibcRand* rng);   // bits/stl_algo.h random_shuffle

// ---------------------------------------------------------------- data model
struct Frame; struct Feature; struct Point;

struct Point {  // point.h:121-146 (fields the path reads)
  int id = 0;
  bool del = false;           // delete_
  bool fixed = false;         // fixed_
  V3 p3d;                     // p3d_
  double rho = 1.0;           // rho_
  double sigma2 = 1.0;        // sigma2_
  int n_successful = 0, n_failed = 0;
  double b_ = 10;
  int last_frame = -1;
  int status = 0;
  std::shared_ptr<Feature> feature;  // init feature
  V3 GetPosition() const;     // point.cc:128-142
  double GetStd() const { return std::sqrt(sigma2); }
  bool Unpromote(int max_failed) { n_failed++; b_++; return n_failed > max_failed; }  // point.cc:108-115
  void Promote() { n_successful++; n_failed = 0; }   // point.cc:102-106
};

struct Feature {  // feature.h:97-104
  std::shared_ptr<Frame> frame;
  std::shared_ptr<Point> point;
  V2 p2d;
  V3 v;
  int level = 0;
  std::vector<uint8_t> descriptor;   // descriptor_ (32 bytes once set; ORB mode only)
};

struct Frame {  // frame.h:148-172
  int id = 0;
  const Camera* cam = nullptr;
  std::vector<Mat8> pyramid;
  std::vector<Corner> corners;
  std::vector<int> corner_scores;
  std::vector<std::vector<uint8_t>> descriptors;   // descriptors_: per corner, filled lazily (ORB mode only)
  std::vector<std::shared_ptr<Feature>> features;
  std::vector<V2> outliers;
  SE3 pose;  // world -> camera
  bool is_keyframe = false;
  V3 GetWorldPosition() const { return pose.Inverse().t; }  // frame.h:93
  SE3 GetWorldPose() const { return pose.Inverse(); }       // frame.h:90
  bool Project(const V3& p3d, V2* p2d) const;               // frame.cc:93-102
};

std::shared_ptr<Frame> MakeFrame(const sdvlb_params& P, const Camera* cam, const uint8_t* img, int w, int h,
                                 bool corners, int id);   // frame.cc:34-56
std::shared_ptr<Feature> MakeFeature(const std::shared_ptr<Frame>& f, const V2& p, int level);  // feature.cc:28-36

// ---------------------------------------------------------------- ImageAlign (image_align.cc)
class ImageAlign {
 public:
  explicit ImageAlign(const sdvlb_params& P) : P_(P) {}
  int ComputePose(const std::shared_ptr<Frame>& f1, const std::shared_ptr<Frame>& f2, bool fast = false);
  double GetError() const { return error_; }
  std::vector<sdvlb_gn_iter> trace;   // one record per Optimize iteration (test hook)
 private:
  void Optimize(SE3* se3, int level);
  double ComputeResiduals(const SE3& se3, int level, bool linearize, bool patches);
  void PrecomputePatches(int level);
  sdvlb_params P_;
  std::shared_ptr<Frame> frame1_, frame2_;
  double chi2_ = 1e10;
  std::vector<float> patch_cache_;
  std::vector<bool> visible_fts_;
  size_t n_meas_ = 0;
  bool stop_ = false;
  double error_ = 1e10;
  Mat6 H_;
  Vec6 Jres_;
  std::vector<double> jacobian_cache_;  // 6 x (N*area), column-major
};

// ---------------------------------------------------------------- Matcher (matcher.cc)
struct SearchDebug { int slevel = -1, zmssd = -1, n_in_range = 0; };
class Matcher {
 public:
  Matcher(const sdvlb_params& P, int size) : P_(P), patch_size_(size) {}
  bool SearchPoint(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Feature>& feature, double idepth,
                   double idepth_std, bool fixed, V2* px, int* flevel, SearchDebug* dbg = nullptr);
  // exposed for unit tests
  void WarpMatrixAffine(const Camera* cam, const V2& px, const V3& v, double depth, const SE3& pose, int level,
                        double A[2][2]);
  int GetSearchLevel(const double A[2][2]);
  void CreatePatch(const double A[2][2], const Mat8& img, const V2& px, int level, int search_level,
                   uint8_t* border_patch, uint8_t* patch);
  bool AlignPatch(const Mat8& img, uint8_t* border_patch, uint8_t* patch, V2* px);
 private:
  void GetCornersInRange(const std::shared_ptr<Frame>& frame, const V2& pxa, const V2& pxb, int level, double range,
                         std::vector<int>* indices);
  void GetCornersInRange(const std::shared_ptr<Frame>& frame, const V2& cpos, int level, double range,
                         std::vector<int>* indices);
  bool SearchFeatures(const std::shared_ptr<Frame>& frame, const std::vector<int>& indices, uint8_t* patch, V2* px,
                      int* best);
  sdvlb_params P_;
  int patch_size_;
  uint8_t patch_[64 * 4];
  uint8_t border_patch_[100 * 4];
  const uint8_t* desc_ = nullptr;   // feature->GetDescriptor() of the current SearchPoint (ORB mode)
};

// ---------------------------------------------------------------- FeatureAlign (feature_align.cc)
typedef std::pair<std::shared_ptr<Point>, V2> PointInfo;
typedef std::list<PointInfo> GridCell;

class FeatureAlign {
 public:
  FeatureAlign(const sdvlb_params& P, const Camera* cam, int max_matches, GlibcRand* rng,
               std::vector<std::shared_ptr<Point>>* trash);
  void Reproject(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Frame>& last_frame, bool reloc = false);
  bool OptimizePose(const std::shared_ptr<Frame>& frame);
  int GetMatches() const { return matches_; }
  int GetAttempts() const { return num_attempts_; }
  // test hooks
  bool ConvergePose(const std::shared_ptr<Frame>& frame, const std::vector<std::shared_ptr<Feature>>& features, SE3* se3);
  int n_inliers() const { return int(inliers_.size()); }
  int n_outliers() const { return int(outliers_.size()); }
  std::vector<std::shared_ptr<Feature>> selected_;   // fs_found of the last Reproject
  // test hooks for the pose-refinement parity tests: run SelectInliers on a given fs_found / preset the two lists
  void SelectInliersHook(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>& fs_found) {
    SelectInliers(frame, fs_found, &inliers_, &outliers_);
  }
  void SetRng(GlibcRand* rng) { rng_ = rng; }
  std::vector<std::shared_ptr<Feature>>& inliers() { return inliers_; }
  std::vector<std::shared_ptr<Feature>>& outliers() { return outliers_; }
 private:
  void SelectPoints(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Frame>& last_frame,
                    std::vector<std::shared_ptr<Feature>>* fs_found);
  void SelectInliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>& fs_found,
                     std::vector<std::shared_ptr<Feature>>* inliers, std::vector<std::shared_ptr<Feature>>* outliers);
  void OptimizePose(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>* features,
                    std::vector<std::shared_ptr<Feature>>* outliers);
  bool RescueOutliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>* inliers,
                      std::vector<std::shared_ptr<Feature>>* outliers);
  void RemoveOutliers(const std::shared_ptr<Frame>& frame, std::vector<std::shared_ptr<Feature>>* outliers);
  int CheckReprojectionError(const std::vector<std::shared_ptr<Feature>>& features, const SE3& se3, double threshold,
                             std::vector<std::shared_ptr<Feature>>* inliers = nullptr,
                             std::vector<std::shared_ptr<Feature>>* outliers = nullptr);
  void ResetGrid();
  void ProjectPoints(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Frame>& last_frame);
  bool ProjectPoint(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Point>& point);
  static double GetTukeyValue(double x);

  sdvlb_params P_;
  const Camera* cam_;
  GlibcRand* rng_;
  std::vector<std::shared_ptr<Point>>* trash_;   // Map::DeletePoint target (map.cc:165-168)
  int cell_size_, max_matches_, grid_width_, grid_height_;
  std::vector<GridCell> grid_;
  std::vector<int> cell_order_;
  int matches_ = 0, num_attempts_ = 0;
  bool relocalizing_ = false;
  std::vector<std::shared_ptr<Feature>> inliers_, outliers_;
};

// ---------------------------------------------------------------- tracking driver
// Stand-in for SDVL::HandleFrame in STATE_RUNNING (sdvl.cc:90-95,179-203,266-281) plus a
// ground-truth seeded map instead of HomographyInit/Map (out of scope).
struct SeedPlane {       // world plane n.X = d, used to give seeded points their depth
  double n[3]; double d;
};
struct TrackStats { int n_tracked = 0, matches = 0, attempts = 0, inliers = 0, outliers = 0, n_feats = 0, gn_iters = 0, keyframe = 0; };

class Tracker {
 public:
  Tracker(const sdvlb_params& P, const Camera& cam, const SeedPlane& plane, int max_points, int kf_every);
  ~Tracker();   // breaks the Point <-> init Feature -> keyframe cycles of every point it created
  // First frame: pose given (ground truth), becomes keyframe and seeds the map.
  // Later frames: motion model prior -> ImageAlign -> FeatureAlign -> motion model update.
  // gt_pose is used only to place seeded points when this frame becomes a keyframe.
  void HandleFrame(const uint8_t* img, int w, int h, const SE3& gt_pose, SE3* est_pose, TrackStats* st);
  std::shared_ptr<Frame> last_frame() const { return last_frame_; }
 private:
  void SeedKeyframe(const std::shared_ptr<Frame>& f, const SE3& gt_pose);
  void EmptyTrash();
  sdvlb_params P_;
  Camera cam_;
  SeedPlane plane_;
  int max_points_, kf_every_;
  GlibcRand rng_;
  std::vector<std::shared_ptr<Point>> trash_;
  std::unique_ptr<FeatureAlign> feature_align_;
  std::shared_ptr<Frame> last_frame_, last_kf_;
  Vec6 vel_;
  int frame_counter_ = 0, point_counter_ = 0, last_matches_ = 0;
  std::vector<std::weak_ptr<Point>> all_points_;
};

}  // namespace oracle
#endif  // SDVL_ORACLE_H_
