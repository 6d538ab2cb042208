// OpenCV stand-in (TEST INFRASTRUCTURE): the cv:: types and functions the reference's hot-path sources use, so that they
// compile unmodified from /root/reference without OpenCV installed (oracle/Makefile, target _ref).  The three image
// operations the path calls -- cv::pyrDown, cv::FAST, cv::KeyPointsFilter::retainBest (and cv::undistort) -- are the
// oracle's restatements, which are pinned bit-exactly against OpenCV 4.13 (tests/test_oracle_cpu.py, tests/golden).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6
#define CV_PI 3.1415926535897932384626433832795

inline int cvRound(double v) { return int(std::nearbyint(v)); }   // round half to even (SSE2 cvtsd2si)
inline int cvFloor(double v) { return int(std::floor(v)); }
inline int cvCeil(double v) { return int(std::ceil(v)); }

namespace cv {

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<int> Point;
typedef Point_<float> Point2f;
struct KeyPoint {
  Point2f pt;
  float size = 7.f, angle = -1.f, response = 0.f;
  int octave = 0, class_id = -1;
  KeyPoint() {}
  KeyPoint(float x, float y, float s, float a = -1.f, float r = 0.f) : pt(x, y), size(s), angle(a), response(r) {}
};

class Mat {
 public:
  struct Step { size_t p[2]; operator size_t() const { return p[0]; } };
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  Step step = {{0, 0}};
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, double fill) {
    create(r, c, type);
    if (type == CV_64F) for (int i = 0; i < r * c; i++) reinterpret_cast<double*>(data)[i] = fill;
    else if (type == CV_32F) for (int i = 0; i < r * c; i++) reinterpret_cast<float*>(data)[i] = float(fill);
    else std::memset(data, int(fill), size_t(r) * c);
  }
  Mat(int r, int c, int type, void* ptr, size_t stp = 0) : rows(r), cols(c), data(static_cast<uchar*>(ptr)), type_(type) {
    step.p[1] = esz(type);
    step.p[0] = stp ? stp : size_t(c) * esz(type);
  }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step.p[1] = esz(type);
    step.p[0] = size_t(c) * esz(type);
    owner_.reset(new uchar[std::max<size_t>(1, size_t(r) * step.p[0])], std::default_delete<uchar[]>());
    data = owner_.get();
  }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step.p[0] == size_t(cols) * esz(type_); }
  size_t step1() const { return step.p[0] / esz(type_); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; y++) std::memcpy(m.data + size_t(y) * m.step.p[0], data + size_t(y) * step.p[0], size_t(cols) * esz(type_));
    return m;
  }
  template <typename T> T& at(int y, int x) { return *reinterpret_cast<T*>(data + size_t(y) * step.p[0] + size_t(x) * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *reinterpret_cast<const T*>(data + size_t(y) * step.p[0] + size_t(x) * sizeof(T)); }
  template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + size_t(y) * step.p[0]); }
  template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + size_t(y) * step.p[0]); }
  Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + size_t(a) * step.p[0]; m.rows = b - a; return m; }
  Mat colRange(int a, int b) const { Mat m = *this; m.data = data + size_t(a) * esz(type_); m.cols = b - a; return m; }
  // only Triangulate (extra/utils.cc:160-191, not on the path) uses these: never executed
  Mat row(int) const { std::abort(); }
  Mat t() const { std::abort(); }
 private:
  static size_t esz(int type) { return type == CV_64F ? 8 : type == CV_32F ? 4 : 1; }
  std::shared_ptr<uchar> owner_;
  int type_ = CV_8U;
};
inline Mat operator*(float, const Mat&) { std::abort(); }
inline Mat operator-(const Mat&, const Mat&) { std::abort(); }
inline Mat operator/(const Mat&, float) { std::abort(); }
template <typename T> class Mat_ : public Mat {
 public:
  Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_64F) {}
  struct Init {
    Mat_* m; int k;
    Init& operator,(T v) { reinterpret_cast<T*>(m->data)[k++] = v; return *this; }
    operator Mat() const { return *m; }
  };
  Init operator<<(T v) { Init i{this, 0}; reinterpret_cast<T*>(data)[i.k++] = v; return i; }
};
struct SVD {
  enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
  static void compute(const Mat&, Mat&, Mat&, Mat&, int = 0) { std::abort(); }
};

// cv::FileStorage for Config::ReadParameters (config.cc:88-163): flat "key: value" YAML as in the reference's *.cfg files.
struct Exception : std::exception { const char* what() const noexcept override { return "cv::Exception"; } };
class FileNode {
 public:
  FileNode() {}
  explicit FileNode(const std::string* v) : v_(v) {}
  bool isNamed() const { return v_ != nullptr; }
  const std::string& str() const { return *v_; }
 private:
  const std::string* v_ = nullptr;
};
inline void operator>>(const FileNode& n, int& v) { v = int(std::strtod(n.str().c_str(), nullptr)); }
inline void operator>>(const FileNode& n, double& v) { v = std::strtod(n.str().c_str(), nullptr); }
inline void operator>>(const FileNode& n, bool& v) { v = std::strtod(n.str().c_str(), nullptr) != 0.0; }
inline void operator>>(const FileNode& n, std::string& v) { v = n.str(); }
class FileStorage {
 public:
  enum { READ = 0 };
  bool open(const char* filename, int) {
    kv_.clear();
    std::ifstream f(filename);
    opened_ = bool(f);
    std::string line;
    while (std::getline(f, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      if (line.empty() || line[0] == '%') continue;
      const size_t colon = line.find(':');
      if (colon == std::string::npos) continue;
      std::string k = trim(line.substr(0, colon)), v = trim(line.substr(colon + 1));
      if (v.size() >= 2 && v.front() == '"' && v.back() == '"') v = v.substr(1, v.size() - 2);
      if (!k.empty()) kv_[k] = v;
    }
    return opened_;
  }
  bool isOpened() const { return opened_; }
  void release() { kv_.clear(); opened_ = false; }
  FileNode operator[](const char* key) const {
    auto it = kv_.find(key);
    return it == kv_.end() ? FileNode() : FileNode(&it->second);
  }
 private:
  static std::string trim(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
  std::map<std::string, std::string> kv_;
  bool opened_ = false;
};

void pyrDown(const Mat& src, Mat& dst, const Size& dstsize = Size());
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
struct KeyPointsFilter { static void retainBest(std::vector<KeyPoint>& keypoints, int npoints); };
void undistort(const Mat& src, Mat& dst, const Mat& cameraMatrix, const Mat& distCoeffs);
float fastAtan2(float y, float x);

}  // namespace cv
