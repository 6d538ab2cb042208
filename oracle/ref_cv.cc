// ref_cv.cc — the OpenCV entry points declared by oracle/ref_shim/opencv2/core/core.hpp (TEST INFRASTRUCTURE ONLY).
//
// The reference's own sources are compiled unmodified from /root/reference into oracle/_ref/libsdvlref.so; OpenCV is
// not installed, so the four image operations they call are supplied here by the oracle's restatements, each of which
// is pinned bit-exactly against OpenCV 4.13 (tests/test_oracle_cpu.py, tests/golden/):
//   cv::pyrDown (frame.cc:119), cv::FAST (extra/fast_detector.cc:95), cv::KeyPointsFilter::retainBest
//   (extra/fast_detector.cc:140,148), cv::undistort (camera.cc:102), cv::fastAtan2 (extra/orb_detector.cc:436).
#include <opencv2/core/core.hpp>

#include "oracle.h"

namespace cv {

static oracle::Mat8 ToMat8(const Mat& m) {
  oracle::Mat8 o(m.cols, m.rows);
  for (int y = 0; y < m.rows; y++) std::memcpy(o.data.data() + size_t(y) * m.cols, m.ptr<uchar>(y), size_t(m.cols));
  return o;
}
static void FromMat8(const oracle::Mat8& s, Mat* d) {
  d->create(s.rows, s.cols, CV_8UC1);
  std::memcpy(d->data, s.data.data(), s.data.size());
}

void pyrDown(const Mat& src, Mat& dst, const Size& dstsize) {
  oracle::Mat8 out;
  oracle::PyrDown(ToMat8(src), &out);   // dst = (cols/2, rows/2), the only size the reference asks for
  if (dstsize.width && (dstsize.width != out.cols || dstsize.height != out.rows)) std::abort();
  FromMat8(out, &dst);
}

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  if (!nonmaxSuppression) std::abort();
  std::vector<oracle::KeyPoint> k;
  oracle::FastRoi(image.data, int(size_t(image.step)), image.cols, image.rows, threshold, &k);
  keypoints.clear();
  for (const auto& p : k) keypoints.push_back(KeyPoint(p.x, p.y, 7.f, -1.f, p.response));
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
  std::vector<oracle::KeyPoint> k(keypoints.size());
  for (size_t i = 0; i < k.size(); i++) { k[i].x = keypoints[i].pt.x; k[i].y = keypoints[i].pt.y; k[i].response = keypoints[i].response; }
  oracle::RetainBest(&k, npoints);
  keypoints.clear();
  for (const auto& p : k) keypoints.push_back(KeyPoint(p.x, p.y, 7.f, -1.f, p.response));
}

void undistort(const Mat& src, Mat& dst, const Mat& K, const Mat& D) {
  oracle::Camera cam;
  cam.width = src.cols; cam.height = src.rows;
  cam.fx = K.at<double>(0, 0); cam.fy = K.at<double>(1, 1); cam.u0 = K.at<double>(0, 2); cam.v0 = K.at<double>(1, 2);
  double d[5];
  for (int i = 0; i < 5; i++) d[i] = D.at<double>(i, 0);
  oracle::Mat8 out;
  oracle::UndistortImage(cam, d, ToMat8(src), &out);
  FromMat8(out, &dst);
}

float fastAtan2(float y, float x) {   // extra/orb_detector.cc:436; the oracle's restatement, pinned against cv2.fastAtan2
  return oracle::FastAtan2(y, x);
}

}  // namespace cv
