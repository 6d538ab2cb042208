// frame.cc — pyramid, FAST, corner selection, Frame/Feature/Point (oracle; test infrastructure only).
// Follows frame.cc:34-56,93-131, extra/fast_detector.cc:58-175, feature.cc:28-36, point.cc:128-142.
// cv::pyrDown / cv::FAST / cv::KeyPointsFilter::retainBest are un-vendored OpenCV: restated from the
// published algorithms and pinned bit-exactly against cv2 4.13 in tests.
#include <algorithm>
#include <cassert>

#include "oracle.h"

namespace oracle {

static inline int Reflect101(int t, int n) {  // cv::BORDER_REFLECT_101
  if (t < 0) return -t;
  if (t >= n) return 2 * n - 2 - t;
  return t;
}

// cv::pyrDown for CV_8UC1 with dst = (cols/2, rows/2): separable [1 4 6 4 1], (sum + 128) >> 8.
void PyrDown(const Mat8& src, Mat8* dst) {
  const int dw = src.cols / 2, dh = src.rows / 2;
  *dst = Mat8(dw, dh);
  static const int k[5] = {1, 4, 6, 4, 1};
  std::vector<int> row(size_t(5) * dw);
  for (int y = 0; y < dh; y++) {
    for (int i = 0; i < 5; i++) {
      const uint8_t* s = src.ptr(Reflect101(2 * y + i - 2, src.rows));
      for (int x = 0; x < dw; x++) {
        int acc = 0;
        for (int j = 0; j < 5; j++) acc += k[j] * s[Reflect101(2 * x + j - 2, src.cols)];
        row[size_t(i) * dw + x] = acc;
      }
    }
    uint8_t* d = dst->data.data() + size_t(y) * dw;
    for (int x = 0; x < dw; x++) {
      int acc = 0;
      for (int i = 0; i < 5; i++) acc += k[i] * row[size_t(i) * dw + x];
      d[x] = uint8_t((acc + 128) >> 8);
    }
  }
}

// OpenCV ring order for FAST 9_16 (makeOffsets, patternSize 16).
static const int kRing[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                 {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// cornerScore<16>: max over the 16 arcs of 9 contiguous ring pixels of min(d) and of min(-d), minus 1.
static inline int CornerScore16(const uint8_t* p, const int* off, int threshold) {
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 25; k++) d[k] = v - p[off[k & 15]];
  int a0 = threshold;
  for (int k = 0; k < 16; k++) {
    int a = d[k];
    for (int j = 1; j < 9; j++) a = std::min(a, d[k + j]);
    a0 = std::max(a0, a);
  }
  int b0 = -a0;
  for (int k = 0; k < 16; k++) {
    int b = d[k];
    for (int j = 1; j < 9; j++) b = std::max(b, d[k + j]);
    b0 = std::min(b0, b);
  }
  return -b0 - 1;
}

void FastRoi(const uint8_t* roi, int stride, int cols, int rows, int threshold, std::vector<KeyPoint>* out) {
  out->clear();
  if (cols < 7 || rows < 7) return;
  threshold = std::min(std::max(threshold, 0), 255);
  int off[16];
  for (int k = 0; k < 16; k++) off[k] = kRing[k][0] + kRing[k][1] * stride;
  std::vector<int> score(size_t(cols) * rows, 0);
  for (int y = 3; y < rows - 3; y++) {
    const uint8_t* p = roi + size_t(y) * stride;
    for (int x = 3; x < cols - 3; x++) {
      const int s = CornerScore16(p + x, off, threshold);
      // a pixel is a corner iff some 9-arc is entirely brighter/darker than +-threshold,
      // i.e. iff the un-clamped arc score exceeds the threshold: score = S-1 >= threshold
      if (s >= threshold) score[size_t(y) * cols + x] = s;
      else score[size_t(y) * cols + x] = 0;
    }
  }
  // A corner needs S > threshold; CornerScore16 returns max(threshold, S) - 1, so S > threshold <=> s >= threshold
  // except S == threshold... handled: S == threshold gives s = threshold-1 < threshold.
  for (int y = 3; y < rows - 3; y++) {
    for (int x = 3; x < cols - 3; x++) {
      const int s = score[size_t(y) * cols + x];
      if (s == 0 && threshold > 0) continue;
      if (s < threshold) continue;
      const int* c = &score[size_t(y) * cols + x];
      if (s > c[-1] && s > c[1] && s > c[-cols - 1] && s > c[-cols] && s > c[-cols + 1] && s > c[cols - 1] &&
          s > c[cols] && s > c[cols + 1]) {
        KeyPoint kp;
        kp.x = float(x); kp.y = float(y); kp.response = float(s);
        out->push_back(kp);
      }
    }
  }
}

void RetainBest(std::vector<KeyPoint>* kps, int n) {  // KeyPointsFilter::retainBest (OpenCV 4.x)
  if (n >= 0 && kps->size() > size_t(n)) {
    if (n == 0) { kps->clear(); return; }
    std::nth_element(kps->begin(), kps->begin() + n - 1, kps->end(),
                     [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    const float amb = (*kps)[n - 1].response;
    auto new_end = std::partition(kps->begin() + n, kps->end(), [amb](const KeyPoint& k) { return k.response >= amb; });
    kps->resize(new_end - kps->begin());
  }
}

// fast_detector.cc:58-152
void SelectPixels(const sdvlb_params& P, const Mat8& src, int level, int nfeatures, std::vector<Corner>* corners,
                  std::vector<int>* scores) {
  const int cell = P.cell_size;
  const int margin = 1 + P.patch_size / 2;  // use_orb == 0 (fast_detector.cc:66)
  const int wcells = int(std::ceil(double(src.cols) / double(cell)));
  const int hcells = int(std::ceil(double(src.rows) / double(cell)));
  std::vector<std::vector<KeyPoint>> cell_fts(size_t(hcells) * wcells);
  std::vector<int> nleft(size_t(hcells) * wcells, 0), nselected(size_t(hcells) * wcells, 0);
  int nempty = 0;
  for (int i = 0; i < hcells; i++) {
    const int inity = std::max(margin, i * cell);
    const int maxy = std::min(src.rows - margin, i * cell + cell);
    if (maxy <= inity) continue;
    for (int j = 0; j < wcells; j++) {
      const int initx = std::max(margin, j * cell);
      const int maxx = std::min(src.cols - margin, j * cell + cell);
      if (maxx <= initx) continue;
      std::vector<KeyPoint>& fts = cell_fts[size_t(i) * wcells + j];
      FastRoi(src.ptr(inity) + initx, src.cols, maxx - initx, maxy - inity, P.fast_threshold, &fts);
      if (!fts.empty()) {
        for (auto& k : fts) { k.x += initx; k.y += inity; }
        nleft[size_t(i) * wcells + j] = int(fts.size());
      } else {
        nempty++;
      }
    }
  }
  const int ncells = hcells * wcells;
  int selected = 0;
  int cells_left = ncells - nempty;
  while ((nfeatures - selected) > 0 && cells_left > 0) {
    const int npercell = int(std::ceil(double(nfeatures - selected) / double(cells_left)));
    cells_left = 0;
    for (int c = 0; c < ncells; c++) {
      if (nleft[c] > 0) {
        if (nleft[c] > npercell) {
          nselected[c] += npercell; selected += npercell; nleft[c] -= npercell; cells_left++;
        } else {
          nselected[c] += nleft[c]; selected += nleft[c]; nleft[c] = 0;
        }
      }
    }
  }
  std::vector<KeyPoint> fts;
  for (int c = 0; c < ncells; c++) {
    RetainBest(&cell_fts[c], nselected[c]);
    for (auto& k : cell_fts[c]) fts.push_back(k);
  }
  if (int(fts.size()) > nfeatures) RetainBest(&fts, nfeatures);
  for (auto& k : fts) {
    corners->push_back(Corner{int(k.x), int(k.y), level});
    if (scores) scores->push_back(int(k.response));
  }
}

// fast_detector.cc:154-175
void DetectPyramid(const sdvlb_params& P, const std::vector<Mat8>& pyr, int nfeatures, std::vector<Corner>* corners,
                   std::vector<int>* scores) {
  assert(int(pyr.size()) >= P.max_fast_levels);
  const double scale = 1.2;
  double factor = 1.0, val = 0.0;
  for (int i = 0; i < P.max_fast_levels; i++) { val += factor; factor /= scale; }
  int levelfeatures = int(nfeatures / val);
  for (int i = 0; i < P.max_fast_levels; i++) {
    SelectPixels(P, pyr[i], i, levelfeatures, corners, scores);
    levelfeatures = int(levelfeatures / scale);
  }
}

// frame.cc:34-56,114-131
std::shared_ptr<Frame> MakeFrame(const sdvlb_params& P, const Camera* cam, const uint8_t* img, int w, int h, bool corners,
                                 int id) {
  auto f = std::make_shared<Frame>();
  f->id = id;
  f->cam = cam;
  f->pyramid.resize(P.pyramid_levels);
  f->pyramid[0] = Mat8(w, h);
  std::memcpy(f->pyramid[0].data.data(), img, size_t(w) * h);
  for (int i = 1; i < P.pyramid_levels; i++) PyrDown(f->pyramid[i - 1], &f->pyramid[i]);
  if (corners) DetectPyramid(P, f->pyramid, P.num_features, &f->corners, &f->corner_scores);
  return f;
}

bool Frame::Project(const V3& p3d, V2* p2d) const {  // frame.cc:93-102
  const V3 rel = pose * p3d;
  if (rel.z < 0.0) return false;
  cam->Project(rel, p2d);
  return true;
}

std::shared_ptr<Feature> MakeFeature(const std::shared_ptr<Frame>& f, const V2& p, int level) {  // feature.cc:28-36
  auto ft = std::make_shared<Feature>();
  ft->frame = f;
  ft->p2d = p;
  ft->v = f->cam->Unproject(p);
  ft->level = level;
  return ft;
}

V3 Point::GetPosition() const {  // point.cc:128-142
  if (fixed) return p3d;
  const SE3 se3 = feature->frame->GetWorldPose();
  return se3 * (feature->v * (1.0 / rho));
}

}  // namespace oracle
