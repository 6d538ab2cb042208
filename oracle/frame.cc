// frame.cc — pyramid, FAST, corner selection, Frame/Feature/Point (oracle; test infrastructure only).
// Follows frame.cc:34-56,93-131, extra/fast_detector.cc:58-175, feature.cc:28-36, point.cc:128-142.
// cv::pyrDown / cv::FAST / cv::KeyPointsFilter::retainBest are un-vendored OpenCV: restated from the
// published algorithms and pinned bit-exactly against cv2 4.13 in tests.
#include <algorithm>
#include <cassert>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "oracle.h"

namespace oracle {

static inline int Reflect101(int t, int n) {  // cv::BORDER_REFLECT_101
  if (t < 0) return -t;
  if (t >= n) return 2 * n - 2 - t;
  return t;
}

// cv::pyrDown for CV_8UC1 with dst = (cols/2, rows/2): separable [1 4 6 4 1], (sum + 128) >> 8.
// Integer arithmetic, so the column pass can run first (contiguous, auto-vectorised) and the strided row pass second:
// the result is bit-identical to OpenCV's rows-then-columns order.  Written to run at a speed comparable to OpenCV's
// SIMD pyrDown, because this oracle is also the timed CPU baseline.
void PyrDown(const Mat8& src, Mat8* dst) {
  const int sw = src.cols, sh = src.rows;
  const int dw = sw / 2, dh = sh / 2;
  *dst = Mat8(dw, dh);
  std::vector<uint16_t> col(size_t(sw) + 8);
  uint16_t* v = col.data() + 4;   // v[-2 .. sw+1] addressable
  for (int y = 0; y < dh; y++) {
    const uint8_t* s0 = src.ptr(Reflect101(2 * y - 2, sh));
    const uint8_t* s1 = src.ptr(Reflect101(2 * y - 1, sh));
    const uint8_t* s2 = src.ptr(2 * y);
    const uint8_t* s3 = src.ptr(Reflect101(2 * y + 1, sh));
    const uint8_t* s4 = src.ptr(Reflect101(2 * y + 2, sh));
    for (int x = 0; x < sw; x++) v[x] = uint16_t(s0[x] + 4 * s1[x] + 6 * s2[x] + 4 * s3[x] + s4[x]);
    v[-1] = v[1]; v[-2] = v[2];                          // BORDER_REFLECT_101 in x
    v[sw] = v[sw - 2]; v[sw + 1] = v[sw - 3];
    uint8_t* d = dst->data.data() + size_t(y) * dw;
    int x = 0;
#if defined(__SSE2__)
    {   // four outputs per step: the five taps as three pairwise multiply-adds on 16-bit lanes (sums <= 16 * 4080)
      const __m128i k14 = _mm_set1_epi32(0x00040001), k64 = _mm_set1_epi32(0x00040006), k10 = _mm_set1_epi32(0x00000001);
      const __m128i half = _mm_set1_epi32(128);
      for (; x + 4 <= dw && 2 * x + 9 <= sw + 2; x += 4) {   // reads v[2x-2 .. 2x+9] <= v[sw+1]
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(v + 2 * x - 2));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(v + 2 * x));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(v + 2 * x + 2));
        __m128i sum = _mm_add_epi32(_mm_madd_epi16(a, k14), _mm_madd_epi16(b, k64));
        sum = _mm_srli_epi32(_mm_add_epi32(_mm_add_epi32(sum, _mm_madd_epi16(c, k10)), half), 8);
        const __m128i p16 = _mm_packs_epi32(sum, sum);
        const int packed = _mm_cvtsi128_si32(_mm_packus_epi16(p16, p16));
        std::memcpy(d + x, &packed, 4);
      }
    }
#endif
    for (; x < dw; x++) {
      const uint16_t* c = v + 2 * x;
      d[x] = uint8_t((c[-2] + 4 * c[-1] + 6 * c[0] + 4 * c[1] + c[2] + 128) >> 8);
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
  for (int k = 0; k < 16; k += 2) {   // same pairing of arcs as OpenCV's scalar cornerScore<16>
    int a = std::min(d[k + 1], d[k + 2]);
    a = std::min(a, d[k + 3]);
    if (a <= a0) continue;
    a = std::min(a, d[k + 4]); a = std::min(a, d[k + 5]); a = std::min(a, d[k + 6]);
    a = std::min(a, d[k + 7]); a = std::min(a, d[k + 8]);
    a0 = std::max(a0, std::min(a, d[k]));
    a0 = std::max(a0, std::min(a, d[k + 9]));
  }
  int b0 = -a0;
  for (int k = 0; k < 16; k += 2) {
    int b = std::max(d[k + 1], d[k + 2]);
    b = std::max(b, d[k + 3]); b = std::max(b, d[k + 4]); b = std::max(b, d[k + 5]);
    if (b >= b0) continue;
    b = std::max(b, d[k + 6]); b = std::max(b, d[k + 7]); b = std::max(b, d[k + 8]);
    b0 = std::min(b0, std::max(b, d[k]));
    b0 = std::min(b0, std::max(b, d[k + 9]));
  }
  return -b0 - 1;
}

// cv::FAST(roi, threshold, nonmax = true), TYPE_9_16: the structure of OpenCV's FAST_t<16> (threshold table, early
// rejection on opposite ring pixels, run-length test, score only for corners, 3-row NMS), so that its cost on the CPU
// is representative of the library the reference calls.
void FastRoi(const uint8_t* roi, int stride, int cols, int rows, int threshold, std::vector<KeyPoint>* out) {
  out->clear();
  if (cols < 7 || rows < 7) return;
  threshold = std::min(std::max(threshold, 0), 255);
  int off[25];
  for (int k = 0; k < 16; k++) off[k] = kRing[k][0] + kRing[k][1] * stride;
  for (int k = 16; k < 25; k++) off[k] = off[k - 16];
  uint8_t tab[512];   // tab[d + 255]: 1 = ring pixel darker than centre - t, 2 = brighter than centre + t
  for (int i = -255; i <= 255; i++) tab[i + 255] = uint8_t(i < -threshold ? 1 : i > threshold ? 2 : 0);
  std::vector<uint8_t> score(size_t(cols) * rows, 0);   // corner scores fit a byte (<= 254); 0 = no corner
  for (int y = 3; y < rows - 3; y++) {
    const uint8_t* p = roi + size_t(y) * stride;
    uint8_t* srow = &score[size_t(y) * cols];
#if defined(__SSE2__)
    if (cols >= 22) {
      // 16 pixels at a time, as OpenCV's FAST_t<16> does with its universal intrinsics: compass-point rejection, then
      // the longest darker / brighter run over the 25 ring positions with saturating byte counters; the score is
      // computed only for the pixels that pass.  Two blocks cover the 26 tested columns of a 32-pixel cell (the
      // second one overlaps the first; a score is a pure function of the pixel, so rewriting it is harmless).
      const __m128i delta = _mm_set1_epi8(char(-128)), t8 = _mm_set1_epi8(char(threshold)), k8 = _mm_set1_epi8(8);
      for (int x = 3;; x += 16) {
        if (x > cols - 19) x = cols - 19;
        const uint8_t* c = p + x;
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(c));
        const __m128i vlo = _mm_xor_si128(_mm_subs_epu8(v, t8), delta), vhi = _mm_xor_si128(_mm_adds_epu8(v, t8), delta);
        auto ring = [&](int k) { return _mm_xor_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(c + off[k])), delta); };
        const __m128i x0 = ring(0), x1 = ring(4), x2 = ring(8), x3 = ring(12);
        __m128i m0 = _mm_and_si128(_mm_cmpgt_epi8(vlo, x0), _mm_cmpgt_epi8(vlo, x1));
        __m128i m1 = _mm_and_si128(_mm_cmpgt_epi8(x0, vhi), _mm_cmpgt_epi8(x1, vhi));
        m0 = _mm_or_si128(m0, _mm_and_si128(_mm_cmpgt_epi8(vlo, x1), _mm_cmpgt_epi8(vlo, x2)));
        m1 = _mm_or_si128(m1, _mm_and_si128(_mm_cmpgt_epi8(x1, vhi), _mm_cmpgt_epi8(x2, vhi)));
        m0 = _mm_or_si128(m0, _mm_and_si128(_mm_cmpgt_epi8(vlo, x2), _mm_cmpgt_epi8(vlo, x3)));
        m1 = _mm_or_si128(m1, _mm_and_si128(_mm_cmpgt_epi8(x2, vhi), _mm_cmpgt_epi8(x3, vhi)));
        m0 = _mm_or_si128(m0, _mm_and_si128(_mm_cmpgt_epi8(vlo, x3), _mm_cmpgt_epi8(vlo, x0)));
        m1 = _mm_or_si128(m1, _mm_and_si128(_mm_cmpgt_epi8(x3, vhi), _mm_cmpgt_epi8(x0, vhi)));
        if (_mm_movemask_epi8(_mm_or_si128(m0, m1)) != 0) {
          __m128i c0 = _mm_setzero_si128(), c1 = c0, max0 = c0, max1 = c0;
          for (int k = 0; k < 25; k++) {
            const __m128i r = ring(k);
            const __m128i d = _mm_cmpgt_epi8(vlo, r), b = _mm_cmpgt_epi8(r, vhi);
            c0 = _mm_and_si128(_mm_sub_epi8(c0, d), d);
            c1 = _mm_and_si128(_mm_sub_epi8(c1, b), b);
            max0 = _mm_max_epu8(max0, c0);
            max1 = _mm_max_epu8(max1, c1);
          }
          int m = _mm_movemask_epi8(_mm_cmpgt_epi8(_mm_max_epu8(max0, max1), k8));
          for (; m; m &= m - 1) {
            const int j = __builtin_ctz(unsigned(m));
            srow[x + j] = uint8_t(CornerScore16(c + j, off, threshold));
          }
        }
        if (x >= cols - 19) break;
      }
      continue;
    }
#endif
    for (int x = 3; x < cols - 3; x++) {
      const uint8_t* c = p + x;
      const uint8_t* t = tab + 255 - int(c[0]);   // t[ring] classifies ring - centre
      int d = t[c[off[0]]] | t[c[off[8]]];
      if (d == 0) continue;
      d &= t[c[off[2]]] | t[c[off[10]]];
      d &= t[c[off[4]]] | t[c[off[12]]];
      d &= t[c[off[6]]] | t[c[off[14]]];
      if (d == 0) continue;
      d &= t[c[off[1]]] | t[c[off[9]]];
      d &= t[c[off[3]]] | t[c[off[11]]];
      d &= t[c[off[5]]] | t[c[off[13]]];
      d &= t[c[off[7]]] | t[c[off[15]]];
      bool corner = false;
      if (d & 1) {
        const int vt = int(c[0]) - threshold;
        int count = 0;
        for (int k = 0; k < 25; k++) {
          if (int(c[off[k]]) < vt) { if (++count > 8) { corner = true; break; } }
          else count = 0;
        }
      }
      if (!corner && (d & 2)) {
        const int vt = int(c[0]) + threshold;
        int count = 0;
        for (int k = 0; k < 25; k++) {
          if (int(c[off[k]]) > vt) { if (++count > 8) { corner = true; break; } }
          else count = 0;
        }
      }
      if (corner) srow[x] = uint8_t(CornerScore16(c, off, threshold));
    }
  }
  // keep a corner iff its score is strictly greater than its 8 neighbours' (non-corners count as 0); raster order
  for (int y = 3; y < rows - 3; y++) {
    const uint8_t* c0 = &score[size_t(y) * cols];
    for (int x = 3; x < cols - 3; x++) {
      const int s = c0[x];
      if (s == 0 || s < threshold) continue;
      const uint8_t* c = c0 + x;
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
  const int margin = BorderMargin(P);  // fast_detector.cc:63-66
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

// ---------------------------------------------------------------- FilterCorners
// extra/utils.cc:61-97.  The sums are integer-valued floats below 2^24 (exact in any order); the divisions are
// float / double -> float; the last line is float arithmetic with the float sqrt overload (g++ >= 6, <math.h> wrapper)
// and a final double product.  Whether a*b - c*d is contracted into an FMA depends on the reference's compiler flags
// (-O3 -march=native, CMakeLists.txt:20); the oracle pins the uncontracted IEEE reading.
double FindShiTomasiScoreAtPoint(const Mat8& img, int px, int py) {
  float dXX = 0.0, dYY = 0.0, dXY = 0.0;
  const int halfbox_size = 4;
  const int box_size = 2 * halfbox_size;
  const int box_area = box_size * box_size;
  const int x_min = px - halfbox_size, x_max = px + halfbox_size;
  const int y_min = py - halfbox_size, y_max = py + halfbox_size;
  if (x_min < 1 || x_max >= img.cols - 1 || y_min < 1 || y_max >= img.rows - 1) return 0.0;
  for (int y = y_min; y < y_max; y++) {
    const uint8_t* row = img.ptr(y);
    const uint8_t* top = img.ptr(y - 1);
    const uint8_t* bot = img.ptr(y + 1);
    for (int x = 0; x < box_size; x++) {
      const float dx = float(int(row[x_min + 1 + x]) - int(row[x_min - 1 + x]));
      const float dy = float(int(bot[x_min + x]) - int(top[x_min + x]));
      dXX += dx * dx;
      dYY += dy * dy;
      dXY += dx * dy;
    }
  }
  dXX = float(dXX / (2.0 * box_area));
  dYY = float(dYY / (2.0 * box_area));
  dXY = float(dXY / (2.0 * box_area));
  // every product rounded to float on its own (no FMA contraction)
  volatile float tr = dXX + dYY;
  volatile float p1 = dXX * dYY, p2 = dXY * dXY, t2 = tr * tr;
  volatile float det = p1 - p2;
  volatile float f4 = 4 * det;
  volatile float disc = t2 - f4;
  volatile float root = sqrtf(disc);
  const float diff = tr - root;
  return 0.5 * diff;
}

// fast_detector.cc:177-218 + LockCell (:46-49) + InitGrid (:38-41): cgrid_ holds (index, int score)
void FilterCorners(const sdvlb_params& P, const std::vector<Mat8>& pyr, const std::vector<Corner>& corners,
                   const std::vector<V2>& locked, int min_feature_score, std::vector<int>* indices) {
  const int cell = P.cell_size;
  const int gw = int(std::ceil(double(pyr[0].cols) / cell)), gh = int(std::ceil(double(pyr[0].rows) / cell));
  std::vector<std::pair<int, int>> cgrid(size_t(gw) * gh, std::make_pair(0, min_feature_score));
  std::vector<bool> mask(size_t(gw) * gh, false);
  for (const V2& p : locked) mask.at(size_t(int(p.y / cell) * gw + int(p.x / cell))) = true;
  const int margin = BorderMargin(P);   // fast_detector.cc:183-186
  int index = 0;
  for (auto it = corners.begin(); it != corners.end(); it++, index++) {
    const int px = it->x, py = it->y, level = it->level, scale = 1 << level;
    if (px < margin || py < margin || px >= pyr[level].cols - margin || py >= pyr[level].rows - margin) continue;
    const int pos = int((py * scale) / cell) * gw + int((px * scale) / cell);
    if (mask[pos]) continue;
    const double score = FindShiTomasiScoreAtPoint(pyr[level], px, py);
    if (score > cgrid.at(pos).second) cgrid.at(pos) = std::make_pair(index, int(score));
  }
  for (auto& c : cgrid)
    if (c.second > min_feature_score) indices->push_back(c.first);
}

}  // namespace oracle

namespace oracle {

// cv::undistort as Camera::UndistortImage calls it (camera.cc:100-105; main.cc:133 is the call site).  OpenCV builds a
// fixed-point map per destination pixel (initUndistortRectifyMap, m1type CV_16SC2: integer source pixel + 5-bit
// fractions, INTER_TAB_SIZE 32) and resamples with cv::remap's fixed-point bilinear kernel: weights
// (32-fx)(32-fy)*32 ... with 15 fractional bits (INTER_REMAP_COEF_SCALE 32768; exact products, so the table's
// normalisation step never fires), result (sum + 2^14) >> 15, taps outside the image read the border value 0.
void UndistortImage(const Camera& cam, const double d[5], const Mat8& in, Mat8* out) {
  const int w = in.cols, h = in.rows;
  *out = Mat8(w, h);
  const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
  const double ir0 = 1.0 / cam.fx, ir2 = -cam.u0 / cam.fx, ir4 = 1.0 / cam.fy, ir5 = -cam.v0 / cam.fy;   // (K)^-1
  for (int i = 0; i < h; i++) {
    const double y = i * ir4 + ir5;
    uint8_t* dst = out->data.data() + size_t(i) * w;
    for (int j = 0; j < w; j++) {
      const double x = j * ir0 + ir2;
      const double x2 = x * x, y2 = y * y;
      const double r2 = x2 + y2, _2xy = 2 * x * y;
      const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((0 * r2 + 0) * r2 + 0) * r2);
      const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2));
      const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy);
      const double u = cam.fx * xd + cam.u0, v = cam.fy * yd + cam.v0;
      const int iu = int(std::nearbyint(u * 32)), iv = int(std::nearbyint(v * 32));   // saturate_cast<int> = cvRound
      const int sx = int16_t(iu >> 5), sy = int16_t(iv >> 5);                         // stored as short
      const int fx = iu & 31, fy = iv & 31;
      const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
      auto tap = [&](int yy, int xx) -> int {
        return (xx >= 0 && xx < w && yy >= 0 && yy < h) ? int(in.ptr(yy)[xx]) : 0;
      };
      const int acc = tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11;
      dst[j] = uint8_t((acc + (1 << 14)) >> 15);
    }
  }
}

}  // namespace oracle
