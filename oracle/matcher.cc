// matcher.cc — restatement of Matcher (matcher.cc:45-476, non-ORB branches) (oracle; test infrastructure only).
#include <cassert>

#include "oracle.h"

namespace oracle {

static const double MAX_SSD_PER_PIXEL = 500;  // matcher.h:36

static inline V2 CamProject(const Camera* cam, const V3& p) { V2 r; cam->Project(p, &r); return r; }

bool Matcher::SearchPoint(const std::shared_ptr<Frame>& frame, const std::shared_ptr<Feature>& feature, double idepth,
                          double idepth_std, bool fixed, V2* px, int* flevel, SearchDebug* dbg) {  // :45-121
  double A[2][2];
  double range, zmin, zmax;
  int slevel, level = feature->level;
  V2 pxa, pxb;

  std::shared_ptr<Frame> ref_frame = feature->frame;
  assert(ref_frame);
  const SE3 pose = frame->pose * ref_frame->pose.Inverse();

  if (fixed) {
    zmin = 1.0 / (idepth + 2.0 * idepth_std);
    const V3 p3d_min = ref_frame->GetWorldPose() * (feature->v * zmin);
    if (!frame->Project(p3d_min, &pxa)) return false;
  } else {
    zmin = 1.0 / (idepth + 2.0 * idepth_std);
    zmax = 1.0 / (std::max(idepth - 2.0 * idepth_std, 0.00000001));
    const V3 p3d_min = ref_frame->GetWorldPose() * (feature->v * zmin);
    const V3 p3d_max = ref_frame->GetWorldPose() * (feature->v * zmax);
    if (!frame->Project(p3d_min, &pxa)) return false;
    if (!frame->Project(p3d_max, &pxb)) return false;
  }

  if (UseOrb()) {   // matcher.cc:79-80: assert(feature->HasDescriptor())
    assert(feature->descriptor.size() == 32);
    desc_ = feature->descriptor.data();
  }

  // feature->GetLevelPosition().cast<int>() (feature.h:93-95): truncation of p2d / 2^level
  const int lx = int(feature->p2d.x / (1 << level));
  const int ly = int(feature->p2d.y / (1 << level));
  if (!ref_frame->cam->IsInsideImage(lx, ly, patch_size_ / 2 + 2, level)) return false;  // :83

  const Mat8& img = ref_frame->pyramid[level];
  WarpMatrixAffine(frame->cam, feature->p2d, feature->v, 1.0 / idepth, pose, level, A);
  slevel = GetSearchLevel(A);
  CreatePatch(A, img, feature->p2d, level, slevel, border_patch_, patch_);
  if (dbg) dbg->slevel = slevel;

  range = P_.search_size;
  for (int i = 1; i <= slevel; i++) range *= 1.2;

  std::vector<int> indices;
  if (fixed) {
    const V2 cpos = *px;
    GetCornersInRange(frame, cpos, level, range, &indices);
  } else {
    GetCornersInRange(frame, pxa, pxb, level, range, &indices);
  }
  if (dbg) dbg->n_in_range = int(indices.size());

  int best = -1;
  const bool ok = SearchFeatures(frame, indices, patch_, px, &best);
  if (dbg) dbg->zmssd = best;
  if (!ok) return false;

  V2 px_scaled;
  px_scaled.x = px->x / (1 << slevel);
  px_scaled.y = px->y / (1 << slevel);
  if (AlignPatch(frame->pyramid[slevel], border_patch_, patch_, &px_scaled)) {
    px->x = px_scaled.x * (1 << slevel);
    px->y = px_scaled.y * (1 << slevel);
    *flevel = slevel;
    return true;
  }
  return false;
}

void Matcher::GetCornersInRange(const std::shared_ptr<Frame>& frame, const V2& pxa, const V2& pxb, int level,
                                double range, std::vector<int>* indices) {  // :123-192
  const double range2 = range * range;
  const int margin = BorderMargin(P_);   // matcher.cc:131-134,199-202
  const std::vector<Corner>& cc = frame->corners;

  double ex = pxa.x - pxb.x, ey = pxa.y - pxb.y;
  const double en2 = ex * ex + ey * ey;   // Vector2d::normalize(): divide by norm when squaredNorm > 0
  if (en2 > 0) { const double en = std::sqrt(en2); ex /= en; ey /= en; }
  const double nx = ey, ny = -ex;
  const double normdist = pxa.x * nx + pxa.y * ny;
  const double xdiff = pxb.x - pxa.x;
  const double ydiff = pxb.y - pxa.y;
  const double vline = xdiff * xdiff + ydiff * ydiff;

  int index = 0;
  for (auto it = cc.begin(); it != cc.end(); ++it, ++index) {
    const int clevel = it->level;
    if (std::abs(clevel - level) > 1) continue;
    if (it->x - margin < 0 || it->y - margin < 0) continue;
    if (it->y + margin >= frame->pyramid[clevel].rows || it->x + margin >= frame->pyramid[clevel].cols) continue;
    const double posx = it->x * (1 << clevel), posy = it->y * (1 << clevel);
    const double dist = normdist - (posx * nx + posy * ny);
    if (std::fabs(dist) > range) continue;
    const double u = ((posx - pxa.x) * xdiff + (posy - pxa.y) * ydiff) / vline;
    if (u > 1) {
      const double dx = posx - pxb.x, dy = posy - pxb.y;
      if ((dx * dx + dy * dy) > range2) continue;
    }
    if (u < 0) {
      const double dx = posx - pxa.x, dy = posy - pxa.y;
      if ((dx * dx + dy * dy) > range2) continue;
    }
    indices->push_back(index);
  }
}

void Matcher::GetCornersInRange(const std::shared_ptr<Frame>& frame, const V2& cpos, int level, double range,
                                std::vector<int>* indices) {  // :194-230
  const double range2 = range * range;
  const int margin = BorderMargin(P_);   // matcher.cc:131-134,199-202
  const std::vector<Corner>& cc = frame->corners;
  int index = 0;
  for (auto it = cc.begin(); it != cc.end(); ++it, ++index) {
    const int clevel = it->level;
    if (std::abs(clevel - level) > 1) continue;
    if (it->x - margin < 0 || it->y - margin < 0) continue;
    if (it->y + margin >= frame->pyramid[clevel].rows || it->x + margin >= frame->pyramid[clevel].cols) continue;
    const double posx = it->x * (1 << clevel), posy = it->y * (1 << clevel);
    const double dx = cpos.x - posx, dy = cpos.y - posy;
    if (dx * dx + dy * dy > range2) continue;
    indices->push_back(index);
  }
}

static void GetZMSSDScore(const uint8_t* patch, int parea, int* sumA, int* sumAA) {  // :447-457
  uint32_t a = 0, aa = 0;
  for (int r = 0; r < parea; r++) {
    const uint8_t n = patch[r];
    a += n;
    aa += n * n;
  }
  *sumA = int(a);
  *sumAA = int(aa);
}

static double CompareZMSSDScore(const uint8_t* ref_patch, const uint8_t* patch, int ps, int sumA, int sumAA, int cols) {  // :459-476
  uint32_t sumB = 0, sumBB = 0, sumAB = 0;
  for (int y = 0, r = 0; y < ps; y++) {
    const uint8_t* p = patch + y * cols;
    for (int x = 0; x < ps; x++, r++) {
      const uint8_t pixel = p[x];
      sumB += pixel;
      sumBB += pixel * pixel;
      sumAB += pixel * ref_patch[r];
    }
  }
  const int B = int(sumB), BB = int(sumBB), AB = int(sumAB);
  return sumAA - 2 * AB + BB - (sumA * sumA - 2 * sumA * B + B * B) / (ps * ps);
}

bool Matcher::SearchFeatures(const std::shared_ptr<Frame>& frame, const std::vector<int>& indices, uint8_t* patch, V2* px,
                             int* best) {  // :232-291
  int sumA = 0, sumAA = 0;
  V2 best_px;
  const std::vector<Corner>& corners = frame->corners;
  const bool orb = UseOrb();
  const int threshold = orb ? 100 /* MIN_ORB_THRESHOLD, matcher.h:37 */ : int(patch_size_ * patch_size_ * MAX_SSD_PER_PIXEL);
  int best_score = threshold + 1;
  if (!orb) GetZMSSDScore(patch, patch_size_ * patch_size_, &sumA, &sumAA);
  for (auto it = indices.begin(); it != indices.end(); ++it) {
    const Corner& corner = corners[*it];
    const int level = corner.level;
    const Mat8& cimg = frame->pyramid[level];
    int score;
    if (orb) {   // matcher.cc:264-273: lazily computed corner descriptor vs the feature's
      score = OrbDetector::Distance(desc_, CornerDescriptor(frame.get(), *it).data());
    } else {
      const uint8_t* cur_patch = cimg.data.data() + (corner.y - patch_size_ / 2) * cimg.cols + (corner.x - patch_size_ / 2);
      score = int(CompareZMSSDScore(patch, cur_patch, patch_size_, sumA, sumAA, cimg.cols));
    }
    if (score < best_score) {
      best_score = score;
      best_px.x = corner.x * (1 << level);
      best_px.y = corner.y * (1 << level);
    }
  }
  if (best) *best = indices.empty() ? -1 : best_score;
  if (best_score >= threshold) return false;
  *px = best_px;
  return true;
}

void Matcher::WarpMatrixAffine(const Camera* cam, const V2& px, const V3& v, double depth, const SE3& pose, int level,
                               double A[2][2]) {  // :293-312
  const int half_size = 5;
  const V3 p3d = v * depth;
  V2 pdu; pdu.x = px.x + double(half_size) * (1 << level); pdu.y = px.y + 0.0 * (1 << level);
  V2 pdv; pdv.x = px.x + 0.0 * (1 << level); pdv.y = px.y + double(half_size) * (1 << level);
  V3 xyz_du = cam->Unproject(pdu);
  V3 xyz_dv = cam->Unproject(pdv);
  xyz_du = xyz_du * (p3d.z / xyz_du.z);
  xyz_dv = xyz_dv * (p3d.z / xyz_dv.z);
  const V2 px_cur = CamProject(cam, pose * p3d);
  const V2 px_du = CamProject(cam, pose * xyz_du);
  const V2 px_dv = CamProject(cam, pose * xyz_dv);
  A[0][0] = (px_du.x - px_cur.x) / half_size;
  A[1][0] = (px_du.y - px_cur.y) / half_size;
  A[0][1] = (px_dv.x - px_cur.x) / half_size;
  A[1][1] = (px_dv.y - px_cur.y) / half_size;
}

int Matcher::GetSearchLevel(const double A[2][2]) {  // :314-323
  int search_level = 0;
  double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  const int max = P_.max_fast_levels - 1;
  while (det > 3.0 && search_level < max) {
    search_level += 1;
    det *= 0.25;
  }
  return search_level;
}

void Matcher::CreatePatch(const double A[2][2], const Mat8& img, const V2& px, int level, int search_level,
                          uint8_t* border_patch, uint8_t* patch) {  // :325-357
  const int bpatch_size = patch_size_ + 2;
  const int half_size = bpatch_size / 2;
  // Eigen Matrix2d::inverse(): adjugate scaled by 1/det
  const double det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  const double invdet = 1.0 / det;
  double Ai[2][2];
  Ai[0][0] = A[1][1] * invdet;
  Ai[0][1] = -A[0][1] * invdet;
  Ai[1][0] = -A[1][0] * invdet;
  Ai[1][1] = A[0][0] * invdet;
  if (std::isnan(Ai[0][0])) return;  // :331-334 (patches keep their previous contents)

  uint8_t* patch_ptr = patch;
  uint8_t* bpatch_ptr = border_patch;
  const double pyrx = px.x / (1 << level), pyry = px.y / (1 << level);
  for (int y = 0; y < bpatch_size; y++) {
    for (int x = 0; x < bpatch_size; x++, bpatch_ptr++) {
      double ppx = x - half_size, ppy = y - half_size;
      ppx *= (1 << search_level);
      ppy *= (1 << search_level);
      const double p0 = Ai[0][0] * ppx + Ai[0][1] * ppy + pyrx;
      const double p1 = Ai[1][0] * ppx + Ai[1][1] * ppy + pyry;
      if (p0 < 0 || p1 < 0 || p0 >= img.cols - 1 || p1 >= img.rows - 1)
        *bpatch_ptr = 0;
      else
        *bpatch_ptr = uint8_t(Interpolate8U(img, float(p0), float(p1)));
      if (y >= 1 && y < bpatch_size - 1 && x >= 1 && x < bpatch_size - 1) {
        *patch_ptr = *bpatch_ptr;
        patch_ptr++;
      }
    }
  }
}

bool Matcher::AlignPatch(const Mat8& img, uint8_t* border_patch, uint8_t* patch, V2* px) {  // :359-445
  const int ps = patch_size_;
  const int half_size = ps / 2;
  const int patch_area = ps * ps;
  bool converged = false;
  std::vector<float> patch_dx(patch_area), patch_dy(patch_area);

  float H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  const int ref_step = ps + 2;
  float* it_dx = patch_dx.data();
  float* it_dy = patch_dy.data();
  for (int y = 0; y < ps; y++) {
    const uint8_t* it = border_patch + (y + 1) * ref_step + 1;
    for (int x = 0; x < ps; x++, it++, it_dx++, it_dy++) {
      float J[3];
      J[0] = float(0.5 * (it[1] - it[-1]));
      J[1] = float(0.5 * (it[ref_step] - it[-ref_step]));
      J[2] = 1;
      *it_dx = J[0];
      *it_dy = J[1];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) H[r][c] += J[r] * J[c];
    }
  }
  // Eigen Matrix3f::inverse(): cofactors / determinant, fp32
  float Hinv[3][3];
  {
    // cofactor_3x3<i,j>(m) = m(i1,j1)*m(i2,j2) - m(i1,j2)*m(i2,j1), i1=(i+1)%3, i2=(i+2)%3 (Eigen/src/LU/InverseImpl.h)
    auto cof = [&H](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return H[i1][j1] * H[i2][j2] - H[i1][j2] * H[i2][j1];
    };
    const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const float det = c0 * H[0][0] + c1 * H[1][0] + c2 * H[2][0];
    const float invdet = 1.0f / det;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Hinv[i][j] = cof(j, i) * invdet;
  }
  float mean_diff = 0;
  float u = float(px->x);
  float v = float(px->y);
  const float min_update_squared = float(0.03 * 0.03);
  const int cur_step = img.cols;
  for (int iter = 0; iter < P_.max_align_its; iter++) {
    const int u_r = int(std::floor(u));
    const int v_r = int(std::floor(v));
    if (u_r < half_size || v_r < half_size || u_r >= img.cols - half_size || v_r >= img.rows - half_size) break;
    if (std::isnan(u) || std::isnan(v)) return false;
    const float subpix_x = u - u_r;
    const float subpix_y = v - v_r;
    const float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
    const float wTR = subpix_x * (1.0 - subpix_y);
    const float wBL = (1.0 - subpix_x) * subpix_y;
    const float wBR = subpix_x * subpix_y;
    const uint8_t* it_ref = patch;
    const float* it_ref_dx = patch_dx.data();
    const float* it_ref_dy = patch_dy.data();
    float Jres[3] = {0, 0, 0};
    for (int y = 0; y < ps; y++) {
      const uint8_t* it = img.data.data() + (v_r + y - half_size) * cur_step + u_r - half_size;
      for (int x = 0; x < ps; x++, it++, it_ref++, it_ref_dx++, it_ref_dy++) {
        const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        const float res = search_pixel - *it_ref + mean_diff;
        Jres[0] -= res * (*it_ref_dx);
        Jres[1] -= res * (*it_ref_dy);
        Jres[2] -= res;
      }
    }
    float update[3];
    for (int r = 0; r < 3; r++) update[r] = Hinv[r][0] * Jres[0] + Hinv[r][1] * Jres[1] + Hinv[r][2] * Jres[2];
    u += update[0];
    v += update[1];
    mean_diff += update[2];
    if (update[0] * update[0] + update[1] * update[1] < min_update_squared) {
      converged = true;
      break;
    }
  }
  px->x = u;
  px->y = v;
  return converged;
}

}  // namespace oracle
