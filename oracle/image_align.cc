// image_align.cc — restatement of ImageAlign (image_align.cc:35-267) (oracle; test infrastructure only).
// Keeps the reference's numeric types (fp32 pixels/weights/residual/chi2, fp64 geometry/J/H/b) and its
// sticky-state quirks (SURVEY.md §8a notes): visible_fts_/patch_cache_ persist across levels, stop_ and
// chi2_ are never reset, fx scales both Jacobian rows.
#include <cassert>

#include "oracle.h"

namespace oracle {

int ImageAlign::ComputePose(const std::shared_ptr<Frame>& frame1, const std::shared_ptr<Frame>& frame2, bool fast) {
  frame1_ = frame1;
  frame2_ = frame2;
  assert(int(frame1_->pyramid.size()) >= P_.max_align_level);  // image_align.cc:52 (off by one, kept)
  const int size = int(frame1_->features.size());
  if (size == 0) return 0;  // image_align.cc:55-58

  const int path_area = P_.align_patch_size * P_.align_patch_size;
  patch_cache_.assign(size_t(size) * path_area, 0.0f);  // cv::Mat(size, area, CV_32F) is uninitialised; 0 here
  jacobian_cache_.assign(size_t(6) * size * path_area, 0.0);
  visible_fts_.assign(size, false);
  trace.clear();

  SE3 current_se3 = frame2_->pose * frame1_->pose.Inverse();  // :66
  for (int level = P_.max_align_level; level >= P_.min_align_level; level--) {
    std::fill(jacobian_cache_.begin(), jacobian_cache_.end(), 0.0);  // :69
    Optimize(&current_se3, level);
    if (fast && error_ > 0.01) {  // :73-76
      error_ = 1e10;
      break;
    }
  }
  frame2_->pose = current_se3 * frame1_->pose;  // :79
  frame1_ = nullptr;
  frame2_ = nullptr;
  return int(n_meas_ / path_area);  // :83
}

void ImageAlign::Optimize(SE3* se3, int level) {  // :86-125
  Vec6 x;
  SE3 se3_bk = *se3;
  for (int i = 0; i < P_.max_img_align_its; i++) {
    for (int r = 0; r < 6; r++) {
      Jres_[r] = 0;
      for (int c = 0; c < 6; c++) H_[r][c] = 0;
    }
    n_meas_ = 0;
    sdvlb_gn_iter rec;
    std::memset(&rec, 0, sizeof(rec));
    rec.level = level;
    rec.iter = i;
    se3->ToArray(rec.T_in);

    const double new_chi2 = ComputeResiduals(*se3, level, true, i == 0);
    if (n_meas_ == 0) stop_ = true;

    LdltSolve6(H_, Jres_, x);  // :102
    bool nan = false;
    if (std::isnan(x[0])) { stop_ = true; nan = true; }

    rec.n_meas = int(n_meas_);
    for (int r = 0; r < 6; r++) {
      rec.b[r] = Jres_[r];
      rec.x[r] = x[r];
      for (int c = 0; c < 6; c++) rec.H[r * 6 + c] = H_[r][c];
    }
    rec.chi2 = new_chi2;
    rec.flags = (nan ? 2 : 0) | (n_meas_ == 0 ? 4 : 0);

    if ((i > 0 && new_chi2 > chi2_) || stop_) {  // :109-112
      *se3 = se3_bk;
      rec.flags |= 1;
      trace.push_back(rec);
      break;
    }
    trace.push_back(rec);

    se3_bk = *se3;
    Vec6 mx;
    for (int r = 0; r < 6; r++) mx[r] = -x[r];
    *se3 = (*se3) * SE3::Exp(mx);  // :116
    chi2_ = new_chi2;
    error_ = AbsMax6(x);
    if (error_ <= 1e-10) break;
  }
}

double ImageAlign::ComputeResiduals(const SE3& se3, int level, bool linearize, bool patches) {  // :127-206
  const int ps = P_.align_patch_size;
  const int half_patch = ps / 2;
  const int path_area = ps * ps;
  const Mat8& last_img = frame2_->pyramid.at(level);
  if (patches) PrecomputePatches(level);

  const int stride = last_img.cols;
  const int border = half_patch + 1;
  const float scale = 1.0f / (1 << level);
  const V3 first_pos = frame1_->GetWorldPosition();
  float chi2 = 0.0;
  size_t counter = 0;
  for (auto it = frame1_->features.begin(); it != frame1_->features.end(); ++it, ++counter) {
    const std::shared_ptr<Feature>& feature = *it;
    if (!visible_fts_[counter]) continue;
    assert(feature->point);
    if (feature->point->del) continue;

    const double depth = (feature->point->GetPosition() - first_pos).norm();
    const V3 xyz_ref = feature->v * depth;
    const V3 xyz_cur = se3 * xyz_ref;
    V2 proj;
    frame2_->cam->Project(xyz_cur, &proj);
    const double upx = proj.x * scale, upy = proj.y * scale;  // Vector2d * float
    const float u_cur = float(upx);
    const float v_cur = float(upy);
    const int u_last_i = int(floorf(u_cur));
    const int v_last_i = int(floorf(v_cur));
    if (u_last_i < 0 || v_last_i < 0 || u_last_i - border < 0 || v_last_i - border < 0 ||
        u_last_i + border >= last_img.cols || v_last_i + border >= last_img.rows)
      continue;

    const float subpix_u_cur = u_cur - u_last_i;
    const float subpix_v_cur = v_cur - v_last_i;
    const float w_last_tl = (1.0 - subpix_u_cur) * (1.0 - subpix_v_cur);
    const float w_last_tr = subpix_u_cur * (1.0 - subpix_v_cur);
    const float w_last_bl = (1.0 - subpix_u_cur) * subpix_v_cur;
    const float w_last_br = subpix_u_cur * subpix_v_cur;
    const float* patch_cache_ptr = patch_cache_.data() + path_area * counter;
    size_t pixel_counter = 0;
    for (int y = 0; y < ps; y++) {
      const uint8_t* p = last_img.data.data() + (v_last_i + y - half_patch) * stride + (u_last_i - half_patch);
      for (int x = 0; x < ps; x++, pixel_counter++, p++, patch_cache_ptr++) {
        const float intensity_cur = w_last_tl * p[0] + w_last_tr * p[1] + w_last_bl * p[stride] + w_last_br * p[stride + 1];
        const float res = intensity_cur - (*patch_cache_ptr);
        const float weight = 1.0;
        chi2 += res * res * weight;
        n_meas_++;
        if (linearize) {
          const double* J = &jacobian_cache_[6 * (counter * path_area + pixel_counter)];
          for (int r = 0; r < 6; r++) {
            for (int c = 0; c < 6; c++) H_[r][c] += J[r] * J[c] * weight;
            Jres_[r] -= J[r] * res * weight;
          }
        }
      }
    }
  }
  return chi2 / n_meas_;  // float / size_t -> float (NaN when n_meas_ == 0), widened to double
}

void ImageAlign::PrecomputePatches(int level) {  // :208-267
  const int ps = P_.align_patch_size;
  const int half_patch = ps / 2;
  const int path_area = ps * ps;
  const int border = half_patch + 1;
  const Mat8& first_img = frame1_->pyramid.at(level);
  const int stride = first_img.cols;
  const float scale = 1.0f / (1 << level);
  const V3 first_pos = frame1_->GetWorldPosition();
  const double focal_length = frame1_->cam->fx;
  size_t counter = 0;
  double frame_jac[2][6];
  for (auto it = frame1_->features.begin(); it != frame1_->features.end(); ++it, ++counter) {
    const std::shared_ptr<Feature>& feature = *it;
    const float u_ref = float(feature->p2d.x * scale);
    const float v_ref = float(feature->p2d.y * scale);
    const int u_first_i = int(floorf(u_ref));
    const int v_first_i = int(floorf(v_ref));
    if (!feature->point || feature->point->del || u_first_i - border < 0 || v_first_i - border < 0 ||
        u_first_i + border >= first_img.cols || v_first_i + border >= first_img.rows)
      continue;
    visible_fts_[counter] = true;

    const double depth = (feature->point->GetPosition() - first_pos).norm();
    const V3 xyz_ref = feature->v * depth;
    Jacobian3DToPlane(xyz_ref, frame_jac);

    const float subpix_u_ref = u_ref - u_first_i;
    const float subpix_v_ref = v_ref - v_first_i;
    const float w_tl = (1.0 - subpix_u_ref) * (1.0 - subpix_v_ref);
    const float w_tr = subpix_u_ref * (1.0 - subpix_v_ref);
    const float w_bl = (1.0 - subpix_u_ref) * subpix_v_ref;
    const float w_br = subpix_u_ref * subpix_v_ref;
    size_t pixel_counter = 0;
    float* cache_ptr = patch_cache_.data() + path_area * counter;
    for (int y = 0; y < ps; y++) {
      const uint8_t* p = first_img.data.data() + (v_first_i + y - half_patch) * stride + (u_first_i - half_patch);
      for (int x = 0; x < ps; x++, p++, cache_ptr++, pixel_counter++) {
        *cache_ptr = w_tl * p[0] + w_tr * p[1] + w_bl * p[stride] + w_br * p[stride + 1];
        const float dx = 0.5f * ((w_tl * p[1] + w_tr * p[2] + w_bl * p[stride + 1] + w_br * p[stride + 2]) -
                                 (w_tl * p[-1] + w_tr * p[0] + w_bl * p[stride - 1] + w_br * p[stride]));
        const float dy = 0.5f * ((w_tl * p[stride] + w_tr * p[1 + stride] + w_bl * p[stride * 2] + w_br * p[stride * 2 + 1]) -
                                 (w_tl * p[-stride] + w_tr * p[1 - stride] + w_bl * p[0] + w_br * p[1]));
        double* J = &jacobian_cache_[6 * (counter * path_area + pixel_counter)];
        const double s = focal_length / (1 << level);
        for (int r = 0; r < 6; r++) J[r] = (dx * frame_jac[0][r] + dy * frame_jac[1][r]) * s;
      }
    }
  }
}

}  // namespace oracle
