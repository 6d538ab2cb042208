// map_candidates.cc — CPU restatement (TEST INFRASTRUCTURE ONLY) of the mapping thread's per-frame pass over its
// depth-filter candidates: the loop body of Map::UpdateCandidates (map.cc:397-498) with
// Point::Update / ComputeTau / PDFNormal / HasConverged / Unpromote (point.cc:63-100,109-116,162-216),
// GetDepthFromTriangulation and GetParallax (extra/utils.cc:193-213), Frame::IsPointVisible (frame.cc:104-112).
// PINNED against the reference's own Map::UpdateCandidates run from oracle/_ref (tests/test_oracle_vs_ref.py::
// test_update_candidates_vs_reference: same converged / deleted / kept sets, filter state to 1e-10); the per-corner
// part of Map::InitCandidates (SDVLB_SEEDS_INIT) shares every step but the list handling and is checked by the
// known-answer tests of tests/test_oracle_cpu.py (the filter converges to the true depth of a synthetic plane).
#include <algorithm>
#include <cmath>

#include "oracle.h"

namespace oracle {

// extra/utils.cc:193-205.  A = [R v_ref, v_cur]; depth = |(-(A'A)^-1 A' t)[0]|, Eigen's 2x2 inverse is the adjugate / det.
static bool GetDepthFromTriangulation(const SE3& pose, const V3& v_ref, const V3& v_cur, double* depth) {
  const M3 R = pose.Rotation();
  const V3 rv(R.m[0][0] * v_ref.x + R.m[0][1] * v_ref.y + R.m[0][2] * v_ref.z,
              R.m[1][0] * v_ref.x + R.m[1][1] * v_ref.y + R.m[1][2] * v_ref.z,
              R.m[2][0] * v_ref.x + R.m[2][1] * v_ref.y + R.m[2][2] * v_ref.z);
  const double a00 = rv.dot(rv), a01 = rv.dot(v_cur), a11 = v_cur.dot(v_cur);
  const double det = a00 * a11 - a01 * a01;
  if (det < 0.000001) return false;
  const double at0 = rv.dot(pose.t), at1 = v_cur.dot(pose.t);
  const double d0 = -((a11 / det) * at0 + (-a01 / det) * at1);
  *depth = std::fabs(d0);
  return true;
}

static double GetParallax(const V3& src1, const V3& src2, const V3& p3d) {  // extra/utils.cc:207-213
  V3 v1 = src1 - p3d, v2 = src2 - p3d;
  const double n1 = v1.norm(), n2 = v2.norm();
  v1 = V3(v1.x / n1, v1.y / n1, v1.z / n1);
  v2 = V3(v2.x / n2, v2.y / n2, v2.z / n2);
  return v1.dot(v2);
}

static double ComputeTau(const SE3& pose, const V3& v, double depth, double px_error_angle) {  // point.cc:186-198
  const double PI = 3.14159265;
  const V3 t = pose.t;
  const V3 a = v * depth - t;
  const double t_norm = t.norm(), a_norm = a.norm();
  const double alpha = std::acos(v.dot(t) / t_norm);
  const double beta = std::acos(a.dot(t * -1.0) / (t_norm * a_norm));
  const double beta_plus = beta + px_error_angle;
  const double gamma_plus = PI - alpha - beta_plus;
  const double depth_plus = t_norm * std::sin(beta_plus) / std::sin(gamma_plus);
  return depth_plus - depth;
}

static double PDFNormal(double mean, double sd, double x) {  // point.cc:200-216
  const double PI = 3.14159265;
  if (sd <= 0) return 0.0;
  double exponent = x - mean;
  exponent *= -exponent;
  exponent /= 2 * sd * sd;
  double result = std::exp(exponent);
  result /= sd * std::sqrt(2.0 * PI);
  return result;
}

// One candidate of Map::UpdateCandidates.  `cur` has corners and its pose set; `ref` = init feature's frame.
void UpdateCandidate(const sdvlb_params& P, const std::shared_ptr<Frame>& cur, const std::shared_ptr<Frame>& ref,
                     const sdvlb_seed_params& sp, sdvlb_seed* S) {
  const Camera* cam = cur->cam;
  auto ft = std::make_shared<Feature>();
  ft->frame = ref;
  ft->p2d.x = S->ref_px[0]; ft->p2d.y = S->ref_px[1];
  ft->v = V3(S->ref_v[0], S->ref_v[1], S->ref_v[2]);
  ft->level = S->ref_level;
  if (UseOrb()) {   // the descriptor the candidate's init feature got at its keyframe (frame.cc:148-161, map.cc:319-323);
                    // a feature outside ORBDetector::IsInsideLimits never gets one: the constructor's 32 zero bytes
    static const OrbDetector det;
    ft->descriptor.assign(32, 0);
    const int lx = int(S->ref_px[0] / (1 << S->ref_level)), ly = int(S->ref_px[1] / (1 << S->ref_level));
    if (det.IsInsideLimits(ref->pyramid[size_t(S->ref_level)], lx, ly))
      det.GetDescriptor(ref->pyramid[size_t(S->ref_level)], lx, ly, ft->descriptor.data());
  }
  S->depth = 0.0;

  // map.cc:426-437: pos = point->GetPosition(); frame->IsPointVisible(pos)
  V3 pos = ref->GetWorldPose() * (ft->v * (1.0 / S->rho));
  const V3 rel = cur->pose * pos;
  const bool init = sp.mode == SDVLB_SEEDS_INIT;   // the per-corner part of Map::InitCandidates (map.cc:306-378)
  bool visible = init || !(rel.z < 0.0);
  if (visible && !init) {
    V2 ip;
    cam->Project(rel, &ip);
    visible = cam->IsInsideImage(int(ip.x), int(ip.y));
  }
  if (!visible) {
    S->status = S->last_kf_id < sp.min_kf_id ? SDVLB_SEED_DELETE_OLD : SDVLB_SEED_NOT_VISIBLE;
    return;
  }
  // map.cc:441-445
  const double distance = (cur->GetWorldPosition() - ref->GetWorldPosition()).norm();
  if (distance / sp.depth_mean < 0.01) { S->status = SDVLB_SEED_SHORT_BASELINE; return; }
  // map.cc:448-453
  Matcher matcher(P, P.patch_size);
  V2 imgpos;
  int level = 0;
  if (!matcher.SearchPoint(cur, ft, S->rho, std::sqrt(S->sigma2), false, &imgpos, &level)) {
    if (init) { S->status = SDVLB_SEED_NOT_FOUND; return; }   // map.cc:320-321
    S->n_failed++;   // Point::Unpromote (point.cc:109-116)
    S->b++;
    S->status = S->n_failed > P.max_failed ? SDVLB_SEED_DELETE_FAILED : SDVLB_SEED_NOT_FOUND;
    return;
  }
  S->px[0] = imgpos.x; S->px[1] = imgpos.y; S->level = level;
  // map.cc:456-461
  const SE3 pose = cur->pose * ref->pose.Inverse();
  const V3 v3d = cam->Unproject(imgpos);
  double depth = 0.0;
  if (!GetDepthFromTriangulation(pose, ft->v, v3d, &depth)) { S->status = SDVLB_SEED_NO_DEPTH; return; }
  S->depth = depth;
  // map.cc:464-478
  const V3 p3d = ref->GetWorldPose() * (ft->v * depth);
  const double cos_alpha = GetParallax(ref->GetWorldPosition(), cur->GetWorldPosition(), p3d);
  if (cos_alpha >= 0.999999) { S->status = SDVLB_SEED_NO_PARALLAX; return; }
  if (depth < sp.map_scale * sp.scale_min_dist || depth < sp.depth_mean * sp.scale_min_dist) {
    S->status = SDVLB_SEED_TOO_CLOSE;
    return;
  }
  S->status = SDVLB_SEED_UPDATED;
  if (init) return;   // map.cc:381-388: candidate->InitCandidate(feature, depth) is the caller's
  // Point::Update (point.cc:63-100)
  const double px_error_angle = std::atan(1.0 / (2.0 * cam->fx)) * 2.0;   // camera.h:104-107
  const SE3 pose_u = ref->pose * cur->pose.Inverse();
  const double tau = ComputeTau(pose_u, ft->v, depth, px_error_angle);
  const double tau_inverse = 0.5 * (1.0 / std::max(0.0000001, depth - tau) - 1.0 / (depth + tau));
  const double tau2 = tau_inverse * tau_inverse;
  const double x = 1. / depth;
  const double norm_scale = std::sqrt(S->sigma2 + tau2);
  if (std::isnan(norm_scale)) return;
  double a_ = S->a, b_ = S->b, rho_ = S->rho, sigma2_ = S->sigma2;
  const double s2 = 1. / (1. / sigma2_ + 1. / tau2);
  const double m = s2 * (rho_ / sigma2_ + x / tau2);
  double C1 = a_ / (a_ + b_) * PDFNormal(rho_, norm_scale, x);
  double C2 = b_ / (a_ + b_) * 1. / S->z_range;
  const double normalization_constant = C1 + C2;
  C1 /= normalization_constant;
  C2 /= normalization_constant;
  const double f = C1 * (a_ + 1.) / (a_ + b_ + 1.) + C2 * a_ / (a_ + b_ + 1.);
  const double e = C1 * (a_ + 1.) * (a_ + 2.) / ((a_ + b_ + 1.) * (a_ + b_ + 2.)) +
                   C2 * a_ * (a_ + 1.0f) / ((a_ + b_ + 1.0f) * (a_ + b_ + 2.0f));
  const double rho_new = C1 * m + C2 * rho_;
  sigma2_ = C1 * (s2 + m * m) + C2 * (sigma2_ + rho_ * rho_) - rho_new * rho_new;
  rho_ = rho_new;
  a_ = (e - f) / (f - e / f);
  b_ = a_ * (1.0f - f) / f;
  S->rho = rho_; S->sigma2 = sigma2_; S->a = a_; S->b = b_;
  pos = ref->GetWorldPose() * (ft->v * (1.0 / rho_));
  S->cos_alpha = GetParallax(ref->GetWorldPosition(), cur->GetWorldPosition(), pos);
  S->last_distance = (cur->GetWorldPosition() - pos).norm();
  S->n_failed = 0;
  // Point::HasConverged (point.cc:162-176)
  const double std_d = std::sqrt(sigma2_) / (rho_ * rho_);
  const double l = 4 * std_d * S->cos_alpha / S->last_distance;
  if (l < 0.1) {
    S->p3d[0] = pos.x; S->p3d[1] = pos.y; S->p3d[2] = pos.z;
    S->status = SDVLB_SEED_CONVERGED;
  }
}

}  // namespace oracle

using namespace oracle;

extern "C" {

// seeds[i].ref_frame carries the index of the reference image in ref_imgs (as orc_search_points does for candidates).
int orc_update_candidates(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* cur_img, int w, int h,
                          const double T_cur[7], const uint8_t* const* ref_imgs, int n_refs, sdvlb_seed* seeds, int n,
                          const sdvlb_seed_params* sp) {
  Camera cam;
  cam.width = cam_->width; cam.height = cam_->height; cam.fx = cam_->fx; cam.fy = cam_->fy; cam.u0 = cam_->u0; cam.v0 = cam_->v0;
  auto cur = MakeFrame(*P, &cam, cur_img, w, h, true, 1000);
  cur->pose = SE3::FromArray(T_cur);
  std::vector<std::shared_ptr<Frame>> refs(n_refs);
  for (int i = 0; i < n_refs; i++) refs[i] = MakeFrame(*P, &cam, ref_imgs[i], w, h, false, i);
  for (int i = 0; i < n; i++) {
    const int ri = int(reinterpret_cast<intptr_t>(seeds[i].ref_frame));
    if (ri < 0 || ri >= n_refs) return -2;
    refs[ri]->pose = SE3::FromArray(seeds[i].ref_T);
    UpdateCandidate(*P, cur, refs[ri], *sp, &seeds[i]);
  }
  return 0;
}

}  // extern "C"
