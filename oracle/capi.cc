// capi.cc — flat C entry points of the oracle for ctypes (test infrastructure only).
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "oracle.h"

using namespace oracle;

static Camera CamFrom(const sdvlb_camera* c) {
  Camera k;
  k.width = c->width; k.height = c->height; k.fx = c->fx; k.fy = c->fy; k.u0 = c->u0; k.v0 = c->v0;
  return k;
}

extern "C" {

void orc_params_default(sdvlb_params* p) {  // config.cc:55-85
  p->pyramid_levels = 5; p->cell_size = 32; p->max_matches = 150; p->max_align_level = 4; p->min_align_level = 2;
  p->max_img_align_its = 30; p->align_patch_size = 4; p->patch_size = 8; p->max_align_its = 10; p->search_size = 6;
  p->max_fast_levels = 3; p->fast_threshold = 10; p->num_features = 1000; p->max_failed = 15;
  p->max_optim_pose_its = 10; p->max_ransac_points = 5; p->max_ransac_its = 100; p->min_matches = 20;
  p->inlier_error_threshold = 2.0;
}

// levels 0..L-1 concatenated (level 0 is a copy of img). Returns total bytes.
int64_t orc_pyramid(const uint8_t* img, int w, int h, int levels, uint8_t* out) {
  Mat8 cur(w, h);
  std::memcpy(cur.data.data(), img, size_t(w) * h);
  int64_t off = 0;
  for (int l = 0; l < levels; l++) {
    if (out) std::memcpy(out + off, cur.data.data(), cur.data.size());
    off += int64_t(cur.data.size());
    if (l + 1 < levels) { Mat8 nxt; PyrDown(cur, &nxt); cur = nxt; }
  }
  return off;
}

int orc_fast_roi(const uint8_t* roi, int stride, int cols, int rows, int thr, int32_t* xy, int32_t* score, int cap) {
  std::vector<KeyPoint> k;
  FastRoi(roi, stride, cols, rows, thr, &k);
  const int n = std::min<int>(cap, int(k.size()));
  for (int i = 0; i < n; i++) { xy[2 * i] = int(k[i].x); xy[2 * i + 1] = int(k[i].y); score[i] = int(k[i].response); }
  return int(k.size());
}

// in/out: xyr = n x 3 floats (x, y, response). Returns the new count; survivors are left in the std:: order.
int orc_retain_best(float* xyr, int n, int keep) {
  std::vector<KeyPoint> k(n);
  for (int i = 0; i < n; i++) { k[i].x = xyr[3 * i]; k[i].y = xyr[3 * i + 1]; k[i].response = xyr[3 * i + 2]; }
  RetainBest(&k, keep);
  for (size_t i = 0; i < k.size(); i++) { xyr[3 * i] = k[i].x; xyr[3 * i + 1] = k[i].y; xyr[3 * i + 2] = k[i].response; }
  return int(k.size());
}

// Camera::UndistortImage
int orc_undistort(const sdvlb_camera* cam_, const double d[5], const uint8_t* img, int w, int h, uint8_t* out) {
  const Camera cam = CamFrom(cam_);
  Mat8 in(w, h), o;
  std::memcpy(in.data.data(), img, size_t(w) * h);
  UndistortImage(cam, d, in, &o);
  std::memcpy(out, o.data.data(), size_t(w) * h);
  return 0;
}

// Frame(img, corners=true) then GetCorners(): returns the number of corners.
int orc_detect(const sdvlb_params* P, const uint8_t* img, int w, int h, int nfeatures, int32_t* xyl, int32_t* score,
               int cap) {
  sdvlb_params p = *P;
  p.num_features = nfeatures;
  Camera cam{double(w), double(h), 1, 1, 0, 0};
  auto f = MakeFrame(p, &cam, img, w, h, true, 0);
  const int n = std::min<int>(cap, int(f->corners.size()));
  for (int i = 0; i < n; i++) {
    xyl[3 * i] = f->corners[i].x; xyl[3 * i + 1] = f->corners[i].y; xyl[3 * i + 2] = f->corners[i].level;
    if (score) score[i] = f->corner_scores[i];
  }
  return int(f->corners.size());
}

// ImageAlign::ComputePose on two images. pos3 = n x 3 world positions of the features' points.
int orc_image_align(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* ref_img, const uint8_t* cur_img,
                    int w, int h, const sdvlb_align_feat* feats, const double* pos3, int n, const double T_ref[7],
                    double T_cur[7], int fast, int* n_tracked, double* error, sdvlb_gn_iter* trace, int trace_cap,
                    int* trace_n) {
  Camera cam = CamFrom(cam_);
  auto f1 = MakeFrame(*P, &cam, ref_img, w, h, false, 0);
  auto f2 = MakeFrame(*P, &cam, cur_img, w, h, false, 1);
  f1->pose = SE3::FromArray(T_ref);
  f2->pose = SE3::FromArray(T_cur);
  for (int i = 0; i < n; i++) {
    auto ft = std::make_shared<Feature>();
    ft->frame = f1;
    ft->p2d.x = feats[i].px[0]; ft->p2d.y = feats[i].px[1];
    ft->v = V3(feats[i].v[0], feats[i].v[1], feats[i].v[2]);
    if (feats[i].valid) {
      auto pt = std::make_shared<Point>();
      pt->fixed = true;
      pt->p3d = V3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
      pt->feature = ft;
      ft->point = pt;
    }
    f1->features.push_back(ft);
  }
  ImageAlign ia(*P);
  const int r = ia.ComputePose(f1, f2, fast != 0);
  f2->pose.ToArray(T_cur);
  if (n_tracked) *n_tracked = r;
  if (error) *error = ia.GetError();
  const int tn = int(ia.trace.size());
  if (trace) for (int i = 0; i < std::min(tn, trace_cap); i++) trace[i] = ia.trace[i];
  if (trace_n) *trace_n = tn;
  f1->features.clear();
  return 0;
}

// Matcher::SearchPoint (+ FeatureAlign::ProjectPoint when SDVLB_CAND_PROJECT) for n candidates.
// cands[i].ref_frame is reinterpreted as an index into ref_imgs. Corners of the current frame come from
// DetectPyramid(P->num_features).
int orc_search_points(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* cur_img, int w, int h,
                      const double T_cur[7], const uint8_t* const* ref_imgs, int n_refs, const sdvlb_candidate* cands,
                      int n, sdvlb_match* out) {
  Camera cam = CamFrom(cam_);
  auto cur = MakeFrame(*P, &cam, cur_img, w, h, true, 1000);
  cur->pose = SE3::FromArray(T_cur);
  std::vector<std::shared_ptr<Frame>> refs(n_refs);
  for (int i = 0; i < n_refs; i++) refs[i] = MakeFrame(*P, &cam, ref_imgs[i], w, h, false, i);
  Matcher m(*P, P->patch_size);
  for (int i = 0; i < n; i++) {
    const sdvlb_candidate& c = cands[i];
    const int ri = int(reinterpret_cast<intptr_t>(c.ref_frame));
    if (ri < 0 || ri >= n_refs) return -2;
    auto rf = refs[ri];
    rf->pose = SE3::FromArray(c.ref_T);
    auto ft = std::make_shared<Feature>();
    ft->frame = rf;
    ft->p2d.x = c.ref_px[0]; ft->p2d.y = c.ref_px[1];
    ft->v = V3(c.ref_v[0], c.ref_v[1], c.ref_v[2]);
    ft->level = c.ref_level;
    if (UseOrb()) {   // the descriptor the feature got where it was created (frame.cc:148-161, map.cc:319-323)
      static const OrbDetector det;
      const int lx = int(c.ref_px[0] / (1 << c.ref_level)), ly = int(c.ref_px[1] / (1 << c.ref_level));
      if (!det.IsInsideLimits(rf->pyramid[size_t(c.ref_level)], lx, ly)) return -3;
      ft->descriptor.resize(32);
      det.GetDescriptor(rf->pyramid[size_t(c.ref_level)], lx, ly, ft->descriptor.data());
    }
    sdvlb_match& o = out[i];
    std::memset(&o, 0, sizeof(o));
    o.zmssd = -1;
    V2 px; px.x = c.px[0]; px.y = c.px[1];
    if (c.flags & SDVLB_CAND_PROJECT) {  // feature_align.cc:323-339
      V2 p;
      const V3 pos(c.pos[0], c.pos[1], c.pos[2]);
      if (!cur->Project(pos, &p) || !cam.IsInsideImage(int(p.x), int(p.y), P->patch_size)) {
        o.status = SDVLB_MATCH_UNSEEN;
        continue;
      }
      px = p;
    }
    o.proj[0] = px.x; o.proj[1] = px.y;
    int level = 0;
    SearchDebug dbg;
    const bool found = m.SearchPoint(cur, ft, c.idepth, c.idepth_std, (c.flags & SDVLB_CAND_FIXED) != 0, &px, &level, &dbg);
    o.status = found ? SDVLB_MATCH_FOUND : SDVLB_MATCH_NOT_FOUND;
    o.zmssd = dbg.zmssd;
    o.n_in_range = dbg.n_in_range;
    if (found) { o.px[0] = px.x; o.px[1] = px.y; o.level = level; }
  }
  return 0;
}

// SDVL::Relocalize's body for one keyframe (sdvl.cc:209-237): ComputePose(kf, cur, fast = true) from the keyframe's
// pose, the GetError() gate, FeatureAlign::Reproject(cur, kf, kf, reloc = true).  feats / pos3 / levels: the keyframe's
// features with their fixed points.  out[0] = matches (-1 when the error gate rejected the keyframe), out[1] =
// attempts, out[2] = features the current frame gained, out[3] = sum of the points' Score() afterwards.
int orc_relocalize(const sdvlb_params* P, const sdvlb_camera* cam_, const uint8_t* kf_img, const uint8_t* cur_img, int w,
                   int h, const sdvlb_align_feat* feats, const double* pos3, const int32_t* levels, int n,
                   const double T_kf[7], double T_out[7], double* error, int32_t out[4]) {
  Camera cam = CamFrom(cam_);
  auto kf = MakeFrame(*P, &cam, kf_img, w, h, false, 0);
  auto cur = MakeFrame(*P, &cam, cur_img, w, h, true, 1);
  kf->pose = SE3::FromArray(T_kf);
  cur->pose = kf->pose;
  const V3 C = kf->GetWorldPosition();
  std::vector<std::shared_ptr<Point>> pts;
  for (int i = 0; i < n; i++) {
    auto ft = std::make_shared<Feature>();
    ft->frame = kf;
    ft->p2d.x = feats[i].px[0]; ft->p2d.y = feats[i].px[1];
    ft->v = V3(feats[i].v[0], feats[i].v[1], feats[i].v[2]);
    ft->level = levels[i];
    auto pt = std::make_shared<Point>();
    pt->id = i;
    pt->fixed = true;
    pt->p3d = V3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
    pt->rho = 1.0 / (pt->p3d - C).norm();
    pt->sigma2 = (0.05 * pt->rho) * (0.05 * pt->rho);
    pt->feature = ft;
    ft->point = pt;
    kf->features.push_back(ft);
    pts.push_back(pt);
  }
  ImageAlign ia(*P);
  ia.ComputePose(kf, cur, true);
  *error = ia.GetError();
  cur->pose.ToArray(T_out);
  out[0] = -1; out[1] = 0; out[2] = 0; out[3] = 0;
  if (!(ia.GetError() >= 0.001)) {
    GlibcRand rng;
    std::vector<std::shared_ptr<Point>> trash;
    FeatureAlign fa(*P, &cam, P->max_matches, &rng, &trash);
    fa.Reproject(cur, kf, true);
    out[0] = fa.GetMatches();
    out[1] = fa.GetAttempts();
  }
  out[2] = int(cur->features.size());
  for (auto& p : pts) { out[3] += p->n_successful; p->feature = nullptr; }
  kf->features.clear();
  cur->features.clear();
  return 0;
}

// ---- primitives exposed for known-answer tests -------------------------------------------------
void orc_se3_exp(const double u[6], double T[7]) { SE3::Exp(u).ToArray(T); }
void orc_se3_log(const double T[7], double u[6]) { SE3::Log(SE3::FromArray(T), u); }
void orc_se3_mul(const double A[7], const double B[7], double C[7]) { (SE3::FromArray(A) * SE3::FromArray(B)).ToArray(C); }
void orc_se3_inv(const double A[7], double C[7]) { SE3::FromArray(A).Inverse().ToArray(C); }
void orc_se3_apply(const double A[7], const double p[3], double q[3]) {
  const V3 r = SE3::FromArray(A) * V3(p[0], p[1], p[2]);
  q[0] = r.x; q[1] = r.y; q[2] = r.z;
}
void orc_ldlt6(const double H[36], const double b[6], double x[6]) {
  Mat6 M;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) M[i][j] = H[i * 6 + j];
  LdltSolve6(M, b, x);
}
void orc_rand_seq(unsigned seed, int n, int32_t* out) {
  GlibcRand r;
  r.Seed(seed);
  for (int i = 0; i < n; i++) out[i] = r.Next();
}
void orc_rand_state(unsigned seed, sdvlb_rand* out) {   // the stream's state right after srand(seed)
  GlibcRand r;
  r.Seed(seed);
  std::memcpy(out->r, r.r, sizeof(r.r));
  out->n = r.n;
}
void orc_shuffle(int n, int32_t* v, unsigned seed) {
  GlibcRand r;
  r.Seed(seed);
  std::vector<int> a(v, v + n);
  RandomShuffle(&a, &r);
  std::copy(a.begin(), a.end(), v);
}
// AlignPatch alone: template = 10x10 border patch (u8), search image, start px (in level coords).
int orc_align_patch(const sdvlb_params* P, const uint8_t* img, int w, int h, const uint8_t* border_patch, double px[2]) {
  Matcher m(*P, P->patch_size);
  Mat8 im(w, h);
  std::memcpy(im.data.data(), img, size_t(w) * h);
  uint8_t bp[100], patch[64];
  std::memcpy(bp, border_patch, 100);
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) patch[y * 8 + x] = bp[(y + 1) * 10 + x + 1];
  V2 p; p.x = px[0]; p.y = px[1];
  const bool ok = m.AlignPatch(im, bp, patch, &p);
  px[0] = p.x; px[1] = p.y;
  return ok ? 1 : 0;
}

// Frame::FilterCorners on a frame built from img (corners detected with nfeatures); locked: n_locked x (x, y).
int orc_filter_corners(const sdvlb_params* P, const uint8_t* img, int w, int h, int nfeatures, const double* locked,
                       int n_locked, int min_feature_score, int32_t* indices, int cap) {
  sdvlb_params p = *P;
  p.num_features = nfeatures;
  Camera cam{double(w), double(h), 1, 1, 0, 0};
  auto f = MakeFrame(p, &cam, img, w, h, true, 0);
  std::vector<V2> lk(n_locked);
  for (int i = 0; i < n_locked; i++) { lk[i].x = locked[2 * i]; lk[i].y = locked[2 * i + 1]; }
  std::vector<int> out;
  FilterCorners(p, f->pyramid, f->corners, lk, min_feature_score, &out);
  for (int i = 0; i < int(out.size()) && i < cap; i++) indices[i] = out[i];
  return int(out.size());
}
double orc_shi_tomasi(const uint8_t* img, int w, int h, int px, int py) {
  Mat8 m(w, h);
  std::memcpy(m.data.data(), img, size_t(w) * h);
  return FindShiTomasiScoreAtPoint(m, px, py);
}

// ---- FeatureAlign pose refinement alone (parity of sdvlb_select_inliers / sdvlb_optimize_pose) -----------------
// obs: n x {v[3], pos[3], level, flags} (sdvlb_pose_obs).  mode 0: SelectInliers (rng = glibc state in/out),
// mode 1: OptimizePose(frame) including RemoveOutliers; flags out: 1 inlier, 2 outlier.
int orc_pose_refine(const sdvlb_params* P, const sdvlb_camera* cam_, sdvlb_pose_obs* obs, int n, double T[7],
                    sdvlb_rand* rng, int mode) {
  const Camera cam = CamFrom(cam_);
  GlibcRand r;
  if (rng) { std::memcpy(r.r, rng->r, sizeof(r.r)); r.n = rng->n; }
  std::vector<std::shared_ptr<Point>> trash;
  // the constructor shuffles cell_order_ with the stream: give it a scratch stream so `r` only sees RANSAC draws
  GlibcRand scratch;
  FeatureAlign fa(*P, &cam, P->max_matches, &scratch, &trash);
  auto frame = std::make_shared<Frame>();
  frame->cam = &cam;
  frame->pose = SE3::FromArray(T);
  std::vector<std::shared_ptr<Feature>> fs;
  for (int i = 0; i < n; i++) {
    auto ft = std::make_shared<Feature>();
    ft->frame = frame;
    ft->v = V3(obs[i].v[0], obs[i].v[1], obs[i].v[2]);
    ft->level = obs[i].level;
    auto pt = std::make_shared<Point>();
    pt->fixed = true;
    pt->p3d = V3(obs[i].pos[0], obs[i].pos[1], obs[i].pos[2]);
    ft->point = pt;
    fs.push_back(ft);
  }
  if (mode == 0) {
    fa.SetRng(&r);   // from here on the draws come from the caller's stream
    fa.SelectInliersHook(frame, fs);
    for (int i = 0; i < n; i++) obs[i].flags = SDVLB_OBS_OUTLIER;
    for (auto& f : fa.inliers())
      for (int i = 0; i < n; i++) if (fs[i] == f) obs[i].flags = SDVLB_OBS_INLIER;
    if (rng) { std::memcpy(rng->r, r.r, sizeof(r.r)); rng->n = r.n; }
    return int(fa.inliers().size());
  }
  for (int i = 0; i < n; i++) {
    if (obs[i].flags == SDVLB_OBS_INLIER) fa.inliers().push_back(fs[i]);
    else if (obs[i].flags == SDVLB_OBS_OUTLIER) fa.outliers().push_back(fs[i]);
  }
  fa.OptimizePose(frame);
  frame->pose.ToArray(T);
  for (int i = 0; i < n; i++) obs[i].flags = SDVLB_OBS_OUTLIER;
  for (auto& f : fa.inliers())
    for (int i = 0; i < n; i++) if (fs[i] == f) obs[i].flags = SDVLB_OBS_INLIER;
  return int(fa.inliers().size());
}

// ---- sequence driver -----------------------------------------------------------------------------
struct OrcTracker {
  Tracker* t;
  double sec_total = 0;
  long long gn_iters = 0;
};

void* orc_tracker_create(const sdvlb_params* P, const sdvlb_camera* cam, const double plane[4], int max_points,
                         int kf_every) {
  SeedPlane pl;
  pl.n[0] = plane[0]; pl.n[1] = plane[1]; pl.n[2] = plane[2]; pl.d = plane[3];
  OrcTracker* o = new OrcTracker;
  o->t = new Tracker(*P, CamFrom(cam), pl, max_points, kf_every);
  return o;
}
void orc_tracker_destroy(void* h) {
  OrcTracker* o = static_cast<OrcTracker*>(h);
  delete o->t;
  delete o;
}
// stats: n_tracked, matches, attempts, inliers, outliers, n_feats, gn_iters, keyframe
int orc_tracker_step(void* h, const uint8_t* img, int w, int h_, const double gt_pose[7], double est_pose[7],
                     int32_t stats[8]) {
  OrcTracker* o = static_cast<OrcTracker*>(h);
  SE3 est;
  TrackStats st;
  const auto t0 = std::chrono::steady_clock::now();
  o->t->HandleFrame(img, w, h_, SE3::FromArray(gt_pose), &est, &st);
  const auto t1 = std::chrono::steady_clock::now();
  o->sec_total += std::chrono::duration<double>(t1 - t0).count();
  o->gn_iters += st.gn_iters;
  est.ToArray(est_pose);
  if (stats) {
    stats[0] = st.n_tracked; stats[1] = st.matches; stats[2] = st.attempts; stats[3] = st.inliers;
    stats[4] = st.outliers; stats[5] = st.n_feats; stats[6] = st.gn_iters; stats[7] = st.keyframe;
  }
  return 0;
}
// Runs n frames (contiguous w*h images); returns wall seconds spent inside HandleFrame.
double orc_tracker_run(void* h, const uint8_t* imgs, int n, int w, int h_, const double* gt_poses, double* est_poses,
                       int32_t* stats) {
  OrcTracker* o = static_cast<OrcTracker*>(h);
  const double before = o->sec_total;
  for (int i = 0; i < n; i++)
    orc_tracker_step(h, imgs + size_t(i) * w * h_, w, h_, gt_poses + 7 * i, est_poses + 7 * i, stats ? stats + 8 * i : nullptr);
  return o->sec_total - before;
}
// Features of the last frame that carry a point (for inspection): px(2), level, point id.
int orc_tracker_last_features(void* h, double* px, int32_t* level, int32_t* pid, int cap) {
  OrcTracker* o = static_cast<OrcTracker*>(h);
  auto f = o->t->last_frame();
  int n = 0;
  if (!f) return 0;
  for (auto& ft : f->features) {
    if (!ft->point || ft->point->del) continue;
    if (n < cap) { px[2 * n] = ft->p2d.x; px[2 * n + 1] = ft->p2d.y; level[n] = ft->level; pid[n] = ft->point->id; }
    n++;
  }
  return n;
}

}  // extern "C"
