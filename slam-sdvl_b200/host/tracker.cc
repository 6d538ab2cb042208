// tracker.cc — sequence drivers over the host mirror + their C entry points (used by tests and bench.py).
//
// A sequence is driven like the RUNNING branch of SDVL::HandleFrame (sdvl.cc:55-130): Frame construction
// (pyramid + FAST), motion-model prior (sdvl.cc:278-281), ProcessFrame = ImageAlign::ComputePose +
// FeatureAlign::Reproject + OptimizePose (sdvl.cc:179-203), motion-model update (sdvl.cc:266-276), keyframe decision
// (Map::NeedKeyframe, map.cc:170-188) and EmptyTrash (sdvl.cc:127).  HomographyInit and the mapping thread are out of
// scope: keyframes get fixed map points seeded from ground-truth depth on a known plane (one per free 32-px cell).
//
// Two ways to run the same work:
//   classic : the reference's own call sequence through the class interfaces, one sequence at a time
//             (Frame ctor, ImageAlign::ComputePose, FeatureAlign::Reproject, FeatureAlign::OptimizePose);
//   batched : many sequences in lock-step; per step ONE sdvlb_track_batch submission per group builds all frames,
//             aligns them and searches every candidate point, then the host replays FeatureAlign's logic.
//             Groups (one context + one host thread each) overlap one group's host work with another's GPU work.
// Both give identical results for a sequence (same kernels, same host code, same order of operations).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>

#include "sdvl_host.h"

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
static inline void CpuRelax() { _mm_pause(); }
#else
static inline void CpuRelax() { std::this_thread::yield(); }
#endif

using std::shared_ptr;
using std::vector;

namespace sdvl {

struct SeedPlane { double n[3]; double d; };

struct Sequence {
  Map map;
  std::unique_ptr<FeatureAlign> fa;
  shared_ptr<Frame> last_frame, last_kf;
  double vel[6] = {0, 0, 0, 0, 0, 0};
  int frame_counter = 0;
  int last_matches = 0;
  // staging for the batched path
  vector<sdvlb_align_feat> feats;
  vector<sdvlb_candidate> cands;
  vector<shared_ptr<Point>> cand_points;
  vector<unsigned char> cand_descs;   // ORB mode: 32 bytes per candidate
  vector<sdvlb_match> matches;
};

// Host side of a resident sequence (sdvlb_seq): the tracked state lives on the device; the host keeps the frame handles
// the device still references and plays the mapping thread (seeding when the device's Map::NeedKeyframe fires).
constexpr int kFrameRing = 16;   // frames of a sequence between "result seen" and "build submitted" (prefetch + depth + 2)
struct ResidentSeq {
  sdvlb_seq* h = nullptr;
  sdvlb_frame* last_frame = nullptr;
  sdvlb_frame* kf_frames[SDVLB_SEQ_KF_CAP] = {};
  int kf_added_at[SDVLB_SEQ_KF_CAP] = {};   // frame_counter of the keyframe that filled the slot
  int frame_counter = 0, last_kf_id = 0;
  int64_t next_point_id = 0;
  vector<sdvlb_seq_point> seeds;
  // pipelined run (indices are steps of the current run)
  sdvlb_frame* ring[kFrameRing] = {};   // built (or being built) frames not yet consumed, by step % kFrameRing
  int next_build = 0;    // next step to submit for frame construction
  int next_track = 0;    // next step to submit for tracking
  int n_done = 0;        // steps whose result the host has processed
};

class SequenceDriver {
 public:
  SequenceDriver(Camera* cam, const SeedPlane& plane, int max_points, int kf_every)
      : cam_(cam), plane_(plane), max_points_(max_points), kf_every_(kf_every) {}

  void InitSequence(Sequence* s) { s->fa.reset(new FeatureAlign(&s->map, cam_, Config::MaxMatches())); }

  void SeedKeyframe(Sequence* s, const shared_ptr<Frame>& f, const SE3& gt_pose) {
    f->SetKeyframe();
    const int cell = Config::CellSize();
    const int gw = int(std::ceil(cam_->GetWidth() / cell)), gh = int(std::ceil(cam_->GetHeight() / cell));
    vector<char> occupied(size_t(gw) * gh, 0);
    int n_points = 0;
    for (auto& ft : f->GetFeatures()) {
      if (!ft->GetPoint() || ft->GetPoint()->ToDelete()) continue;
      n_points++;
      const int cx = int(ft->GetPosition()(0) / cell), cy = int(ft->GetPosition()(1) / cell);
      if (cx >= 0 && cx < gw && cy >= 0 && cy < gh) occupied[size_t(cy) * gw + cx] = 1;
    }
    const SE3 gt_wc = gt_pose.Inverse();
    double Rwc[9];
    gt_wc.GetRotation(Rwc);
    const Eigen::Vector3d C = gt_wc.GetTranslation();
    const Eigen::Vector3d est_C = f->GetWorldPosition();
    const vector<Eigen::Vector3i>& corners = f->GetCorners();
    const int n = int(corners.size());
    const int margin = Config::PatchSize() / 2 + 2;
    for (int i = 0; i < n && n_points < max_points_; i++) {
      const Eigen::Vector3i& c = corners[size_t((long long)i * 7919 % n)];
      const int lw = int(cam_->GetWidth()) >> c(2), lh = int(cam_->GetHeight()) >> c(2);
      if (c(0) < margin || c(1) < margin || c(0) >= lw - margin || c(1) >= lh - margin) continue;
      const Eigen::Vector2d px(double(c(0) * (1 << c(2))), double(c(1) * (1 << c(2))));
      const int cx = int(px(0) / cell), cy = int(px(1) / cell);
      if (occupied[size_t(cy) * gw + cx]) continue;
      shared_ptr<Feature> ft = std::make_shared<Feature>(f, px, c(2));
      if (Config::UseORB()) {   // the descriptor an init feature gets at its keyframe (frame.cc:148-161, map.cc:319-323)
        if (c(0) < 19 || c(1) < 19 || c(0) >= lw - 19 || c(1) >= lh - 19) continue;   // ORBDetector::IsInsideLimits
        ft->SetDescriptor(f->GetDescriptors()[size_t((long long)i * 7919 % n)]);
      }
      const Eigen::Vector3d& v = ft->GetVector();
      const Eigen::Vector3d dir(Rwc[0] * v(0) + Rwc[1] * v(1) + Rwc[2] * v(2), Rwc[3] * v(0) + Rwc[4] * v(1) + Rwc[5] * v(2),
                                Rwc[6] * v(0) + Rwc[7] * v(1) + Rwc[8] * v(2));
      const Eigen::Vector3d nrm(plane_.n[0], plane_.n[1], plane_.n[2]);
      const double denom = nrm.dot(dir);
      if (std::fabs(denom) < 1e-9) continue;
      const double sdist = (plane_.d - nrm.dot(C)) / denom;
      if (sdist <= 0) continue;
      shared_ptr<Point> pt = std::make_shared<Point>();
      const Eigen::Vector3d p3d = C + dir * sdist;
      const double depth = (p3d - est_C).norm();
      const double rho = 1.0 / depth;
      pt->InitFixed(ft, p3d, rho, (0.05 * rho) * (0.05 * rho));
      ft->SetPoint(pt);
      f->AddFeature(ft);
      occupied[size_t(cy) * gw + cx] = 1;
      n_points++;
    }
    s->last_kf = f;
  }

  // Everything SDVL::HandleFrame does after ProcessFrame's GPU part. stats: n_tracked, matches, attempts, inliers,
  // outliers, n_feats, gn_iters, keyframe.
  void FinishFrame(Sequence* s, const shared_ptr<Frame>& frame, const SE3& gt_pose, bool first, int32_t* stats) {
    if (first) {
      frame->SetPose(gt_pose);
      SeedKeyframe(s, frame, gt_pose);
      stats[7] = 1;
    } else {
      stats[1] = s->fa->GetMatches();
      stats[2] = s->fa->GetAttempts();
      s->fa->OptimizePose(frame);                                   // sdvl.cc:200
      stats[3] = s->fa->GetInliers();
      stats[4] = s->fa->GetOutliers();
      const SE3 mov = frame->GetPose() * s->last_frame->GetPose().Inverse();   // sdvl.cc:266-276
      double vel[6];
      SE3::Log(mov, vel);
      for (int i = 0; i < 6; i++) s->vel[i] = 0.9 * (0.5 * vel[i] + 0.5 * s->vel[i]);
      const int npoints = frame->GetNumPoints();                   // map.cc:170-188
      const bool enough_its = (frame->GetID() - s->last_kf->GetID()) >= kf_every_;
      const bool lost_many = npoints < s->last_matches * 0.7;
      const bool lost_some = npoints < s->last_matches * 0.9;
      s->last_matches = std::max(s->last_matches, npoints);
      if ((enough_its && lost_some) || lost_many) {
        s->last_matches = npoints;
        SeedKeyframe(s, frame, gt_pose);
        stats[7] = 1;
      }
    }
    int nf = 0;
    for (auto& ft : frame->GetFeatures())
      if (ft->GetPoint() && !ft->GetPoint()->ToDelete()) nf++;
    stats[5] = nf;
    // Frame <-> Feature shared_ptr cycle (the reference breaks it in Map::EmptyTrash / RemoveFeatures): once a frame
    // stops being the alignment reference nothing on this path reads its feature list again; keyframes stay alive
    // through the init features their live points hold.
    if (s->last_frame) s->last_frame->RemoveFeatures();
    s->last_frame = frame;
    s->frame_counter++;
    s->map.EmptyTrash();                                            // sdvl.cc:127
  }

  // The reference's call sequence through the class interfaces (sdvl.cc:59,92-94,185-200).
  void ClassicStep(Sequence* s, const uint8_t* img, int w, int h, const SE3& gt_pose, double est[7], int32_t* stats) {
    std::memset(stats, 0, 8 * sizeof(int32_t));
    cv::Mat m(h, w, CV_8UC1, const_cast<uint8_t*>(img));
    shared_ptr<Frame> frame = std::make_shared<Frame>(cam_, static_cast<ORBDetector*>(nullptr), m, true);
    frame->SetID(s->frame_counter);
    const bool first = !s->last_frame;
    if (!first) {
      frame->SetPose(SE3::Exp(s->vel) * s->last_frame->GetPose());   // SetMotionModel
      ImageAlign image_align;
      stats[0] = image_align.ComputePose(s->last_frame, frame);
      stats[6] = image_align.GetIterations();
      s->fa->Reproject(frame, s->last_frame, s->last_kf);
    }
    FinishFrame(s, frame, gt_pose, first, stats);
    frame->GetPose().ToArray(est);
  }

  Camera* cam() { return cam_; }

  // SeedKeyframe for a resident sequence: same rule, fed from the device's feature list of the frame and the pinned
  // corner mirror; the new points go to the device with sdvlb_seq_add_points.  Returns the number of seeded points.
  int SeedResident(sdvlb_ctx* ctx, ResidentSeq* s, sdvlb_frame* f, const SE3& est_pose, const SE3& gt_pose,
                   const sdvlb_seq_feat* feats, int n_feats) {
    const int cell = Config::CellSize();
    const int gw = int(std::ceil(cam_->GetWidth() / cell)), gh = int(std::ceil(cam_->GetHeight() / cell));
    vector<char> occupied(size_t(gw) * gh, 0);
    int n_points = 0;
    for (int i = 0; i < n_feats; i++) {
      if (!(feats[i].flags & SDVLB_FEAT_HAS_POINT)) continue;
      n_points++;
      const int cx = int(feats[i].px[0] / cell), cy = int(feats[i].px[1] / cell);
      if (cx >= 0 && cx < gw && cy >= 0 && cy < gh) occupied[size_t(cy) * gw + cx] = 1;
    }
    const SE3 gt_wc = gt_pose.Inverse();
    double Rwc[9];
    gt_wc.GetRotation(Rwc);
    const Eigen::Vector3d C = gt_wc.GetTranslation();
    const Eigen::Vector3d est_C = est_pose.Inverse().GetTranslation();
    const int32_t* xyls = nullptr;
    int n = 0;
    if (sdvlb_frame_corners(f, &xyls, &n)) throw std::runtime_error(std::string("sdvl-b200: corners: ") + sdvlb_last_error());
    const int margin = Config::PatchSize() / 2 + 2;
    s->seeds.clear();
    const int before = n_points;
    // visiting order i * 7919 mod n, kept incrementally (this loop is on the critical path of the sequence's next frame)
    const int stride = n > 0 ? 7919 % n : 0;
    int at = 0;
    for (int i = 0; i < n && n_points < max_points_; i++, at = (at + stride >= n ? at + stride - n : at + stride)) {
      const int32_t* c = xyls + 4 * size_t(at);
      const int lw = int(cam_->GetWidth()) >> c[2], lh = int(cam_->GetHeight()) >> c[2];
      if (c[0] < margin || c[1] < margin || c[0] >= lw - margin || c[1] >= lh - margin) continue;
      const Eigen::Vector2d px(double(c[0] * (1 << c[2])), double(c[1] * (1 << c[2])));
      const int cx = int(px(0) / cell), cy = int(px(1) / cell);
      if (occupied[size_t(cy) * gw + cx]) continue;
      // ORB mode: the init feature gets its descriptor at the keyframe (on the device, sdvlb_seq_add_points)
      if (Config::UseORB() && (c[0] < 19 || c[1] < 19 || c[0] >= lw - 19 || c[1] >= lh - 19)) continue;
      const Eigen::Vector3d v = cam_->Unproject(px);
      const Eigen::Vector3d dir(Rwc[0] * v(0) + Rwc[1] * v(1) + Rwc[2] * v(2), Rwc[3] * v(0) + Rwc[4] * v(1) + Rwc[5] * v(2),
                                Rwc[6] * v(0) + Rwc[7] * v(1) + Rwc[8] * v(2));
      const Eigen::Vector3d nrm(plane_.n[0], plane_.n[1], plane_.n[2]);
      const double denom = nrm.dot(dir);
      if (std::fabs(denom) < 1e-9) continue;
      const double sdist = (plane_.d - nrm.dot(C)) / denom;
      if (sdist <= 0) continue;
      const Eigen::Vector3d p3d = C + dir * sdist;
      const double depth = (p3d - est_C).norm();
      const double rho = 1.0 / depth;
      sdvlb_seq_point p;
      std::memset(&p, 0, sizeof(p));
      p.pos[0] = p3d(0); p.pos[1] = p3d(1); p.pos[2] = p3d(2);
      p.ref_px[0] = p.cur_px[0] = px(0); p.ref_px[1] = p.cur_px[1] = px(1);
      p.idepth = rho;
      p.idepth_std = std::sqrt((0.05 * rho) * (0.05 * rho));
      p.user_id = s->next_point_id++;
      p.ref_level = p.cur_level = c[2];
      p.flags = SDVLB_CAND_FIXED;
      s->seeds.push_back(p);
      occupied[size_t(cy) * gw + cx] = 1;
      n_points++;
    }
    double T[7];
    est_pose.ToArray(T);
    int slot = -1;
    if (sdvlb_seq_add_points(ctx, s->h, f, T, s->seeds.data(), int(s->seeds.size()), &slot))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_add_points failed: ") + sdvlb_last_error());
    s->kf_frames[slot] = f;
    s->kf_added_at[slot] = s->frame_counter;
    s->last_kf_id = s->frame_counter;
    return n_points - before;
  }

 private:
  Camera* cam_;
  SeedPlane plane_;
  int max_points_, kf_every_;
};

// ------------------------------------------------------------------------------------------------ one group
// One context, one host thread, a fixed subset of the sequences.
class Group {
 public:
  Group(int device, const SeedPlane& plane, int max_points, int kf_every, int n_seq, bool timing, bool resident)
      : cam_(), driver_(&cam_, plane, max_points, kf_every), seqs_(resident ? 0 : n_seq), jobs_(n_seq), kf_every_(kf_every),
        resident_(resident), rseqs_(resident ? n_seq : 0) {
    const int rc = sdvlb_ctx_create(device, &Config::Params(), &Config::CameraParams(), &ctx_);
    if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_ctx_create failed: ") + sdvlb_last_error());
    if (Config::UseORB() && sdvlb_ctx_set_orb(ctx_, 1))   // before any frame slot or sequence exists
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_ctx_set_orb failed: ") + sdvlb_last_error());
    if (timing) sdvlb_timing_enable(ctx_, 1);
    // current + prefetched + reference frame and a handful of live keyframes per sequence: no allocation while tracking
    if (sdvlb_ctx_reserve_frames(ctx_, 20 * n_seq))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_ctx_reserve_frames failed: ") + sdvlb_last_error());
    for (auto& s : seqs_) driver_.InitSequence(&s);
    for (auto& r : rseqs_) {
      if (sdvlb_seq_create(ctx_, std::max(256, 2 * std::max(max_points, Config::MaxMatches())), &r.h))
        throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_create failed: ") + sdvlb_last_error());
      seq_handles_.push_back(r.h);
      // SDVL::HandleFrame's own decision after a frame (Map::NeedKeyframe, map.cc:170-188, MinKeyframeIts = kf_every,
      // LostRatio = 0.7) runs on the device, so that later frames can be queued before the host has seen the result
      sdvlb_seq_policy pol;
      std::memset(&pol, 0, sizeof(pol));
      pol.keyframe_rule = 1;
      pol.min_keyframe_its = kf_every;
      pol.lost_ratio = 0.7;
      if (sdvlb_seq_set_policy(ctx_, r.h, &pol))
        throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_set_policy failed: ") + sdvlb_last_error());
    }
    results_.resize(rseqs_.size());
  }
  ~Group() {
    seqs_.clear();   // frames go back to the pool before the context dies
    sdvlb_ctx_destroy(ctx_);
  }
  int size() const { return resident_ ? int(rseqs_.size()) : int(seqs_.size()); }
  bool resident() const { return resident_; }

  // ---- resident sequences: one frame for every sequence, synchronously
  void StepResident(const uint8_t* const* images, int on_device, const double* gt, double* est, int32_t* stats) {
    Device::SetCurrent(ctx_);
    const int n = size();
    vector<sdvlb_frame*> built(n, nullptr);
    const auto tA = std::chrono::steady_clock::now();
    if (sdvlb_frames_submit(ctx_, images, n, on_device, 1, Config::NumFeatures(), built.data()))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_frames_submit failed: ") + sdvlb_last_error());
    const auto tB = std::chrono::steady_clock::now();
    const bool first = rseqs_[0].last_frame == nullptr;   // the sequences of a group start together
    auto tC = tB;
    if (first) {
      InitResident(built, gt, 1, est, stats);
    } else {
      if (sdvlb_seq_track_submit(ctx_, seq_handles_.data(), built.data(), n))
        throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_track_submit failed: ") + sdvlb_last_error());
      tC = std::chrono::steady_clock::now();
      if (sdvlb_seq_track_collect(ctx_, results_.data()))
        throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_track_collect failed: ") + sdvlb_last_error());
      AccumulateSlowest(n);
      for (int i = 0; i < n; i++) {
        if (results_[i].status != SDVLB_SEQ_TRACKED) throw std::runtime_error("sdvl-b200: a sequence was not tracked in a lock-step step");
        FinishTracked(i, built[i], results_[i], gt + 7 * size_t(i), est + 7 * size_t(i), stats + 8 * size_t(i));
      }
    }
    const auto tD = std::chrono::steady_clock::now();
    phase_s_[0] += std::chrono::duration<double>(tB - tA).count();
    phase_s_[1] += std::chrono::duration<double>(tC - tB).count();
    phase_s_[2] += std::chrono::duration<double>(tD - tC).count();
  }
  sdvlb_ctx* ctx() { return ctx_; }

  void StepClassic(const uint8_t* const* images, const double* gt, double* est, int32_t* stats) {
    Device::SetCurrent(ctx_);
    const int w = int(cam_.GetWidth()), h = int(cam_.GetHeight());
    for (int i = 0; i < size(); i++)
      driver_.ClassicStep(&seqs_[i], images[i], w, h, SE3(gt + 7 * i), est + 7 * i, stats + 8 * i);
  }

  void StepBatched(const uint8_t* const* images, int on_device, const double* gt, double* est, int32_t* stats) {
    Device::SetCurrent(ctx_);
    const int n = size();
    const int w = int(cam_.GetWidth()), h = int(cam_.GetHeight());
    std::memset(stats, 0, size_t(n) * 8 * sizeof(int32_t));
    const auto tA = std::chrono::steady_clock::now();
    // ---- phase A: marshal
    for (int i = 0; i < n; i++) MarshalJob(i, images[i], on_device);
    // ---- phase B: one submission
    const auto tB = std::chrono::steady_clock::now();
    const int rc = sdvlb_track_batch(ctx_, jobs_.data(), n, w, h, 1);
    if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_track_batch failed: ") + sdvlb_last_error());
    const auto tC = std::chrono::steady_clock::now();
    // ---- phase C: host replay
    for (int i = 0; i < n; i++) ReplayJob(i, SE3(gt + 7 * i), est + 7 * i, stats + 8 * i);
    const auto tD = std::chrono::steady_clock::now();
    phase_s_[0] += std::chrono::duration<double>(tB - tA).count();
    phase_s_[1] += std::chrono::duration<double>(tC - tB).count();
    phase_s_[2] += std::chrono::duration<double>(tD - tC).count();
  }
  const double* phase_seconds() {
    phase_s_[4] = 0;
    for (auto& s : seqs_) phase_s_[4] += s.fa->ransac_seconds;
    return phase_s_;
  }
  void reset_phases() {
    for (double& v : phase_s_) v = 0;
    for (auto& s : seqs_) s.fa->ransac_seconds = 0;
  }

  // ---- pipelined run: frame batches are built one step ahead of the tracking they feed (sdvlb_frames_submit on the
  // context's build stream), tracking is submitted asynchronously and collected when the device is done, so a host
  // thread can interleave several groups.  Tables are [sequence][step] with row stride `table_stride`.
  void BeginRun(const uint8_t* const* images, int on_device, int n_steps, int table_stride, const double* gt,
                double* est, int32_t* stats) {
    images_ = images; on_device_ = on_device; n_steps_ = n_steps; stride_ = table_stride; gt_ = gt; est_ = est;
    stats_ = stats;
    step_ = 0; in_flight_ = false;
    if (resident_) {
      for (auto& r : rseqs_) {
        r.next_build = r.next_track = r.n_done = 0;
        for (auto& f : r.ring) f = nullptr;
      }
      subs_.clear();
      return;
    }
    built_.assign(size(), nullptr);
    next_built_.assign(size(), nullptr);
    ahead_.clear();
    if (n_steps_ > 0) SubmitBuild(0, &built_);
    // frame batches are built `prefetch_` steps ahead of the tracking they feed, so that the PCIe link (uploads) and
    // the build stream never wait for the host between steps
    for (int k = 1; k < prefetch_ && k < n_steps_; k++) {
      ahead_.emplace_back(size(), nullptr);
      SubmitBuild(k, &ahead_.back());
    }
  }
  bool RunDone() const {
    if (!resident_) return step_ >= n_steps_;
    if (!subs_.empty()) return false;
    for (const auto& r : rseqs_) if (r.n_done < n_steps_) return false;
    return true;
  }
  // One scheduling round of the pipelined run; returns whether anything happened.  `block`: nothing else to do on this
  // thread, so waiting inside a collect is fine.
  bool Pump(bool block) {
    if (resident_) return PumpResident(block);
    if (!InFlight()) { SubmitStep(); return true; }
    if (block || Poll()) { FinishStep(); return true; }
    return false;
  }
  void AddIdle(double s) { phase_s_[7] += s; }
  bool InFlight() const { return in_flight_; }

  void SubmitStep() {   // build(step+1) then track(step)
    Device::SetCurrent(ctx_);
    const auto tA = std::chrono::steady_clock::now();
    if (step_ + prefetch_ < n_steps_) {
      ahead_.emplace_back(size(), nullptr);
      SubmitBuild(step_ + prefetch_, &ahead_.back());
    }
    const int n = size();
    const int w = int(cam_.GetWidth()), h = int(cam_.GetHeight());
    for (int i = 0; i < n; i++) {
      MarshalJob(i, nullptr);
      jobs_[i].cur = built_[i];
    }
    const auto tB = std::chrono::steady_clock::now();
    const int rc = sdvlb_track_submit(ctx_, jobs_.data(), n, w, h, 1);
    if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_track_submit failed: ") + sdvlb_last_error());
    in_flight_ = true;
    const auto tC = std::chrono::steady_clock::now();
    phase_s_[0] += std::chrono::duration<double>(tB - tA).count();
    phase_s_[1] += std::chrono::duration<double>(tC - tB).count();
  }
  bool Poll() {
    const int rc = sdvlb_track_poll(ctx_);
    if (rc < 0) throw std::runtime_error(std::string("sdvl-b200: sdvlb_track_poll failed: ") + sdvlb_last_error());
    return rc == 1;
  }
  void FinishStep() {   // waits if the device is not done yet
    Device::SetCurrent(ctx_);
    const auto tA = std::chrono::steady_clock::now();
    const int rc = sdvlb_track_collect(ctx_);
    if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_track_collect failed: ") + sdvlb_last_error());
    in_flight_ = false;
    const auto tB = std::chrono::steady_clock::now();
    const int n = size();
    for (int i = 0; i < n; i++) {
      const size_t o = size_t(i) * stride_ + step_;
      int32_t* st = stats_ + 8 * o;
      std::memset(st, 0, 8 * sizeof(int32_t));
      ReplayJob(i, SE3(gt_ + 7 * o), est_ + 7 * o, st);
    }
    AdvanceBuilt();
    step_++;
    const auto tC = std::chrono::steady_clock::now();
    phase_s_[1] += std::chrono::duration<double>(tB - tA).count();
    phase_s_[6] += std::chrono::duration<double>(tB - tA).count();
    phase_s_[2] += std::chrono::duration<double>(tC - tB).count();
  }

 private:
  // Fills job i from sequence i's state: motion-model prior, ImageAlign features, FeatureAlign candidates.
  void MarshalJob(int i, const uint8_t* image, int on_device = 0) {
    Sequence& s = seqs_[i];
    sdvlb_track_job& j = jobs_[i];
    std::memset(&j, 0, sizeof(j));
    j.image = image;
    j.image_on_device = on_device;
    j.want_corners = 1;
    j.nfeatures = Config::NumFeatures();
    if (s.last_frame) {
      (SE3::Exp(s.vel) * s.last_frame->GetPose()).ToArray(j.T_cur);   // sdvl.cc:278-281
      s.last_frame->GetPose().ToArray(j.T_ref);
      ImageAlign::CollectFeatures(s.last_frame, &s.feats);
      s.fa->CollectCandidates(s.frame_counter, s.last_frame, false, &s.cands, &s.cand_points,
                              Config::UseORB() ? &s.cand_descs : nullptr);
      s.matches.resize(s.cands.size());
      j.cand_desc = Config::UseORB() ? s.cand_descs.data() : nullptr;
      j.ref = s.last_frame->Handle();
      j.feats = s.feats.data();
      j.n_feats = int(s.feats.size());
      j.cands = s.cands.data();
      j.n_cands = int(s.cands.size());
      j.matches = s.matches.data();
    }
  }
  // Host half of ProcessFrame for job i once the device results are in.
  void ReplayJob(int i, const SE3& gt_pose, double* est, int32_t* st) {
    Sequence& s = seqs_[i];
    sdvlb_track_job& j = jobs_[i];
    shared_ptr<Frame> frame = std::make_shared<Frame>(&cam_, ctx_, j.cur, s.frame_counter);
    const bool first = !s.last_frame;
    const auto t0 = std::chrono::steady_clock::now();
    if (!first) {
      frame->SetPose(SE3(j.T_cur));
      st[0] = j.n_tracked;
      st[6] = j.gn_iters;
      s.fa->ApplyMatches(frame, s.cand_points, s.matches.data());
    }
    const auto t1 = std::chrono::steady_clock::now();
    driver_.FinishFrame(&s, frame, gt_pose, first, st);
    frame->GetPose().ToArray(est);
    phase_s_[3] += std::chrono::duration<double>(t1 - t0).count();
    phase_s_[5] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
  }
  // ---------------------------------------------------------------------------------------------- resident sequences
  // First frame of every sequence of `built` (one per sequence, in order): ground-truth pose + map seeding (stand-in
  // for SDVL's initialisation).  Tables are indexed [i * stride].
  void InitResident(const vector<sdvlb_frame*>& built, const double* gt, int stride, double* est, int32_t* stats) {
    const int n = size();
    if (sdvlb_frames_wait(ctx_, built.data(), n))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_frames_wait failed: ") + sdvlb_last_error());
    for (int i = 0; i < n; i++) InitOne(i, built[i], gt + 7 * size_t(i) * stride, est + 7 * size_t(i) * stride, stats + 8 * size_t(i) * stride);
  }
  void InitOne(int i, sdvlb_frame* frame, const double* gt7, double* est7, int32_t* st) {
    ResidentSeq& s = rseqs_[i];
    std::memset(st, 0, 8 * sizeof(int32_t));
    const SE3 gt_pose(gt7);
    if (sdvlb_seq_reset(ctx_, s.h, frame, gt7))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_reset failed: ") + sdvlb_last_error());
    for (auto& k : s.kf_frames) k = nullptr;
    s.frame_counter = 0;
    st[5] = driver_.SeedResident(ctx_, &s, frame, gt_pose, gt_pose, nullptr, 0);
    st[7] = 1;
    gt_pose.ToArray(est7);
    if (s.last_frame) sdvlb_frame_destroy(ctx_, s.last_frame);
    s.last_frame = frame;
    s.frame_counter++;
  }
  // Host half of a tracked frame: results, seeding when the device's keyframe rule fired (the sequence is on hold until
  // the new points -- or an empty list -- reach it), frame lifetimes.
  void FinishTracked(int i, sdvlb_frame* frame, const sdvlb_seq_result& r, const double* gt7, double* est7, int32_t* st) {
    ResidentSeq& s = rseqs_[i];
    std::memset(st, 0, 8 * sizeof(int32_t));
    for (int k = 0; k < 8; k++) post_cycles_[k] += r.phase_cycles[k];
    post_cycles_[8] += 1;
    for (int k = 0; k < 4; k++) align_cycles_[k] += r.align_cycles[k];
    st[0] = r.n_tracked; st[1] = r.matches; st[2] = r.attempts; st[3] = r.inliers; st[4] = r.outliers;
    st[6] = r.gn_iters;
    std::memcpy(est7, r.pose, 7 * sizeof(double));
    // keyframe slots nothing references any more give their frames back (a result only speaks for the slots that
    // were filled before its frame)
    for (int k = 0; k < SDVLB_SEQ_KF_CAP; k++)
      if (s.kf_frames[k] && s.frame_counter > s.kf_added_at[k] && r.kf_live[k] == 0) {
        if (s.kf_frames[k] != s.last_frame) sdvlb_frame_destroy(ctx_, s.kf_frames[k]);
        s.kf_frames[k] = nullptr;
      }
    int seeded = 0;
    if (r.need_keyframe) {
      seeded = driver_.SeedResident(ctx_, &s, frame, SE3(r.pose), SE3(gt7), r.feats, r.n_feats);
      st[7] = 1;
    }
    st[5] = r.n_points + seeded;
    if (s.last_frame) {
      bool held = false;
      for (int k = 0; k < SDVLB_SEQ_KF_CAP && !held; k++) held = s.kf_frames[k] == s.last_frame;
      if (!held) sdvlb_frame_destroy(ctx_, s.last_frame);
    }
    s.last_frame = frame;
    s.frame_counter++;
  }

  // Latency breakdown of the SLOWEST sequence of a submission (the kernels are one CTA per sequence: a step lasts as
  // long as its slowest CTA), separately for the FeatureAlign and the ImageAlign kernel.
  void AccumulateSlowest(int n) {
    int bp = -1, ba = -1;
    long long tp = -1, ta = -1;
    for (int j = 0; j < n; j++) {
      const sdvlb_seq_result& r = results_[j];
      if (r.status != SDVLB_SEQ_TRACKED) continue;
      long long p = 0, a = 0;
      for (int k = 0; k < 7; k++) p += r.phase_cycles[k];
      for (int k = 0; k < 4; k++) a += r.align_cycles[k];
      if (p > tp) { tp = p; bp = j; }
      if (a > ta) { ta = a; ba = j; }
    }
    if (bp < 0) return;
    for (int k = 0; k < 8; k++) slowest_[k] += results_[bp].phase_cycles[k];
    for (int k = 0; k < 4; k++) slowest_[9 + k] += results_[ba].align_cycles[k];
    slowest_[8] += 1;
  }

  struct TrackEntry { int seq; int step; };
  void FinishOldest() {
    const auto t0 = std::chrono::steady_clock::now();
    if (sdvlb_seq_track_collect(ctx_, results_.data()))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_track_collect failed: ") + sdvlb_last_error());
    const auto t1 = std::chrono::steady_clock::now();
    const vector<TrackEntry> entries = std::move(subs_.front());
    subs_.pop_front();
    AccumulateSlowest(int(entries.size()));
    for (size_t j = 0; j < entries.size(); j++) {
      const int i = entries[j].seq, k = entries[j].step;
      ResidentSeq& s = rseqs_[i];
      const sdvlb_seq_result& r = results_[j];
      if (r.status == SDVLB_SEQ_HELD) continue;   // queued behind a keyframe decision: the frame is submitted again
      if (r.status != SDVLB_SEQ_TRACKED || k != s.n_done)
        throw std::runtime_error("sdvl-b200: resident sequence result out of order");
      const size_t o = size_t(i) * stride_ + k;
      sdvlb_frame* frame = s.ring[k % kFrameRing];
      s.ring[k % kFrameRing] = nullptr;
      FinishTracked(i, frame, r, gt_ + 7 * o, est_ + 7 * o, stats_ + 8 * o);
      s.n_done = k + 1;
      if (r.need_keyframe) s.next_track = k + 1;   // whatever was queued behind this frame has been skipped
    }
    const auto t2 = std::chrono::steady_clock::now();
    phase_s_[1] += std::chrono::duration<double>(t1 - t0).count();
    phase_s_[6] += std::chrono::duration<double>(t1 - t0).count();
    phase_s_[2] += std::chrono::duration<double>(t2 - t1).count();
    phase_s_[5] += std::chrono::duration<double>(t2 - t1).count();
  }

  bool PumpResident(bool block) {
    Device::SetCurrent(ctx_);
    bool progressed = false;
    const int n = size();
    // 1. finished submissions, oldest first
    while (!subs_.empty()) {
      const int rc = sdvlb_seq_track_poll(ctx_);
      if (rc < 0) throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_track_poll failed: ") + sdvlb_last_error());
      if (rc != 1) break;
      FinishOldest();
      progressed = true;
    }
    const auto tA = std::chrono::steady_clock::now();
    // 2. tracking: every sequence whose next frame has been handed to the build stream, up to depth_ submissions in
    //    flight.  The critical path of a sequence is this chain, so it is queued before any prefetch work.
    while (int(subs_.size()) < depth_) {
      track_seqs_.clear(); track_frames_.clear();
      vector<TrackEntry> entries;
      for (int i = 0; i < n; i++) {
        ResidentSeq& s = rseqs_[i];
        if (!s.last_frame || s.next_track >= n_steps_ || s.next_track >= s.next_build) continue;
        track_seqs_.push_back(s.h);
        track_frames_.push_back(s.ring[s.next_track % kFrameRing]);
        entries.push_back(TrackEntry{i, s.next_track});
      }
      if (entries.empty()) break;
      if (sdvlb_seq_track_submit(ctx_, track_seqs_.data(), track_frames_.data(), int(entries.size())))
        throw std::runtime_error(std::string("sdvl-b200: sdvlb_seq_track_submit failed: ") + sdvlb_last_error());
      for (const TrackEntry& e : entries) rseqs_[e.seq].next_track++;
      subs_.push_back(std::move(entries));
      progressed = true;
    }
    const auto tB = std::chrono::steady_clock::now();
    // 3. frame construction, prefetch_ steps ahead of each sequence's tracking
    {
      build_imgs_.clear(); build_entries_.clear();
      for (int pass = 0; pass < prefetch_; pass++)
        for (int i = 0; i < n; i++) {
          ResidentSeq& s = rseqs_[i];
          if (s.next_build >= n_steps_ || s.next_build >= s.next_track + prefetch_) continue;
          build_imgs_.push_back(images_[size_t(i) * stride_ + s.next_build]);
          build_entries_.push_back(TrackEntry{i, s.next_build});
          s.next_build++;
        }
      if (!build_entries_.empty()) {
        build_out_.assign(build_entries_.size(), nullptr);
        if (sdvlb_frames_submit(ctx_, build_imgs_.data(), int(build_imgs_.size()), on_device_, 1, Config::NumFeatures(), build_out_.data()))
          throw std::runtime_error(std::string("sdvl-b200: sdvlb_frames_submit failed: ") + sdvlb_last_error());
        for (size_t j = 0; j < build_entries_.size(); j++)
          rseqs_[build_entries_[j].seq].ring[build_entries_[j].step % kFrameRing] = build_out_[j];
        progressed = true;
      }
    }
    // 4. sequences without a track yet: their first frame initialises them on the host (needs the built frame)
    for (int i = 0; i < n; i++) {
      ResidentSeq& s = rseqs_[i];
      if (s.n_done != 0 || s.next_track != 0 || s.next_build == 0 || s.last_frame) continue;
      sdvlb_frame* f = s.ring[0];
      if (sdvlb_frames_wait(ctx_, &f, 1))
        throw std::runtime_error(std::string("sdvl-b200: sdvlb_frames_wait failed: ") + sdvlb_last_error());
      const size_t o = size_t(i) * stride_;
      s.ring[0] = nullptr;
      InitOne(i, f, gt_ + 7 * o, est_ + 7 * o, stats_ + 8 * o);
      s.n_done = 1;
      s.next_track = 1;
      progressed = true;
    }
    const auto tC = std::chrono::steady_clock::now();
    phase_s_[1] += std::chrono::duration<double>(tB - tA).count();
    phase_s_[0] += std::chrono::duration<double>(tC - tB).count();
    if (!progressed && block && !subs_.empty()) {   // nothing else to do on this thread: wait inside the collect
      FinishOldest();
      progressed = true;
    }
    return progressed;
  }
  void AdvanceBuilt() {
    if (!ahead_.empty()) {
      built_.swap(ahead_.front());
      ahead_.pop_front();
    }
  }
  void SubmitBuild(int step, vector<sdvlb_frame*>* out) {
    const int n = size();
    img_ptrs_.resize(n);
    for (int i = 0; i < n; i++) img_ptrs_[i] = images_[size_t(i) * stride_ + step];
    const int rc = sdvlb_frames_submit(ctx_, img_ptrs_.data(), n, on_device_, 1, Config::NumFeatures(), out->data());
    if (rc) throw std::runtime_error(std::string("sdvl-b200: sdvlb_frames_submit failed: ") + sdvlb_last_error());
  }

  Camera cam_;
  SequenceDriver driver_;
  sdvlb_ctx* ctx_ = nullptr;
  vector<Sequence> seqs_;
  vector<sdvlb_track_job> jobs_;
  int kf_every_ = 0;
  bool resident_ = false;
  // resident pipelined run: tracking submissions in flight (oldest first) and per-call staging
  std::deque<vector<TrackEntry>> subs_;
  vector<sdvlb_seq*> track_seqs_;
  vector<sdvlb_frame*> track_frames_, build_out_;
  vector<const uint8_t*> build_imgs_;
  vector<TrackEntry> build_entries_;
 public:
  double slowest_[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // as post_cycles_ + align_cycles_, [8] = submissions
  double align_cycles_[4] = {0, 0, 0, 0};
  double post_cycles_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // seq_post_kernel phase cycles summed over frames, [8] = frames
 private:
  vector<ResidentSeq> rseqs_;
  vector<sdvlb_seq*> seq_handles_;
  vector<sdvlb_seq_result> results_;
  // thread-seconds: [0] marshal (+ frame-batch submission), [1] tracking submission + wait, [2] host replay total,
  // of which [3] FeatureAlign::ApplyMatches, [4] its RANSAC (SelectInliers), [5] FinishFrame (OptimizePose, motion
  // model, keyframe seeding), [6] waiting only
  double phase_s_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // pipelined run state
  const uint8_t* const* images_ = nullptr;
  int on_device_ = 0, n_steps_ = 0, stride_ = 0, step_ = 0;
  const double* gt_ = nullptr;
  double* est_ = nullptr;
  int32_t* stats_ = nullptr;
  bool in_flight_ = false;
  vector<sdvlb_frame*> built_, next_built_;
  std::deque<vector<sdvlb_frame*>> ahead_;   // frame batches of the steps after the current one, oldest first
 public:
  int prefetch_ = 2;
  int depth_ = 2;    // tracking submissions in flight per group (resident sequences)
 private:
  vector<const uint8_t*> img_ptrs_;
};

// ------------------------------------------------------------------------------------------------ all groups
// `n_groups` contexts (each with its own streams, staging and sequences) driven by `n_threads` host threads; thread t
// owns groups t, t + n_threads, ...
class BatchTracker {
 public:
  BatchTracker(const SeedPlane& plane, int max_points, int kf_every, int n_seq, int n_groups, int n_threads, int device,
               bool timing, bool resident)
      : n_seq_(n_seq), resident_(resident) {
    n_groups = std::max(1, std::min(n_groups, n_seq));
    n_threads_ = n_threads <= 0 ? n_groups : std::max(1, std::min(n_threads, n_groups));
    int left = n_seq;
    for (int g = 0; g < n_groups; g++) {
      const int take = (left + (n_groups - g) - 1) / (n_groups - g);
      offsets_.push_back(n_seq - left);
      groups_.emplace_back(new Group(device, plane, max_points, kf_every, take, timing, resident));
      left -= take;
    }
    if (n_groups > 1) {
      for (int t = 0; t < n_threads_; t++) workers_.emplace_back([this, t] { WorkerLoop(t); });
    }
  }
  ~BatchTracker() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      quit_ = true;
      generation_++;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
    groups_.clear();
  }

  // One frame for every sequence, all groups in lock-step.
  void Step(const uint8_t* const* images, int on_device, int classic, const double* gt, double* est, int32_t* stats) {
    if (resident_ && classic) throw std::runtime_error("sdvl-b200: a resident tracker has no classic mode");
    mode_ = resident_ ? 3 : (classic ? 1 : 0);
    images_ = images; on_device_ = on_device; gt_ = gt; est_ = est; stats_ = stats;
    Dispatch();
  }

  // n_steps frames for every sequence; groups run free (no barrier between steps), frame batches one step ahead.
  // images: n_seq x n_steps pointers ([sequence][step]); gt/est: n_seq x n_steps x 7; stats: n_seq x n_steps x 8.
  void Run(const uint8_t* const* images, int on_device, int n_steps, const double* gt, double* est, int32_t* stats) {
    mode_ = 2;
    images_ = images; on_device_ = on_device; n_steps_ = n_steps; gt_ = gt; est_ = est; stats_ = stats;
    Dispatch();
  }

  void TimingRead(double ms[SDVLB_K_COUNT], int64_t launches[SDVLB_K_COUNT], int reset) {
    for (int k = 0; k < SDVLB_K_COUNT; k++) { ms[k] = 0; launches[k] = 0; }
    for (auto& g : groups_) {
      double m[SDVLB_K_COUNT];
      int64_t l[SDVLB_K_COUNT];
      sdvlb_timing_read(g->ctx(), m, l, reset);
      for (int k = 0; k < SDVLB_K_COUNT; k++) { ms[k] += m[k]; launches[k] += l[k]; }
    }
  }
  void Counters(int64_t* launches, int64_t* h2d, int64_t* d2h, int reset) {
    *launches = 0; *h2d = 0; *d2h = 0;
    for (auto& g : groups_) {
      int64_t a, b, c;
      sdvlb_ctx_counters(g->ctx(), &a, &b, &c, reset);
      *launches += a; *h2d += b; *d2h += c;
    }
  }
  void Phases(double out[8], int reset) {   // summed over groups (thread-seconds)
    for (int i = 0; i < 8; i++) out[i] = 0;
    for (auto& g : groups_) {
      const double* ph = g->phase_seconds();
      for (int i = 0; i < 8; i++) out[i] += ph[i];
      if (reset) g->reset_phases();
    }
  }
  void SetPrefetch(int depth) { for (auto& g : groups_) g->prefetch_ = std::max(1, std::min(depth, 6)); }
  void SetDepth(int depth) { for (auto& g : groups_) g->depth_ = std::max(1, std::min(depth, SDVLB_SEQ_DEPTH)); }
  void PostCycles(double out[13], int reset) {
    for (int i = 0; i < 13; i++) out[i] = 0;
    for (auto& g : groups_) {
      for (int i = 0; i < 9; i++) { out[i] += g->post_cycles_[i]; if (reset) g->post_cycles_[i] = 0; }
      for (int i = 0; i < 4; i++) { out[9 + i] += g->align_cycles_[i]; if (reset) g->align_cycles_[i] = 0; }
    }
  }
  void SlowestCycles(double out[13], int reset) {
    for (int i = 0; i < 13; i++) out[i] = 0;
    for (auto& g : groups_)
      for (int i = 0; i < 13; i++) { out[i] += g->slowest_[i]; if (reset) g->slowest_[i] = 0; }
  }
  sdvlb_ctx* ctx0() { return groups_[0]->ctx(); }
  int n_seq() const { return n_seq_; }
  int n_groups() const { return int(groups_.size()); }
  int n_threads() const { return workers_.empty() ? 1 : n_threads_; }

 private:
  void Dispatch() {
    error_.clear();
    if (workers_.empty()) {
      RunThread(0, 1);
    } else {
      {
        std::unique_lock<std::mutex> lk(mu_);
        pending_ = int(workers_.size());
        generation_++;
      }
      cv_.notify_all();
      std::unique_lock<std::mutex> lk(mu_);
      done_cv_.wait(lk, [this] { return pending_ == 0; });
    }
    if (!error_.empty()) throw std::runtime_error(error_);
  }

  void RunThread(int t, int stride) {
    try {
      vector<int> mine;
      for (int g = t; g < int(groups_.size()); g += stride) mine.push_back(g);
      if (mode_ != 2) {
        for (int g : mine) {
          const int o = offsets_[g];
          if (mode_ == 1) groups_[g]->StepClassic(images_ + o, gt_ + 7 * o, est_ + 7 * o, stats_ + 8 * o);
          else if (mode_ == 3) groups_[g]->StepResident(images_ + o, on_device_, gt_ + 7 * o, est_ + 7 * o, stats_ + 8 * o);
          else groups_[g]->StepBatched(images_ + o, on_device_, gt_ + 7 * o, est_ + 7 * o, stats_ + 8 * o);
        }
        return;
      }
      // free-running pipelined groups, interleaved on this thread
      for (int g : mine) {
        const size_t o = size_t(offsets_[g]) * n_steps_;
        groups_[g]->BeginRun(images_ + o, on_device_, n_steps_, n_steps_, gt_ + 7 * o, est_ + 7 * o, stats_ + 8 * o);
      }
      int active = 0;
      for (int g : mine) if (!groups_[g]->RunDone()) active++;
      while (active > 0) {
        bool progressed = false;
        for (int g : mine) {
          Group& G = *groups_[g];
          if (G.RunDone()) continue;
          if (G.Pump(active == 1)) progressed = true;   // the only unfinished group may block in its collect
          if (G.RunDone()) active--;
        }
        if (!progressed) {   // everything in flight: nothing to do but wait (plain memory polls, no driver calls)
          const auto t0 = std::chrono::steady_clock::now();
          for (int k = 0; k < 32; k++) CpuRelax();   // ~1 us: leaves the core's issue slots to a sibling hardware thread
          groups_[mine[0]]->AddIdle(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
      }
    } catch (const std::exception& e) {
      std::unique_lock<std::mutex> lk(mu_);
      error_ = e.what();
    }
  }
  void WorkerLoop(int t) {
    long seen = 0;
    while (true) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (quit_) return;
      }
      RunThread(t, n_threads_);
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }

  int n_seq_, n_threads_ = 1;
  bool resident_ = false;
  vector<std::unique_ptr<Group>> groups_;
  vector<int> offsets_;
  vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  long generation_ = 0;
  int pending_ = 0;
  bool quit_ = false;
  int mode_ = 0;   // 0 batched lock-step, 1 classic lock-step, 2 pipelined run
  const uint8_t* const* images_ = nullptr;
  int on_device_ = 0, n_steps_ = 0;
  const double* gt_ = nullptr;
  double* est_ = nullptr;
  int32_t* stats_ = nullptr;
  std::string error_;
};

}  // namespace sdvl

// ================================================================================================ C entry points
static thread_local std::string g_host_error;

extern "C" {

const char* sdvlh_last_error(void) { return g_host_error.c_str(); }

// Config is process-wide in the reference (singleton, config.h:56); set it before creating trackers.
void sdvlh_config_set(const sdvlb_params* p, const sdvlb_camera* cam) { sdvl::Config::Set(*p, *cam); }
// Config::UseORB() for the trackers / frames created from now on (contexts pick it up when they are created)
void sdvlh_config_set_orb(int on) { sdvl::Config::SetUseORB(on != 0); }

// A context of its own for the test hooks below (the process-wide default context keeps the camera it was made with).
struct HookContext {
  sdvlb_ctx* ctx = nullptr;
  HookContext() {
    if (sdvlb_ctx_create(0, &sdvl::Config::Params(), &sdvl::Config::CameraParams(), &ctx))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_ctx_create failed: ") + sdvlb_last_error());
    if (sdvl::Config::UseORB() && sdvlb_ctx_set_orb(ctx, 1))
      throw std::runtime_error(std::string("sdvl-b200: sdvlb_ctx_set_orb failed: ") + sdvlb_last_error());
    sdvl::Device::SetCurrent(ctx);
  }
  ~HookContext() {
    sdvl::Device::SetCurrent(nullptr);
    sdvlb_ctx_destroy(ctx);
  }
};

// Camera::SetDistortions + Camera::UndistortImage of the host mirror (test hook).  Config must have been set.
int sdvlh_camera_undistort(const double d[5], const uint8_t* in, int w, int h, uint8_t* out) {
  try {
    using namespace sdvl;
    HookContext hook_ctx;
    Camera cam;
    cam.SetDistortions(d[0], d[1], d[2], d[3], d[4]);
    cv::Mat src(h, w, CV_8UC1, const_cast<uint8_t*>(in)), dst;
    cam.UndistortImage(src, &dst);
    std::memcpy(out, dst.data, size_t(w) * h);
    return 0;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return -1;
  }
}

// Test hook for the class-level mapping pass: builds a keyframe with n candidate points (state taken from `seeds`),
// runs Map::UpdateCandidates against each of the n_frames images (poses: n_frames x 7) and writes the candidates'
// final state back into `seeds` (status = outcome of the last pass a candidate took part in).  *n_left = candidates
// still in the list.  Config must have been set (sdvlh_config_set).
int sdvlh_map_update_candidates(const uint8_t* ref_img, const double ref_T[7], const uint8_t* const* cur_imgs,
                                const double* cur_poses, int n_frames, int w, int h, sdvlb_seed* seeds, int n,
                                double depth_mean, int min_kf_id, int* n_left) {
  try {
    using namespace sdvl;
    HookContext hook_ctx;   // declared first: destroyed after every frame below
    Camera cam;
    cv::Mat rm(h, w, CV_8UC1, const_cast<uint8_t*>(ref_img));
    shared_ptr<Frame> ref = std::make_shared<Frame>(&cam, static_cast<ORBDetector*>(nullptr), rm, false);
    ref->SetPose(SE3(ref_T));
    ref->SetKeyframe();
    Map map;
    vector<shared_ptr<Point>> pts(n);
    for (int i = 0; i < n; i++) {
      auto ft = std::make_shared<Feature>(ref, Eigen::Vector2d(seeds[i].ref_px[0], seeds[i].ref_px[1]), seeds[i].ref_level);
      pts[i] = std::make_shared<Point>();
      pts[i]->InitCandidate(ft, 1.0 / seeds[i].rho);
      sdvlb_seed s0 = seeds[i];
      s0.status = SDVLB_SEED_UPDATED;
      pts[i]->FromSeed(s0);
      pts[i]->SetLastKeyframeID(seeds[i].last_kf_id);
      ft->SetPoint(pts[i]);
      map.AddCandidate(pts[i]);
      seeds[i].status = -1;
    }
    for (int k = 0; k < n_frames; k++) {
      cv::Mat m(h, w, CV_8UC1, const_cast<uint8_t*>(cur_imgs[k]));
      shared_ptr<Frame> cur = std::make_shared<Frame>(&cam, static_cast<ORBDetector*>(nullptr), m, true);
      cur->SetPose(SE3(cur_poses + 7 * k));
      vector<shared_ptr<Point>> before = map.GetCandidates();
      map.UpdateCandidates(cur, depth_mean, min_kf_id);
      const vector<sdvlb_seed>& out = map.LastSeeds();
      for (size_t c = 0; c < before.size(); c++) {
        if (out[c].status < 0) continue;
        for (int i = 0; i < n; i++)
          if (pts[i] == before[c]) {
            const sdvlb_frame* keep = seeds[i].ref_frame;
            seeds[i] = out[c];
            seeds[i].ref_frame = keep;
          }
      }
      map.EmptyTrash();
    }
    if (n_left) *n_left = int(map.GetCandidates().size());
    return 0;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return -1;
  }
}

// Test hook for Map::InitCandidates + UpdateCandidates + AddConnectionsPoints of the host mirror.  A new keyframe
// (img_new, T_new) is matched against an older keyframe (img_old, T_old): InitCandidates creates candidates;
// UpdateCandidates then runs on every image of upd_imgs (poses upd_T, n_upd x 7); finally the points that converged
// are searched in a last frame (img_last, T_last) through AddConnectionsPoints.
// out[0] = candidates created, out[1] = list entries after InitCandidates, out[2] = entries left after the updates,
// out[3] = points that became fixed, out[4] = links made by AddConnectionsPoints.  cand_px / cand_rho (cap entries):
// position in the new keyframe and inverse depth of every created candidate after the updates.
int sdvlh_map_init_candidates(const uint8_t* img_new, const double T_new[7], const uint8_t* img_old, const double T_old[7],
                              const uint8_t* const* upd_imgs, const double* upd_T, int n_upd, const uint8_t* img_last,
                              const double T_last[7], int w, int h, double depth_mean, int32_t out[5], double* cand_px,
                              double* cand_rho, int32_t* cand_fixed, int cap) {
  try {
    using namespace sdvl;
    HookContext hook_ctx;   // declared first: destroyed after every frame below
    Camera cam;
    auto mk = [&](const uint8_t* img, const double* T, bool corners) {
      cv::Mat m(h, w, CV_8UC1, const_cast<uint8_t*>(img));
      shared_ptr<Frame> f = std::make_shared<Frame>(&cam, static_cast<ORBDetector*>(nullptr), m, corners);
      f->SetPose(SE3(T));
      return f;
    };
    shared_ptr<Frame> kf_new = mk(img_new, T_new, true), kf_old = mk(img_old, T_old, true);
    kf_new->SetKeyframe();
    kf_old->SetKeyframe();
    Map map;
    out[0] = map.InitCandidates(kf_new, {kf_old}, depth_mean);
    out[1] = int(map.GetCandidates().size());
    vector<shared_ptr<Point>> created;
    for (auto& p : map.GetCandidates())
      if (created.empty() || created.back() != p) created.push_back(p);
    for (int k = 0; k < n_upd; k++) {
      shared_ptr<Frame> cur = mk(upd_imgs[k], upd_T + 7 * k, true);
      map.UpdateCandidates(cur, depth_mean, -1000);
      map.EmptyTrash();
    }
    out[2] = int(map.GetCandidates().size());
    out[3] = 0;
    for (size_t i = 0; i < created.size(); i++) {
      if (created[i]->IsFixed()) out[3]++;
      if (int(i) < cap) {
        cand_px[2 * i] = created[i]->GetInitFeature()->GetPosition()(0);
        cand_px[2 * i + 1] = created[i]->GetInitFeature()->GetPosition()(1);
        cand_rho[i] = created[i]->GetInverseDepth();
        cand_fixed[i] = created[i]->IsFixed() ? 1 : 0;
      }
    }
    // keep only the fixed points on the keyframes, then look for them in the last frame
    shared_ptr<Frame> last = mk(img_last, T_last, true);
    out[4] = map.AddConnectionsPoints(last, {kf_new});
    return 0;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return -1;
  }
}

// resident != 0: the sequences live on the device (sdvlb_seq_*): the host neither marshals features nor replays matches.
void* sdvlh_tracker_create2(const double plane[4], int max_points, int kf_every, int n_seq, int n_groups, int n_threads,
                            int device, int timing, int resident) {
  try {
    sdvl::SeedPlane pl;
    pl.n[0] = plane[0]; pl.n[1] = plane[1]; pl.n[2] = plane[2]; pl.d = plane[3];
    return new sdvl::BatchTracker(pl, max_points, kf_every, n_seq, n_groups, n_threads, device, timing != 0, resident != 0);
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return nullptr;
  }
}

void* sdvlh_tracker_create(const double plane[4], int max_points, int kf_every, int n_seq, int n_groups, int n_threads,
                           int device, int timing) {
  return sdvlh_tracker_create2(plane, max_points, kf_every, n_seq, n_groups, n_threads, device, timing, 0);
}

void sdvlh_tracker_destroy(void* t) { delete static_cast<sdvl::BatchTracker*>(t); }

// images: n_seq pointers (host, or device when on_device); gt/est: n_seq x 7; stats: n_seq x 8.
// classic != 0 runs the reference call sequence through the class interfaces instead of the batched submission.
int sdvlh_tracker_step(void* t, const uint8_t* const* images, int on_device, int classic, const double* gt_poses,
                       double* est_poses, int32_t* stats) {
  try {
    static_cast<sdvl::BatchTracker*>(t)->Step(images, on_device, classic, gt_poses, est_poses, stats);
    return 0;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return -1;
  }
}

// Pipelined run of n_steps frames per sequence (tables are [sequence][step]); see BatchTracker::Run.
int sdvlh_tracker_run(void* t, const uint8_t* const* images, int on_device, int n_steps, const double* gt_poses,
                      double* est_poses, int32_t* stats) {
  try {
    static_cast<sdvl::BatchTracker*>(t)->Run(images, on_device, n_steps, gt_poses, est_poses, stats);
    return 0;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    return -1;
  }
}

int sdvlh_tracker_timing_read(void* t, double ms[SDVLB_K_COUNT], int64_t launches[SDVLB_K_COUNT], int reset) {
  static_cast<sdvl::BatchTracker*>(t)->TimingRead(ms, launches, reset);
  return 0;
}

int sdvlh_tracker_counters(void* t, int64_t* launches, int64_t* h2d, int64_t* d2h, int reset) {
  static_cast<sdvl::BatchTracker*>(t)->Counters(launches, h2d, d2h, reset);
  return 0;
}

// Thread-seconds spent in: [0] host marshalling (+ frame-batch submission), [1] tracking submission + wait, [2] host
// replay, of which [3] ApplyMatches, [4] its RANSAC, [5] FinishFrame; [6] blocked waiting for the device; [7] idle
// polling (every group of the thread in flight).
int sdvlh_tracker_phases(void* t, double out[8], int reset) {
  static_cast<sdvl::BatchTracker*>(t)->Phases(out, reset);
  return 0;
}

// Latency breakdown of the device-side FeatureAlign kernel, summed over tracked frames (SM cycles): cell ranks,
// SelectPoints, RANSAC hypotheses / supporters / replay, OptimizePose, the rest, (unused); out[8] = number of frames.
// out[9..12]: the ImageAlign kernel (PrecomputePatches, residuals, reduction, solve + update).
int sdvlh_tracker_post_cycles(void* t, double out[13], int reset) {
  static_cast<sdvl::BatchTracker*>(t)->PostCycles(out, reset);
  return 0;
}

// The same breakdown for the slowest sequence of every submission (out[8] = number of submissions).
int sdvlh_tracker_slowest_cycles(void* t, double out[13], int reset) {
  static_cast<sdvl::BatchTracker*>(t)->SlowestCycles(out, reset);
  return 0;
}

// How many steps ahead of the tracking the frame batches (upload + pyramid + FAST) are submitted in sdvlh_tracker_run.
void sdvlh_tracker_set_prefetch(void* t, int depth) { static_cast<sdvl::BatchTracker*>(t)->SetPrefetch(depth); }
void sdvlh_tracker_set_depth(void* t, int depth) { static_cast<sdvl::BatchTracker*>(t)->SetDepth(depth); }

void* sdvlh_tracker_ctx(void* t) { return static_cast<sdvl::BatchTracker*>(t)->ctx0(); }
int sdvlh_tracker_groups(void* t) { return static_cast<sdvl::BatchTracker*>(t)->n_groups(); }
int sdvlh_tracker_threads(void* t) { return static_cast<sdvl::BatchTracker*>(t)->n_threads(); }

}  // extern "C"
