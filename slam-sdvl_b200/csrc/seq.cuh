// seq.cuh — device-side state of a resident sequence and the launch descriptors of its kernels (align.cu, search.cu,
// seq.cu), shared with the host side of the C-ABI (seq_api.cu).
//
// A sequence's tracked state is the feature list of its last frame, every feature carrying the map point it observes
// inline (the reference keeps Frame::features_ -> Feature -> Point objects on the heap, frame.h:148-172,
// feature.h:97-104, point.h:121-146).  Two lists ping-pong: tracking frame k+1 reads list k and writes list k+1.
//
// A tracked frame is three launches: seq_align_kernel (commands of the mapping thread, motion-model prior,
// ImageAlign), search_seq_kernel (one warp per feature that observes a point: ProjectPoint + SearchPoint) and
// seq_post_kernel (SelectPoints, RANSAC, OptimizePose, motion model, tracking quality / keyframe decision, result).
// Several steps may be queued behind each other: a sequence that needs its mapping thread (new keyframe) raises `hold`
// and the steps already queued behind it leave it untouched until the caller has answered.
#pragma once
#include "common.cuh"

// One feature of a frame + the map point it observes.
struct SeqFeat {
  double px[2];        // Feature::p2d_ (level-0 pixels, in the frame the list belongs to)
  double v[3];         // Feature::v_ (unit bearing)
  double pos[3];       // Point::GetPosition()
  double ref_px[2];    // init feature of the point: position in its keyframe
  double ref_v[3];     //   bearing
  double idepth, idepth_std;
  int64_t user_id;
  int32_t level;       // Feature::level_
  int32_t ref_level;
  int32_t kf;          // keyframe slot of the init feature
  int32_t flags;       // SEQF_*
  int32_t n_successful, n_failed;   // Point::n_successful_ / n_failed_ (point.cc:102-115)
  int32_t status;      // Point::PointStatus
  int32_t n_unpromoted;// times Unpromote ran since the point was handed over (Point::b_ increments)
};
enum { SEQF_HAS_POINT = 1, SEQF_FIXED = 2 };
enum { SEQP_FOUND = 0, SEQP_NOT_FOUND = 1, SEQP_SEEN = 2, SEQP_UNSEEN = 3 };   // Point::PointStatus order (point.h)

struct SeqKf {
  const uint8_t* pyr;  // pyramid of the keyframe (FrameDev::pyr)
  double T[7];         // its pose
};

// Pinned, device-visible result block of one sequence and one submission slot (written by seq_post_kernel).
struct SeqResultHost {
  double pose[7];
  int32_t stats[8];    // n_tracked(n_meas), matches, attempts, inliers, outliers, n_points, gn_iters, n_feats
  int32_t kf_live[SDVLB_SEQ_KF_CAP];
  int32_t error;       // != 0: a capacity was exceeded
  int32_t status;      // SDVLB_SEQ_TRACKED / _HELD / _IDLE
  int32_t quality;     // SDVL::CalcTrackingQuality: SDVLB_TRACKING_*
  int32_t need_keyframe;   // Map::NeedKeyframe said yes (the sequence is on hold until the caller answers)
  int32_t lost_frames;
  int32_t phase_cycles[8];   // seq_post_kernel latency breakdown (SM cycles of the sequence's CTA)
  int32_t align_cycles[4];   // ImageAlign kernel: PrecomputePatches, residuals, reduction, solve + update
  int32_t pad_[3];
  // sdvlb_seq_feat feats[max_feats] follows
};

// Header of a sequence's device allocation; the arrays it points to live in the same allocation.
struct SeqState {
  // ---- dynamic state
  double T_last[7];    // pose of the frame the current list belongs to
  double vel[6];       // SDVL::vel_ (sdvl.cc:266-276)
  FrameDev last;       // that frame
  int32_t has_last;
  int32_t frame_id;    // frames tracked since the last reset (the frame the current list belongs to)
  int32_t n_list;      // features in list[cur]
  int32_t cur;
  int32_t overflow;    // a command tried to append beyond max_feats (sticky until the next reset)
  int32_t hold;        // != 0: waiting for the caller (keyframe / relocalisation); queued steps skip the sequence
  // Map::NeedKeyframe (map.cc:170-188) / SDVL::CalcTrackingQuality (sdvl.cc:240-264) state
  int32_t last_matches;
  int32_t last_kf_frame;   // frame_id at which points were last appended (last_kf_->GetID())
  int32_t lost_frames;
  int32_t pad0_;
  sdvlb_seq_policy policy;
  sdvlb_rand rng;
  SeqKf kf[SDVLB_SEQ_KF_CAP];
  // ImageAlign outputs of the step in flight
  double align_pose[7];
  double align_error;
  int32_t align_info[2];
  int32_t align_cycles[4];
  // ---- static layout
  int32_t max_feats, n_cells;
  SeqFeat* list[2];
  int32_t* cell_order;             // FeatureAlign::cell_order_ (already shuffled for the next frame)
  sdvlb_match* matches;            // SearchPoint result per feature of list[cur]
  uint8_t* align_scratch;          // per-feature H_f and cache overflow of the ImageAlign kernel
  // FeatureAlign scratch, one entry per feature / found feature
  int32_t* c_cell; int32_t* c_score; int32_t* c_rank; int32_t* c_next;
  double* o_a; double* o_pos; double* o_scale; double* o_err; int32_t* o_flag;
  SeqResultHost* result[SDVLB_SEQ_DEPTH];   // pinned host memory, one block per submission slot
  // Config::UseORB(): Feature::descriptor_ of the init feature of every feature's point (8 words each), parallel to
  // list[0] / list[1]; nullptr when the sequence was created outside ORB mode
  uint32_t* fdesc[2];
};

// Pose-refinement problem (FeatureAlign lists as arrays): obs i is (a = SimpleProject(v), pos, scale = 2^-level).
struct PoseProblem {
  double* o_a; double* o_pos; double* o_scale; double* o_err; int32_t* o_flag;
  int n;
};

#define SDVLB_SEQ_BATCH 64
struct SeqStepArgs {
  int n;
  int max_feats;
  SeqState* seq[SDVLB_SEQ_BATCH];
  FrameDev cur[SDVLB_SEQ_BATCH];
  const struct SeqCmd* cmds;            // commands of this step's sequences, applied by the align kernel
  int2 cmd_range[SDVLB_SEQ_BATCH];      // (first, count) per sequence of the step
  uint32_t* d_done;      // device counter of CTAs that finished the post kernel
  uint32_t* h_flag;      // pinned completion word
  uint32_t seq_no;       // value to publish
  int32_t slot;          // result slot of this submission (seq_no % SDVLB_SEQ_DEPTH)
  int32_t use_orb;       // Config::UseORB(): SearchPoint scored by descriptor distance
  int32_t pad_;
  DevParams dp;
  PyrGeom g;
};

// Host -> device commands applied before a step: restart a track / append points / release a hold / set the policy.
enum { SEQC_RESET = 0, SEQC_ADD_POINTS = 1, SEQC_RELEASE = 2, SEQC_POLICY = 3 };
struct SeqCmd {
  SeqState* seq;
  int32_t kind;          // SEQC_*
  int32_t n;             // points
  int32_t kf_slot;
  int32_t pad_;
  FrameDev frame;        // reset: the frame the track starts at
  const uint8_t* kf_pyr;
  double T[7];           // reset: pose; add: keyframe pose
  const sdvlb_seq_point* pts;   // device copy
  uint32_t* desc;        // ORB mode: 8 words per point, filled by seq_orb_points_kernel before the commands are applied
  sdvlb_seq_policy policy;
};

// One standalone FeatureAlign pose-refinement call (sdvlb_select_inliers / sdvlb_optimize_pose).
struct PoseCallArgs {
  sdvlb_pose_obs* obs;   // device copy, flags updated in place
  int n;
  int mode;              // 0 SelectInliers, 1 OptimizePose
  double* T;             // 7: frame pose in, refined pose out (mode 1)
  sdvlb_rand* rng;       // mode 0, in/out
  double* scratch;       // 7 * n doubles
  int32_t* iscratch;     // n ints
  DevParams dp;
};

#if defined(__CUDACC__)
// The commands of one sequence (contiguous, applied in order) by the threads of one CTA (any block size).
__device__ inline void seq_apply_commands(const SeqCmd* __restrict__ cmds, int2 range, const DevParams& dp, int* s_base) {
  const int tid = threadIdx.x;
  for (int ci = range.x; ci < range.x + range.y; ci++) {
    const SeqCmd& C = cmds[ci];
    SeqState* S = C.seq;
    if (C.kind != SEQC_ADD_POINTS) {
      if (tid == 0) {
        if (C.kind == SEQC_RESET) {
          for (int i = 0; i < 7; i++) S->T_last[i] = C.T[i];
          for (int i = 0; i < 6; i++) S->vel[i] = 0.0;
          S->last = C.frame;
          S->has_last = 1;
          S->n_list = 0;
          S->frame_id = 0;
          S->overflow = 0;
          S->hold = 0;
          S->last_matches = 0;
          S->last_kf_frame = 0;
          S->lost_frames = 0;
          for (int i = 0; i < 7; i++) C.frame.pose[i] = C.T[i];
        } else if (C.kind == SEQC_RELEASE) {
          S->hold = 0;
        } else if (C.kind == SEQC_POLICY) {
          S->policy = C.policy;
        }
      }
      __syncthreads();
      continue;
    }
    // append points to the current list (keyframe seeding / mapping thread output); answers a keyframe hold
    if (tid == 0) {
      *s_base = S->n_list;
      SeqKf& K = S->kf[C.kf_slot];
      K.pyr = C.kf_pyr;
      for (int i = 0; i < 7; i++) K.T[i] = C.T[i];
    }
    __syncthreads();
    const int base = *s_base;
    SeqFeat* L = S->list[S->cur];
    for (int k = tid; k < C.n; k += blockDim.x) {
      if (base + k >= S->max_feats) break;
      const sdvlb_seq_point p = C.pts[k];
      SeqFeat f;
      f.px[0] = p.cur_px[0]; f.px[1] = p.cur_px[1];
      cam_unproject_unit(dp.cam, p.cur_px[0], p.cur_px[1], f.v);
      f.pos[0] = p.pos[0]; f.pos[1] = p.pos[1]; f.pos[2] = p.pos[2];
      f.ref_px[0] = p.ref_px[0]; f.ref_px[1] = p.ref_px[1];
      cam_unproject_unit(dp.cam, p.ref_px[0], p.ref_px[1], f.ref_v);
      f.idepth = p.idepth; f.idepth_std = p.idepth_std;
      f.user_id = p.user_id;
      f.level = p.cur_level; f.ref_level = p.ref_level;
      f.kf = C.kf_slot;
      f.flags = SEQF_HAS_POINT | ((p.flags & SDVLB_CAND_FIXED) ? SEQF_FIXED : 0);
      f.n_successful = p.n_successful; f.n_failed = p.n_failed;
      f.status = SEQP_FOUND;
      f.n_unpromoted = 0;
      L[base + k] = f;
      if (C.desc && S->fdesc[0]) {   // Feature::descriptor_ of the init feature (ORB mode)
        const uint4* src = reinterpret_cast<const uint4*>(C.desc + size_t(k) * 8);
        uint4* dst = reinterpret_cast<uint4*>(S->fdesc[S->cur] + size_t(base + k) * 8);
        dst[0] = src[0]; dst[1] = src[1];
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (base + C.n > S->max_feats) S->overflow = 1;   // reported with every later result until the next reset
      S->n_list = min(base + C.n, S->max_feats);
      S->last_kf_frame = S->frame_id;
      S->hold = 0;
    }
    __syncthreads();
  }
}
#endif

cudaError_t sdvlb_launch_seq_apply(const SeqCmd* d_cmds, const int2* d_ranges, int n_ranges, const DevParams& dp,
                                   cudaStream_t stream);
cudaError_t sdvlb_launch_seq_align(const SeqStepArgs& A, int n_bound, cudaStream_t stream);
cudaError_t sdvlb_launch_search_seq(const SeqStepArgs& A, cudaStream_t stream);
cudaError_t sdvlb_launch_seq_orb_points(const SeqCmd* d_cmds, int n_cmds, int max_points, const PyrGeom& g, cudaStream_t stream);
cudaError_t sdvlb_launch_seq_post(const SeqStepArgs& A, cudaStream_t stream);
cudaError_t sdvlb_launch_pose_call(const PoseCallArgs& A, cudaStream_t stream);
