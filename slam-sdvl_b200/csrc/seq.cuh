// seq.cuh — device-side state of a resident sequence and the launch descriptors of its kernels (seq.cu), shared with the
// host side of the C-ABI (seq_api.cu).
//
// A sequence's tracked state is the feature list of its last frame, every feature carrying the map point it observes
// inline (the reference keeps Frame::features_ -> Feature -> Point objects on the heap, frame.h:148-172,
// feature.h:97-104, point.h:121-146).  Two lists ping-pong: tracking frame k+1 reads list k and writes list k+1.
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

// Pinned, device-visible result block of one sequence (written by seq_post_kernel).
struct SeqResultHost {
  double pose[7];
  int32_t stats[8];    // n_tracked(n_meas), matches, attempts, inliers, outliers, n_points, gn_iters, n_feats
  int32_t kf_live[SDVLB_SEQ_KF_CAP];
  int32_t error;       // != 0: a capacity was exceeded
  int32_t phase_cycles[8];   // seq_post_kernel latency breakdown (SM cycles of the sequence's CTA)
  int32_t align_cycles[4];   // image_align_kernel: PrecomputePatches, residuals, reduction, solve + update
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
  int32_t frame_id;
  int32_t n_list;      // features in list[cur]
  int32_t cur;
  int32_t n_cands;     // candidates of the step in flight
  int32_t pad0_;
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
  sdvlb_align_feat* afeat;         // ImageAlign features of the step in flight
  SearchCandDev* cands;            // SearchPoint candidates
  int32_t* cand_feat;              // candidate -> index in list[cur]
  sdvlb_match* matches;
  uint8_t* align_scratch;          // patch / Jacobian caches of image_align_kernel
  // FeatureAlign scratch, one entry per candidate / found feature
  int32_t* c_cell; int32_t* c_score; int32_t* c_rank;
  double* o_a; double* o_pos; double* o_scale; double* o_err; int32_t* o_flag;
  SeqResultHost* result;           // pinned host memory
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
  AlignJobDev* jobs;     // n entries, written by the prep kernel
  FrameDev* frames;      // n entries (SearchCandDev::cur_index)
  const struct SeqCmd* cmds;            // commands of this step's sequences, applied by the prep kernel
  int2 cmd_range[SDVLB_SEQ_BATCH];      // (first, count) per sequence of the step
  uint32_t* d_done;      // device counter of CTAs that finished the post kernel
  uint32_t* h_flag;      // pinned completion word (nullptr: the caller launches signal_kernel itself)
  uint32_t seq_no;       // value to publish
  uint32_t pad_;
  DevParams dp;
  PyrGeom g;
};

// Host -> device commands applied before a step (seq_apply_kernel): restart a track / append points.
struct SeqCmd {
  SeqState* seq;
  int32_t kind;          // 0 reset, 1 add points
  int32_t n;             // points
  int32_t kf_slot;
  int32_t pad_;
  FrameDev frame;        // reset: the frame the track starts at
  const uint8_t* kf_pyr;
  double T[7];           // reset: pose; add: keyframe pose
  const sdvlb_seq_point* pts;   // device copy
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

cudaError_t sdvlb_launch_seq_apply(const SeqCmd* d_cmds, const int2* d_ranges, int n_ranges, const DevParams& dp,
                                   cudaStream_t stream);
cudaError_t sdvlb_launch_seq_prep(const SeqStepArgs& A, cudaStream_t stream);
cudaError_t sdvlb_launch_seq_post(const SeqStepArgs& A, cudaStream_t stream);
cudaError_t sdvlb_launch_pose_call(const PoseCallArgs& A, cudaStream_t stream);
cudaError_t sdvlb_launch_search_seq(const SeqStepArgs& A, cudaStream_t stream);
