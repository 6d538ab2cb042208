/*
 * sdvl_b200.h — C-ABI of the B200-native SDVL tracking front-end.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * The reference (JdeRobot/slam-SDVL) has no FFI; its boundary for this path is
 * four C++ classes.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree):
 *
 *   sdvlb_frame_create / _detect / _level / _corners / sdvlb_frames_submit
 *        -> Frame::Frame, Frame::CreatePyramid, Frame::CreateCorners,
 *           Frame::GetPyramid, Frame::GetCorners
 *           (frame.h:45,58-60,139; frame.cc:34-56,114-131;
 *            extra/fast_detector.cc:58-175)
 *   sdvlb_image_align
 *        -> ImageAlign::ComputePose / GetError
 *           (image_align.h:41,43; image_align.cc:46-267)
 *   sdvlb_search_points
 *        -> Matcher::SearchPoint (matcher.h:45-46; matcher.cc:45-121) as it is
 *           driven by FeatureAlign::SelectPoints/ProjectPoint
 *           (feature_align.cc:88-150,323-339)
 *   sdvlb_update_candidates
 *        -> the loop body of Map::UpdateCandidates (map.cc:397-498) with
 *           Point::Update / HasConverged (point.cc:63-100,162-176)
 *   sdvlb_track_batch
 *        -> one SDVL::ProcessFrame front half (sdvl.cc:59,185-193) for many
 *           independent sequences in one submission (no reference equivalent;
 *           it is how one GPU is kept busy).
 *
 * All functions return 0 on success or a negative sdvlb_status.  Nothing
 * throws across this boundary.  There is NO CPU fallback: if no CUDA device is
 * usable sdvlb_ctx_create fails with SDVLB_ERR_CUDA.
 *
 * Poses are world->camera (frame.h:90) packed as double[7] =
 * {q0(w), q1(x), q2(y), q3(z), tx, ty, tz} (extra/se3.h:76-77).
 */
#ifndef SDVL_B200_H_
#define SDVL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sdvlb_status {
  SDVLB_OK = 0,
  SDVLB_ERR_CUDA = -1,        /* CUDA runtime error (see sdvlb_last_error) */
  SDVLB_ERR_ARG = -2,         /* invalid argument / unsupported parameter */
  SDVLB_ERR_NOMEM = -3,
  SDVLB_ERR_OVERFLOW = -4,    /* a fixed device-side capacity was exceeded */
  SDVLB_ERR_STATE = -5        /* call order violated (e.g. corners not detected) */
} sdvlb_status;

/* Tunables of the path; defaults are config.cc:55-85 of the reference. */
typedef struct sdvlb_params {
  int32_t pyramid_levels;      /* kPyramidLevels_   5  */
  int32_t cell_size;           /* kCellSize_        32 (only 32 supported) */
  int32_t max_matches;         /* kMaxMatches_      150 */
  int32_t max_align_level;     /* kMaxAlignLevel_   4  */
  int32_t min_align_level;     /* kMinAlignLevel_   2  */
  int32_t max_img_align_its;   /* kMaxImgAlignIts_  30 */
  int32_t align_patch_size;    /* kAlignPatchSize_  4  (only 4 supported) */
  int32_t patch_size;          /* kPatchSize_       8  (only 8 supported) */
  int32_t max_align_its;       /* kMaxAlignIts_     10 */
  int32_t search_size;         /* kSearchSize_      6  */
  int32_t max_fast_levels;     /* kMaxFastLevels_   3  */
  int32_t fast_threshold;      /* kFastThreshold_   10 */
  int32_t num_features;        /* kNumFeatures_     1000 */
  int32_t max_failed;          /* kMaxFailed_       15 */
  int32_t max_optim_pose_its;  /* kMaxOptimPoseIts_ 10 */
  int32_t max_ransac_points;   /* kMaxRansacPoints_ 5  */
  int32_t max_ransac_its;      /* kMaxRansacIts_    100 */
  int32_t min_matches;         /* kMinMatches_      20 */
  double inlier_error_threshold; /* kInlierErrorThreshold_ 2.0 */
} sdvlb_params;

/* Pinhole intrinsics without distortion (camera.h:119-126). */
typedef struct sdvlb_camera {
  double width, height, fx, fy, u0, v0;
} sdvlb_camera;

/* Fills *p with the reference defaults (config.cc:55-85). */
void sdvlb_params_default(sdvlb_params* p);

typedef struct sdvlb_ctx sdvlb_ctx;     /* device, stream, scratch; one per host thread */
typedef struct sdvlb_frame sdvlb_frame; /* device pyramid + corners, pinned host mirror */

/* ---- context ------------------------------------------------------------ */
int sdvlb_ctx_create(int device, const sdvlb_params* params, const sdvlb_camera* cam,
                     sdvlb_ctx** out);
int sdvlb_ctx_destroy(sdvlb_ctx* ctx);
/* Pre-sizes the context's frame pool (device slots + pinned corner mirrors)
 * so that a steady-state tracker performs no CUDA allocation. */
int sdvlb_ctx_reserve_frames(sdvlb_ctx* ctx, int n_frames);
/* Blocks until everything submitted on this context has finished. */
int sdvlb_ctx_sync(sdvlb_ctx* ctx);
/* The cudaStream_t every kernel of this context is launched on (for external
 * CUDA-event timing). */
void* sdvlb_ctx_stream(sdvlb_ctx* ctx);
const char* sdvlb_last_error(void);
/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
int sdvlb_host_alloc(void** p, uint64_t bytes);
int sdvlb_host_free(void* p);
/* Device buffers for callers that keep inputs resident in HBM. */
int sdvlb_dev_alloc(sdvlb_ctx* ctx, void** p, uint64_t bytes);
int sdvlb_dev_free(sdvlb_ctx* ctx, void* p);
int sdvlb_dev_upload(sdvlb_ctx* ctx, void* dst, const void* src, uint64_t bytes);

/* ---- per-kernel timing (CUDA events on the context stream) -------------- */
enum {
  SDVLB_K_PYRAMID = 0, SDVLB_K_FAST = 1, SDVLB_K_SELECT = 2, SDVLB_K_ALIGN = 3,
  SDVLB_K_SEARCH = 4, SDVLB_K_ORB = 5 /* corner descriptors of a frame batch (ORB mode) */, SDVLB_K_POSE = 6,
  SDVLB_K_COUNT = 7
};
int sdvlb_timing_enable(sdvlb_ctx* ctx, int on);
/* Accumulated milliseconds and launch counts per kernel since the last reset.
 * Synchronises the context. */
int sdvlb_timing_read(sdvlb_ctx* ctx, double ms[SDVLB_K_COUNT], int64_t launches[SDVLB_K_COUNT],
                      int reset);

/* Counters since the last reset: kernels launched by this context, bytes copied host->device and device->host. */
int sdvlb_ctx_counters(sdvlb_ctx* ctx, int64_t* kernel_launches, int64_t* h2d_bytes, int64_t* d2h_bytes, int reset);

/* ---- input pre-processing: Camera::UndistortImage ----
 * Camera::SetDistortions (camera.cc:38-67): d = (k1, k2, p1, p2, k3), the five cv::undistort coefficients the
 * reference reads from Camera.d1..d5; all zero switches distortion off (camera.cc:45-48).  When set, every level-0
 * image handed to sdvlb_frame_create / sdvlb_frames_submit / sdvlb_track_batch is first undistorted on the device,
 * as main.cc:133 does before SDVL::HandleFrame -- bit-exact with cv::undistort(in, out, K, D) (OpenCV 4.13:
 * initUndistortRectifyMap CV_16SC2 + remap INTER_LINEAR, BORDER_CONSTANT). */
int sdvlb_ctx_set_distortion(sdvlb_ctx* ctx, const double d[5]);
/* Camera::UndistortImage(in, out) (camera.cc:100-105) on its own: w*h u8 host buffers of the context's camera size.
 * With distortion off it copies, as the reference does. */
int sdvlb_undistort(sdvlb_ctx* ctx, const uint8_t* in, uint8_t* out);

/* ---- ORB descriptor mode ----------------------------------------------------------
 * Config::UseORB() (config.h:139, `SDVL.use_orb`), orb_size 31 (the size the learned pattern exists for).
 * sdvlb_ctx_set_orb(ctx, 1) switches the context to the reference's ORB mode, end to end:
 *   - FAST cells and Frame::FilterCorners use the border margin 4 + orb_size/2 = 19
 *     (extra/fast_detector.cc:63-64,183-184) for every frame built from then on;
 *   - every frame built with corners also gets Frame::descriptors_ -- the reference fills them lazily
 *     (matcher.cc:265-269, frame.cc:148-161), here one kernel per frame batch computes all of them when the frame is
 *     built (sdvlb_frame_descriptors reads them);
 *   - Matcher::SearchPoint scores corners by ORBDetector::Distance against the descriptor of the point's init feature
 *     (matcher.cc:243-277): sdvlb_search_points_orb / sdvlb_track_job.cand_desc for the class API and the batched
 *     tracker, and inside the resident-sequence chain (sdvlb_seq_*), where the descriptor of every map point's init
 *     feature is computed on the device from its keyframe when the point is handed over (sdvlb_seq_add_points) and
 *     travels with the point from frame to frame;
 *   - sdvlb_update_candidates (Map::UpdateCandidates / InitCandidates) scores by descriptor distance too; the
 *     descriptor of a candidate's init feature is recomputed from its keyframe (ref_frame, ref_px, ref_level).
 * A feature's descriptor is ORBDetector::GetDescriptor(keyframe pyramid[level], int(position / 2^level)) wherever the
 * reference creates one (frame.cc:159 + map.cc:322; homography_init.cc:141-149), so it is a function of the keyframe
 * and the feature; a feature outside ORBDetector::IsInsideLimits keeps the zero descriptor of the Feature constructor.
 * Call it before creating frames (frame slots allocated outside ORB mode have no room for descriptors: switching it on
 * regrows the pool, which fails with SDVLB_ERR_STATE while frames are alive) and before sdvlb_seq_create.  It
 * synchronises the context. */
int sdvlb_ctx_set_orb(sdvlb_ctx* ctx, int on);
/* Frame::GetDescriptors(): 32 bytes per corner in Frame::GetCorners() order, at most cap corners; *n = corners the frame
 * has.  The frame must have been built with corners in ORB mode. */
int sdvlb_frame_descriptors(sdvlb_ctx* ctx, const sdvlb_frame* f, uint8_t* desc, int cap, int* n);
/* ORBDetector::GetDescriptor / GetOrientation (extra/orb_detector.cc:350-437) at n positions xyl = n x (x, y, level),
 * level coordinates, of frame f: what Frame::FilterCorners stores in Frame::descriptors_ for the filtered corners
 * (frame.cc:148-161), Map::InitCandidates copies into Feature::descriptor_ (map.cc:319-323) and Matcher::SearchFeatures
 * fills in lazily (matcher.cc:265-269).  desc receives n x 32 bytes; angle (optional) the orientation in degrees
 * (cv::fastAtan2).  Positions must satisfy ORBDetector::IsInsideLimits (19 px from the border of their level), else
 * SDVLB_ERR_ARG is returned and their descriptor is zero (the reference asserts). */
int sdvlb_frame_orb_descriptors(sdvlb_ctx* ctx, const sdvlb_frame* f, const int32_t* xyl, int n, uint8_t* desc,
                                float* angle);

/* ---- Frame -------------------------------------------------------------- */
/* Frame::Frame(camera, detector, img, corners): uploads `img` (u8, `stride`
 * bytes per row), builds the pyramid on the device, optionally runs FAST with
 * budget `nfeatures` (Config::NumFeatures()), mirrors levels and corners into
 * pinned host memory, and synchronises.  `img` may be pageable. */
int sdvlb_frame_create(sdvlb_ctx* ctx, const uint8_t* img, int w, int h, int stride,
                       int want_corners, int nfeatures, sdvlb_frame** out);
/* Frame::CreateCorners(levels, nfeatures) (frame.cc:122-131). Replaces corners. */
int sdvlb_frame_detect(sdvlb_ctx* ctx, sdvlb_frame* f, int nfeatures);
/* Frame::GetPyramid()[l]: continuous u8 host mirror (stride == *w). */
int sdvlb_frame_level(const sdvlb_frame* f, int level, const uint8_t** data, int* w, int* h);
/* Frame::GetCorners(): n records (x, y, level, score), 4 x int32 each, in level
 * coordinates and in the reference's order; score = cv::KeyPoint::response
 * (not kept by the reference, exposed for parity checks).  The pointer aims
 * into the frame's pinned host mirror and stays valid until the frame is
 * destroyed or re-detected. */
int sdvlb_frame_corners(const sdvlb_frame* f, const int32_t** xyls, int* n);
int sdvlb_frame_destroy(sdvlb_ctx* ctx, sdvlb_frame* f);
/* Frame::FilterCorners() (frame.cc:133-146) -> FastDetector::FilterCorners (extra/fast_detector.cc:177-218):
 * per free 32-px cell the corner with the best Shi-Tomasi score (FindShiTomasiScoreAtPoint, extra/utils.cc:61-97)
 * above min_feature_score (Config::MinFeatureScore(), 50).  locked_px: level-0 positions of the frame's features
 * (FastDetector::LockCell), n_locked x (x, y).  indices receives up to cap indices into the frame's corner list, in
 * cell order (Frame::filtered_corners_); *n = how many there are. */
int sdvlb_frame_filter_corners(sdvlb_ctx* ctx, const sdvlb_frame* f, const double* locked_px, int n_locked,
                               int min_feature_score, int32_t* indices, int cap, int* n);

/* ---- ImageAlign --------------------------------------------------------- */
/* One feature of frame1 (frame1->GetFeatures() order). */
typedef struct sdvlb_align_feat {
  double px[2];    /* Feature::GetPosition(), level-0 pixels */
  double v[3];     /* Feature::GetVector(), unit bearing */
  double depth;    /* |Point::GetPosition() - frame1 world position| (image_align.cc:159,234) */
  int32_t valid;   /* GetPoint() != nullptr && !GetPoint()->ToDelete() (image_align.cc:154,229) */
  int32_t pad_;
} sdvlb_align_feat;

/* One Gauss-Newton iteration as seen by ImageAlign::Optimize (image_align.cc:91-124). */
typedef struct sdvlb_gn_iter {
  int32_t level;
  int32_t iter;
  int32_t n_meas;     /* n_meas_ after ComputeResiduals */
  int32_t flags;      /* bit0: iteration rejected (rollback), bit1: NaN solve, bit2: n_meas==0 */
  double T_in[7];     /* relative pose T (frame2 <- frame1) the residuals were evaluated at */
  double H[36];       /* row-major, full symmetric */
  double b[6];        /* Jres_ */
  double x[6];        /* H.ldlt().solve(Jres_) */
  double chi2;        /* chi2 / n_meas */
} sdvlb_gn_iter;

/* Optional teacher forcing, used by parity tests: iteration k is evaluated at
 * T[k] and level l runs exactly iters[l] iterations; the kernel's own
 * accept/rollback decisions are bypassed. */
typedef struct sdvlb_gn_forced {
  const double* T;        /* n_total x 7 */
  const int32_t* iters;   /* indexed by pyramid level, pyramid_levels entries */
  int32_t n_total;
  int32_t pad_;
} sdvlb_gn_forced;

/* ImageAlign::ComputePose(frame1=ref, frame2=cur, fast).
 * T_ref: pose of ref; T_cur: in = prior pose of cur, out = aligned pose
 * (frame2->SetPose).  *n_tracked = return value (n_meas_/patch_area),
 * *error = GetError().  n == 0 returns 0 tracked and leaves T_cur untouched
 * (image_align.cc:55-58).  trace (optional) receives up to trace_cap
 * iterations; *trace_n = number that ran. */
int sdvlb_image_align(sdvlb_ctx* ctx, const sdvlb_frame* ref, sdvlb_frame* cur,
                      const sdvlb_align_feat* feats, int n, const double T_ref[7],
                      double T_cur[7], int fast, int* n_tracked, double* error,
                      sdvlb_gn_iter* trace, int trace_cap, int* trace_n,
                      const sdvlb_gn_forced* forced);

/* ---- Matcher::SearchPoint batch ---------------------------------------- */
enum { SDVLB_CAND_FIXED = 1, SDVLB_CAND_PROJECT = 2 };
typedef struct sdvlb_candidate {
  const sdvlb_frame* ref_frame; /* feature->GetFrame(): frame of the point's init feature */
  double ref_T[7];     /* ref_frame->GetPose() */
  double ref_px[2];    /* feature->GetPosition() */
  double ref_v[3];     /* feature->GetVector() */
  double idepth;       /* point->GetInverseDepth() */
  double idepth_std;   /* point->GetStd() */
  double px[2];        /* in: *px (predicted position) unless SDVLB_CAND_PROJECT */
  double pos[3];       /* point->GetPosition(), used with SDVLB_CAND_PROJECT */
  int32_t ref_level;   /* feature->GetLevel() */
  int32_t flags;       /* SDVLB_CAND_FIXED = point->IsFixed();
                          SDVLB_CAND_PROJECT = run FeatureAlign::ProjectPoint first */
} sdvlb_candidate;

enum { SDVLB_MATCH_UNSEEN = 0, SDVLB_MATCH_NOT_FOUND = 1, SDVLB_MATCH_FOUND = 2 };
typedef struct sdvlb_match {
  double px[2];        /* refined position, level-0 pixels (valid when FOUND) */
  double proj[2];      /* projected position (SDVLB_CAND_PROJECT), else copy of input px */
  int32_t level;       /* search level (valid when FOUND) */
  int32_t status;      /* SDVLB_MATCH_* ; UNSEEN only with SDVLB_CAND_PROJECT */
  int32_t zmssd;       /* best ZMSSD score, -1 if no corner in range */
  int32_t n_in_range;  /* corners that passed GetCornersInRange */
} sdvlb_match;

/* Evaluates Matcher::SearchPoint for n candidates against `cur` (which must
 * have corners).  T_cur == NULL uses the pose the last sdvlb_image_align left
 * on `cur` in device memory. */
int sdvlb_search_points(sdvlb_ctx* ctx, const sdvlb_frame* cur, const sdvlb_candidate* cands,
                        int n, const double T_cur[7], sdvlb_match* out);
/* The same with Config::UseORB() (matcher.cc:79-80,131-134,243-277): corners are gated with the ORB margin and scored
 * by ORBDetector::Distance between cand_desc[i] (n x 32 bytes: feature->GetDescriptor() of candidate i) and the
 * corner's descriptor, accepted below MIN_ORB_THRESHOLD = 100 (matcher.h:37); sdvlb_match.zmssd then carries that
 * distance.  Warp, search level, patch and the Lucas-Kanade refinement are unchanged.  Needs sdvlb_ctx_set_orb(ctx, 1)
 * and T_cur. */
int sdvlb_search_points_orb(sdvlb_ctx* ctx, const sdvlb_frame* cur, const sdvlb_candidate* cands, int n,
                            const double T_cur[7], const uint8_t* cand_desc, sdvlb_match* out);

/* ---- batched front half of ProcessFrame -------------------------------- */
/* Where a level-0 image lives.  HOST: any host memory (one cudaMemcpyAsync per
 * frame); DEVICE: device memory; PINNED: page-locked host memory the device
 * can read (cudaHostAlloc / cudaHostRegister): a whole batch is uploaded by
 * one kernel reading it over PCIe. */
enum { SDVLB_IMG_HOST = 0, SDVLB_IMG_DEVICE = 1, SDVLB_IMG_PINNED = 2 };
typedef struct sdvlb_track_job {
  const uint8_t* image;     /* w*h u8, continuous; host (pinned preferred) or device.
                               NULL: `cur` is a frame made by sdvlb_frames_submit */
  int32_t image_on_device;  /* SDVLB_IMG_* : where `image` lives */
  int32_t want_corners;
  int32_t nfeatures;
  int32_t n_feats;
  int32_t n_cands;
  int32_t n_tracked;        /* out: ImageAlign::ComputePose return value */
  int32_t gn_iters;         /* out: Gauss-Newton iterations ImageAlign ran */
  int32_t pad_;
  const sdvlb_frame* ref;   /* last frame (NULL: only build the frame) */
  sdvlb_frame* cur;         /* out: new frame handle (created by the call); in: prebuilt frame when image == NULL */
  const sdvlb_align_feat* feats;
  const sdvlb_candidate* cands;
  sdvlb_match* matches;     /* out, n_cands entries */
  double T_ref[7];
  double T_cur[7];          /* in: prior, out: ImageAlign result */
  double error;             /* out */
  const uint8_t* cand_desc; /* Config::UseORB() (sdvlb_ctx_set_orb): n_cands x 32 bytes, feature->GetDescriptor() of
                               every candidate's init feature (matcher.cc:109); NULL outside ORB mode */
} sdvlb_track_job;

/* For every job: upload image, pyramid, FAST, ImageAlign(ref,cur), then
 * SearchPoint for all candidates at the aligned pose — submitted as batched
 * launches with a single synchronisation at the end.  `mirror` != 0 also
 * copies pyramids and corners to the pinned host mirrors of the new frames. */
int sdvlb_track_batch(sdvlb_ctx* ctx, sdvlb_track_job* jobs, int n_jobs, int w, int h, int mirror);

/* ---- asynchronous forms ("frame batches" + overlapped tracking) ----------
 * Frames of one sequence depend on each other only through ImageAlign and
 * FeatureAlign (sdvl.cc:93,119,278-281); Frame construction (frame.cc:34-56)
 * does not.  sdvlb_frames_submit enqueues upload + pyramid (+ FAST when
 * want_corners) for n images on the context's BUILD stream and returns at
 * once; the handles can be given to sdvlb_track_batch / _submit as jobs with
 * image == NULL (ordered after the build on the device), or waited for on the
 * host with sdvlb_frames_wait / any sdvlb_frame_* accessor.  Images must stay
 * valid until then.  All jobs of one batch either carry images or prebuilt
 * frames. */
int sdvlb_frames_submit(sdvlb_ctx* ctx, const uint8_t* const* images, int n, int images_on_device /* SDVLB_IMG_* */,
                        int want_corners, int nfeatures, sdvlb_frame** out);
int sdvlb_frames_wait(sdvlb_ctx* ctx, sdvlb_frame* const* frames, int n);
/* sdvlb_track_batch split in two: _submit enqueues everything on the TRACKING
 * stream and returns; _poll returns 1 when the device has finished (0 while
 * running, <0 on error); _collect waits and fills the jobs' outputs.  `jobs`
 * must stay valid until _collect; one submission in flight per context. */
int sdvlb_track_submit(sdvlb_ctx* ctx, sdvlb_track_job* jobs, int n_jobs, int w, int h, int mirror);
int sdvlb_track_poll(sdvlb_ctx* ctx);
int sdvlb_track_collect(sdvlb_ctx* ctx);

/* ---- FeatureAlign pose refinement (feature_align.cc:152-283,341-431) ------
 * RANSAC inlier selection and the Tukey-weighted Gauss-Newton on SE3 that
 * FeatureAlign runs on the matches of a frame, as device kernels
 * (SURVEY.md section 8(f), row 1). */
typedef struct sdvlb_pose_obs {
  double v[3];      /* feature->GetVector() (unit bearing in the frame) */
  double pos[3];    /* feature->GetPoint()->GetPosition() (world) */
  int32_t level;    /* feature->GetLevel() */
  int32_t flags;    /* SDVLB_OBS_*: which of the two FeatureAlign lists the feature is in */
} sdvlb_pose_obs;
enum { SDVLB_OBS_INLIER = 1, SDVLB_OBS_OUTLIER = 2 };

/* glibc rand() (random_r TYPE_3) as an explicit state.  The reference draws
 * from the process-wide rand() (feature_align.cc:53,103,180, never seeded);
 * here every FeatureAlign owns one stream that starts like srand(1). */
typedef struct sdvlb_rand { uint32_t r[34]; int32_t n; int32_t pad_; } sdvlb_rand;
void sdvlb_rand_seed(sdvlb_rand* s, unsigned seed);
int sdvlb_rand_next(sdvlb_rand* s);
/* std::random_shuffle(v, v + n) of libstdc++ driven by this stream. */
void sdvlb_rand_shuffle(sdvlb_rand* s, int32_t* v, int n);

/* FeatureAlign::SelectInliers(frame, fs_found, inliers, outliers)
 * (feature_align.cc:152-216): obs = fs_found in order; on return every
 * obs[i].flags is SDVLB_OBS_INLIER or SDVLB_OBS_OUTLIER and *rng has advanced
 * by the number of rand() calls the reference would have made. */
int sdvlb_select_inliers(sdvlb_ctx* ctx, sdvlb_pose_obs* obs, int n, const double T_frame[7], sdvlb_rand* rng);
/* FeatureAlign::OptimizePose(frame) (feature_align.cc:73-82) up to, not
 * including, RemoveOutliers: OptimizePose(inliers) + RescueOutliers +
 * OptimizePose again when something was rescued.  T_frame: in = frame pose,
 * out = refined pose; flags in/out. */
int sdvlb_optimize_pose(sdvlb_ctx* ctx, sdvlb_pose_obs* obs, int n, double T_frame[7]);

/* ---- resident sequences ----------------------------------------------------
 * The whole per-frame chain of SDVL::ProcessFrame (sdvl.cc:179-203) plus the
 * motion model (sdvl.cc:266-281) with the tracked state of a sequence -- the
 * features of its last frame and the map points they observe -- kept in HBM:
 *   prior pose -> ImageAlign::ComputePose -> FeatureAlign::Reproject
 *   (ProjectPoints, SelectPoints, Matcher::SearchPoint, SelectInliers) ->
 *   FeatureAlign::OptimizePose -> RemoveOutliers -> GetMotionModel
 * runs as one chain of kernels per submission for many sequences; the host
 * only reads the result (pose, statistics, the new frame's feature list) and
 * plays the mapping thread's role (sdvlb_seq_add_points). */
typedef struct sdvlb_seq sdvlb_seq;
#define SDVLB_SEQ_KF_CAP 64   /* keyframes that may be referenced by live points of one sequence */
#define SDVLB_SEQ_DEPTH 4     /* tracking submissions that may be in flight on one context (results are kept in a ring) */

/* What the tracking thread decides by itself after a frame (SDVL::HandleFrame, sdvl.cc:93-120), evaluated by the
 * FeatureAlign kernel so that the next frames can be queued without waiting for the host.  All zero (the default):
 * every frame is adopted and nothing is ever held. */
typedef struct sdvlb_seq_policy {
  int32_t keyframe_rule;     /* 1: Map::NeedKeyframe (map.cc:170-188).  When it fires the result carries need_keyframe
                                and the sequence is HELD: tracking steps already queued behind it skip the sequence
                                (status SDVLB_SEQ_HELD, its frame is not consumed) until sdvlb_seq_add_points or
                                sdvlb_seq_release answers. */
  int32_t min_keyframe_its;  /* Config::MinKeyframeIts() */
  double lost_ratio;         /* Config::LostRatio() */
  int32_t tracking_quality;  /* 1: SDVL::CalcTrackingQuality (sdvl.cc:240-264, Config::MinMatches() = params.min_matches):
                                a TRACKING_BAD frame is not adopted as the next alignment reference (sdvl.cc:99,119);
                                after 3 lost frames the sequence is held for relocalisation (sdvl.cc:74-91). */
  int32_t pad_;
} sdvlb_seq_policy;
enum { SDVLB_SEQ_TRACKED = 0, SDVLB_SEQ_HELD = 1, SDVLB_SEQ_IDLE = 2 };
enum { SDVLB_TRACKING_GOOD = 0, SDVLB_TRACKING_INSUFFICIENT = 1, SDVLB_TRACKING_BAD = 2 };

typedef struct sdvlb_seq_point {   /* a map point handed to the tracker with its observation in the current frame */
  double pos[3];       /* Point::GetPosition() */
  double ref_px[2];    /* Point::GetInitFeature()->GetPosition() in the keyframe, level-0 pixels */
  double cur_px[2];    /* position of the feature observing it in the sequence's current frame */
  double idepth;       /* Point::GetInverseDepth() */
  double idepth_std;   /* Point::GetStd() */
  int64_t user_id;     /* caller's handle of the point; returned in sdvlb_seq_feat */
  int32_t ref_level;   /* init feature level */
  int32_t cur_level;   /* level of the observing feature */
  int32_t flags;       /* SDVLB_CAND_FIXED = Point::IsFixed() */
  int32_t n_successful;/* Point::Score() so far */
  int32_t n_failed;
  int32_t pad_;
} sdvlb_seq_point;

enum { SDVLB_FEAT_HAS_POINT = 1 };
typedef struct sdvlb_seq_feat {    /* one feature of the sequence's current frame (frame->GetFeatures() order) */
  double px[2];        /* Feature::GetPosition() */
  int64_t user_id;
  int32_t level;       /* Feature::GetLevel() */
  int32_t flags;       /* SDVLB_FEAT_HAS_POINT: cleared for outliers (RemoveOutliers, feature_align.cc:245-256) */
} sdvlb_seq_feat;

typedef struct sdvlb_seq_result {
  double pose[7];      /* frame->GetPose() after OptimizePose */
  int32_t n_tracked;   /* ImageAlign::ComputePose return value */
  int32_t matches;     /* FeatureAlign::GetMatches() */
  int32_t attempts;    /* FeatureAlign::GetAttempts() */
  int32_t inliers, outliers;
  int32_t n_points;    /* Frame::GetNumPoints() */
  int32_t gn_iters;    /* ImageAlign Gauss-Newton iterations */
  int32_t n_feats;     /* entries in feats */
  const sdvlb_seq_feat* feats;             /* pinned host memory, valid until SDVLB_SEQ_DEPTH later submissions of
                                              the context have been made */
  int32_t status;      /* SDVLB_SEQ_TRACKED; _HELD: the sequence was on hold, the frame was not consumed (submit it
                          again once the hold is answered); _IDLE: no track yet (no sdvlb_seq_reset) */
  int32_t quality;     /* SDVLB_TRACKING_* (GOOD unless policy.tracking_quality) */
  int32_t need_keyframe;   /* policy.keyframe_rule fired on this frame: the sequence is now held */
  int32_t lost_frames; /* SDVL::lost_frames_ */
  int32_t kf_live[SDVLB_SEQ_KF_CAP];       /* live points per keyframe slot (see sdvlb_seq_add_points) */
  int32_t phase_cycles[8];                 /* latency breakdown of the FeatureAlign kernel for this sequence, SM cycles:
                                              cell ranks, SelectPoints, RANSAC hypotheses, RANSAC supporters, RANSAC
                                              replay + inlier flags, OptimizePose, the rest, (unused) */
  int32_t align_cycles[4];                 /* same for the ImageAlign kernel: PrecomputePatches, residuals,
                                              reduction, solve + pose update */
} sdvlb_seq_result;

/* FeatureAlign(map, camera, max_matches) + an empty track: RNG seeded like srand(1), cell order shuffled once
 * (feature_align.cc:33-54).  max_feats = capacity of a frame's feature list (0: 2 * max_matches, at least 256). */
int sdvlb_seq_create(sdvlb_ctx* ctx, int max_feats, sdvlb_seq** out);
int sdvlb_seq_destroy(sdvlb_ctx* ctx, sdvlb_seq* seq);
/* (Re)starts the track at `frame` with pose T (world->camera), no features, zero velocity (what SDVL's
 * initialisation leaves behind, sdvl.cc:132-177).  Applied in order with later calls on the context's stream. */
int sdvlb_seq_reset(sdvlb_ctx* ctx, sdvlb_seq* seq, const sdvlb_frame* frame, const double T[7]);
/* Mapping thread -> tracker: n points whose init features live in keyframe `kf` (pose T_kf) are appended, in order,
 * to the feature list of the sequence's current frame.  *kf_slot receives the keyframe slot the points reference;
 * `kf` must stay alive until a result reports kf_live[slot] == 0.  Takes effect before the next tracked frame.
 * ORB mode: Feature::descriptor_ of every init feature is computed on the device from `kf` at (ref_px, ref_level). */
int sdvlb_seq_add_points(sdvlb_ctx* ctx, sdvlb_seq* seq, const sdvlb_frame* kf, const double T_kf[7],
                         const sdvlb_seq_point* pts, int n, int* kf_slot);
/* Sets the sequence's policy; takes effect before the next tracked frame (queued like sdvlb_seq_add_points). */
int sdvlb_seq_set_policy(sdvlb_ctx* ctx, sdvlb_seq* seq, const sdvlb_seq_policy* policy);
/* Answers a hold without adding points (sdvlb_seq_add_points releases it too). */
int sdvlb_seq_release(sdvlb_ctx* ctx, sdvlb_seq* seq);
/* One new frame for each of n sequences (frames built by sdvlb_frames_submit / sdvlb_frame_create with corners): three
 * launches on the context's tracking stream, chained with programmatic dependent launch.  Asynchronous, and up to
 * SDVLB_SEQ_DEPTH submissions may be in flight: the frames of one sequence are serially dependent (sdvl.cc:93,119,
 * 278-281), but that dependency lives in the sequence's device state, so frame k+1 can be queued while frame k is
 * being tracked and the host's turnaround leaves the critical path.  _poll returns 1 when the OLDEST submission in
 * flight has finished, _collect waits for it and fills its n results in submission order; _inflight returns how many
 * are outstanding.  A frame must stay alive until the sequence has tracked (not skipped) the next one. */
int sdvlb_seq_track_submit(sdvlb_ctx* ctx, sdvlb_seq* const* seqs, sdvlb_frame* const* frames, int n);
int sdvlb_seq_track_poll(sdvlb_ctx* ctx);
int sdvlb_seq_track_collect(sdvlb_ctx* ctx, sdvlb_seq_result* results);
int sdvlb_seq_track_inflight(sdvlb_ctx* ctx);

/* ---- depth-filter seeds: Map::UpdateCandidates (map.cc:397-498) -------------
 * The mapping thread's per-frame pass over its candidate points (SURVEY.md
 * section 8(f), row 2): visibility and baseline checks, Matcher::SearchPoint
 * along the epipolar segment of the current depth estimate (matcher.cc:45-121,
 * non-fixed branch), GetDepthFromTriangulation and GetParallax
 * (extra/utils.cc:193-213), and the Bayesian inverse-depth update Point::Update
 * / ComputeTau / PDFNormal / HasConverged (point.cc:63-100,162-216).  One warp
 * per candidate, all candidates of a frame in one launch. */
typedef struct sdvlb_seed {
  const sdvlb_frame* ref_frame; /* point->GetInitFeature()->GetFrame() (a keyframe) */
  double ref_T[7];              /* its pose */
  double ref_px[2];             /* init feature position, level-0 pixels */
  double ref_v[3];              /* init feature unit bearing */
  double rho, sigma2;           /* Point::rho_, sigma2_ (inverse depth and its variance) */
  double a, b, z_range;         /* Point::a_, b_, z_range_ (Beta inlier model, uniform outlier range) */
  double cos_alpha;             /* Point::cos_alpha_ */
  double last_distance;         /* Point::last_distance_ */
  double p3d[3];                /* out, SDVLB_SEED_CONVERGED: Point::p3d_ (the point becomes fixed) */
  double depth;                 /* out: triangulated depth of this frame's observation (when one was made) */
  double px[2];                 /* out: matched position in the current frame, level-0 pixels (when found) */
  int32_t ref_level;            /* init feature level */
  int32_t n_failed;             /* Point::n_failed_ */
  int32_t last_kf_id;           /* point->GetLastFeature()->GetFrame()->GetKeyframeID() */
  int32_t status;               /* out: SDVLB_SEED_* */
  int32_t level;                /* out: pyramid level the match was refined at (SearchPoint's *flevel, when found) */
  int32_t pad_;
} sdvlb_seed;

enum {
  SDVLB_SEED_NOT_VISIBLE = 0,   /* map.cc:428-436, kept */
  SDVLB_SEED_DELETE_OLD = 1,    /* not visible and last seen before min_kf_id: DeletePoint */
  SDVLB_SEED_SHORT_BASELINE = 2,/* map.cc:440-445 */
  SDVLB_SEED_NOT_FOUND = 3,     /* SearchPoint failed, Unpromote (n_failed, b updated) */
  SDVLB_SEED_DELETE_FAILED = 4, /* Unpromote returned true: DeletePoint */
  SDVLB_SEED_NO_DEPTH = 5,      /* GetDepthFromTriangulation failed */
  SDVLB_SEED_NO_PARALLAX = 6,   /* cos_alpha >= 0.999999 */
  SDVLB_SEED_TOO_CLOSE = 7,     /* map.cc:474-478 */
  SDVLB_SEED_UPDATED = 8,       /* Point::Update ran */
  SDVLB_SEED_CONVERGED = 9      /* ... and Point::HasConverged(): remove from the candidates */
};

typedef struct sdvlb_seed_params {
  double depth_mean;       /* frame->GetSceneDepth() */
  double map_scale;        /* Config::MapScale()      1.0  */
  double scale_min_dist;   /* Config::ScaleMinDist()  0.25 */
  int32_t min_kf_id;       /* last_kf->GetKeyframeID() - 2 * Config::MaxSearchKeyframes() */
  int32_t mode;            /* SDVLB_SEEDS_UPDATE or SDVLB_SEEDS_INIT */
} sdvlb_seed_params;
/* SDVLB_SEEDS_INIT: the per-corner part of Map::InitCandidates (map.cc:262-395) -- a corner of the new keyframe
 * (ref_frame; rho = 1/depth_mean, sigma2 = 1) is searched in a connected keyframe (`cur`), triangulated and screened
 * (parallax, minimum distance) exactly as above, but there is no visibility test and no filter update: status
 * SDVLB_SEED_UPDATED means "InitCandidate(feature, depth) may be called", with `depth` and `px` (imgpos) filled in.
 * The 1-px link test against cframe's features (map.cc:325-343) and the list handling stay with the caller. */
enum { SDVLB_SEEDS_UPDATE = 0, SDVLB_SEEDS_INIT = 1 };

/* The loop body of Map::UpdateCandidates for n candidates against `cur` (with corners) at pose T_cur.  Seeds are
 * updated in place; the caller applies the list surgery (erase / DeletePoint) that the statuses call for. */
int sdvlb_update_candidates(sdvlb_ctx* ctx, const sdvlb_frame* cur, const double T_cur[7], sdvlb_seed* seeds, int n,
                            const sdvlb_seed_params* sp);

#ifdef __cplusplus
}
#endif
#endif  /* SDVL_B200_H_ */
