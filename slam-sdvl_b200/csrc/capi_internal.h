// capi_internal.h — host-side objects behind the opaque handles of include/sdvl_b200.h, shared by the translation
// units that implement the C-ABI (capi.cu: contexts, frames, tracking batches; seq_api.cu: resident sequences).
#pragma once
#include <cstddef>
#include <cstdint>
#include <deque>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "seq.cuh"


inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// NVTX ranges around the entry points of the path (tracing): header-only NVTX v3, a couple of
// loads per call unless a tool (Nsight Systems / Compute) is attached, in which case the ranges appear on the calling
// thread's timeline: sdvlb.frames_submit, sdvlb.track_submit, sdvlb.track_collect, sdvlb.seq_track_submit, ...
#include <nvtx3/nvToolsExt.h>
struct SdvlbRange {
  explicit SdvlbRange(const char* name) { nvtxRangePushA(name); }
  ~SdvlbRange() { nvtxRangePop(); }
  SdvlbRange(const SdvlbRange&) = delete;
  SdvlbRange& operator=(const SdvlbRange&) = delete;
};
#define SDVLB_RANGE(name) SdvlbRange sdvlb_range_(name)

struct Arena {   // bump allocator over a pinned host buffer and (optionally) a device buffer with identical layout
  uint8_t* h = nullptr;
  uint8_t* d = nullptr;
  size_t cap = 0, used = 0;
  size_t take(size_t bytes) {
    const size_t off = align_up(used, 256);
    used = off + bytes;
    return off;
  }
};

struct TimerSlot { cudaEvent_t a, b; int kind; };

constexpr int kBuildEvents = 8;   // ring of "frame batch enqueued" events; a frame borrows the one of its batch
constexpr int kSlabFrames = 32;
constexpr int kStageSets = 4;     // staging sets for pageable host images

struct BatchOut {   // per-job results in the `out` arena
  double pose[7];
  double error;
  int32_t info[2];
  int32_t pad[2];
};


struct sdvlb_frame {
  sdvlb_ctx* ctx = nullptr;
  FrameDev dev{};                // host_mirror left null here; set per submission when a mirror is wanted
  uint8_t* d_block = nullptr;    // slot inside one of the context's device slabs
  size_t off_hdr = 0, off_pose = 0;   // corner header (count) + int4 corner list; pose
  uint8_t* h_pyr = nullptr;      // pinned host mirror of the pyramid, allocated on first sdvlb_frame_level()
  bool pyr_mirrored = false;
  uint8_t* h_corners = nullptr;  // pinned, device-visible: 16-byte header (count) + the first corner_copy corners
  std::vector<int32_t> h_more;   // whole corner list, only when it is longer than corner_copy
  cudaEvent_t built = nullptr;   // borrowed from the context's ring: recorded after the frame's batch was enqueued
  bool build_pending = false;    // submitted with sdvlb_frames_submit, completion not yet observed by the host
  bool build_corners = false, build_mirror = false;
  bool build_desc = false;       // ORB mode: the build also writes the corner descriptors (FrameDev::desc)
  bool has_corners = false;
  bool has_desc = false;
  bool overflowed = false;       // the frame's corner selection exceeded a capacity (its list is truncated)
  int corners_mirrored = -1;     // number of corners valid in the host mirror, -1 = not mirrored
  int n_corners = 0;
};

// Everything sdvlb_track_collect needs to finish a submission made by submit_batch.
struct PendingTrack {
  bool active = false;
  sdvlb_track_job* jobs = nullptr;
  int n = 0;
  bool build_frames = false;
  sdvlb_gn_iter* trace = nullptr;
  int trace_cap = 0;
  int* trace_n = nullptr;
  size_t o_res = 0, o_match = 0, o_trace = 0;
};

struct sdvlb_seq {   // host handle of a resident sequence
  sdvlb_ctx* ctx = nullptr;
  uint8_t* d_block = nullptr;        // SeqState followed by its arrays
  uint8_t* h_result = nullptr;       // pinned, device-visible: SDVLB_SEQ_DEPTH blocks of result_stride bytes
  size_t result_stride = 0;          // SeqResultHost + max_feats feature records, 256-byte aligned
  int max_feats = 0, n_cells = 0;
  int n_bound = 0;                   // upper bound of the device-side feature count at the next submission
  bool has_desc = false;             // created in ORB mode: the feature lists carry init-feature descriptors
  sdvlb_seq_policy policy = {};
  int kf_state[SDVLB_SEQ_KF_CAP] = {};   // 0 free, 1 points queued, 2 queued points submitted, 3 live
  uint32_t kf_seq[SDVLB_SEQ_KF_CAP] = {};   // submission that carried the slot's points to the device
};

struct SeqSubmission {   // one sdvlb_seq_track_submit in flight
  uint32_t seq_no = 0;
  int slot = 0;
  std::vector<sdvlb_seq*> seqs;
  std::vector<sdvlb_frame*> frames;
};

struct sdvlb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;    // tracking stream (ImageAlign, SearchPoint, synchronous frame construction)
  cudaStream_t bstream = nullptr;   // build stream (asynchronous frame batches: upload, pyramid, FAST)
  cudaEvent_t bevents[kBuildEvents] = {};
  int bevent_next = 0;
  cudaStream_t ustream = nullptr;   // upload stream (shared by the contexts of the device): level 0 of asynchronous
                                    // frame batches (PCIe), ahead of bstream
  std::mutex* umutex = nullptr;     // serialises launch + event record on the shared stream
  // Camera::UndistortImage: raw (distorted) level-0 images land in a ring of scratch sets, one per frame batch in
  // flight, and the undistortion kernel writes level 0 of the frame slots
  bool has_dist = false;
  UndistortArgs und = {};
  uint8_t* raw_scratch = nullptr;   // kBuildEvents sets x SDVLB_BATCH_MAX images
  size_t raw_stride = 0;            // bytes per image slot
  int raw_next = 0;
  cudaEvent_t raw_done[kBuildEvents] = {};
  bool raw_used[kBuildEvents] = {};
  // pageable host images: pinned, device-visible staging sets the calling thread copies into (one set per frame
  // batch in flight; a set is reused once the event of its last upload has fired)
  uint8_t* stage[4] = {};
  size_t stage_cap[4] = {};
  cudaEvent_t stage_done[4] = {};
  bool stage_used[4] = {};
  int stage_next = 0;
  bool use_orb = false;             // Config::UseORB(): ORB margins for FAST / FilterCorners, sdvlb_search_points_orb
  uint8_t* orb_buf = nullptr;       // device scratch of the ORB entry points (positions / descriptors / candidates)
  size_t orb_cap = 0;
  sdvlb_seed* d_seeds = nullptr;    // sdvlb_update_candidates staging (device + pinned host), grown on demand
  sdvlb_seed* h_seeds = nullptr;
  int seeds_cap = 0;
  cudaEvent_t uevents[kBuildEvents] = {};
  int uevent_next = 0;
  cudaEvent_t last_build = nullptr; // event of the most recent asynchronous build (null: none yet)
  cudaEvent_t track_build_done = nullptr;   // recorded after a tracking batch built its own frames (shared FAST scratch)
  bool track_build_pending = false;
  int32_t* h_overflow = nullptr;    // pinned, device-visible: [0] overflow flag written by the selector,
                                    // [16] sequence number of the last finished tracking submission (signal kernel)
  uint32_t track_seq = 0;
  PendingTrack pending;
  sdvlb_params params{};
  sdvlb_camera cam{};
  PyrGeom geom{};
  DevParams dp{};
  int w = 0, h = 0;
  int corner_cap = 0;
  int corner_copy = 0;           // corners the pinned host mirror of a frame holds
  std::vector<sdvlb_frame*> pool;      // free frames (device slot attached)
  std::vector<sdvlb_frame*> all_frames;
  std::vector<uint8_t*> slabs;         // device slabs of kSlabFrames frame slots each
  std::vector<uint8_t*> mirror_slabs;  // pinned host slabs: one corner mirror per frame slot
  size_t block_bytes = 0;
  std::vector<FastPlan> plans;   // one per nfeatures budget seen
  // FAST scratch (sized for `fast_frames` frames)
  int fast_frames = 0;
  uint32_t* cell_kp = nullptr;
  int32_t* cell_cnt = nullptr;
  int32_t* cell_kept = nullptr;   // survivors of the per-cell retainBest
  uint32_t* level_kp = nullptr;
  int32_t* level_cnt = nullptr;
  int32_t* frame_ticket = nullptr;
  size_t level_kp_total = 0;
  // staging
  Arena in;                      // host->device descriptors (pinned + device copy)
  Arena out;                     // results: pinned, device-visible host memory the kernels write directly
  uint8_t* scratch = nullptr;    // device-only scratch for ImageAlign caches
  size_t scratch_cap = 0;
  // resident sequences (seq_api.cu)
  Arena seq_in[SDVLB_SEQ_DEPTH];   // commands + new points (pinned + device copy), one staging per submission slot
  std::vector<SeqCmd> seq_cmds;  // queued by sdvlb_seq_reset / _add_points until the next submission
  std::vector<sdvlb_seq_point> seq_pts;
  std::vector<const sdvlb_frame*> seq_cmd_frames;
  uint32_t* d_seq_done = nullptr;
  std::deque<SeqSubmission> seq_queue;   // submissions in flight, oldest first (at most SDVLB_SEQ_DEPTH)
  std::vector<sdvlb_seq*> seqs;  // every sequence created on this context
  // counters
  int64_t n_launches = 0, h2d_bytes = 0, d2h_bytes = 0;
  // timing
  bool timing = false;
  std::vector<TimerSlot> timers;
  size_t timers_used = 0;
  cudaStream_t timer_stream = nullptr;
  double t_ms[SDVLB_K_COUNT] = {};
  int64_t t_launches[SDVLB_K_COUNT] = {};
};

// helpers implemented in capi.cu
namespace sdvlb_detail {
int ensure_arena(Arena* a, size_t bytes, bool need_device);
void timer_begin(sdvlb_ctx* c, int kind, cudaStream_t stream = nullptr);
void timer_end(sdvlb_ctx* c);
int wait_frame_built(sdvlb_ctx* c, const sdvlb_frame* f, cudaStream_t stream);   // orders `stream` after f's batch
void finalize_build(sdvlb_ctx* c, sdvlb_frame* f);
int check_overflow(sdvlb_ctx* c);
int wait_signal(sdvlb_ctx* c, uint32_t seq_no);   // spins on the pinned completion word until submission seq_no is done
}  // namespace sdvlb_detail
