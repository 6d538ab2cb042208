// capi.cu — implementation of the C-ABI in include/sdvl_b200.h: contexts, frame storage pool, pinned staging,
// batched submission of the K1..K4 kernels and per-kernel CUDA-event timing.  No CPU fallback: every entry point
// either runs the sm_100a kernels or fails with a negative status.
//
// Submission cost is part of the design (a tracked frame is a handful of short kernels): frame batches reach the build
// kernels BY VALUE in kernel parameter space (no descriptor upload), level 0 of a whole batch is uploaded by ONE kernel
// reading pinned host memory, and every result the host needs (corner lists, poses, matches, overflow flag) is written
// by the kernels straight into pinned, device-visible host memory, so the steady state issues no D2H copy at all.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include "common.cuh"

// ---- kernel launchers (other translation units)
cudaError_t sdvlb_launch_upload(const FrameBatch& B, const ImageBatch& I, int bytes, bool src_on_device, cudaStream_t stream);
cudaError_t sdvlb_launch_undistort(const FrameBatch& B, const ImageBatch& raw, const UndistortArgs& U, cudaStream_t stream);
cudaError_t sdvlb_launch_seed_update(sdvlb_seed* d_seeds, int n, const FrameDev& cur, const PyrGeom& g,
                                     const DevParams& dp, const sdvlb_seed_params& sp, cudaStream_t stream, bool orb);
cudaError_t sdvlb_launch_pyramid(const FrameBatch& B, const PyrGeom& g, cudaStream_t stream);
int sdvlb_pyramid_launches(const PyrGeom& g);
void sdvlb_fast_plan(const PyrGeom& g, const sdvlb_params& p, int nfeatures, int corner_cap, FastPlan* plan);
cudaError_t sdvlb_launch_orb_positions(const uint8_t* pyr, const PyrGeom& g, int levels, const int32_t* d_xyl, int n,
                                       uint8_t* d_desc, float* d_angle, cudaStream_t stream);
cudaError_t sdvlb_launch_orb_frames(const FrameBatch& B, const PyrGeom& g, int levels, int cap, cudaStream_t stream);
cudaError_t sdvlb_launch_fast_cells(const FrameBatch& B, const FastPlan& plan, uint32_t* cell_kp, int32_t* cell_cnt,
                                    cudaStream_t stream);
cudaError_t sdvlb_launch_fast_select(const FrameBatch& B, const FastPlan& plan, uint32_t* cell_kp, int32_t* cell_cnt,
                                     int32_t* cell_kept, uint32_t* level_kp, int32_t* level_cnt, int32_t* tickets,
                                     cudaStream_t stream);
cudaError_t sdvlb_launch_align(const void* d_jobs, int n_jobs, int n_bound, const PyrGeom& g, const DevParams& dp,
                               cudaStream_t stream);
cudaError_t sdvlb_launch_search(const SearchCandDev* d_cands, int n, const FrameDev* d_frames, sdvlb_match* d_out,
                                const PyrGeom& g, const DevParams& dp, cudaStream_t stream, const uint32_t* d_qdesc);
cudaError_t sdvlb_launch_signal(uint32_t* h_flag, uint32_t seq, cudaStream_t stream);

// ---- error reporting
static thread_local std::string g_last_error;
int sdvlb_set_cuda_error(cudaError_t e, const char* expr, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", int(e), cudaGetErrorString(e), file, line, expr);
  g_last_error = buf;
  return SDVLB_ERR_CUDA;
}
int sdvlb_set_error(int code, const char* msg) {
  g_last_error = msg;
  return code;
}

#include "capi_internal.h"

// Per (device, kernel) launch preparation (declared in common.cuh).  A thread-local record answers the steady state
// without a lock; the first call per thread / a larger request goes through the process-wide table.
cudaError_t sdvlb_kernel_prepare_ptr(const void* kernel, int dyn_smem_bytes) {
  struct Rec { const void* kernel; int device; int dyn; };
  static std::mutex mu;
  static std::vector<Rec> table;
  static int carveout = -2;
  thread_local std::vector<Rec> seen;
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  for (const Rec& r : seen)
    if (r.kernel == kernel && r.device == device && r.dyn >= dyn_smem_bytes) return cudaSuccess;
  std::lock_guard<std::mutex> lk(mu);
  if (carveout == -2) {
    const char* env = getenv("SDVLB_CARVEOUT");
    carveout = env ? atoi(env) : 100;
  }
  Rec* rec = nullptr;
  for (Rec& r : table)
    if (r.kernel == kernel && r.device == device) rec = &r;
  if (!rec) {
    if (carveout >= 0) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
      if (e != cudaSuccess) return e;
    }
    table.push_back(Rec{kernel, device, 0});
    rec = &table.back();
  }
  if (dyn_smem_bytes > 48 * 1024 && dyn_smem_bytes > rec->dyn) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem_bytes);
    if (e != cudaSuccess) return e;
  }
  if (dyn_smem_bytes > rec->dyn) rec->dyn = dyn_smem_bytes;
  bool found = false;
  for (Rec& r : seen)
    if (r.kernel == kernel && r.device == device) { r.dyn = rec->dyn; found = true; }
  if (!found) seen.push_back(*rec);
  return cudaSuccess;
}


namespace sdvlb_detail {

int ensure_arena(Arena* a, size_t bytes, bool need_device) {
  if (bytes <= a->cap) return 0;
  const size_t ncap = align_up(std::max(bytes, a->cap * 2), 1 << 20);
  if (a->h) cudaFreeHost(a->h);
  if (a->d) cudaFree(a->d);
  a->h = nullptr; a->d = nullptr; a->cap = 0;
  SDVLB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&a->h), ncap, cudaHostAllocDefault));
  if (need_device) SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&a->d), ncap));
  a->cap = ncap;
  return 0;
}

int ensure_scratch(sdvlb_ctx* c, size_t bytes) {
  if (bytes <= c->scratch_cap) return 0;
  const size_t ncap = align_up(std::max(bytes, c->scratch_cap * 2), 1 << 20);
  if (c->scratch) cudaFree(c->scratch);
  c->scratch = nullptr; c->scratch_cap = 0;
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->scratch), ncap));
  c->scratch_cap = ncap;
  return 0;
}

// Growing the FAST scratch frees device memory that work in flight on either stream may still use: drain first.
int ensure_fast_scratch(sdvlb_ctx* c, int n_frames) {
  if (n_frames <= c->fast_frames) return 0;
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->bstream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const int nf = std::max(n_frames, c->fast_frames * 2);
  cudaFree(c->cell_kp); cudaFree(c->cell_cnt); cudaFree(c->cell_kept); cudaFree(c->level_kp); cudaFree(c->level_cnt);
  cudaFree(c->frame_ticket);
  c->cell_kp = nullptr; c->cell_cnt = nullptr; c->cell_kept = nullptr; c->level_kp = nullptr; c->level_cnt = nullptr;
  c->frame_ticket = nullptr;
  c->fast_frames = 0;
  const size_t cells = size_t(c->geom.total_cells) * nf;
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->cell_kp), cells * SDVLB_CELL_CAP * sizeof(uint32_t)));
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->cell_cnt), cells * sizeof(int32_t)));
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->cell_kept), cells * sizeof(int32_t)));
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->level_kp), c->level_kp_total * nf * sizeof(uint32_t)));
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->level_cnt), size_t(nf) * SDVLB_MAX_LEVELS * sizeof(int32_t)));
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->frame_ticket), size_t(nf) * SDVLB_TICKET_STRIDE * sizeof(int32_t)));
  SDVLB_CUDA_TRY(cudaMemset(c->frame_ticket, 0, size_t(nf) * SDVLB_TICKET_STRIDE * sizeof(int32_t)));
  c->fast_frames = nf;
  return 0;
}

inline size_t corner_mirror_bytes(const sdvlb_ctx* c) { return align_up(16 + size_t(c->corner_copy) * sizeof(int4), 256); }

// One more slab: kSlabFrames device slots (one cudaMalloc) + their pinned corner mirrors (one cudaHostAlloc).
int grow_pool(sdvlb_ctx* c) {
  size_t off = align_up(size_t(c->geom.total) + 256, 256);
  const size_t off_hdr = off;   off = align_up(off + 16 + size_t(c->corner_cap) * sizeof(int4), 256);
  const size_t off_pose = off;  off = align_up(off + 7 * sizeof(double), 256);
  const size_t gcells = size_t(c->geom.wcells[0]) * c->geom.hcells[0];
  const size_t off_grid = off;  off = align_up(off + (2 * gcells + 1 + size_t(c->corner_cap)) * sizeof(int32_t), 256);
  // Config::UseORB(): Frame::descriptors_, 32 bytes per corner.  Slots carry the region when the context was in ORB mode
  // when the slab was allocated (sdvlb_ctx_set_orb regrows the pool, which is why it wants no live frames).
  const bool with_desc = c->use_orb;
  const size_t off_desc = off;
  if (with_desc) off = align_up(off + size_t(c->corner_cap) * 32, 256);
  c->block_bytes = off;
  uint8_t* slab = nullptr;
  uint8_t* mslab = nullptr;
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&slab), c->block_bytes * kSlabFrames));
  c->slabs.push_back(slab);
  SDVLB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&mslab), corner_mirror_bytes(c) * kSlabFrames, cudaHostAllocDefault));
  c->mirror_slabs.push_back(mslab);
  for (int i = kSlabFrames - 1; i >= 0; i--) {
    sdvlb_frame* f = new sdvlb_frame;
    f->ctx = c;
    f->d_block = slab + size_t(i) * c->block_bytes;
    f->h_corners = mslab + size_t(i) * corner_mirror_bytes(c);
    f->off_hdr = off_hdr; f->off_pose = off_pose;
    f->dev.pyr = f->d_block;
    f->dev.n_corners = reinterpret_cast<int32_t*>(f->d_block + off_hdr);
    f->dev.corners = reinterpret_cast<int4*>(f->d_block + off_hdr + 16);
    f->dev.pose = reinterpret_cast<double*>(f->d_block + off_pose);
    f->dev.grid = reinterpret_cast<int32_t*>(f->d_block + off_grid);
    f->dev.desc = with_desc ? reinterpret_cast<uint32_t*>(f->d_block + off_desc) : nullptr;
    f->dev.host_mirror = nullptr;
    f->dev.mirror_cap = c->corner_copy;
    c->pool.push_back(f);
    c->all_frames.push_back(f);
  }
  return 0;
}

// Frames are carved out of slabs and recycled through a free list, so the steady state of a tracker performs no CUDA
// allocation at all (sdvlb_ctx_reserve_frames pre-sizes the pool).
int frame_alloc(sdvlb_ctx* c, sdvlb_frame** out) {
  if (c->pool.empty()) {
    const int rc = grow_pool(c);
    if (rc) return rc;
  }
  sdvlb_frame* f = c->pool.back();
  c->pool.pop_back();
  f->has_corners = false; f->pyr_mirrored = false; f->corners_mirrored = -1; f->n_corners = 0;
  f->build_pending = false; f->build_corners = false; f->build_mirror = false;
  f->build_desc = false; f->has_desc = false; f->overflowed = false;
  f->built = nullptr;
  f->h_more.clear();
  *out = f;
  return 0;
}

void frame_release(sdvlb_ctx* c, sdvlb_frame* f) { c->pool.push_back(f); }

// ---- timing helpers
void timer_begin(sdvlb_ctx* c, int kind, cudaStream_t stream) {
  if (!c->timing) return;
  if (!stream) stream = c->stream;
  c->timer_stream = stream;
  if (c->timers_used == c->timers.size()) {
    TimerSlot s;
    cudaEventCreate(&s.a);
    cudaEventCreate(&s.b);
    c->timers.push_back(s);
  }
  TimerSlot& s = c->timers[c->timers_used];
  s.kind = kind;
  cudaEventRecord(s.a, stream);
}
void timer_end(sdvlb_ctx* c) {
  if (!c->timing) return;
  cudaEventRecord(c->timers[c->timers_used].b, c->timer_stream);
  c->timers_used++;
}
void timer_collect(sdvlb_ctx* c) {   // streams must be idle
  for (size_t i = 0; i < c->timers_used; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->timers[i].a, c->timers[i].b) == cudaSuccess) {
      c->t_ms[c->timers[i].kind] += ms;
      c->t_launches[c->timers[i].kind] += 1;
    }
  }
  c->timers_used = 0;
}

void build_geom(const sdvlb_params& p, int w, int h, PyrGeom* g) {
  memset(g, 0, sizeof(*g));
  g->levels = p.pyramid_levels;
  int cw = w, ch = h;
  size_t off = 0;
  for (int l = 0; l < p.pyramid_levels; l++) {
    g->w[l] = cw; g->h[l] = ch;
    g->off[l] = int(off);
    off = align_up(off + size_t(cw) * ch, 256);
    cw /= 2; ch /= 2;   // frame.cc:119
  }
  g->total = int(off);
  int cells = 0;
  for (int l = 0; l < p.max_fast_levels; l++) {
    g->wcells[l] = int(std::ceil(double(g->w[l]) / double(p.cell_size)));   // fast_detector.cc:72-73
    g->hcells[l] = int(std::ceil(double(g->h[l]) / double(p.cell_size)));
    g->cell_off[l] = cells;
    cells += g->wcells[l] * g->hcells[l];
  }
  g->total_cells = cells;
}

const FastPlan* get_plan(sdvlb_ctx* c, int nfeatures) {
  for (const FastPlan& p : c->plans)
    if (p.nfeatures == nfeatures) return &p;
  c->plans.reserve(16);
  FastPlan p;
  sdvlb_fast_plan(c->geom, c->params, nfeatures, c->corner_cap, &p);
  if (c->use_orb) p.args.margin = 4 + 31 / 2;   // Config::UseORB(): fast_detector.cc:63-64
  p.args.overflow_flag = c->h_overflow;
  c->plans.push_back(p);
  return &c->plans.back();
}

int check_overflow(sdvlb_ctx* c) {   // the stream that ran the selector must have been synchronised
  volatile int32_t* flag = c->h_overflow;
  if (*flag) {
    *flag = 0;
    return sdvlb_set_error(SDVLB_ERR_OVERFLOW, "corner capacity exceeded in FAST selection");
  }
  return 0;
}

// Host bookkeeping once a frame's build commands are known to have completed.
void finalize_build(sdvlb_ctx* c, sdvlb_frame* f) {
  f->has_corners = f->build_corners;
  f->has_desc = f->build_corners && f->build_desc;
  f->corners_mirrored = -1;
  if (f->build_corners && f->build_mirror) {
    memcpy(&f->n_corners, f->h_corners, sizeof(int32_t));
    int32_t ovf = 0;
    memcpy(&ovf, f->h_corners + sizeof(int32_t), sizeof(int32_t));   // header word 1: the frame's own overflow flag
    f->overflowed = ovf != 0;
    f->corners_mirrored = std::min(f->n_corners, c->corner_copy);
    c->d2h_bytes += 16 + int64_t(f->corners_mirrored) * 16;
  }
  f->build_pending = false;
}

int wait_frame_built(sdvlb_ctx* c, const sdvlb_frame* f, cudaStream_t stream) {
  if (f->build_pending) SDVLB_CUDA_TRY(cudaStreamWaitEvent(stream, f->built, 0));
  return 0;
}

int ensure_built(sdvlb_frame* f) {
  if (!f->build_pending) return 0;
  SDVLB_CUDA_TRY(cudaSetDevice(f->ctx->device));
  SDVLB_CUDA_TRY(cudaEventSynchronize(f->built));
  finalize_build(f->ctx, f);
  if (f->overflowed) return sdvlb_set_error(SDVLB_ERR_OVERFLOW, "corner capacity exceeded in this frame's FAST selection");
  return check_overflow(f->ctx);
}

// Upload streams are shared by all contexts of a device: the level-0 uploads of every group then run one after the
// other (round-robin over SDVLB_UPLOAD_STREAMS streams, default 4, so that the tail of one overlaps the head of the
// next) and what is outstanding on PCIe stays bounded by the upload kernel's own depth instead of growing with the
// number of contexts.
cudaError_t shared_upload_stream(int device, int prio, cudaStream_t* out, std::mutex** mu) {
  static std::mutex g_mu;
  static std::mutex g_stream_mu[16][4];
  static cudaStream_t g_streams[16][4] = {};
  static int g_next[16] = {};
  static int n_streams = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (n_streams == 0) {
    const char* e = getenv("SDVLB_UPLOAD_STREAMS");
    n_streams = e ? atoi(e) : 4;
    if (n_streams < 1 || n_streams > 4) n_streams = 4;
  }
  if (device < 0 || device >= 16) return cudaErrorInvalidDevice;
  const int k = g_next[device]++ % n_streams;
  if (!g_streams[device][k]) {
    const cudaError_t e = cudaStreamCreateWithPriority(&g_streams[device][k], cudaStreamNonBlocking, prio);
    if (e != cudaSuccess) return e;
  }
  *out = g_streams[device][k];
  *mu = &g_stream_mu[device][k];
  return cudaSuccess;
}

// Copy into a pinned staging slot (256-byte aligned) with non-temporal stores: the destination is read next by the GPU
// over PCIe, never by this core, so it should neither be fetched for ownership nor displace the cache.
static void stream_copy(uint8_t* dst, const uint8_t* src, size_t bytes) {
#if defined(__SSE2__)
  size_t i = 0;
  for (; i + 64 <= bytes; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b2);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c2);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  if (i < bytes) memcpy(dst + i, src + i, bytes - i);
  _mm_sfence();
#else
  memcpy(dst, src, bytes);
#endif
}

// Enqueues level-0 upload + pyramid (+ FAST + selection) for n frames on `stream`, in chunks of SDVLB_BATCH_MAX.
// image_loc: 0 host memory of any kind (one cudaMemcpyAsync per frame), 1 device memory, 2 pinned device-visible host
// memory (both uploaded by one kernel per chunk).
// ustream (optional): the level-0 upload goes on that stream instead and `stream` waits for it, so that the PCIe
// transfer of a later batch overlaps the pyramid / FAST kernels of an earlier one of the same context.
int enqueue_build(sdvlb_ctx* c, sdvlb_frame* const* frames, const uint8_t* const* images, const int32_t* image_loc,
                  int n, bool want_corners, int nfeatures, bool mirror, cudaStream_t stream,
                  cudaStream_t ustream = nullptr) {
  if (!ustream) ustream = stream;
  const size_t img_bytes = size_t(c->w) * c->h;
  const FastPlan* plan = want_corners ? get_plan(c, nfeatures) : nullptr;
  for (int base = 0; base < n; base += SDVLB_BATCH_MAX) {
    const int m = std::min(SDVLB_BATCH_MAX, n - base);
    FrameBatch B;
    ImageBatch I;
    B.n = m;
    B.scratch_base = base;
    bool by_kernel = true, all_device = true;
    for (int i = 0; i < m; i++) {
      if (image_loc[base + i] != 1) all_device = false;
      sdvlb_frame* f = frames[base + i];
      B.f[i] = f->dev;
      B.f[i].host_mirror = (want_corners && mirror) ? reinterpret_cast<int32_t*>(f->h_corners) : nullptr;
      I.src[i] = images[base + i];
      if (image_loc[base + i] == 0) by_kernel = false;
      f->build_corners = want_corners;
      f->build_mirror = want_corners && mirror;
      f->build_desc = want_corners && c->use_orb && f->dev.desc != nullptr;
    }
    // Camera::UndistortImage: the raw images go to a scratch set, the undistortion kernel fills level 0
    FrameBatch Bup = B;
    ImageBatch Iraw;
    int raw_set = -1;
    if (c->has_dist) {
      if (!c->raw_scratch) {
        c->raw_stride = (img_bytes + 255) & ~size_t(255);
        SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->raw_scratch), c->raw_stride * SDVLB_BATCH_MAX * kBuildEvents));
        for (int i = 0; i < kBuildEvents; i++)
          SDVLB_CUDA_TRY(cudaEventCreateWithFlags(&c->raw_done[i], cudaEventDisableTiming));
      }
      raw_set = c->raw_next;
      c->raw_next = (c->raw_next + 1) % kBuildEvents;
      if (c->raw_used[raw_set]) SDVLB_CUDA_TRY(cudaStreamWaitEvent(ustream, c->raw_done[raw_set], 0));
      for (int i = 0; i < m; i++) {
        uint8_t* slot = c->raw_scratch + (size_t(raw_set) * SDVLB_BATCH_MAX + i) * c->raw_stride;
        Bup.f[i].pyr = slot;
        Iraw.src[i] = slot;
      }
    }
    // Pageable host images (the reference hands cv::Mat data to Frame::Frame): copied by the calling thread into a
    // pinned, device-visible staging set of the context and then uploaded like pinned images, by the same kernel.
    // (cudaMemcpyAsync from pageable memory is staged by the driver one frame at a time, synchronously: 9.6 GB/s for
    // the whole process however many threads submit, profiles/r02_a_bench_C2.json e2e_pageable.)
    int stage_set = -1;
    if (!by_kernel && !getenv("SDVLB_PAGEABLE_MEMCPY")) {
      stage_set = c->stage_next;
      c->stage_next = (c->stage_next + 1) % kStageSets;
      const size_t slot_bytes = (img_bytes + 255) & ~size_t(255);
      if (c->stage_cap[stage_set] < size_t(m) * slot_bytes) {
        if (c->stage_used[stage_set]) SDVLB_CUDA_TRY(cudaEventSynchronize(c->stage_done[stage_set]));
        if (c->stage[stage_set]) cudaFreeHost(c->stage[stage_set]);
        c->stage[stage_set] = nullptr; c->stage_cap[stage_set] = 0;
        SDVLB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->stage[stage_set]), size_t(m) * slot_bytes, cudaHostAllocDefault));
        c->stage_cap[stage_set] = size_t(m) * slot_bytes;
        if (!c->stage_done[stage_set]) SDVLB_CUDA_TRY(cudaEventCreateWithFlags(&c->stage_done[stage_set], cudaEventDisableTiming));
        c->stage_used[stage_set] = false;
      }
      if (c->stage_used[stage_set]) SDVLB_CUDA_TRY(cudaEventSynchronize(c->stage_done[stage_set]));   // its last upload has read it
      for (int i = 0; i < m; i++) {
        if (image_loc[base + i] != 0) continue;
        uint8_t* dst = c->stage[stage_set] + size_t(i) * slot_bytes;
        stream_copy(dst, I.src[i], img_bytes);
        I.src[i] = dst;
      }
      by_kernel = true;
      all_device = false;
    }
    // launch + event record are one unit on the shared upload stream
    std::unique_lock<std::mutex> ulock;
    if (ustream != stream && c->umutex) ulock = std::unique_lock<std::mutex>(*c->umutex);
    if (by_kernel) {
      SDVLB_CUDA_TRY(sdvlb_launch_upload(Bup, I, int(img_bytes), all_device, ustream));
      c->n_launches += 1;
      if (stage_set >= 0) {
        SDVLB_CUDA_TRY(cudaEventRecord(c->stage_done[stage_set], ustream));
        c->stage_used[stage_set] = true;
      }
    } else {
      // copy engine into level 0 of the frame slots; with distortion set the raw pixels then move on to the scratch
      // set through the device-to-device path of the upload kernel (level 0 is rewritten by the undistortion)
      for (int i = 0; i < m; i++)
        SDVLB_CUDA_TRY(cudaMemcpyAsync(B.f[i].pyr, I.src[i], img_bytes,
                                       image_loc[base + i] == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ustream));
      if (c->has_dist) {
        ImageBatch I2;
        for (int i = 0; i < m; i++) I2.src[i] = B.f[i].pyr;
        SDVLB_CUDA_TRY(sdvlb_launch_upload(Bup, I2, int(img_bytes), true, ustream));
        c->n_launches += 1;
      }
    }
    if (ustream != stream) {
      cudaEvent_t ev = c->uevents[c->uevent_next];
      c->uevent_next = (c->uevent_next + 1) % kBuildEvents;
      SDVLB_CUDA_TRY(cudaEventRecord(ev, ustream));
      if (ulock.owns_lock()) ulock.unlock();
      SDVLB_CUDA_TRY(cudaStreamWaitEvent(stream, ev, 0));
    }
    for (int i = 0; i < m; i++)
      if (image_loc[base + i] != 1) c->h2d_bytes += int64_t(img_bytes);
    if (c->has_dist) {
      SDVLB_CUDA_TRY(sdvlb_launch_undistort(B, Iraw, c->und, stream));
      SDVLB_CUDA_TRY(cudaEventRecord(c->raw_done[raw_set], stream));
      c->raw_used[raw_set] = true;
      c->n_launches += 1;
    }
    timer_begin(c, SDVLB_K_PYRAMID, stream);
    SDVLB_CUDA_TRY(sdvlb_launch_pyramid(B, c->geom, stream));
    timer_end(c);
    c->n_launches += sdvlb_pyramid_launches(c->geom);
    if (want_corners) {
      timer_begin(c, SDVLB_K_FAST, stream);
      SDVLB_CUDA_TRY(sdvlb_launch_fast_cells(B, *plan, c->cell_kp, c->cell_cnt, stream));
      timer_end(c);
      timer_begin(c, SDVLB_K_SELECT, stream);
      SDVLB_CUDA_TRY(sdvlb_launch_fast_select(B, *plan, c->cell_kp, c->cell_cnt, c->cell_kept, c->level_kp, c->level_cnt,
                                              c->frame_ticket, stream));
      timer_end(c);
      c->n_launches += 2;
      if (c->use_orb) {   // Frame::descriptors_ for every corner, once (the reference fills them lazily)
        timer_begin(c, SDVLB_K_ORB, stream);
        SDVLB_CUDA_TRY(sdvlb_launch_orb_frames(B, c->geom, c->params.pyramid_levels, c->corner_cap, stream));
        timer_end(c);
        c->n_launches += 1;
      }
    }
  }
  return 0;
}

// Submits (upload + pyramid + FAST for jobs that carry an image) + align + search for n jobs on the tracking stream;
// does not wait.  collect_batch() synchronises and copies the results into the jobs.
int submit_batch(sdvlb_ctx* c, sdvlb_track_job* jobs, int n, int mirror, sdvlb_gn_iter* trace, int trace_cap,
                 int* trace_n, const sdvlb_gn_forced* forced, bool build_frames, int fast = 0) {
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  if (c->pending.active || !c->seq_queue.empty()) return sdvlb_set_error(SDVLB_ERR_STATE, "a submission is still in flight on this context");
  const PyrGeom& g = c->geom;

  int n_detect = 0, n_align = 0, n_cands = 0, n_feats = 0, nfeatures = -1, max_feats_job = 0;
  int n_with_desc = 0, n_with_cands = 0;
  for (int i = 0; i < n; i++) {
    if (jobs[i].n_cands > 0) { n_with_cands++; if (jobs[i].cand_desc) n_with_desc++; }
    if (jobs[i].ref) max_feats_job = std::max(max_feats_job, jobs[i].n_feats);
    if (build_frames && jobs[i].want_corners) {
      n_detect++;
      if (nfeatures < 0) nfeatures = jobs[i].nfeatures;
      else if (nfeatures != jobs[i].nfeatures) return sdvlb_set_error(SDVLB_ERR_ARG, "mixed nfeatures in one batch");
    }
    if (jobs[i].ref) { n_align++; n_feats += jobs[i].n_feats; }
    if (jobs[i].n_cands < 0 || jobs[i].n_feats < 0) return sdvlb_set_error(SDVLB_ERR_ARG, "negative count");
    n_cands += jobs[i].n_cands;
  }
  if (build_frames && n_detect != 0 && n_detect != n)
    return sdvlb_set_error(SDVLB_ERR_ARG, "mixed want_corners in one batch");
  // Config::UseORB(): candidates come with the descriptors of their init features and are scored against the corner
  // descriptors of the current frames (matcher.cc:243-277)
  const bool orb_search = n_with_desc > 0;
  if (orb_search) {
    if (n_with_desc != n_with_cands) return sdvlb_set_error(SDVLB_ERR_ARG, "cand_desc on some jobs of a batch only");
    if (!c->use_orb) return sdvlb_set_error(SDVLB_ERR_STATE, "candidate descriptors need sdvlb_ctx_set_orb(ctx, 1)");
    for (int i = 0; i < n; i++)
      if (jobs[i].n_cands > 0 && !build_frames && !(jobs[i].cur && (jobs[i].cur->has_desc || (jobs[i].cur->build_pending && jobs[i].cur->build_desc))))
        return sdvlb_set_error(SDVLB_ERR_STATE, "the current frame was built outside ORB mode: it has no corner descriptors");
  } else if (c->use_orb && n_with_cands > 0) {
    return sdvlb_set_error(SDVLB_ERR_ARG, "ORB mode: candidates need cand_desc (feature->GetDescriptor())");
  }

  // ---- frames
  if (build_frames) {
    if (n_detect > 0) {
      const int rc = ensure_fast_scratch(c, n);
      if (rc) return rc;
    }
    for (int i = 0; i < n; i++) {
      sdvlb_frame* f = nullptr;
      const int rc = frame_alloc(c, &f);
      if (rc) { for (int k = 0; k < i; k++) { frame_release(c, jobs[k].cur); jobs[k].cur = nullptr; } return rc; }
      jobs[i].cur = f;
    }
  }

  // ---- stage inputs
  Arena& in = c->in;
  Arena& out = c->out;
  in.used = 0; out.used = 0;
  const size_t need_in = 4096 + size_t(n) * (sizeof(FrameDev) + sizeof(AlignJobDev) + 512) +
                         size_t(n_feats) * sizeof(sdvlb_align_feat) + size_t(n_cands) * sizeof(SearchCandDev) +
                         (forced ? size_t(forced->n_total) * 56 + 512 : 0) + 32 * 256 + (orb_search ? size_t(n_cands) * 32 + 256 : 0);
  const size_t need_out = 4096 + size_t(n) * sizeof(BatchOut) + size_t(n_cands) * sizeof(sdvlb_match) +
                          size_t(trace ? trace_cap : 0) * sizeof(sdvlb_gn_iter) + 16 * 256;
  int rc = ensure_arena(&in, need_in, true);
  if (rc) return rc;
  rc = ensure_arena(&out, need_out, false);
  if (rc) return rc;
  size_t scratch_need = 0;
  for (int i = 0; i < n; i++)
    if (jobs[i].ref) scratch_need += align_up(SDVLB_ALIGN_SC_BYTES(jobs[i].n_feats) + 1024, 256);
  rc = ensure_scratch(c, scratch_need);
  if (rc) return rc;

  const size_t o_frames = in.take(size_t(n) * sizeof(FrameDev));
  const size_t o_align = in.take(size_t(std::max(n_align, 1)) * sizeof(AlignJobDev));
  const size_t o_feats = in.take(size_t(std::max(n_feats, 1)) * sizeof(sdvlb_align_feat));
  const size_t o_cands = in.take(size_t(std::max(n_cands, 1)) * sizeof(SearchCandDev));
  const size_t o_qdesc = orb_search ? in.take(size_t(std::max(n_cands, 1)) * 32) : 0;
  size_t o_forced_T = 0, o_forced_it = 0;
  if (forced) {
    o_forced_T = in.take(size_t(forced->n_total) * 7 * sizeof(double));
    o_forced_it = in.take(SDVLB_MAX_LEVELS * sizeof(int32_t));
    memcpy(in.h + o_forced_T, forced->T, size_t(forced->n_total) * 7 * sizeof(double));
    memset(in.h + o_forced_it, 0, SDVLB_MAX_LEVELS * sizeof(int32_t));
    memcpy(in.h + o_forced_it, forced->iters, size_t(c->params.pyramid_levels) * sizeof(int32_t));
  }
  const size_t o_prior = in.take(size_t(n) * 7 * sizeof(double));
  for (int i = 0; i < n; i++) memcpy(in.h + o_prior + size_t(i) * 56, jobs[i].T_cur, 56);
  const size_t o_res = out.take(size_t(n) * sizeof(BatchOut));
  const size_t o_match = out.take(size_t(std::max(n_cands, 1)) * sizeof(sdvlb_match));
  const size_t o_trace = trace ? out.take(size_t(trace_cap) * sizeof(sdvlb_gn_iter)) : 0;

  FrameDev* hf = reinterpret_cast<FrameDev*>(in.h + o_frames);
  AlignJobDev* ha = reinterpret_cast<AlignJobDev*>(in.h + o_align);
  sdvlb_align_feat* hfe = reinterpret_cast<sdvlb_align_feat*>(in.h + o_feats);
  SearchCandDev* hc = reinterpret_cast<SearchCandDev*>(in.h + o_cands);
  int ai = 0, fi = 0, cidx = 0;
  size_t sc_off = 0;
  for (int i = 0; i < n; i++) {
    sdvlb_track_job& j = jobs[i];
    if (!j.cur) return sdvlb_set_error(SDVLB_ERR_ARG, "job without current frame");
    hf[i] = j.cur->dev;
    if (j.ref) {
      AlignJobDev& a = ha[ai];
      memset(&a, 0, sizeof(a));
      a.ref = j.ref->dev;
      a.cur = j.cur->dev;
      a.feats = reinterpret_cast<const sdvlb_align_feat*>(in.d + o_feats) + fi;
      a.n = j.n_feats;
      a.fast = fast;
      memcpy(a.T_ref, j.T_ref, sizeof(a.T_ref));
      memcpy(a.T_cur, j.T_cur, sizeof(a.T_cur));
      BatchOut* hres = reinterpret_cast<BatchOut*>(out.h + o_res) + i;   // pinned host memory, written by the kernel
      a.out_pose = hres->pose;
      a.out_info = hres->info;
      a.out_error = &hres->error;
      a.out_cycles = nullptr;
      a.trace = (trace && ai == 0) ? reinterpret_cast<sdvlb_gn_iter*>(out.h + o_trace) : nullptr;
      a.trace_cap = trace ? trace_cap : 0;
      if (forced && ai == 0) {
        a.forced_T = reinterpret_cast<const double*>(in.d + o_forced_T);
        a.forced_iters = reinterpret_cast<const int32_t*>(in.d + o_forced_it);
        a.forced_n = forced->n_total;
      }
      uint8_t* sc = c->scratch + sc_off;
      a.sc_d = reinterpret_cast<double*>(sc);
      a.sc_f = reinterpret_cast<float*>(sc + align_up(size_t(j.n_feats) * SDVLB_ALIGN_SC_DOUBLES * 8, 256));
      a.sc_flags = reinterpret_cast<int32_t*>(sc + align_up(size_t(j.n_feats) * SDVLB_ALIGN_SC_DOUBLES * 8, 256) +
                                              align_up(size_t(j.n_feats) * SDVLB_ALIGN_SC_FLOATS * 4, 256));
      sc_off += align_up(SDVLB_ALIGN_SC_BYTES(j.n_feats) + 1024, 256);
      if (j.n_feats > 0) memcpy(hfe + fi, j.feats, size_t(j.n_feats) * sizeof(sdvlb_align_feat));
      fi += j.n_feats;
      ai++;
    }
    if (orb_search && j.n_cands > 0) memcpy(in.h + o_qdesc + size_t(cidx) * 32, j.cand_desc, size_t(j.n_cands) * 32);
    for (int k = 0; k < j.n_cands; k++) {
      const sdvlb_candidate& s = j.cands[k];
      SearchCandDev& d = hc[cidx++];
      if (!s.ref_frame) return sdvlb_set_error(SDVLB_ERR_ARG, "candidate without reference frame");
      d.ref_pyr = s.ref_frame->dev.pyr;
      memcpy(d.ref_T, s.ref_T, sizeof(d.ref_T));
      d.ref_px[0] = s.ref_px[0]; d.ref_px[1] = s.ref_px[1];
      d.ref_v[0] = s.ref_v[0]; d.ref_v[1] = s.ref_v[1]; d.ref_v[2] = s.ref_v[2];
      d.idepth = s.idepth; d.idepth_std = s.idepth_std;
      d.px[0] = s.px[0]; d.px[1] = s.px[1];
      d.pos[0] = s.pos[0]; d.pos[1] = s.pos[1]; d.pos[2] = s.pos[2];
      d.ref_level = s.ref_level;
      d.flags = s.flags;
      d.cur_index = i;
      d.pad_ = 0;
      if (s.ref_level < 0 || s.ref_level >= c->params.pyramid_levels)
        return sdvlb_set_error(SDVLB_ERR_ARG, "candidate level out of range");
    }
  }

  // ---- ordering against asynchronous frame batches: frames still being built (wait for THEIR batch only, not for
  // batches enqueued later as prefetch), and the FAST scratch both streams share when this call builds frames itself
  {
    cudaEvent_t waited[4] = {nullptr, nullptr, nullptr, nullptr};
    int nw = 0;
    auto wait_for = [&](cudaEvent_t ev) -> cudaError_t {
      for (int k = 0; k < nw; k++) if (waited[k] == ev) return cudaSuccess;
      if (nw < 4) waited[nw++] = ev;
      return cudaStreamWaitEvent(c->stream, ev, 0);
    };
    if (build_frames && c->last_build) SDVLB_CUDA_TRY(wait_for(c->last_build));
    for (int i = 0; i < n; i++) {
      if (!build_frames && jobs[i].cur->build_pending) SDVLB_CUDA_TRY(wait_for(jobs[i].cur->built));
      if (jobs[i].ref && jobs[i].ref->build_pending) SDVLB_CUDA_TRY(wait_for(jobs[i].ref->built));
    }
  }

  // ---- H2D of the descriptors (one copy)
  SDVLB_CUDA_TRY(cudaMemcpyAsync(in.d, in.h, in.used, cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes += int64_t(in.used);
  // prior poses for frames that are not aligned (search may still read cur.pose)
  for (int i = 0; i < n; i++)
    if (!jobs[i].ref && jobs[i].n_cands > 0)
      SDVLB_CUDA_TRY(cudaMemcpyAsync(jobs[i].cur->dev.pose, in.d + o_prior + size_t(i) * 56, 56,
                                     cudaMemcpyDeviceToDevice, c->stream));

  // ---- kernels
  if (build_frames) {
    std::vector<sdvlb_frame*> fr(n);
    std::vector<const uint8_t*> img(n);
    std::vector<int32_t> loc(n);
    for (int i = 0; i < n; i++) { fr[i] = jobs[i].cur; img[i] = jobs[i].image; loc[i] = jobs[i].image_on_device; }
    rc = enqueue_build(c, fr.data(), img.data(), loc.data(), n, n_detect > 0, nfeatures, mirror != 0, c->stream);
    if (rc) return rc;
    // the FAST scratch is shared with the build stream: a frame batch submitted later must not start before this one
    // has read its cell lists
    if (n_detect > 0) {
      if (!c->track_build_done) SDVLB_CUDA_TRY(cudaEventCreateWithFlags(&c->track_build_done, cudaEventDisableTiming));
      SDVLB_CUDA_TRY(cudaEventRecord(c->track_build_done, c->stream));
      c->track_build_pending = true;
    }
  }
  if (n_align > 0) {
    timer_begin(c, SDVLB_K_ALIGN);
    SDVLB_CUDA_TRY(sdvlb_launch_align(in.d + o_align, n_align, max_feats_job, g, c->dp, c->stream));
    timer_end(c);
    c->n_launches += 1;
  }
  if (n_cands > 0) {
    timer_begin(c, SDVLB_K_SEARCH);
    SDVLB_CUDA_TRY(sdvlb_launch_search(reinterpret_cast<const SearchCandDev*>(in.d + o_cands), n_cands,
                                       reinterpret_cast<const FrameDev*>(in.d + o_frames),
                                       reinterpret_cast<sdvlb_match*>(out.h + o_match), g, c->dp, c->stream,
                                       orb_search ? reinterpret_cast<const uint32_t*>(in.d + o_qdesc) : nullptr));
    timer_end(c);
    c->n_launches += 1;
  }
  c->d2h_bytes += int64_t(out.used);   // written over PCIe by the kernels themselves
  c->track_seq++;
  c->n_launches += 1;
  SDVLB_CUDA_TRY(sdvlb_launch_signal(reinterpret_cast<uint32_t*>(c->h_overflow) + 16, c->track_seq, c->stream));

  PendingTrack& P = c->pending;
  P.active = true;
  P.jobs = jobs; P.n = n; P.build_frames = build_frames;
  P.trace = trace; P.trace_cap = trace_cap; P.trace_n = trace_n;
  P.o_res = o_res; P.o_match = o_match; P.o_trace = o_trace;
  return 0;
}

// Spins on the pinned completion word until submission `seq_no` of the tracking stream has finished (submissions
// complete in order); falls back to the driver now and then so that device errors surface.
int wait_signal(sdvlb_ctx* c, uint32_t seq_no) {
  volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(c->h_overflow) + 16;
  uint64_t spins = 0;
  while (int32_t(*done - seq_no) < 0) {
    if ((++spins & 0xFFFFF) == 0) {
      SDVLB_CUDA_TRY(cudaSetDevice(c->device));
      const cudaError_t e = cudaStreamQuery(c->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) return sdvlb_set_cuda_error(e, "cudaStreamQuery", __FILE__, __LINE__);
      if (e == cudaSuccess && int32_t(*done - seq_no) < 0)
        return sdvlb_set_error(SDVLB_ERR_CUDA, "tracking stream drained without publishing its completion word");
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  return 0;
}

int collect_batch(sdvlb_ctx* c) {
  PendingTrack& P = c->pending;
  if (!P.active) return sdvlb_set_error(SDVLB_ERR_STATE, "nothing was submitted on this context");
  P.active = false;
  {
    const int rcw = wait_signal(c, c->track_seq);
    if (rcw) return rcw;
  }
  Arena& out = c->out;
  const int rc = check_overflow(c);
  if (rc) return rc;
  const BatchOut* res = reinterpret_cast<const BatchOut*>(out.h + P.o_res);
  const sdvlb_match* hm = reinterpret_cast<const sdvlb_match*>(out.h + P.o_match);
  int mo = 0;
  bool first_align = true, any_overflow = false;
  for (int i = 0; i < P.n; i++) {
    sdvlb_track_job& j = P.jobs[i];
    // the tracking stream ran after this frame's build (same stream, or ordered by its `built` event)
    if (P.build_frames || j.cur->build_pending) finalize_build(c, j.cur);
    if (j.cur->overflowed) any_overflow = true;
    if (j.ref) {
      memcpy(j.T_cur, res[i].pose, sizeof(j.T_cur));
      j.n_tracked = res[i].info[0] / (c->params.align_patch_size * c->params.align_patch_size);
      j.gn_iters = res[i].info[1];
      j.error = res[i].error;
      if (first_align) {
        if (P.trace_n) *P.trace_n = res[i].info[1];
        if (P.trace) memcpy(P.trace, out.h + P.o_trace, size_t(std::min(res[i].info[1], P.trace_cap)) * sizeof(sdvlb_gn_iter));
        first_align = false;
      }
    }
    if (j.n_cands > 0) memcpy(j.matches, hm + mo, size_t(j.n_cands) * sizeof(sdvlb_match));
    mo += j.n_cands;
  }
  if (any_overflow) return sdvlb_set_error(SDVLB_ERR_OVERFLOW, "corner capacity exceeded in a frame's FAST selection");
  return 0;
}

int run_batch(sdvlb_ctx* c, sdvlb_track_job* jobs, int n, int mirror, sdvlb_gn_iter* trace, int trace_cap,
              int* trace_n, const sdvlb_gn_forced* forced, bool build_frames, int fast = 0) {
  const int rc = submit_batch(c, jobs, n, mirror, trace, trace_cap, trace_n, forced, build_frames, fast);
  if (rc) { c->pending.active = false; return rc; }
  return collect_batch(c);
}

int check_track_args(sdvlb_ctx* ctx, sdvlb_track_job* jobs, int n_jobs, int w, int h, bool* build) {
  if (!ctx || !jobs || n_jobs <= 0) return sdvlb_set_error(SDVLB_ERR_ARG, "bad batch");
  if (w != ctx->w || h != ctx->h) return sdvlb_set_error(SDVLB_ERR_ARG, "image size differs from the context camera");
  int with_image = 0;
  for (int i = 0; i < n_jobs; i++) {
    if (jobs[i].image) with_image++;
    else if (!jobs[i].cur) return sdvlb_set_error(SDVLB_ERR_ARG, "job with neither an image nor a prebuilt frame");
  }
  if (with_image != 0 && with_image != n_jobs)
    return sdvlb_set_error(SDVLB_ERR_ARG, "a batch must either build all its frames or use prebuilt frames only");
  *build = with_image == n_jobs;
  return 0;
}

}  // namespace sdvlb_detail

using namespace sdvlb_detail;

extern "C" {

void sdvlb_params_default(sdvlb_params* p) {   // config.cc:55-85
  p->pyramid_levels = 5; p->cell_size = 32; p->max_matches = 150; p->max_align_level = 4; p->min_align_level = 2;
  p->max_img_align_its = 30; p->align_patch_size = 4; p->patch_size = 8; p->max_align_its = 10; p->search_size = 6;
  p->max_fast_levels = 3; p->fast_threshold = 10; p->num_features = 1000; p->max_failed = 15;
  p->max_optim_pose_its = 10; p->max_ransac_points = 5; p->max_ransac_its = 100; p->min_matches = 20;
  p->inlier_error_threshold = 2.0;
}

const char* sdvlb_last_error(void) { return g_last_error.c_str(); }

int sdvlb_ctx_create(int device, const sdvlb_params* params, const sdvlb_camera* cam, sdvlb_ctx** out) {
  if (!params || !cam || !out) return sdvlb_set_error(SDVLB_ERR_ARG, "null argument");
  const sdvlb_params& p = *params;
  const int w = int(cam->width), h = int(cam->height);
  if (p.cell_size != SDVLB_CELL || p.align_patch_size != 4 || p.patch_size != 8)
    return sdvlb_set_error(SDVLB_ERR_ARG, "only cell_size 32, align_patch_size 4, patch_size 8 are supported");
  if (p.pyramid_levels < 1 || p.pyramid_levels > SDVLB_MAX_LEVELS || p.max_fast_levels < 1 ||
      p.max_fast_levels > p.pyramid_levels || p.max_align_level >= p.pyramid_levels || p.min_align_level < 0 ||
      p.min_align_level > p.max_align_level)
    return sdvlb_set_error(SDVLB_ERR_ARG, "inconsistent pyramid / level parameters");
  if (w < 16 || h < 16 || w > SDVLB_MAX_DIM || h > SDVLB_MAX_DIM)
    return sdvlb_set_error(SDVLB_ERR_ARG, "image size must be within [16, 2048]");
  int ndev = 0;
  SDVLB_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return sdvlb_set_error(SDVLB_ERR_CUDA, "no such CUDA device");
  SDVLB_CUDA_TRY(cudaSetDevice(device));
  sdvlb_ctx* c = new sdvlb_ctx;
  c->device = device;
  c->params = p;
  c->cam = *cam;
  c->dp.p = p;
  c->dp.cam = *cam;
  c->w = w; c->h = h;
  build_geom(p, w, h, &c->geom);
  c->corner_cap = std::max(8192, 8 * p.num_features);
  c->corner_copy = std::min(c->corner_cap, std::max(2048, 2 * p.num_features));
  // tracking is the latency-critical chain of a sequence; frame batches are prefetch work
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (const char* env = getenv("SDVLB_STREAM_PRIO")) {   // experiment knob: 0 = equal priorities, -1 = build stream first
    if (atoi(env) == 0) prio_hi = prio_lo;
    else if (atoi(env) < 0) std::swap(prio_hi, prio_lo);
  }
  cudaError_t e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->bstream, cudaStreamNonBlocking, prio_lo);
  for (int i = 0; i < kBuildEvents && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&c->bevents[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = shared_upload_stream(device, prio_lo, &c->ustream, &c->umutex);
  for (int i = 0; i < kBuildEvents && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&c->uevents[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&c->h_overflow), 128, cudaHostAllocDefault);
  if (e != cudaSuccess) { delete c; return sdvlb_set_cuda_error(e, "context resources", __FILE__, __LINE__); }
  memset(c->h_overflow, 0, 128);
  const FastPlan* plan = get_plan(c, p.num_features);
  c->level_kp_total = size_t(plan->args.level_kp_total);
  *out = c;
  return 0;
}

int sdvlb_ctx_destroy(sdvlb_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->bstream);
  for (sdvlb_frame* f : c->all_frames) {
    if (f->h_pyr) cudaFreeHost(f->h_pyr);
    delete f;
  }
  for (uint8_t* slab : c->slabs) cudaFree(slab);
  for (uint8_t* slab : c->mirror_slabs) cudaFreeHost(slab);
  for (int i = 0; i < kBuildEvents; i++) if (c->bevents[i]) cudaEventDestroy(c->bevents[i]);
  for (int i = 0; i < kBuildEvents; i++) if (c->uevents[i]) cudaEventDestroy(c->uevents[i]);
  if (c->ustream) cudaStreamSynchronize(c->ustream);   // shared by the contexts of the device: never destroyed
  if (c->orb_buf) cudaFree(c->orb_buf);
  for (int i = 0; i < kStageSets; i++) {
    if (c->stage[i]) cudaFreeHost(c->stage[i]);
    if (c->stage_done[i]) cudaEventDestroy(c->stage_done[i]);
  }
  if (c->d_seeds) cudaFree(c->d_seeds);
  if (c->h_seeds) cudaFreeHost(c->h_seeds);
  if (c->raw_scratch) cudaFree(c->raw_scratch);
  for (int i = 0; i < kBuildEvents; i++) if (c->raw_done[i]) cudaEventDestroy(c->raw_done[i]);
  cudaFree(c->cell_kp); cudaFree(c->cell_cnt); cudaFree(c->cell_kept); cudaFree(c->level_kp); cudaFree(c->level_cnt);
  cudaFree(c->frame_ticket); cudaFree(c->scratch);
  if (c->h_overflow) cudaFreeHost(c->h_overflow);
  if (c->track_build_done) cudaEventDestroy(c->track_build_done);
  while (!c->seqs.empty()) sdvlb_seq_destroy(c, c->seqs.back());
  for (Arena& a : c->seq_in) {
    if (a.h) cudaFreeHost(a.h);
    if (a.d) cudaFree(a.d);
  }
  cudaFree(c->d_seq_done);
  if (c->in.h) cudaFreeHost(c->in.h);
  if (c->in.d) cudaFree(c->in.d);
  if (c->out.h) cudaFreeHost(c->out.h);
  for (auto& t : c->timers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->bstream);
  delete c;
  return 0;
}

int sdvlb_ctx_sync(sdvlb_ctx* c) {
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->bstream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}
void* sdvlb_ctx_stream(sdvlb_ctx* c) { return c->stream; }

int sdvlb_ctx_reserve_frames(sdvlb_ctx* c, int n_frames) {
  if (!c || n_frames < 0) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  while (int(c->all_frames.size()) < n_frames) {
    const int rc = grow_pool(c);
    if (rc) return rc;
  }
  return 0;
}

int sdvlb_host_alloc(void** p, uint64_t bytes) { SDVLB_CUDA_TRY(cudaHostAlloc(p, bytes, cudaHostAllocDefault)); return 0; }
int sdvlb_host_free(void* p) { SDVLB_CUDA_TRY(cudaFreeHost(p)); return 0; }
int sdvlb_dev_alloc(sdvlb_ctx* c, void** p, uint64_t bytes) {
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  SDVLB_CUDA_TRY(cudaMalloc(p, bytes));
  return 0;
}
int sdvlb_dev_free(sdvlb_ctx* c, void* p) {
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  SDVLB_CUDA_TRY(cudaFree(p));
  return 0;
}
int sdvlb_dev_upload(sdvlb_ctx* c, void* dst, const void* src, uint64_t bytes) {
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  SDVLB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

int sdvlb_ctx_counters(sdvlb_ctx* c, int64_t* kernel_launches, int64_t* h2d_bytes, int64_t* d2h_bytes, int reset) {
  if (kernel_launches) *kernel_launches = c->n_launches;
  if (h2d_bytes) *h2d_bytes = c->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = c->d2h_bytes;
  if (reset) { c->n_launches = 0; c->h2d_bytes = 0; c->d2h_bytes = 0; }
  return 0;
}

int sdvlb_timing_enable(sdvlb_ctx* c, int on) { c->timing = on != 0; return 0; }
int sdvlb_timing_read(sdvlb_ctx* c, double ms[SDVLB_K_COUNT], int64_t launches[SDVLB_K_COUNT], int reset) {
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->bstream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  timer_collect(c);
  for (int i = 0; i < SDVLB_K_COUNT; i++) {
    if (ms) ms[i] = c->t_ms[i];
    if (launches) launches[i] = c->t_launches[i];
    if (reset) { c->t_ms[i] = 0; c->t_launches[i] = 0; }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- tracking batches
int sdvlb_track_batch(sdvlb_ctx* ctx, sdvlb_track_job* jobs, int n_jobs, int w, int h, int mirror) {
  SDVLB_RANGE("sdvlb.track_batch");
  bool build = false;
  const int rc = check_track_args(ctx, jobs, n_jobs, w, h, &build);
  if (rc) return rc;
  return run_batch(ctx, jobs, n_jobs, mirror, nullptr, 0, nullptr, nullptr, build);
}

int sdvlb_track_submit(sdvlb_ctx* ctx, sdvlb_track_job* jobs, int n_jobs, int w, int h, int mirror) {
  SDVLB_RANGE("sdvlb.track_submit");
  bool build = false;
  int rc = check_track_args(ctx, jobs, n_jobs, w, h, &build);
  if (rc) return rc;
  rc = submit_batch(ctx, jobs, n_jobs, mirror, nullptr, 0, nullptr, nullptr, build);
  if (rc) ctx->pending.active = false;
  return rc;
}

int sdvlb_track_poll(sdvlb_ctx* ctx) {
  if (!ctx || !ctx->pending.active) return sdvlb_set_error(SDVLB_ERR_STATE, "nothing was submitted on this context");
  const volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(ctx->h_overflow) + 16;
  return *done == ctx->track_seq ? 1 : 0;   // a plain read of pinned memory: no driver call, no lock
}

int sdvlb_track_collect(sdvlb_ctx* ctx) {
  SDVLB_RANGE("sdvlb.track_collect");
  if (!ctx) return sdvlb_set_error(SDVLB_ERR_ARG, "null context");
  return collect_batch(ctx);
}

// ---------------------------------------------------------------------------------------------- frame batches
int sdvlb_frames_submit(sdvlb_ctx* c, const uint8_t* const* images, int n, int images_on_device, int want_corners,
                        int nfeatures, sdvlb_frame** out) {
  SDVLB_RANGE("sdvlb.frames_submit");
  if (!c || !images || !out || n <= 0) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (images_on_device < 0 || images_on_device > 2) return sdvlb_set_error(SDVLB_ERR_ARG, "bad image location");
  for (int i = 0; i < n; i++)
    if (!images[i]) return sdvlb_set_error(SDVLB_ERR_ARG, "null image");
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  int rc = want_corners ? ensure_fast_scratch(c, n) : 0;
  if (rc) return rc;
  for (int i = 0; i < n; i++) {
    sdvlb_frame* f = nullptr;
    rc = frame_alloc(c, &f);
    if (rc) { for (int k = 0; k < i; k++) frame_release(c, out[k]); return rc; }
    out[i] = f;
  }
  if (c->track_build_pending) {   // a tracking batch that built its own frames used the FAST scratch on the other stream
    SDVLB_CUDA_TRY(cudaStreamWaitEvent(c->bstream, c->track_build_done, 0));
    c->track_build_pending = false;
  }
  std::vector<int32_t> loc(n, images_on_device);
  rc = enqueue_build(c, out, images, loc.data(), n, want_corners != 0, nfeatures, true, c->bstream,
                     (images_on_device == SDVLB_IMG_DEVICE && !getenv("SDVLB_UPLOAD_MIN_NS_PER_FRAME")) ? nullptr : c->ustream);
  if (rc) return rc;
  cudaEvent_t ev = c->bevents[c->bevent_next];
  c->bevent_next = (c->bevent_next + 1) % kBuildEvents;
  SDVLB_CUDA_TRY(cudaEventRecord(ev, c->bstream));
  for (int i = 0; i < n; i++) { out[i]->build_pending = true; out[i]->built = ev; }
  c->last_build = ev;
  return 0;
}

int sdvlb_frames_wait(sdvlb_ctx* c, sdvlb_frame* const* frames, int n) {
  SDVLB_RANGE("sdvlb.frames_wait");
  if (!c || (n > 0 && !frames)) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  for (int i = 0; i < n; i++) {
    const int rc = ensure_built(frames[i]);
    if (rc) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- Frame
int sdvlb_frame_create(sdvlb_ctx* ctx, const uint8_t* img, int w, int h, int stride, int want_corners, int nfeatures,
                       sdvlb_frame** out) {
  SDVLB_RANGE("sdvlb.frame_create");
  if (!ctx || !img || !out) return sdvlb_set_error(SDVLB_ERR_ARG, "null argument");
  if (w != ctx->w || h != ctx->h) return sdvlb_set_error(SDVLB_ERR_ARG, "image size differs from the context camera");
  std::vector<uint8_t> packed;
  const uint8_t* src = img;
  if (stride != w) {   // the device layout is continuous (image_align.cc:139 assumes stride == cols too)
    packed.resize(size_t(w) * h);
    for (int y = 0; y < h; y++) memcpy(packed.data() + size_t(y) * w, img + size_t(y) * stride, w);
    src = packed.data();
  }
  sdvlb_track_job j;
  memset(&j, 0, sizeof(j));
  j.image = src;
  j.want_corners = want_corners;
  j.nfeatures = nfeatures;
  const int rc = run_batch(ctx, &j, 1, 1, nullptr, 0, nullptr, nullptr, true);
  if (rc) return rc;
  *out = j.cur;
  return 0;
}

int sdvlb_frame_detect(sdvlb_ctx* ctx, sdvlb_frame* f, int nfeatures) {
  if (!ctx || !f) return sdvlb_set_error(SDVLB_ERR_ARG, "null argument");
  SDVLB_CUDA_TRY(cudaSetDevice(ctx->device));
  int rc = ensure_built(f);
  if (rc) return rc;
  rc = ensure_fast_scratch(ctx, 1);
  if (rc) return rc;
  if (ctx->last_build) SDVLB_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->last_build, 0));   // shared FAST scratch
  FrameBatch B;
  B.n = 1;
  B.scratch_base = 0;
  B.f[0] = f->dev;
  B.f[0].host_mirror = reinterpret_cast<int32_t*>(f->h_corners);
  const FastPlan* plan = get_plan(ctx, nfeatures);
  timer_begin(ctx, SDVLB_K_FAST);
  SDVLB_CUDA_TRY(sdvlb_launch_fast_cells(B, *plan, ctx->cell_kp, ctx->cell_cnt, ctx->stream));
  timer_end(ctx);
  timer_begin(ctx, SDVLB_K_SELECT);
  SDVLB_CUDA_TRY(sdvlb_launch_fast_select(B, *plan, ctx->cell_kp, ctx->cell_cnt, ctx->cell_kept, ctx->level_kp,
                                          ctx->level_cnt, ctx->frame_ticket, ctx->stream));
  timer_end(ctx);
  ctx->n_launches += 2;
  f->build_desc = ctx->use_orb && f->dev.desc != nullptr;
  if (ctx->use_orb) {
    SDVLB_CUDA_TRY(sdvlb_launch_orb_frames(B, ctx->geom, ctx->params.pyramid_levels, ctx->corner_cap, ctx->stream));
    ctx->n_launches += 1;
  }
  SDVLB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  rc = check_overflow(ctx);
  if (rc) return rc;
  f->build_corners = true;
  f->build_mirror = true;
  finalize_build(ctx, f);
  f->h_more.clear();
  return 0;
}

int sdvlb_frame_level(const sdvlb_frame* f, int level, const uint8_t** data, int* w, int* h) {
  if (!f || level < 0 || level >= f->ctx->geom.levels) return sdvlb_set_error(SDVLB_ERR_ARG, "bad level");
  sdvlb_frame* mf = const_cast<sdvlb_frame*>(f);
  if (!mf->pyr_mirrored) {   // the host mirror is materialised on first use
    sdvlb_ctx* c = f->ctx;
    SDVLB_CUDA_TRY(cudaSetDevice(c->device));
    const int rc = ensure_built(mf);
    if (rc) return rc;
    if (!mf->h_pyr) SDVLB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&mf->h_pyr), size_t(c->geom.total), cudaHostAllocDefault));
    SDVLB_CUDA_TRY(cudaMemcpyAsync(mf->h_pyr, mf->d_block, size_t(c->geom.total), cudaMemcpyDeviceToHost, c->stream));
    SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += int64_t(c->geom.total);
    mf->pyr_mirrored = true;
  }
  if (data) *data = f->h_pyr + f->ctx->geom.off[level];
  if (w) *w = f->ctx->geom.w[level];
  if (h) *h = f->ctx->geom.h[level];
  return 0;
}

int sdvlb_frame_corners(const sdvlb_frame* f, const int32_t** xyls, int* n) {
  if (!f) return sdvlb_set_error(SDVLB_ERR_ARG, "null frame");
  sdvlb_frame* mf = const_cast<sdvlb_frame*>(f);
  sdvlb_ctx* c = f->ctx;
  const int rc = ensure_built(mf);
  if (rc) return rc;
  if (!f->has_corners) return sdvlb_set_error(SDVLB_ERR_STATE, "corners were not detected on this frame");
  if (mf->corners_mirrored < 0) {   // built without a mirror: fetch header + first block now
    SDVLB_CUDA_TRY(cudaSetDevice(c->device));
    SDVLB_CUDA_TRY(cudaMemcpyAsync(mf->h_corners, mf->d_block + mf->off_hdr, 16 + size_t(c->corner_copy) * 16,
                                   cudaMemcpyDeviceToHost, c->stream));
    SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    mf->build_corners = true;
    mf->build_mirror = true;
    finalize_build(c, mf);
  }
  const int32_t* src = reinterpret_cast<const int32_t*>(mf->h_corners + 16);
  if (mf->corners_mirrored < mf->n_corners) {   // longer than the eager mirror: fetch the whole list once
    if (mf->h_more.size() != size_t(mf->n_corners) * 4) {
      SDVLB_CUDA_TRY(cudaSetDevice(c->device));
      mf->h_more.resize(size_t(mf->n_corners) * 4);
      SDVLB_CUDA_TRY(cudaMemcpyAsync(mf->h_more.data(), mf->d_block + mf->off_hdr + 16, size_t(mf->n_corners) * 16,
                                     cudaMemcpyDeviceToHost, c->stream));
      SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
      c->d2h_bytes += int64_t(mf->n_corners) * 16;
    }
    src = mf->h_more.data();
  }
  if (xyls) *xyls = src;
  if (n) *n = f->n_corners;
  return 0;
}

int sdvlb_frame_destroy(sdvlb_ctx* ctx, sdvlb_frame* f) {
  if (!f) return 0;
  if (!ctx) ctx = f->ctx;
  if (f->build_pending) {   // the slot must not be recycled while its build is in flight
    cudaSetDevice(ctx->device);
    cudaEventSynchronize(f->built);
    f->build_pending = false;
  }
  frame_release(ctx, f);
  return 0;
}

// ---------------------------------------------------------------------------------------------- ImageAlign / Matcher
int sdvlb_image_align(sdvlb_ctx* ctx, const sdvlb_frame* ref, sdvlb_frame* cur, const sdvlb_align_feat* feats, int n,
                      const double T_ref[7], double T_cur[7], int fast, int* n_tracked, double* error,
                      sdvlb_gn_iter* trace, int trace_cap, int* trace_n, const sdvlb_gn_forced* forced) {
  SDVLB_RANGE("sdvlb.image_align");
  if (!ctx || !ref || !cur || !T_ref || !T_cur || n < 0 || (n > 0 && !feats))
    return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (trace_n) *trace_n = 0;
  if (n == 0) {   // image_align.cc:55-58
    if (n_tracked) *n_tracked = 0;
    if (error) *error = 1e10;
    return 0;
  }
  sdvlb_track_job j;
  memset(&j, 0, sizeof(j));
  j.ref = ref;
  j.cur = cur;
  j.feats = feats;
  j.n_feats = n;
  memcpy(j.T_ref, T_ref, sizeof(j.T_ref));
  memcpy(j.T_cur, T_cur, sizeof(j.T_cur));
  const int rc = run_batch(ctx, &j, 1, 0, trace, trace ? trace_cap : 0, trace_n, forced, false, fast);
  if (rc) return rc;
  memcpy(T_cur, j.T_cur, sizeof(j.T_cur));
  if (n_tracked) *n_tracked = j.n_tracked;
  if (error) *error = j.error;
  return 0;
}

int sdvlb_search_points(sdvlb_ctx* ctx, const sdvlb_frame* cur, const sdvlb_candidate* cands, int n,
                        const double T_cur[7], sdvlb_match* out) {
  SDVLB_RANGE("sdvlb.search_points");
  if (!ctx || !cur || n < 0 || (n > 0 && (!cands || !out))) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  const int rcb = ensure_built(const_cast<sdvlb_frame*>(cur));
  if (rcb) return rcb;
  if (!cur->has_corners) return sdvlb_set_error(SDVLB_ERR_STATE, "current frame has no corners");
  if (n == 0) return 0;
  sdvlb_track_job j;
  memset(&j, 0, sizeof(j));
  j.cur = const_cast<sdvlb_frame*>(cur);
  j.cands = cands;
  j.n_cands = n;
  j.matches = out;
  if (T_cur) {
    memcpy(j.T_cur, T_cur, sizeof(j.T_cur));
  } else {   // keep the device-side pose: read it back so the upload below is a no-op in value
    SDVLB_CUDA_TRY(cudaSetDevice(ctx->device));
    SDVLB_CUDA_TRY(cudaMemcpyAsync(j.T_cur, cur->dev.pose, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SDVLB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return run_batch(ctx, &j, 1, 0, nullptr, 0, nullptr, nullptr, false);
}

int sdvlb_ctx_set_distortion(sdvlb_ctx* c, const double d[5]) {
  if (!c || !d) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  c->has_dist = !(d[0] == 0.0 && d[1] == 0.0 && d[2] == 0.0 && d[3] == 0.0 && d[4] == 0.0);
  c->und.w = c->w; c->und.h = c->h;
  c->und.fx = c->cam.fx; c->und.fy = c->cam.fy; c->und.u0 = c->cam.u0; c->und.v0 = c->cam.v0;
  c->und.k1 = d[0]; c->und.k2 = d[1]; c->und.p1 = d[2]; c->und.p2 = d[3]; c->und.k3 = d[4];
  return 0;
}

// Camera::UndistortImage(in, out) (camera.cc:100-105): a frame slot is borrowed for the result
int sdvlb_undistort(sdvlb_ctx* c, const uint8_t* in, uint8_t* out) {
  if (!c || !in || !out) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  const size_t bytes = size_t(c->w) * c->h;
  if (!c->has_dist) { memmove(out, in, bytes); return 0; }
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  sdvlb_frame* f = nullptr;
  int rc = frame_alloc(c, &f);
  if (rc) return rc;
  const int32_t loc = SDVLB_IMG_HOST;
  const uint8_t* img = in;
  // upload + undistortion only: no corners, and the pyramid levels built behind it are simply not read
  rc = enqueue_build(c, &f, &img, &loc, 1, false, 0, false, c->stream);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(out, f->dev.pyr, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = sdvlb_set_cuda_error(e, "undistort", __FILE__, __LINE__);
    c->d2h_bytes += int64_t(bytes);
  }
  frame_release(c, f);
  return rc;
}

// Map::UpdateCandidates (map.cc:397-498) for n seeds against `cur`: one upload, one kernel, one read-back.
int sdvlb_update_candidates(sdvlb_ctx* c, const sdvlb_frame* cur, const double T_cur[7], sdvlb_seed* seeds, int n,
                            const sdvlb_seed_params* sp) {
  SDVLB_RANGE("sdvlb.update_candidates");
  if (!c || !cur || !T_cur || !sp || n < 0 || (n > 0 && !seeds)) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  int rc = ensure_built(const_cast<sdvlb_frame*>(cur));
  if (rc) return rc;
  if (!cur->has_corners) return sdvlb_set_error(SDVLB_ERR_STATE, "current frame has no corners");
  if (n == 0) return 0;
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  for (int i = 0; i < n; i++) {
    if (!seeds[i].ref_frame) return sdvlb_set_error(SDVLB_ERR_ARG, "seed without a reference frame");
    if (seeds[i].ref_level < 0 || seeds[i].ref_level >= c->params.pyramid_levels)
      return sdvlb_set_error(SDVLB_ERR_ARG, "seed level out of range");
    rc = ensure_built(const_cast<sdvlb_frame*>(seeds[i].ref_frame));
    if (rc) return rc;
  }
  if (n > c->seeds_cap) {
    const int cap = std::max(n, std::max(256, 2 * c->seeds_cap));
    if (c->d_seeds) cudaFree(c->d_seeds);
    if (c->h_seeds) cudaFreeHost(c->h_seeds);
    c->d_seeds = nullptr; c->h_seeds = nullptr; c->seeds_cap = 0;
    SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_seeds), sizeof(sdvlb_seed) * cap));
    SDVLB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->h_seeds), sizeof(sdvlb_seed) * cap, cudaHostAllocDefault));
    c->seeds_cap = cap;
  }
  for (int i = 0; i < n; i++) {   // frame handles -> device pyramids
    c->h_seeds[i] = seeds[i];
    c->h_seeds[i].ref_frame = reinterpret_cast<const sdvlb_frame*>(seeds[i].ref_frame->dev.pyr);
  }
  const size_t bytes = sizeof(sdvlb_seed) * size_t(n);
  SDVLB_CUDA_TRY(cudaMemcpyAsync(cur->dev.pose, T_cur, 7 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SDVLB_CUDA_TRY(cudaMemcpyAsync(c->d_seeds, c->h_seeds, bytes, cudaMemcpyHostToDevice, c->stream));
  SDVLB_CUDA_TRY(sdvlb_launch_seed_update(c->d_seeds, n, cur->dev, c->geom, c->dp, *sp, c->stream, c->use_orb));
  SDVLB_CUDA_TRY(cudaMemcpyAsync(c->h_seeds, c->d_seeds, bytes, cudaMemcpyDeviceToHost, c->stream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->n_launches += 1;
  c->h2d_bytes += int64_t(bytes) + 56;
  c->d2h_bytes += int64_t(bytes);
  for (int i = 0; i < n; i++) {
    const sdvlb_frame* ref = seeds[i].ref_frame;
    seeds[i] = c->h_seeds[i];
    seeds[i].ref_frame = ref;
  }
  return check_overflow(c);
}

// ---- ORB descriptor mode (Config::UseORB()) ------------------------------------------
static int orb_scratch(sdvlb_ctx* c, size_t bytes) {
  if (bytes <= c->orb_cap) return 0;
  const size_t cap = std::max(bytes, 2 * c->orb_cap);
  if (c->orb_buf) cudaFree(c->orb_buf);
  c->orb_buf = nullptr; c->orb_cap = 0;
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->orb_buf), cap));
  c->orb_cap = cap;
  return 0;
}

int sdvlb_ctx_set_orb(sdvlb_ctx* c, int on) {
  if (!c) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (on && (c->params.patch_size != 8)) return sdvlb_set_error(SDVLB_ERR_ARG, "ORB mode needs patch_size 8");
  const int rc = sdvlb_ctx_sync(c);
  if (rc) return rc;
  if (on && !c->use_orb && !c->slabs.empty()) {
    // frame slots allocated so far have no room for Frame::descriptors_: regrow the pool (no frame may be alive)
    if (c->pool.size() != c->all_frames.size())
      return sdvlb_set_error(SDVLB_ERR_STATE, "sdvlb_ctx_set_orb(ctx, 1) with frames alive: switch the mode before creating frames");
    SDVLB_CUDA_TRY(cudaSetDevice(c->device));
    const size_t n_frames = c->all_frames.size();
    for (sdvlb_frame* f : c->all_frames) { if (f->h_pyr) cudaFreeHost(f->h_pyr); delete f; }
    c->all_frames.clear(); c->pool.clear();
    for (uint8_t* p : c->slabs) cudaFree(p);
    for (uint8_t* p : c->mirror_slabs) cudaFreeHost(p);
    c->slabs.clear(); c->mirror_slabs.clear();
    c->use_orb = true;
    c->plans.clear();
    return sdvlb_ctx_reserve_frames(c, int(n_frames));
  }
  c->use_orb = on != 0;
  c->plans.clear();          // FAST plans carry the border margin
  return 0;
}

int sdvlb_frame_orb_descriptors(sdvlb_ctx* c, const sdvlb_frame* f, const int32_t* xyl, int n, uint8_t* desc, float* angle) {
  if (!c || !f || n < 0 || (n > 0 && (!xyl || !desc))) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  int rc = ensure_built(const_cast<sdvlb_frame*>(f));
  if (rc) return rc;
  if (n == 0) return 0;
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  const size_t o_xyl = 0, o_desc = align_up(size_t(n) * 12, 256), o_ang = o_desc + align_up(size_t(n) * 32, 256);
  rc = orb_scratch(c, o_ang + size_t(n) * 4);
  if (rc) return rc;
  int32_t* d_xyl = reinterpret_cast<int32_t*>(c->orb_buf + o_xyl);
  uint8_t* d_desc = c->orb_buf + o_desc;
  float* d_ang = reinterpret_cast<float*>(c->orb_buf + o_ang);
  SDVLB_CUDA_TRY(cudaMemcpyAsync(d_xyl, xyl, size_t(n) * 12, cudaMemcpyHostToDevice, c->stream));
  SDVLB_CUDA_TRY(sdvlb_launch_orb_positions(f->dev.pyr, c->geom, c->params.pyramid_levels, d_xyl, n, d_desc, d_ang, c->stream));
  SDVLB_CUDA_TRY(cudaMemcpyAsync(desc, d_desc, size_t(n) * 32, cudaMemcpyDeviceToHost, c->stream));
  if (angle) SDVLB_CUDA_TRY(cudaMemcpyAsync(angle, d_ang, size_t(n) * 4, cudaMemcpyDeviceToHost, c->stream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->n_launches += 1;
  c->h2d_bytes += int64_t(n) * 12;
  c->d2h_bytes += int64_t(n) * (angle ? 36 : 32);
  for (int i = 0; i < n; i++) {   // ORBDetector::GetDescriptor asserts IsInsideLimits (extra/orb_detector.cc:357)
    const int l = xyl[3 * i + 2];
    if (l < 0 || l >= c->params.pyramid_levels || xyl[3 * i] < 19 || xyl[3 * i] >= c->geom.w[l] - 19 || xyl[3 * i + 1] < 19 ||
        xyl[3 * i + 1] >= c->geom.h[l] - 19)
      return sdvlb_set_error(SDVLB_ERR_ARG, "position outside the ORB limits (descriptor zeroed)");
  }
  return 0;
}

int sdvlb_search_points_orb(sdvlb_ctx* c, const sdvlb_frame* cur, const sdvlb_candidate* cands, int n, const double T_cur[7],
                            const uint8_t* cand_desc, sdvlb_match* out) {
  if (!c || !cur || !T_cur || n < 0 || (n > 0 && (!cands || !out || !cand_desc)))
    return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (!c->use_orb) return sdvlb_set_error(SDVLB_ERR_STATE, "sdvlb_ctx_set_orb(ctx, 1) first");
  const int rc = ensure_built(const_cast<sdvlb_frame*>(cur));
  if (rc) return rc;
  if (!cur->has_corners) return sdvlb_set_error(SDVLB_ERR_STATE, "current frame has no corners");
  if (n == 0) return 0;
  sdvlb_track_job j;
  memset(&j, 0, sizeof(j));
  j.cur = const_cast<sdvlb_frame*>(cur);
  j.cands = cands;
  j.n_cands = n;
  j.matches = out;
  j.cand_desc = cand_desc;
  memcpy(j.T_cur, T_cur, sizeof(j.T_cur));
  return run_batch(c, &j, 1, 0, nullptr, 0, nullptr, nullptr, false);
}

// Frame::GetDescriptors() (frame.h): the 32-byte descriptor of every corner, in Frame::GetCorners() order.
int sdvlb_frame_descriptors(sdvlb_ctx* c, const sdvlb_frame* f, uint8_t* desc, int cap, int* n) {
  if (!c || !f || (cap > 0 && !desc)) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  const int rc = ensure_built(const_cast<sdvlb_frame*>(f));
  if (rc) return rc;
  if (!f->has_corners || !f->has_desc)
    return sdvlb_set_error(SDVLB_ERR_STATE, "the frame has no corner descriptors (built without corners or outside ORB mode)");
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  int nc = f->n_corners;
  if (f->corners_mirrored < 0) {   // built without a mirror: the count is still on the device
    SDVLB_CUDA_TRY(cudaMemcpyAsync(&nc, f->dev.n_corners, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  if (n) *n = nc;
  const int m = std::min(nc, cap);
  if (m > 0) {
    SDVLB_CUDA_TRY(cudaMemcpyAsync(desc, f->dev.desc, size_t(m) * 32, cudaMemcpyDeviceToHost, c->stream));
    SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += int64_t(m) * 32;
  }
  return 0;
}

}  // extern "C"
