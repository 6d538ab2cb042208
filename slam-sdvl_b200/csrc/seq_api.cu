// seq_api.cu — host side of the resident-sequence entry points and of the standalone FeatureAlign pose refinement
// (include/sdvl_b200.h: sdvlb_seq_*, sdvlb_select_inliers, sdvlb_optimize_pose, sdvlb_rand_*).
//
// A tracked frame of n sequences is one submission of three kernels on the context's tracking stream (align -- which
// first applies the mapping thread's queued commands --, search, post; each a programmatic dependent launch of the one
// before; in ORB mode a small kernel first computes the descriptors of new map points); nothing is copied to the device
// except those commands (new points), and every result lands in pinned host memory by itself.
#include <algorithm>
#include <cstring>
#include <vector>

#include "capi_internal.h"

using namespace sdvlb_detail;

namespace {

struct SeqLayout {
  size_t list[2], cell_order, matches, align_scratch, c_cell, c_score, c_rank, c_next, o_a, o_pos, o_scale, o_err, o_flag, total;
  size_t fdesc[2];
};

SeqLayout seq_layout(int max_feats, int n_cells, bool orb) {
  SeqLayout L;
  size_t off = align_up(sizeof(SeqState), 256);
  auto take = [&off](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t n = size_t(max_feats);
  L.list[0] = take(n * sizeof(SeqFeat));
  L.list[1] = take(n * sizeof(SeqFeat));
  L.cell_order = take(size_t(n_cells) * sizeof(int32_t));
  L.matches = take(n * sizeof(sdvlb_match));
  L.align_scratch = take(SDVLB_ALIGN_SC_BYTES(n) + 256);
  L.c_cell = take(n * sizeof(int32_t));
  L.c_score = take(n * sizeof(int32_t));
  L.c_rank = take(n * sizeof(int32_t));
  L.c_next = take(n * sizeof(int32_t));
  L.o_a = take(n * 2 * sizeof(double));
  L.o_pos = take(n * 3 * sizeof(double));
  L.o_scale = take(n * sizeof(double));
  L.o_err = take(n * sizeof(double));
  L.o_flag = take(n * sizeof(int32_t));
  L.fdesc[0] = L.fdesc[1] = 0;
  if (orb) { L.fdesc[0] = take(n * 32); L.fdesc[1] = take(n * 32); }
  L.total = off;
  return L;
}

size_t result_stride(int max_feats) { return align_up(sizeof(SeqResultHost) + size_t(max_feats) * sizeof(sdvlb_seq_feat), 256); }

int ensure_step_buffers(sdvlb_ctx* c) {
  if (c->d_seq_done) return 0;
  SDVLB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c->d_seq_done), 256));
  SDVLB_CUDA_TRY(cudaMemsetAsync(c->d_seq_done, 0, 256, c->stream));
  // the command staging of every submission slot up front: an allocation (pinned host + device) synchronises the
  // device and costs milliseconds, and which slot carries the next keyframe's points is a matter of timing
  for (Arena& a : c->seq_in) {
    const int rc = ensure_arena(&a, 1 << 20, true);
    if (rc) return rc;
  }
  return 0;
}

bool seq_busy(const sdvlb_ctx* c) { return !c->seq_queue.empty(); }

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------- rand()
void sdvlb_rand_seed(sdvlb_rand* s, unsigned seed) { rand_seed(s, seed); }
int sdvlb_rand_next(sdvlb_rand* s) { return rand_next(s); }
void sdvlb_rand_shuffle(sdvlb_rand* s, int32_t* v, int n) {   // libstdc++ std::random_shuffle(first, last)
  for (int i = 1; i < n; ++i) {
    const int j = rand_next(s) % (i + 1);
    if (i != j) std::swap(v[i], v[j]);
  }
}

// ---------------------------------------------------------------------------------------------- pose refinement
static int pose_call(sdvlb_ctx* c, sdvlb_pose_obs* obs, int n, double T[7], sdvlb_rand* rng, int mode) {
  if (!c || n < 0 || (n > 0 && !obs) || !T) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (c->pending.active || seq_busy(c)) return sdvlb_set_error(SDVLB_ERR_STATE, "a submission is still in flight on this context");
  if (n == 0) return 0;   // SelectInliers returns at once (feature_align.cc:163-164); ConvergePose has no errors (:373-374)
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  Arena& in = c->in;
  in.used = 0;
  const size_t need = 4096 + size_t(n) * (sizeof(sdvlb_pose_obs) + 7 * 8 + 4) + 1024;
  const int rc = ensure_arena(&in, need, true);
  if (rc) return rc;
  const size_t o_obs = in.take(size_t(n) * sizeof(sdvlb_pose_obs));
  const size_t o_T = in.take(7 * sizeof(double));
  const size_t o_rng = in.take(sizeof(sdvlb_rand));
  const size_t o_sc = in.take(size_t(n) * 7 * sizeof(double));
  const size_t o_isc = in.take(size_t(n) * sizeof(int32_t));
  memcpy(in.h + o_obs, obs, size_t(n) * sizeof(sdvlb_pose_obs));
  memcpy(in.h + o_T, T, 7 * sizeof(double));
  if (rng) memcpy(in.h + o_rng, rng, sizeof(sdvlb_rand));
  const size_t up = o_rng + sizeof(sdvlb_rand);
  SDVLB_CUDA_TRY(cudaMemcpyAsync(in.d, in.h, up, cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes += int64_t(up);
  PoseCallArgs A;
  A.obs = reinterpret_cast<sdvlb_pose_obs*>(in.d + o_obs);
  A.n = n;
  A.mode = mode;
  A.T = reinterpret_cast<double*>(in.d + o_T);
  A.rng = reinterpret_cast<sdvlb_rand*>(in.d + o_rng);
  A.scratch = reinterpret_cast<double*>(in.d + o_sc);
  A.iscratch = reinterpret_cast<int32_t*>(in.d + o_isc);
  A.dp = c->dp;
  timer_begin(c, SDVLB_K_POSE);
  SDVLB_CUDA_TRY(sdvlb_launch_pose_call(A, c->stream));
  timer_end(c);
  c->n_launches += 1;
  SDVLB_CUDA_TRY(cudaMemcpyAsync(in.h, in.d, up, cudaMemcpyDeviceToHost, c->stream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->d2h_bytes += int64_t(up);
  memcpy(obs, in.h + o_obs, size_t(n) * sizeof(sdvlb_pose_obs));
  memcpy(T, in.h + o_T, 7 * sizeof(double));
  if (rng) memcpy(rng, in.h + o_rng, sizeof(sdvlb_rand));
  return 0;
}

int sdvlb_select_inliers(sdvlb_ctx* ctx, sdvlb_pose_obs* obs, int n, const double T_frame[7], sdvlb_rand* rng) {
  SDVLB_RANGE("sdvlb.select_inliers");
  if (!rng || !T_frame) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (ctx && (ctx->params.max_ransac_its > 256 || ctx->params.max_ransac_points > 8))
    return sdvlb_set_error(SDVLB_ERR_ARG, "max_ransac_its <= 256 and max_ransac_points <= 8 are supported");
  double T[7];
  memcpy(T, T_frame, sizeof(T));
  return pose_call(ctx, obs, n, T, rng, 0);
}

int sdvlb_optimize_pose(sdvlb_ctx* ctx, sdvlb_pose_obs* obs, int n, double T_frame[7]) {
  SDVLB_RANGE("sdvlb.optimize_pose");
  return pose_call(ctx, obs, n, T_frame, nullptr, 1);
}

// ---------------------------------------------------------------------------------------------- sequences
int sdvlb_seq_create(sdvlb_ctx* c, int max_feats, sdvlb_seq** out) {
  if (!c || !out) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (c->params.max_ransac_its > 256 || c->params.max_ransac_points > 8)
    return sdvlb_set_error(SDVLB_ERR_ARG, "max_ransac_its <= 256 and max_ransac_points <= 8 are supported");
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  if (max_feats <= 0) max_feats = std::max(256, 2 * c->params.max_matches);
  max_feats = int(align_up(size_t(max_feats), 64));
  // FeatureAlign's grid (feature_align.cc:44-46) is the level-0 FAST grid
  const int n_cells = c->geom.wcells[0] * c->geom.hcells[0];
  if (n_cells > 4096) return sdvlb_set_error(SDVLB_ERR_ARG, "image too large for the FeatureAlign grid");
  const SeqLayout L = seq_layout(max_feats, n_cells, c->use_orb);
  sdvlb_seq* s = new sdvlb_seq;
  s->ctx = c;
  s->max_feats = max_feats;
  s->n_cells = n_cells;
  s->result_stride = result_stride(max_feats);
  s->has_desc = c->use_orb;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->d_block), L.total);
  if (e == cudaSuccess)
    e = cudaHostAlloc(reinterpret_cast<void**>(&s->h_result), s->result_stride * SDVLB_SEQ_DEPTH, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    if (s->d_block) cudaFree(s->d_block);
    delete s;
    return sdvlb_set_cuda_error(e, "sequence storage", __FILE__, __LINE__);
  }
  for (int k = 0; k < SDVLB_SEQ_DEPTH; k++) memset(s->h_result + size_t(k) * s->result_stride, 0, sizeof(SeqResultHost));
  // initial state: FeatureAlign constructor (cell order shuffled once, feature_align.cc:47-53) + the shuffle that
  // opens the first tracked frame's SelectPoints (:103), which the device always holds one frame ahead
  std::vector<uint8_t> init(L.cell_order + size_t(n_cells) * sizeof(int32_t), 0);
  SeqState* st = reinterpret_cast<SeqState*>(init.data());
  st->T_last[0] = 1.0;
  rand_seed(&st->rng, 1);
  int32_t* order = reinterpret_cast<int32_t*>(init.data() + L.cell_order);
  for (int i = 0; i < n_cells; i++) order[i] = i;
  sdvlb_rand_shuffle(&st->rng, order, n_cells);
  sdvlb_rand_shuffle(&st->rng, order, n_cells);
  st->max_feats = max_feats;
  st->n_cells = n_cells;
  uint8_t* d = s->d_block;
  st->list[0] = reinterpret_cast<SeqFeat*>(d + L.list[0]);
  st->list[1] = reinterpret_cast<SeqFeat*>(d + L.list[1]);
  st->cell_order = reinterpret_cast<int32_t*>(d + L.cell_order);
  st->matches = reinterpret_cast<sdvlb_match*>(d + L.matches);
  st->align_scratch = d + L.align_scratch;
  st->c_cell = reinterpret_cast<int32_t*>(d + L.c_cell);
  st->c_score = reinterpret_cast<int32_t*>(d + L.c_score);
  st->c_rank = reinterpret_cast<int32_t*>(d + L.c_rank);
  st->c_next = reinterpret_cast<int32_t*>(d + L.c_next);
  st->o_a = reinterpret_cast<double*>(d + L.o_a);
  st->o_pos = reinterpret_cast<double*>(d + L.o_pos);
  st->o_scale = reinterpret_cast<double*>(d + L.o_scale);
  st->o_err = reinterpret_cast<double*>(d + L.o_err);
  st->o_flag = reinterpret_cast<int32_t*>(d + L.o_flag);
  st->fdesc[0] = c->use_orb ? reinterpret_cast<uint32_t*>(d + L.fdesc[0]) : nullptr;
  st->fdesc[1] = c->use_orb ? reinterpret_cast<uint32_t*>(d + L.fdesc[1]) : nullptr;
  for (int k = 0; k < SDVLB_SEQ_DEPTH; k++)
    st->result[k] = reinterpret_cast<SeqResultHost*>(s->h_result + size_t(k) * s->result_stride);
  e = cudaMemcpyAsync(d, init.data(), init.size(), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    cudaFree(s->d_block);
    cudaFreeHost(s->h_result);
    delete s;
    return sdvlb_set_cuda_error(e, "sequence upload", __FILE__, __LINE__);
  }
  c->seqs.push_back(s);
  *out = s;
  return ensure_step_buffers(c);
}

int sdvlb_seq_destroy(sdvlb_ctx* c, sdvlb_seq* s) {
  if (!s) return 0;
  if (!c) c = s->ctx;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  // drop queued commands of this sequence
  for (size_t i = 0; i < c->seq_cmds.size();) {
    if (c->seq_cmds[i].seq == reinterpret_cast<SeqState*>(s->d_block)) {
      c->seq_cmds.erase(c->seq_cmds.begin() + i);
      c->seq_cmd_frames.erase(c->seq_cmd_frames.begin() + i);
    } else {
      i++;
    }
  }
  c->seqs.erase(std::remove(c->seqs.begin(), c->seqs.end(), s), c->seqs.end());
  cudaFree(s->d_block);
  cudaFreeHost(s->h_result);
  delete s;
  return 0;
}

static void push_cmd(sdvlb_ctx* c, const SeqCmd& cmd, const sdvlb_frame* frame) {
  c->seq_cmds.push_back(cmd);
  c->seq_cmd_frames.push_back(frame);
}

int sdvlb_seq_reset(sdvlb_ctx* c, sdvlb_seq* s, const sdvlb_frame* frame, const double T[7]) {
  if (!c || !s || !frame || !T) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  SeqCmd cmd;
  memset(&cmd, 0, sizeof(cmd));
  cmd.seq = reinterpret_cast<SeqState*>(s->d_block);
  cmd.kind = SEQC_RESET;
  cmd.frame = frame->dev;
  cmd.frame.host_mirror = nullptr;
  memcpy(cmd.T, T, sizeof(cmd.T));
  push_cmd(c, cmd, frame);
  for (int k = 0; k < SDVLB_SEQ_KF_CAP; k++) s->kf_state[k] = 0;
  s->n_bound = 0;
  return 0;
}

int sdvlb_seq_set_policy(sdvlb_ctx* c, sdvlb_seq* s, const sdvlb_seq_policy* policy) {
  if (!c || !s || !policy) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  SeqCmd cmd;
  memset(&cmd, 0, sizeof(cmd));
  cmd.seq = reinterpret_cast<SeqState*>(s->d_block);
  cmd.kind = SEQC_POLICY;
  cmd.policy = *policy;
  push_cmd(c, cmd, nullptr);
  s->policy = *policy;
  return 0;
}

int sdvlb_seq_release(sdvlb_ctx* c, sdvlb_seq* s) {
  if (!c || !s) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  SeqCmd cmd;
  memset(&cmd, 0, sizeof(cmd));
  cmd.seq = reinterpret_cast<SeqState*>(s->d_block);
  cmd.kind = SEQC_RELEASE;
  push_cmd(c, cmd, nullptr);
  return 0;
}

int sdvlb_seq_add_points(sdvlb_ctx* c, sdvlb_seq* s, const sdvlb_frame* kf, const double T_kf[7],
                         const sdvlb_seq_point* pts, int n, int* kf_slot) {
  SDVLB_RANGE("sdvlb.seq_add_points");
  if (!c || !s || !kf || !T_kf || n < 0 || (n > 0 && !pts)) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (n > s->max_feats) return sdvlb_set_error(SDVLB_ERR_OVERFLOW, "more points than the sequence's feature capacity");
  for (int k = 0; k < n; k++)
    if (pts[k].ref_level < 0 || pts[k].ref_level >= c->params.pyramid_levels || pts[k].cur_level < 0 ||
        pts[k].cur_level >= c->params.pyramid_levels)
      return sdvlb_set_error(SDVLB_ERR_ARG, "point level out of range");
  int slot = -1;
  for (int k = 0; k < SDVLB_SEQ_KF_CAP && slot < 0; k++)
    if (s->kf_state[k] == 0) slot = k;
  if (slot < 0) return sdvlb_set_error(SDVLB_ERR_OVERFLOW, "every keyframe slot of the sequence is referenced by live points");
  s->kf_state[slot] = 1;
  SeqCmd cmd;
  memset(&cmd, 0, sizeof(cmd));
  cmd.seq = reinterpret_cast<SeqState*>(s->d_block);
  cmd.kind = SEQC_ADD_POINTS;
  cmd.n = n;
  cmd.kf_slot = slot;
  cmd.kf_pyr = kf->dev.pyr;
  memcpy(cmd.T, T_kf, sizeof(cmd.T));
  cmd.pts = reinterpret_cast<const sdvlb_seq_point*>(c->seq_pts.size());   // index for now, device pointer at submission
  push_cmd(c, cmd, kf);
  c->seq_pts.insert(c->seq_pts.end(), pts, pts + n);
  s->n_bound = std::min(s->max_feats, s->n_bound + n);   // the device truncates at max_feats (and reports it)
  if (kf_slot) *kf_slot = slot;
  return 0;
}

int sdvlb_seq_track_inflight(sdvlb_ctx* c) { return c ? int(c->seq_queue.size()) : 0; }

int sdvlb_seq_track_submit(sdvlb_ctx* c, sdvlb_seq* const* seqs, sdvlb_frame* const* frames, int n) {
  SDVLB_RANGE("sdvlb.seq_track_submit");
  if (!c || !seqs || !frames || n <= 0 || n > SDVLB_SEQ_BATCH)
    return sdvlb_set_error(SDVLB_ERR_ARG, "a sequence submission takes 1..64 sequences");
  if (c->pending.active) return sdvlb_set_error(SDVLB_ERR_STATE, "a tracking batch is still in flight on this context");
  if (int(c->seq_queue.size()) >= SDVLB_SEQ_DEPTH)
    return sdvlb_set_error(SDVLB_ERR_STATE, "SDVLB_SEQ_DEPTH sequence submissions are already in flight: collect one first");
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure_step_buffers(c);
  if (rc) return rc;
  int n_bound = 0;
  for (int i = 0; i < n; i++) {
    if (!seqs[i] || !frames[i] || seqs[i]->ctx != c) return sdvlb_set_error(SDVLB_ERR_ARG, "null / foreign sequence or frame");
    for (int k = 0; k < i; k++)
      if (seqs[k] == seqs[i]) return sdvlb_set_error(SDVLB_ERR_ARG, "a sequence appears twice in one submission");
    if (!(frames[i]->has_corners || (frames[i]->build_pending && frames[i]->build_corners)))
      return sdvlb_set_error(SDVLB_ERR_STATE, "a tracked frame needs corners");
    if (c->use_orb && !(frames[i]->has_desc || (frames[i]->build_pending && frames[i]->build_desc)))
      return sdvlb_set_error(SDVLB_ERR_STATE, "ORB mode: the frame was built without corner descriptors");
    if (c->use_orb && !seqs[i]->has_desc)
      return sdvlb_set_error(SDVLB_ERR_STATE, "ORB mode: the sequence was created before sdvlb_ctx_set_orb(ctx, 1)");
    n_bound = std::max(n_bound, seqs[i]->n_bound);
  }
  // ---- order the tracking stream after the builds of every frame it touches
  {
    cudaEvent_t waited[8];
    int nw = 0;
    auto wait_once = [&](const sdvlb_frame* f) -> int {
      if (!f || !f->build_pending) return 0;
      for (int k = 0; k < nw; k++) if (waited[k] == f->built) return 0;
      if (nw < 8) waited[nw++] = f->built;
      return wait_frame_built(c, f, c->stream);
    };
    for (int i = 0; i < n; i++) { rc = wait_once(frames[i]); if (rc) return rc; }
    for (const sdvlb_frame* f : c->seq_cmd_frames) { rc = wait_once(f); if (rc) return rc; }
  }

  c->track_seq++;
  const int slot = int(c->track_seq % SDVLB_SEQ_DEPTH);

  // ---- queued commands: one upload (the slot's own staging: an earlier submission's copy may still be in flight);
  // applied by the align kernel
  int2 cmd_range[SDVLB_SEQ_BATCH];
  for (int i = 0; i < SDVLB_SEQ_BATCH; i++) { cmd_range[i].x = 0; cmd_range[i].y = 0; }
  const SeqCmd* d_cmds = nullptr;
  const int n_cmds = int(c->seq_cmds.size());
  if (n_cmds > 0) {
    Arena& in = c->seq_in[slot];
    in.used = 0;
    const size_t need = 2560 + size_t(n_cmds) * (sizeof(SeqCmd) + sizeof(int2)) +
                        c->seq_pts.size() * (sizeof(sdvlb_seq_point) + (c->use_orb ? 32 : 0));
    if (need > in.cap) SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));   // growing frees the old buffers
    rc = ensure_arena(&in, need, true);
    if (rc) return rc;
    const size_t o_cmds = in.take(size_t(n_cmds) * sizeof(SeqCmd));
    const size_t o_ranges = in.take(size_t(n_cmds) * sizeof(int2));
    const size_t o_pts = in.take(std::max<size_t>(1, c->seq_pts.size()) * sizeof(sdvlb_seq_point));
    const size_t up_bytes = in.used;   // what follows is device-only scratch
    // ORB mode: Feature::descriptor_ of the new points' init features, written by seq_orb_points_kernel
    const size_t o_pdesc = c->use_orb ? in.take(std::max<size_t>(1, c->seq_pts.size()) * 32) : 0;
    int max_points = 0;
    // commands grouped by sequence, order kept inside a sequence.  Sequences of this step get theirs applied by their
    // own align CTA; the others (not tracked now) by one extra launch.
    std::vector<int> idx(n_cmds);
    for (int k = 0; k < n_cmds; k++) idx[k] = k;
    std::stable_sort(idx.begin(), idx.end(), [c](int a, int b) { return c->seq_cmds[a].seq < c->seq_cmds[b].seq; });
    SeqCmd* hc = reinterpret_cast<SeqCmd*>(in.h + o_cmds);
    int2* hr = reinterpret_cast<int2*>(in.h + o_ranges);
    int n_foreign = 0;
    for (int k = 0; k < n_cmds; k++) {
      const SeqCmd& src = c->seq_cmds[idx[k]];
      hc[k] = src;
      if (src.kind == SEQC_ADD_POINTS) {
        hc[k].pts = reinterpret_cast<const sdvlb_seq_point*>(in.d + o_pts) + reinterpret_cast<size_t>(src.pts);
        hc[k].desc = c->use_orb ? reinterpret_cast<uint32_t*>(in.d + o_pdesc) + reinterpret_cast<size_t>(src.pts) * 8 : nullptr;
        max_points = std::max(max_points, src.n);
      }
      if (k > 0 && src.seq == c->seq_cmds[idx[k - 1]].seq) continue;   // same run as the previous command
      int cnt = 1;
      while (k + cnt < n_cmds && c->seq_cmds[idx[k + cnt]].seq == src.seq) cnt++;
      int owner = -1;
      for (int i = 0; i < n && owner < 0; i++)
        if (reinterpret_cast<SeqState*>(seqs[i]->d_block) == src.seq) owner = i;
      if (owner >= 0) { cmd_range[owner].x = k; cmd_range[owner].y = cnt; }
      else { hr[n_foreign].x = k; hr[n_foreign].y = cnt; n_foreign++; }
    }
    if (!c->seq_pts.empty()) memcpy(in.h + o_pts, c->seq_pts.data(), c->seq_pts.size() * sizeof(sdvlb_seq_point));
    SDVLB_CUDA_TRY(cudaMemcpyAsync(in.d, in.h, up_bytes, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += int64_t(up_bytes);
    d_cmds = reinterpret_cast<const SeqCmd*>(in.d + o_cmds);
    if (c->use_orb && max_points > 0) {
      SDVLB_CUDA_TRY(sdvlb_launch_seq_orb_points(d_cmds, n_cmds, max_points, c->geom, c->stream));
      c->n_launches += 1;
    }
    if (n_foreign > 0) {
      SDVLB_CUDA_TRY(sdvlb_launch_seq_apply(d_cmds, reinterpret_cast<const int2*>(in.d + o_ranges), n_foreign, c->dp, c->stream));
      c->n_launches += 1;
    }
    for (sdvlb_seq* s : c->seqs)
      for (int k = 0; k < SDVLB_SEQ_KF_CAP; k++)
        if (s->kf_state[k] == 1) { s->kf_state[k] = 2; s->kf_seq[k] = c->track_seq; }
    c->seq_cmds.clear();
    c->seq_pts.clear();
    c->seq_cmd_frames.clear();
  }

  // ---- the step: align (+ commands, prior) -> search -> post, each a programmatic dependent of the one before
  SeqStepArgs A;
  A.n = n;
  A.max_feats = std::max(4, n_bound);
  for (int i = 0; i < n; i++) {
    A.seq[i] = reinterpret_cast<SeqState*>(seqs[i]->d_block);
    A.cur[i] = frames[i]->dev;
    A.cur[i].host_mirror = nullptr;
  }
  A.cmds = d_cmds;
  memcpy(A.cmd_range, cmd_range, sizeof(cmd_range));
  A.d_done = c->d_seq_done;
  A.h_flag = reinterpret_cast<uint32_t*>(c->h_overflow) + 16;
  A.seq_no = c->track_seq;
  A.slot = slot;
  A.use_orb = c->use_orb ? 1 : 0;
  A.pad_ = 0;
  A.dp = c->dp;
  A.g = c->geom;
  timer_begin(c, SDVLB_K_ALIGN);
  SDVLB_CUDA_TRY(sdvlb_launch_seq_align(A, n_bound, c->stream));
  timer_end(c);
  timer_begin(c, SDVLB_K_SEARCH);
  SDVLB_CUDA_TRY(sdvlb_launch_search_seq(A, c->stream));
  timer_end(c);
  timer_begin(c, SDVLB_K_POSE);
  SDVLB_CUDA_TRY(sdvlb_launch_seq_post(A, c->stream));
  timer_end(c);
  c->n_launches += 3;   // the post kernel publishes the completion word itself
  // a tracked frame leaves at most max_matches features behind (feature_align.cc:108); a frame that tracking quality
  // rejects leaves the list as it was
  for (int i = 0; i < n; i++)
    if (!seqs[i]->policy.tracking_quality) seqs[i]->n_bound = std::min(seqs[i]->n_bound, c->params.max_matches);
  c->seq_queue.emplace_back();
  SeqSubmission& sub = c->seq_queue.back();
  sub.seq_no = c->track_seq;
  sub.slot = slot;
  sub.seqs.assign(seqs, seqs + n);
  sub.frames.assign(frames, frames + n);
  return 0;
}

int sdvlb_seq_track_poll(sdvlb_ctx* c) {
  if (!c || c->seq_queue.empty()) return sdvlb_set_error(SDVLB_ERR_STATE, "nothing was submitted on this context");
  const volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(c->h_overflow) + 16;
  return int32_t(*done - c->seq_queue.front().seq_no) >= 0 ? 1 : 0;
}

int sdvlb_seq_track_collect(sdvlb_ctx* c, sdvlb_seq_result* results) {
  SDVLB_RANGE("sdvlb.seq_track_collect");
  if (!c || !results) return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (c->seq_queue.empty()) return sdvlb_set_error(SDVLB_ERR_STATE, "nothing was submitted on this context");
  const SeqSubmission sub = std::move(c->seq_queue.front());
  c->seq_queue.pop_front();
  int rc = wait_signal(c, sub.seq_no);
  if (rc) return rc;
  rc = check_overflow(c);
  if (rc) return rc;
  const int pa = c->params.align_patch_size * c->params.align_patch_size;
  for (size_t i = 0; i < sub.seqs.size(); i++) {
    sdvlb_seq* s = sub.seqs[i];
    sdvlb_frame* f = sub.frames[i];
    if (f->build_pending) finalize_build(c, f);   // the tracking stream ran after this frame's build
    if (f->overflowed) return sdvlb_set_error(SDVLB_ERR_OVERFLOW, "corner capacity exceeded in a tracked frame's FAST selection");
    const SeqResultHost* R = reinterpret_cast<const SeqResultHost*>(s->h_result + size_t(sub.slot) * s->result_stride);
    if (R->error) return sdvlb_set_error(SDVLB_ERR_OVERFLOW, "a sequence exceeded its feature capacity (points were dropped)");
    sdvlb_seq_result& o = results[i];
    memset(&o, 0, sizeof(o));
    o.status = R->status;
    o.lost_frames = R->lost_frames;
    if (R->status != SDVLB_SEQ_TRACKED) continue;   // held / idle: the frame was not consumed, nothing else is valid
    memcpy(o.pose, R->pose, sizeof(o.pose));
    o.n_tracked = R->stats[0] / pa;
    o.matches = R->stats[1];
    o.attempts = R->stats[2];
    o.inliers = R->stats[3];
    o.outliers = R->stats[4];
    o.n_points = R->stats[5];
    o.gn_iters = R->stats[6];
    o.n_feats = R->stats[7];
    o.quality = R->quality;
    o.need_keyframe = R->need_keyframe;
    o.feats = reinterpret_cast<const sdvlb_seq_feat*>(reinterpret_cast<const uint8_t*>(R) + sizeof(SeqResultHost));
    memcpy(o.kf_live, R->kf_live, sizeof(o.kf_live));
    memcpy(o.phase_cycles, R->phase_cycles, sizeof(o.phase_cycles));
    memcpy(o.align_cycles, R->align_cycles, sizeof(o.align_cycles));
    // keyframe slots: a result only speaks for the slots whose points had reached the device when its step ran
    for (int k = 0; k < SDVLB_SEQ_KF_CAP; k++) {
      if (s->kf_state[k] < 2 || int32_t(sub.seq_no - s->kf_seq[k]) < 0) continue;
      s->kf_state[k] = R->kf_live[k] > 0 ? 3 : 0;
    }
    c->d2h_bytes += int64_t(sizeof(SeqResultHost)) + int64_t(o.n_feats) * int64_t(sizeof(sdvlb_seq_feat));
    // With nothing else in flight the host now knows the sequence's feature count exactly (the list of an adopted
    // frame is its found features) instead of the conservative max_matches: the next launch is sized for it (the
    // ImageAlign kernel's shared-memory cache is 256 or 512 features, the search grid follows the bound).
    if (c->seq_queue.empty() && R->quality != SDVLB_TRACKING_BAD) {
      int nb = R->stats[7];
      for (const SeqCmd& cmd : c->seq_cmds) {   // queued, not yet submitted
        if (cmd.seq != reinterpret_cast<SeqState*>(s->d_block)) continue;
        if (cmd.kind == SEQC_RESET) nb = 0;
        else if (cmd.kind == SEQC_ADD_POINTS) nb += cmd.n;
      }
      s->n_bound = std::min(s->max_feats, nb);
    }
  }
  return 0;
}

}  // extern "C"
