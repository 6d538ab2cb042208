// filter.cu — Frame::FilterCorners (frame.cc:133-146) -> FastDetector::FilterCorners (extra/fast_detector.cc:177-218) with
// FindShiTomasiScoreAtPoint (extra/utils.cc:61-97), on the frame's device-resident pyramid and corner list
// (the keyframe-time step right after FAST).
//
// One CTA per frame.  Phase 1: one thread per corner computes the reference's margin test, its grid cell and its
// Shi-Tomasi score (integer-valued float sums below 2^24, so any order is exact; then the reference's float / double
// mix).  Phase 2: the reference keeps, per cell, (index, int(score)) and lets a later corner win when its double score
// exceeds the stored *truncated* one — order dependent, so one thread per cell replays its corners in index order.
// Phase 3: cells above min_feature_score are emitted in cell order (ordered compaction).
#include <algorithm>
#include <cstring>
#include <vector>

#include "capi_internal.h"

using namespace sdvlb_detail;

namespace {

constexpr int FC_THREADS = 512;

struct FilterArgs {
  FrameDev f;
  PyrGeom g;
  int margin;              // 1 + patch_size / 2, or 4 + orb_size / 2 with Config::UseORB()
  int min_score;           // Config::MinFeatureScore()
  int n_cells, gw;
  const uint8_t* locked;   // n_cells bytes: FastDetector::grid_mask_
  double* score;           // corner_cap doubles
  int32_t* cell;           // corner_cap ints (-1: skipped)
  int32_t* out;            // out[0] = count, then the indices (pinned host memory)
};

__device__ __forceinline__ double shi_tomasi(const uint8_t* __restrict__ img, int cols, int rows, int px, int py) {
  const int halfbox_size = 4, box_size = 8, box_area = 64;
  const int x_min = px - halfbox_size, x_max = px + halfbox_size;
  const int y_min = py - halfbox_size, y_max = py + halfbox_size;
  if (x_min < 1 || x_max >= cols - 1 || y_min < 1 || y_max >= rows - 1) return 0.0;
  int sxx = 0, syy = 0, sxy = 0;   // exact: |dx|,|dy| <= 255, 64 terms
  for (int y = y_min; y < y_max; y++) {
    const uint8_t* __restrict__ row = img + size_t(y) * cols + x_min;
#pragma unroll
    for (int x = 0; x < box_size; x++) {
      const int dx = int(__ldg(row + x + 1)) - int(__ldg(row + x - 1));
      const int dy = int(__ldg(row + x + cols)) - int(__ldg(row + x - cols));
      sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
    }
  }
  const float dXX = float(double(float(sxx)) / (2.0 * box_area));
  const float dYY = float(double(float(syy)) / (2.0 * box_area));
  const float dXY = float(double(float(sxy)) / (2.0 * box_area));
  const float tr = __fadd_rn(dXX, dYY);
  const float disc = __fsub_rn(__fmul_rn(tr, tr), __fmul_rn(4.0f, __fsub_rn(__fmul_rn(dXX, dYY), __fmul_rn(dXY, dXY))));
  return 0.5 * double(__fsub_rn(tr, __fsqrt_rn(disc)));
}

__global__ void __launch_bounds__(FC_THREADS) filter_corners_kernel(const __grid_constant__ FilterArgs A) {
  __shared__ int s_cnt[FC_THREADS / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = *A.f.n_corners;
  for (int i = tid; i < n; i += FC_THREADS) {
    const int4 c = A.f.corners[i];
    const int cols = A.g.w[c.z], rows = A.g.h[c.z];
    int pos = -1;
    double sc = 0.0;
    if (!(c.x < A.margin || c.y < A.margin || c.x >= cols - A.margin || c.y >= rows - A.margin)) {
      pos = ((c.y << c.z) / SDVLB_CELL) * A.gw + ((c.x << c.z) / SDVLB_CELL);
      if (A.locked[pos]) pos = -1;
      else sc = shi_tomasi(A.f.pyr + A.g.off[c.z], cols, rows, c.x, c.y);
    }
    A.cell[i] = pos;
    A.score[i] = sc;
  }
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < A.n_cells; c0 += FC_THREADS) {
    const int c = c0 + tid;
    int best = 0, stored = A.min_score;
    if (c < A.n_cells) {
      for (int i = 0; i < n; i++) {   // warp-uniform loads, the cell's corners in index order
        if (A.cell[i] != c) continue;
        const double sc = A.score[i];
        if (sc > double(stored)) { best = i; stored = int(sc); }
      }
    }
    const bool emit = c < A.n_cells && stored > A.min_score;
    const unsigned bal = __ballot_sync(0xffffffffu, emit);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; w++) off += s_cnt[w];
    if (emit) A.out[1 + off + __popc(bal & ((1u << lane) - 1))] = best;
    __syncthreads();
    if (tid == 0) { for (int w = 0; w < FC_THREADS / 32; w++) s_base += s_cnt[w]; }
    __syncthreads();
  }
  if (tid == 0) A.out[0] = s_base;
}

}  // namespace

extern "C" int sdvlb_frame_filter_corners(sdvlb_ctx* c, const sdvlb_frame* f, const double* locked_px, int n_locked,
                                          int min_feature_score, int32_t* indices, int cap, int* n_out) {
  if (!c || !f || n_locked < 0 || (n_locked > 0 && !locked_px) || !indices || !n_out)
    return sdvlb_set_error(SDVLB_ERR_ARG, "bad argument");
  if (c->pending.active || !c->seq_queue.empty()) return sdvlb_set_error(SDVLB_ERR_STATE, "a submission is still in flight on this context");
  SDVLB_CUDA_TRY(cudaSetDevice(c->device));
  sdvlb_frame* mf = const_cast<sdvlb_frame*>(f);
  if (mf->build_pending) {
    SDVLB_CUDA_TRY(cudaEventSynchronize(mf->built));
    finalize_build(c, mf);
  }
  if (!f->has_corners) return sdvlb_set_error(SDVLB_ERR_STATE, "corners were not detected on this frame");
  const int gw = c->geom.wcells[0], gh = c->geom.hcells[0], n_cells = gw * gh;
  Arena& in = c->in;
  in.used = 0;
  const size_t need = 4096 + size_t(n_cells) * 5 + size_t(c->corner_cap) * 12 + 1024;
  int rc = ensure_arena(&in, need, true);
  if (rc) return rc;
  const size_t o_mask = in.take(size_t(n_cells));
  const size_t o_out = in.take(size_t(n_cells + 1) * sizeof(int32_t));
  const size_t up = in.used;
  const size_t o_score = in.take(size_t(c->corner_cap) * sizeof(double));
  const size_t o_cell = in.take(size_t(c->corner_cap) * sizeof(int32_t));
  memset(in.h + o_mask, 0, size_t(n_cells));
  for (int i = 0; i < n_locked; i++) {   // FastDetector::LockCell (fast_detector.cc:46-49)
    const int idx = int(locked_px[2 * i + 1] / c->params.cell_size) * gw + int(locked_px[2 * i] / c->params.cell_size);
    if (idx < 0 || idx >= n_cells) return sdvlb_set_error(SDVLB_ERR_ARG, "locked position outside the image");
    in.h[o_mask + idx] = 1;
  }
  SDVLB_CUDA_TRY(cudaMemcpyAsync(in.d + o_mask, in.h + o_mask, size_t(n_cells), cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes += int64_t(n_cells);
  FilterArgs A;
  A.f = f->dev;
  A.g = c->geom;
  A.margin = c->use_orb ? 4 + 31 / 2 : 1 + c->params.patch_size / 2;   // fast_detector.cc:183-186
  A.min_score = min_feature_score;
  A.n_cells = n_cells;
  A.gw = gw;
  A.locked = in.d + o_mask;
  A.score = reinterpret_cast<double*>(in.d + o_score);
  A.cell = reinterpret_cast<int32_t*>(in.d + o_cell);
  A.out = reinterpret_cast<int32_t*>(in.d + o_out);
  SDVLB_PREPARE(filter_corners_kernel, 0);
  filter_corners_kernel<<<1, FC_THREADS, 0, c->stream>>>(A);
  SDVLB_CUDA_TRY(cudaGetLastError());
  c->n_launches += 1;
  SDVLB_CUDA_TRY(cudaMemcpyAsync(in.h + o_out, in.d + o_out, size_t(n_cells + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                 c->stream));
  SDVLB_CUDA_TRY(cudaStreamSynchronize(c->stream));
  (void)up;
  const int32_t* ho = reinterpret_cast<const int32_t*>(in.h + o_out);
  const int n = ho[0];
  c->d2h_bytes += int64_t(n + 1) * 4;
  *n_out = n;
  for (int i = 0; i < n && i < cap; i++) indices[i] = ho[1 + i];
  return 0;
}
