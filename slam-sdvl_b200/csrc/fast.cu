// fast.cu — K2: FAST-9/16 corner detection per 32x32 cell and SDVL's per-cell quota selection, bit-exact with
// FastDetector::DetectPyramid / SelectPixels (extra/fast_detector.cc:58-175), i.e. with cv::FAST(roi, thr, nms=true)
// and cv::KeyPointsFilter::retainBest as SDVL calls them.
//
//  fast_cells_kernel : one CTA per (cell, frame). The cell's ROI [max(m,32i), min(rows-m,32i+32)) is staged in shared
//                      memory (cv::FAST never reads outside the ROI, so cells are independent and the outer 3 px of
//                      every ROI never fire). Phase 1: 16-bit bright/dark ring masks, 9-contiguous test. Phase 2:
//                      exact cornerScore for candidates only. Phase 3: strict 3x3 NMS, raster-ordered compaction.
//  fast_select_kernel: one CTA per (level, frame): water-filling quota (fast_detector.cc:114-135), per-cell
//                      retainBest, level-wide retainBest, and — by the last CTA of a frame — concatenation of the
//                      levels into Frame::corners_ order.
#include "common.cuh"
#include "select_impl.h"

namespace {

constexpr int DET_THREADS = 128;
constexpr int SEL_THREADS = 256;
constexpr int SEL_SMEM_KEYS = 4096;   // level-wide retainBest runs in shared memory up to this many keypoints

__device__ __forceinline__ bool has9(uint32_t m) {  // 9 contiguous set bits in a circular 16-bit mask
  uint32_t mm = m | (m << 16);
  uint32_t x = mm & (mm >> 1);
  x &= x >> 2;
  x &= x >> 4;       // 8 contiguous
  x &= mm >> 8;      // 9 contiguous
  return (x & 0xFFFFu) != 0;
}

// ring offsets in OpenCV order (dx, dy)
__constant__ int8_t c_ring[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                     {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

constexpr int TS = 36;  // shared tile row stride (bytes)

__global__ void __launch_bounds__(DET_THREADS) fast_cells_kernel(const __grid_constant__ FrameBatch B,
                                                                 const __grid_constant__ FastArgs A,
                                                                 uint32_t* __restrict__ cell_kp,
                                                                 int32_t* __restrict__ cell_cnt) {
  __shared__ __align__(16) uint8_t tile[32 * TS];
  __shared__ int16_t s_score[32 * 32];
  __shared__ uint16_t s_cand[32 * 32];
  __shared__ int s_ncand;
  __shared__ int s_warp_cnt[DET_THREADS / 32];
  __shared__ int s_base;

  const int tid = threadIdx.x;
  const int frame = blockIdx.y;
  int cell = blockIdx.x;
  int level = 0;
  while (level + 1 < A.n_fast_levels && cell >= A.g.cell_off[level + 1]) level++;
  cell -= A.g.cell_off[level];
  const int wc = A.g.wcells[level];
  const int ci = cell / wc, cj = cell - ci * wc;
  const int W = A.g.w[level], H = A.g.h[level];
  const int out_idx = (B.scratch_base + frame) * A.g.total_cells + blockIdx.x;

  const int inity = max(A.margin, ci * SDVLB_CELL), maxy = min(H - A.margin, ci * SDVLB_CELL + SDVLB_CELL);
  const int initx = max(A.margin, cj * SDVLB_CELL), maxx = min(W - A.margin, cj * SDVLB_CELL + SDVLB_CELL);
  if (maxy <= inity || maxx <= initx) {   // `continue` in the reference: cell neither empty nor populated
    if (tid == 0) cell_cnt[out_idx] = -1;
    return;
  }
  const int cols = maxx - initx, rows = maxy - inity;
  if (cols < 7 || rows < 7) {             // cv::FAST finds nothing in ROIs narrower than the ring
    if (tid == 0) cell_cnt[out_idx] = 0;
    return;
  }
  const uint8_t* __restrict__ img = B.f[frame].pyr + A.g.off[level];

  // ---- stage ROI
  if (((initx & 3) == 0) && ((W & 3) == 0)) {
    const int wpr = (cols + 3) >> 2;
    for (int i = tid; i < rows * wpr; i += DET_THREADS) {
      const int r = i / wpr, c = i - r * wpr;
      const uint8_t* p = img + size_t(inity + r) * W + initx + 4 * c;
      uint32_t v;
      if (initx + 4 * c + 3 < W) v = __ldg(reinterpret_cast<const uint32_t*>(p));
      else {
        v = 0;
        for (int k = 0; k < 4; k++)
          if (initx + 4 * c + k < W) v |= uint32_t(__ldg(p + k)) << (8 * k);
      }
      *reinterpret_cast<uint32_t*>(&tile[r * TS + 4 * c]) = v;
    }
  } else {
    for (int i = tid; i < rows * cols; i += DET_THREADS) {
      const int r = i / cols, c = i - r * cols;
      tile[r * TS + c] = __ldg(img + size_t(inity + r) * W + initx + c);
    }
  }
  for (int i = tid; i < 32 * 32; i += DET_THREADS) s_score[i] = 0;
  if (tid == 0) s_ncand = 0;
  __syncthreads();

  // ---- phase 1: candidate test on the ROI interior
  const int tw = cols - 6, th = rows - 6;
  const int npix = tw * th;
  int off[16];
#pragma unroll
  for (int k = 0; k < 16; k++) off[k] = c_ring[k][1] * TS + c_ring[k][0];
  const int t = A.threshold;
  for (int p = tid; p < npix; p += DET_THREADS) {
    const int y = p / tw + 3, x = p - (p / tw) * tw + 3;
    const uint8_t* c = &tile[y * TS + x];
    const int v = c[0];
    uint32_t bright = 0, dark = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int r = c[off[k]];
      bright |= uint32_t(r > v + t) << k;
      dark |= uint32_t(r < v - t) << k;
    }
    if (has9(bright) || has9(dark)) {
      const int slot = atomicAdd(&s_ncand, 1);
      s_cand[slot] = uint16_t(y * 32 + x);
    }
  }
  __syncthreads();

  // ---- phase 2: exact score for candidates (cornerScore<16>: max over arcs of min(d) / min(-d), minus 1)
  const int ncand = s_ncand;
  for (int i = tid; i < ncand; i += DET_THREADS) {
    const int yx = s_cand[i];
    const int y = yx >> 5, x = yx & 31;
    const uint8_t* c = &tile[y * TS + x];
    const int v = c[0];
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - int(c[off[k]]);
    int lo2[16], hi2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
    int lo4[16], hi4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
    int a0 = t, b0 = -t;   // a0 = max(thr, S+), b0 = min(-thr, -S-)
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int lo9 = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]);
      const int hi9 = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
      a0 = max(a0, lo9);
      b0 = min(b0, hi9);
    }
    const int score = max(a0, -b0) - 1;
    s_score[yx] = int16_t(score);
  }
  __syncthreads();

  // ---- phase 3: strict NMS against the 8 neighbours (non-corners are 0), raster-ordered compaction
  uint32_t* __restrict__ out = cell_kp + size_t(out_idx) * SDVLB_CELL_CAP;
  const int ox = initx - cj * SDVLB_CELL, oy = inity - ci * SDVLB_CELL;   // ROI origin relative to the cell origin
  if (tid == 0) s_base = 0;
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int p0 = 0; p0 < npix; p0 += DET_THREADS) {
    const int p = p0 + tid;
    bool keep = false;
    int s = 0, x = 0, y = 0;
    if (p < npix) {
      y = p / tw + 3; x = p - (p / tw) * tw + 3;
      const int16_t* c = &s_score[y * 32 + x];
      s = c[0];
      keep = s > 0 && s >= t && s > c[-1] && s > c[1] && s > c[-33] && s > c[-32] && s > c[-31] && s > c[31] &&
             s > c[32] && s > c[33];
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int base = s_base;
    for (int w = 0; w < warp; w++) base += s_warp_cnt[w];
    if (keep) {
      const int slot = base + __popc(bal & ((1u << lane) - 1));
      if (slot < SDVLB_CELL_CAP) out[slot] = (uint32_t(s) << 10) | (uint32_t(y + oy) << 5) | uint32_t(x + ox);
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < DET_THREADS / 32; w++) tot += s_warp_cnt[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) cell_cnt[out_idx] = s_base;   // <= 169 by construction of NMS
}

// ------------------------------------------------------------------------------------------------ selection
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* s_tmp) {   // all threads get the result
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_tmp[warp] = v;
  __syncthreads();
  T r = 0;
  for (int w = 0; w < SEL_THREADS / 32; w++) r += s_tmp[w];
  return r;
}

__global__ void __launch_bounds__(SEL_THREADS) fast_select_kernel(const __grid_constant__ FrameBatch B,
                                                                  const __grid_constant__ FastArgs A,
                                                                  uint32_t* __restrict__ cell_kp,
                                                                  const int32_t* __restrict__ cell_cnt,
                                                                  uint32_t* __restrict__ level_kp,
                                                                  int32_t* __restrict__ level_cnt,
                                                                  int32_t* __restrict__ frame_ticket) {
  extern __shared__ int s_dyn[];   // nleft[ncells], nsel[ncells], kept_off[ncells+1]
  __shared__ int s_tmp[SEL_THREADS / 32];
  __shared__ uint32_t s_keys[SEL_SMEM_KEYS];
  __shared__ int s_final, s_ticket;

  const int tid = threadIdx.x;
  const int level = blockIdx.x, frame = B.scratch_base + blockIdx.y;   // `frame` indexes the scratch arrays
  const int ncells = A.g.wcells[level] * A.g.hcells[level];
  int* nleft = s_dyn;
  int* nsel = s_dyn + ncells;
  int* koff = s_dyn + 2 * ncells;
  const int cbase = frame * A.g.total_cells + A.g.cell_off[level];
  const int nfeatures = A.nfeat[level];

  // ---- counts (fast_detector.cc:97-105)
  int my_empty = 0;
  for (int c = tid; c < ncells; c += SEL_THREADS) {
    const int n = cell_cnt[cbase + c];
    nleft[c] = n > 0 ? n : 0;
    nsel[c] = 0;
    if (n == 0) my_empty++;
  }
  const int nempty = block_sum<int>(my_empty, s_tmp);

  // ---- water-filling (fast_detector.cc:108-135)
  int selected = 0;
  int cells_left = ncells - nempty;
  while ((nfeatures - selected) > 0 && cells_left > 0) {
    const int npercell = int(ceil(double(nfeatures - selected) / double(cells_left)));
    int add = 0, cl = 0;
    for (int c = tid; c < ncells; c += SEL_THREADS) {
      const int nl = nleft[c];
      if (nl > 0) {
        if (nl > npercell) { nsel[c] += npercell; add += npercell; nleft[c] = nl - npercell; cl++; }
        else { nsel[c] += nl; add += nl; nleft[c] = 0; }
      }
    }
    selected += block_sum<int>(add, s_tmp);
    cells_left = block_sum<int>(cl, s_tmp);
  }
  __syncthreads();

  // ---- per-cell retainBest (fast_detector.cc:138-140), in place in the cell scratch
  for (int c = tid; c < ncells; c += SEL_THREADS) {
    const int n = cell_cnt[cbase + c];
    int kept = 0;
    if (n > 0) kept = sdvlb_sel::retain_best<10>(cell_kp + size_t(cbase + c) * SDVLB_CELL_CAP, n, nsel[c]);
    nleft[c] = kept;   // reuse as kept count
  }
  __syncthreads();

  // ---- exclusive scan of kept counts in cell order
  {
    const int per = (ncells + SEL_THREADS - 1) / SEL_THREADS;
    const int c0 = tid * per, c1 = min(ncells, c0 + per);
    int local = 0;
    for (int c = c0; c < c1; c++) local += nleft[c];
    // block exclusive scan of `local`
    const int lane = tid & 31, warp = tid >> 5;
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; w++) wbase += s_tmp[w];
    int run = wbase + incl - local;
    for (int c = c0; c < c1; c++) { koff[c] = run; run += nleft[c]; }
    if (tid == SEL_THREADS - 1) koff[ncells] = run;
    __syncthreads();
  }
  const int total = koff[ncells];
  uint32_t* __restrict__ lk = level_kp + size_t(frame) * A.level_kp_total + A.level_kp_off[level];
  const bool use_smem = total <= SEL_SMEM_KEYS;
  const bool fits = total <= A.level_cap[level];
  if (!fits && tid == 0) *reinterpret_cast<volatile int32_t*>(A.overflow_flag) = 1;

  // ---- gather to the level list (fast_detector.cc:141-142): key = score<<22 | y<<11 | x (level coordinates)
  const int wc = A.g.wcells[level];
  if (fits) {
    for (int c = tid; c < ncells; c += SEL_THREADS) {
      const int k = nleft[c];
      const uint32_t* src = cell_kp + size_t(cbase + c) * SDVLB_CELL_CAP;
      const int ci = c / wc, cj = c - ci * wc;
      for (int i = 0; i < k; i++) {
        const uint32_t v = src[i];
        const uint32_t key = ((v >> 10) << 22) | (uint32_t(ci * SDVLB_CELL + ((v >> 5) & 31)) << 11) |
                             uint32_t(cj * SDVLB_CELL + (v & 31));
        if (use_smem) s_keys[koff[c] + i] = key;
        else lk[koff[c] + i] = key;
      }
    }
  }
  __syncthreads();

  // ---- level-wide retainBest (fast_detector.cc:146-148)
  if (tid == 0) {
    int fin = fits ? total : 0;
    if (fits && total > nfeatures) fin = sdvlb_sel::retain_best<22>(use_smem ? s_keys : lk, total, nfeatures);
    s_final = fin;
  }
  __syncthreads();
  const int fin = s_final;
  if (use_smem)
    for (int i = tid; i < fin; i += SEL_THREADS) lk[i] = s_keys[i];
  if (tid == 0) level_cnt[frame * SDVLB_MAX_LEVELS + level] = fin;

  // ---- last CTA of this frame concatenates the levels (fast_detector.cc:150-151,170-173)
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(&frame_ticket[frame], 1);
  __syncthreads();
  if (s_ticket != A.n_fast_levels - 1) return;
  __threadfence();
  const FrameDev& fr = B.f[blockIdx.y];
  int4* const mirror = reinterpret_cast<int4*>(fr.host_mirror);   // header at [0], records from [1]
  int base = 0;
  for (int l = 0; l < A.n_fast_levels; l++) {
    const int n = *reinterpret_cast<volatile int32_t*>(&level_cnt[frame * SDVLB_MAX_LEVELS + l]);
    const uint32_t* src = level_kp + size_t(frame) * A.level_kp_total + A.level_kp_off[l];
    if (base + n > A.corner_cap) {
      if (tid == 0) *reinterpret_cast<volatile int32_t*>(A.overflow_flag) = 1;
      break;
    }
    for (int i = tid; i < n; i += SEL_THREADS) {
      const uint32_t key = __ldcg(src + i);
      const int4 rec = make_int4(int32_t(key & 2047), int32_t((key >> 11) & 2047), l, int32_t(key >> 22));
      fr.corners[base + i] = rec;
      if (mirror && base + i < fr.mirror_cap) mirror[1 + base + i] = rec;
    }
    base += n;
  }
  if (tid == 0) {
    *fr.n_corners = base;
    if (mirror) mirror[0] = make_int4(base, 0, 0, 0);
    frame_ticket[frame] = 0;
  }
}

}  // namespace

// Fills per-level budgets and scratch layout. nfeatures = Frame::CreateCorners budget.
void sdvlb_fast_plan(const PyrGeom& g, const sdvlb_params& p, int nfeatures, int corner_cap, FastPlan* plan) {
  FastArgs& A = plan->args;
  A.g = g;
  A.n_fast_levels = p.max_fast_levels;
  A.margin = 1 + p.patch_size / 2;
  A.threshold = p.fast_threshold < 0 ? 0 : (p.fast_threshold > 255 ? 255 : p.fast_threshold);
  A.corner_cap = corner_cap;
  // fast_detector.cc:160-173
  const double scale = 1.2;
  double factor = 1.0, val = 0.0;
  for (int i = 0; i < p.max_fast_levels; i++) { val += factor; factor /= scale; }
  int lf = int(nfeatures / val);
  int off = 0;
  plan->max_cells_level = 0;
  for (int l = 0; l < p.max_fast_levels; l++) {
    A.nfeat[l] = lf;
    lf = int(lf / scale);
    const int nc = g.wcells[l] * g.hcells[l];
    A.level_cap[l] = nc * SDVLB_CELL_CAP;
    A.level_kp_off[l] = off;
    off += A.level_cap[l];
    if (nc > plan->max_cells_level) plan->max_cells_level = nc;
  }
  A.level_kp_total = off;
  plan->nfeatures = nfeatures;
}

cudaError_t sdvlb_launch_fast_cells(const FrameBatch& B, const FastPlan& plan, uint32_t* cell_kp, int32_t* cell_cnt,
                                    cudaStream_t stream) {
  const FastArgs& A = plan.args;
  dim3 g1(A.g.total_cells, B.n);
  fast_cells_kernel<<<g1, DET_THREADS, 0, stream>>>(B, A, cell_kp, cell_cnt);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_fast_select(const FrameBatch& B, const FastPlan& plan, uint32_t* cell_kp, int32_t* cell_cnt,
                                     uint32_t* level_kp, int32_t* level_cnt, int32_t* frame_ticket,
                                     cudaStream_t stream) {
  const FastArgs& A = plan.args;
  dim3 g2(A.n_fast_levels, B.n);
  const size_t dyn = size_t(3 * plan.max_cells_level + 1) * sizeof(int);
  fast_select_kernel<<<g2, SEL_THREADS, dyn, stream>>>(B, A, cell_kp, cell_cnt, level_kp, level_cnt, frame_ticket);
  return cudaGetLastError();
}
