// fast.cu — K2: FAST-9/16 corner detection per 32x32 cell and SDVL's per-cell quota selection, bit-exact with
// FastDetector::DetectPyramid / SelectPixels (extra/fast_detector.cc:58-175), i.e. with cv::FAST(roi, thr, nms=true)
// and cv::KeyPointsFilter::retainBest as SDVL calls them.
//
//  fast_cells_kernel : one warp per (cell, frame). The cell's ROI [max(m,32i), min(rows-m,32i+32)) is staged in shared
//                      memory (cv::FAST never reads outside the ROI, so cells are independent and the outer 3 px of
//                      every ROI never fire). One warp per cell: exact cornerScore of every tested pixel with 16-bit
//                      SIMD min/max (two pixels per register), strict 3x3 NMS on the corners, raster-ordered output.
//  fast_select_kernel: one CTA per (level, frame): water-filling quota (fast_detector.cc:114-135), per-cell
//                      retainBest, level-wide retainBest, and — by the last CTA of a frame — concatenation of the
//                      levels into Frame::corners_ order.
#include <algorithm>

#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "select_warp.cuh"

namespace {

constexpr int DET_WARPS = 4;          // one warp per 32x32 cell, no block-level synchronisation at all
constexpr int DET_THREADS = DET_WARPS * 32;
constexpr int SEL_THREADS = 256;
constexpr int SEL_SMEM_KEYS = 2048;   // level-wide retainBest runs in shared memory up to this many keypoints
constexpr int SEL_PART_CELLS = 32;    // cells of one level per CTA of the selection kernel (4 per warp)
constexpr int TSE = 40;               // score tile row stride in elements: one BYTE per pixel (a score is at most 254), so
                                      // that a CTA needs 36.6 instead of 41.7 KB and six instead of five share an SM
constexpr int TSH = TSE / 2;          // ... in 16-bit halves (one half = one horizontally adjacent pixel pair)
// Pixel tile row stride in words.  A row holds 34 pixels = 18 words with the reach of the ring; 23 is the smallest odd
// stride at which the quads of four consecutive tile rows that one warp pass of the quick test touches spread best
// over the banks (1.67 wavefronts per load; 45, the stride of the pair-per-lane quick test, gives 2.0) -- and, what
// matters more, a CTA then needs 25.3 instead of 36.6 KB, so eight instead of six share an SM: 68.2 -> 64.5 us per
// 64 frames (stride 39, same banks, six CTAs: 68.0).  Ten CTAs per SM (candidates aliased into the tail of the corner
// list, score rows of 32 bytes, 48 registers) measured 64.2: not kept.
constexpr int PSW = 23;
constexpr int PSE = 2 * PSW;
constexpr int LIST_CAP = 26 * 26 + 28;

// upper half of a | lower half of b << 16: the pixel pair that starts one element to the right of word a
__device__ __forceinline__ uint32_t mid_pair(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5432); }

// FAST-9/16 on 32x32 cells, one WARP per cell.  Scoring is branch-free: every lane scores one horizontally adjacent
// PIXEL PAIR held as two 16-bit lanes of a register (VIMNMX3.S16x2 / VIADD.16x2 are native on sm_100a):
//   S1 = v - min over the 16 arcs of (max of the 9 arc pixels)      centre brighter than a whole arc
//   S2 = (max over the 16 arcs of (min of the 9 arc pixels)) - v    centre darker than a whole arc
//   corner <=> max(S1, S2) > threshold, score = max(S1, S2) - 1     == cv::FAST's cornerScore<16> for corners
// arc minima/maxima come from 3-input min/max: x3[k] = op(r[k], r[k+1], r[k+2]), x9[k] = op(x3[k], x3[k+3], x3[k+6]).
// The cell is staged in shared memory as 16-bit pixels at element x+1 (tested pixels start at x = 3, so pairs are
// word-aligned).  Corners are appended to a raster-ordered list with warp ballots; the strict 3x3 NMS then only visits
// listed corners and compacts the survivors, again in raster order (the order cv::FAST emits keypoints in).
__global__ void __launch_bounds__(DET_THREADS) fast_cells_kernel(const __grid_constant__ FrameBatch B,
                                                                 const __grid_constant__ FastArgs A,
                                                                 uint32_t* __restrict__ cell_kp,
                                                                 int32_t* __restrict__ cell_cnt) {
  __shared__ __align__(16) uint32_t s_tile[DET_WARPS][32 * PSW];
  __shared__ __align__(16) uint16_t s_score[DET_WARPS][32 * TSH];
  __shared__ uint16_t s_list[DET_WARPS][LIST_CAP];
  __shared__ uint16_t s_cand[DET_WARPS][13 * 26 + 14];   // pixel pairs that pass the quick test, raster order

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int frame = blockIdx.y;
  const int gcell = blockIdx.x * DET_WARPS + warp;
  if (gcell >= A.g.total_cells) return;
  int cell = gcell;
  int level = 0;
  while (level + 1 < A.n_fast_levels && cell >= A.g.cell_off[level + 1]) level++;
  cell -= A.g.cell_off[level];
  const int wc = A.g.wcells[level];
  const int ci = cell / wc, cj = cell - ci * wc;
  const int W = A.g.w[level], H = A.g.h[level];
  const int out_idx = (B.scratch_base + frame) * A.g.total_cells + gcell;

  const int inity = max(A.margin, ci * SDVLB_CELL), maxy = min(H - A.margin, ci * SDVLB_CELL + SDVLB_CELL);
  const int initx = max(A.margin, cj * SDVLB_CELL), maxx = min(W - A.margin, cj * SDVLB_CELL + SDVLB_CELL);
  if (maxy <= inity || maxx <= initx) {   // `continue` in the reference: cell neither empty nor populated
    if (lane == 0) cell_cnt[out_idx] = -1;
    return;
  }
  const int cols = maxx - initx, rows = maxy - inity;
  if (cols < 7 || rows < 7) {             // cv::FAST finds nothing in ROIs narrower than the ring
    if (lane == 0) cell_cnt[out_idx] = 0;
    return;
  }
  const uint8_t* __restrict__ img = B.f[frame].pyr + A.g.off[level];
  uint32_t* const tile = s_tile[warp];
  uint16_t* const score = s_score[warp];
  uint16_t* const list = s_list[warp];
  uint16_t* const tile16 = reinterpret_cast<uint16_t*>(tile);

  // ---- stage the ROI as 16-bit pixels (element x + 1 of row y); scores start at zero
#pragma unroll
  for (int i = lane; i < 32 * TSH * 2 / 16; i += 32) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
  if (((initx & 3) == 0) && ((W & 3) == 0)) {
    const int wpr = (cols + 3) >> 2;                      // <= 8 words per row
    const int lr = lane >> 3, c4 = (lane & 7) * 4;        // 4 rows x 8 words per pass
    if (c4 < cols) {
      const uint8_t* p0 = img + size_t(inity + lr) * W + initx + c4;
      const bool whole = (initx + c4 + 3 < W);
      for (int r0 = 0; r0 < rows; r0 += 16) {
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int r = r0 + 4 * k + lr;
          v[k] = 0;
          if (r < rows) {
            const uint8_t* p = p0 + size_t(r0 + 4 * k) * W;
            if (whole) v[k] = __ldg(reinterpret_cast<const uint32_t*>(p));
            else
              for (int b = 0; b < 4; b++)
                if (initx + c4 + b < W) v[k] |= uint32_t(__ldg(p + b)) << (8 * b);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int r = r0 + 4 * k + lr;
          if (r < rows) {
            uint16_t* d = tile16 + r * PSE + c4 + 1;      // elements c4+1 .. c4+4: the middle two are word-aligned
            d[0] = uint16_t(v[k] & 255u);
            *reinterpret_cast<uint32_t*>(d + 1) = __byte_perm(v[k], 0, 0x4241);
            d[3] = uint16_t(v[k] >> 24);
          }
        }
      }
    }
    (void)wpr;
  } else {
    for (int i = lane; i < rows * cols; i += 32) {
      const int r = i / cols, c = i - r * cols;
      tile16[r * PSE + c + 1] = __ldg(img + size_t(inity + r) * W + initx + c);
    }
  }
  __syncwarp();

  // ---- scores: task t = (q, p) is the pixel pair x = 3 + 2p, x + 1 of row y = 3 + q
  const int np = (cols - 6 + 1) >> 1;     // pairs per row
  const int nq = rows - 6;                // tested rows
  const uint32_t t2 = uint32_t(A.threshold) * 0x00010001u;
  const uint32_t lt = (1u << lane) - 1u;
  // ---- quick test (cv::FAST's own early rejection, fast.cpp: opposite ring pixels): an arc of 9 contains ring pixel
  // 0 or 8 and ring pixel 4 or 12, so a corner needs min(max(r0, r8), max(r4, r12)) > v + t (brighter arc) or
  // max(min(r0, r8), min(r4, r12)) < v - t (darker arc).  ~20 instructions per pair instead of ~130; the pairs that
  // pass (15 % at level 0, 28 % at level 1 of the synthetic scenes) are compacted, in raster order, for full scoring.
  uint16_t* const cand = s_cand[warp];
  int ncand = 0;
  // Task = two horizontally adjacent pairs (a QUAD, pixels x .. x + 3) per lane: the six row words T[-2..3] serve the
  // centre and the left / right ring pixels of both pairs (10 shared loads per quad instead of 14 for two separate
  // pairs).  __vcmpgts2 / __vsub2 are 5- / 3-instruction emulations on sm_100a, hence the biased 32-bit form below.
  const int nq4 = (np + 1) >> 1;          // quads per row (<= 7); the last one may hold a single pair
  const int ntask4 = nq4 * nq;
  const uint32_t rcp4 = (65536u + uint32_t(nq4) - 1u) / uint32_t(nq4);   // t / nq4 == (t * rcp4) >> 16 for t < 182
  const uint32_t kq = 0x80008000u - (uint32_t(A.threshold) + 1u) * 0x00010001u;
  for (int t0 = 0; t0 < ntask4; t0 += 32) {
    const int t = t0 + lane;
    const bool active = t < ntask4;
    const uint32_t tt = active ? uint32_t(t) : 0u;
    const int qq = int((tt * rcp4) >> 16), pp = 2 * (int(tt) - qq * nq4);
    const uint32_t* T = tile + (3 + qq) * PSW + 2 + pp;
    const uint32_t u0 = T[3 * PSW], u1 = T[3 * PSW + 1], d0 = T[-3 * PSW], d1 = T[-3 * PSW + 1];
    const uint32_t m2 = T[-2], m1 = T[-1], c0 = T[0], c1 = T[1], p2 = T[2], p3 = T[3];
    bool pass[2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const uint32_t r0 = hh ? u1 : u0, r8 = hh ? d1 : d0, v2 = hh ? c1 : c0;
      const uint32_t r4 = hh ? mid_pair(p2, p3) : mid_pair(c1, p2), r12 = hh ? mid_pair(m1, c0) : mid_pair(m2, m1);
      const uint32_t br = __vmins2(__vmaxs2(r0, r8), __vmaxs2(r4, r12));
      const uint32_t dk = __vmaxs2(__vmins2(r0, r8), __vmins2(r4, r12));
      // br > v + t  <=>  br - v - (t + 1) >= 0;   v - t > dk  <=>  v - dk - (t + 1) >= 0.  Both lanes of a register at
      // once with ONE plain 32-bit add each: biased by 0x8000 a lane stays within 0x8000 +- 600, so no borrow crosses
      // into the upper lane and bit 15 of a lane is the outcome.
      const uint32_t ge = (br + kq - v2) | (v2 + kq - dk);
      pass[hh] = active && (ge & 0x80008000u) != 0;
    }
    pass[1] = pass[1] && pp + 1 < np;
    const uint32_t b0 = __ballot_sync(0xffffffffu, pass[0]), b1 = __ballot_sync(0xffffffffu, pass[1]);
    int pos = ncand + __popc(b0 & lt) + __popc(b1 & lt);
    const int code = (qq << 4) | pp;
    if (pass[0]) cand[pos++] = uint16_t(code);
    if (pass[1]) cand[pos] = uint16_t(code + 1);
    ncand += __popc(b0) + __popc(b1);
  }
  __syncwarp();
  int nlist = 0;
  for (int t0 = 0; t0 < ncand; t0 += 32) {
    const bool active = t0 + lane < ncand;
    const int qp = active ? cand[t0 + lane] : 0;
    const int qq = qp >> 4, pp = qp & 15;
    const int y = 3 + qq, x = 3 + 2 * pp;
    const int wi = y * TSH + 2 + pp;       // score half-word of the pair (elements x + 1, x + 2)
    const uint32_t* T = tile + y * PSW + 2 + pp;
    uint32_t r[16];
    {
      const uint32_t a = T[3 * PSW - 1], b = T[3 * PSW], c = T[3 * PSW + 1];          // row +3: dx -1, 0, 1
      r[15] = mid_pair(a, b); r[0] = b; r[1] = mid_pair(b, c);
    }
    r[14] = T[2 * PSW - 1]; r[2] = T[2 * PSW + 1];                                      // row +2: dx -2, 2
    r[13] = mid_pair(T[PSW - 2], T[PSW - 1]); r[3] = mid_pair(T[PSW + 1], T[PSW + 2]);  // row +1: dx -3, 3
    r[12] = mid_pair(T[-2], T[-1]); r[4] = mid_pair(T[1], T[2]);                        // row  0
    r[11] = mid_pair(T[-PSW - 2], T[-PSW - 1]); r[5] = mid_pair(T[-PSW + 1], T[-PSW + 2]);   // row -1
    r[10] = T[-2 * PSW - 1]; r[6] = T[-2 * PSW + 1];                                    // row -2
    {
      const uint32_t a = T[-3 * PSW - 1], b = T[-3 * PSW], c = T[-3 * PSW + 1];        // row -3: dx -1, 0, 1
      r[9] = mid_pair(a, b); r[8] = b; r[7] = mid_pair(b, c);
    }
    const uint32_t v2 = T[0];
    uint32_t lo3[16], hi3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      lo3[k] = __vimin3_s16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
      hi3[k] = __vimax3_s16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    uint32_t lo9[16], hi9[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      lo9[k] = __vimin3_s16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
      hi9[k] = __vimax3_s16x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
    }
    uint32_t m = __vimax3_s16x2(lo9[0], lo9[1], lo9[2]);
    uint32_t M = __vimin3_s16x2(hi9[0], hi9[1], hi9[2]);
#pragma unroll
    for (int k = 3; k < 15; k += 2) {
      m = __vimax3_s16x2(m, lo9[k], lo9[k + 1]);
      M = __vimin3_s16x2(M, hi9[k], hi9[k + 1]);
    }
    m = __vmaxs2(m, lo9[15]);
    M = __vmins2(M, hi9[15]);
    const uint32_t S = __vmaxs2(__vsub2(v2, M), __vsub2(m, v2));
    uint32_t sc = __vsub2(S, 0x00010001u) & __vcmpgts2(S, t2);
    if (x + 1 >= cols - 3) sc &= 0x0000FFFFu;   // second pixel of the pair is outside the tested columns
    if (!active) sc = 0;
    if (active) score[wi] = uint16_t((sc & 0xFFu) | ((sc >> 16) << 8));
    // raster-ordered list of corners (score > 0)
    const bool c0 = (sc & 0xFFFFu) != 0, c1 = (sc >> 16) != 0;
    const uint32_t b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
    int pos = nlist + __popc(b0 & lt) + __popc(b1 & lt);
    if (c0) list[pos++] = uint16_t((y << 5) | x);
    if (c1) list[pos] = uint16_t((y << 5) | (x + 1));
    nlist += __popc(b0) + __popc(b1);
  }
  __syncwarp();

  // ---- strict NMS against the 8 neighbours (non-corners are 0) for listed corners, raster-ordered compaction
  uint32_t* __restrict__ out = cell_kp + size_t(out_idx) * SDVLB_CELL_CAP;
  const int ox = initx - cj * SDVLB_CELL, oy = inity - ci * SDVLB_CELL;   // ROI origin relative to the cell origin
  const uint8_t* sc8 = reinterpret_cast<const uint8_t*>(score);
  int nout = 0;
  for (int i0 = 0; i0 < nlist; i0 += 32) {
    const int i = i0 + lane;
    bool keep = false;
    int s = 0, x = 0, y = 0;
    if (i < nlist) {
      const int yx = list[i];
      y = yx >> 5; x = yx & 31;
      const uint8_t* c = sc8 + y * TSE + x + 1;
      s = c[0];
      keep = s > c[-1] && s > c[1] && s > c[-TSE - 1] && s > c[-TSE] && s > c[-TSE + 1] && s > c[TSE - 1] &&
             s > c[TSE] && s > c[TSE + 1];
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int slot = nout + __popc(bal & lt);
      if (slot < SDVLB_CELL_CAP) out[slot] = (uint32_t(s) << 10) | (uint32_t(y + oy) << 5) | uint32_t(x + ox);
    }
    nout += __popc(bal);
  }
  if (lane == 0) cell_cnt[out_idx] = nout;   // <= 169 by construction of NMS
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

// Exclusive scan of vals[0..n) in place (n <= a few thousand) by the CTA; vals[n] receives the total.
__device__ __forceinline__ void block_exclusive_scan(int* vals, int n, int* s_tmp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + SEL_THREADS - 1) / SEL_THREADS;
  const int c0 = min(n, tid * per), c1 = min(n, c0 + per);
  int local = 0;
  for (int c = c0; c < c1; c++) local += vals[c];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __syncthreads();
  if (lane == 31) s_tmp[warp] = incl;
  __syncthreads();
  int run = incl - local;
  for (int w = 0; w < warp; w++) run += s_tmp[w];
  for (int c = c0; c < c1; c++) { const int v = vals[c]; vals[c] = run; run += v; }
  if (tid == SEL_THREADS - 1) vals[n] = run;
  __syncthreads();
}

// fast_select_kernel: grid (parts, frames); a part is SEL_PART_CELLS consecutive cells of one level.
//   1. every CTA runs the level's water-filling quota (fast_detector.cc:108-135) -- a few block-wide sums over the
//      level's cell counts -- and keeps the quotas of its own cells;
//   2. per-cell retainBest (fast_detector.cc:138-140), one WARP per cell (select_warp.cuh), in shared memory;
//   3. the last CTA of a (level, frame) gathers the survivors into the level list and runs the level-wide retainBest
//      (fast_detector.cc:146-148) on warp 0;
//   4. the last CTA of a frame concatenates the levels into Frame::corners_ order (fast_detector.cc:150-151,170-173),
//      mirrors the list to the host and builds the matcher's grid index.
__global__ void __launch_bounds__(SEL_THREADS) fast_select_kernel(const __grid_constant__ FrameBatch B,
                                                                  const __grid_constant__ FastArgs A,
                                                                  uint32_t* __restrict__ cell_kp,
                                                                  const int32_t* __restrict__ cell_cnt,
                                                                  int32_t* __restrict__ cell_kept,
                                                                  uint32_t* __restrict__ level_kp,
                                                                  int32_t* __restrict__ level_cnt,
                                                                  int32_t* __restrict__ tickets, int max_cells) {
  extern __shared__ __align__(16) int s_dyn[];   // nleft[max_cells + 1], nsel[max_cells + 1], then the work area
  __shared__ int s_tmp[SEL_THREADS / 32];
  __shared__ int s_final, s_ticket;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = B.scratch_base + blockIdx.y;   // index into the scratch arrays
#ifdef SDVLB_SELECT_DEBUG
  unsigned long long t_dbg[12];
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_dbg[0]));
#define SEL_MARK(k) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_dbg[k]))
#else
#define SEL_MARK(k)
#endif
  int level = 0;
  while (level + 1 < A.n_fast_levels && int(blockIdx.x) >= A.part_off[level + 1]) level++;
  const int part = int(blockIdx.x) - A.part_off[level];
  const int n_parts = A.part_off[level + 1] - A.part_off[level];
  const int ncells = A.g.wcells[level] * A.g.hcells[level];
  int* nleft = s_dyn;
  int* nsel = s_dyn + (max_cells + 1);
  unsigned char* work = reinterpret_cast<unsigned char*>(s_dyn + 2 * (max_cells + 1));
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
  SEL_MARK(1);

  // ---- per-cell retainBest (fast_detector.cc:138-140): warp w takes cells c0 + w, c0 + w + 8, ... of this part.  The
  // keypoints of the warp's NEXT cell are fetched (into registers) while the current one is being selected.
  {
    uint32_t* const kbuf = reinterpret_cast<uint32_t*>(work) + warp * SDVLB_CELL_CAP;
    uint8_t* const pbuf = work + (SEL_THREADS / 32) * SDVLB_CELL_CAP * sizeof(uint32_t) + warp * 2 * SDVLB_CELL_CAP;
    constexpr int KPL = (SDVLB_CELL_CAP + 31) / 32;   // keypoints per lane
    const int c0 = part * SEL_PART_CELLS, c1 = min(ncells, c0 + SEL_PART_CELLS);
    uint32_t nxt[KPL];
    int n_nxt = 0;
    auto fetch = [&](int c) {
      n_nxt = 0;
      if (c < c1) {
        n_nxt = max(cell_cnt[cbase + c], 0);
        if (n_nxt <= nsel[c]) return;   // nothing to select: the list stays as it is
        const uint32_t* g = cell_kp + size_t(cbase + c) * SDVLB_CELL_CAP;
#pragma unroll
        for (int k = 0; k < KPL; k++) nxt[k] = (lane + 32 * k < n_nxt) ? g[lane + 32 * k] : 0u;
      }
    };
    fetch(c0 + warp);
    for (int c = c0 + warp; c < c1; c += SEL_THREADS / 32) {
      const int n = n_nxt;
      const int quota = nsel[c];
      int kept = n;
      if (n > quota) {
#pragma unroll
        for (int k = 0; k < KPL; k++)
          if (lane + 32 * k < n) kbuf[lane + 32 * k] = nxt[k];
      }
      __syncwarp();
      fetch(c + SEL_THREADS / 32);
      if (n > quota) {
        uint32_t* const g = cell_kp + size_t(cbase + c) * SDVLB_CELL_CAP;
        kept = sdvlb_sel::warp_retain_best<10, uint8_t>(kbuf, n, quota, pbuf);
        for (int i = lane; i < kept; i += 32) g[i] = kbuf[i];
        __syncwarp();
      }
      if (lane == 0) cell_kept[cbase + c] = kept;
    }
  }

  // ---- last CTA of this (level, frame): level list + level-wide retainBest
  SEL_MARK(2);
  __threadfence();
  __syncthreads();
  SEL_MARK(3);
  if (tid == 0) s_ticket = atomicAdd(&tickets[frame * SDVLB_TICKET_STRIDE + 1 + level], 1);
  __syncthreads();
#ifdef SDVLB_SELECT_DEBUG
  if (s_ticket != n_parts - 1) {
    if (false)
      printf("sel part %d lvl %d: waterfill %llu cells(+barrier) %llu ns (start %llu)\n", int(blockIdx.x), level,
             t_dbg[1] - t_dbg[0], t_dbg[3] - t_dbg[1], t_dbg[0] % 1000000ull);
  }
#endif
  if (s_ticket != n_parts - 1) return;
  SEL_MARK(7);
  __threadfence();
  if (tid == 0) tickets[frame * SDVLB_TICKET_STRIDE + 1 + level] = 0;
  int* kept_of = nleft;   // kept count, then its exclusive scan (the water-filling state is no longer needed)
  for (int c = tid; c < ncells; c += SEL_THREADS) kept_of[c] = __ldcg(&cell_kept[cbase + c]);
  __syncthreads();
  for (int c = tid; c < ncells; c += SEL_THREADS) nsel[c] = kept_of[c];   // counts, kept beside their offsets
  __syncthreads();
  SEL_MARK(8);
  block_exclusive_scan(kept_of, ncells, s_tmp);
  SEL_MARK(9);
  const int total = kept_of[ncells];
  uint32_t* __restrict__ lk = level_kp + size_t(frame) * A.level_kp_total + A.level_kp_off[level];
  uint32_t* const s_keys = reinterpret_cast<uint32_t*>(work);
  uint16_t* const s_pos = reinterpret_cast<uint16_t*>(work + SEL_SMEM_KEYS * sizeof(uint32_t));
  const bool use_smem = total <= SEL_SMEM_KEYS;
  const bool fits = total <= A.level_cap[level];
  // a capacity was exceeded: recorded per frame (it reaches the frame's pinned header, so the overflow is reported for
  // THIS frame whichever call observes it first); frames built without a host mirror fall back to the context's flag
  if (!fits && tid == 0) {
    atomicExch(&tickets[frame * SDVLB_TICKET_STRIDE + SDVLB_TICKET_OVERFLOW], 1);
    if (!B.f[blockIdx.y].host_mirror) *reinterpret_cast<volatile int32_t*>(A.overflow_flag) = 1;
  }

  // ---- gather to the level list (fast_detector.cc:141-142): key = score<<22 | y<<11 | x (level coordinates)
  const int wc = A.g.wcells[level];
  if (fits) {
    // a thread per survivor: its cell is found by bisection of the scanned offsets (shared memory), so every global
    // load of the level is in flight at once (a thread per cell walked its survivors one L2 round trip at a time)
    for (int j = tid; j < total; j += SEL_THREADS) {
      int lo = 0, hi = ncells;   // largest c with kept_of[c] <= j; cells without survivors share an offset with the
      while (hi - lo > 1) {      // next one and lose the bisection to it
        const int mid = (lo + hi) >> 1;
        if (kept_of[mid] <= j) lo = mid; else hi = mid;
      }
      const int c = lo, i = j - kept_of[c];
      const uint32_t v = __ldcg(cell_kp + size_t(cbase + c) * SDVLB_CELL_CAP + i);
      const int ci = c / wc, cj = c - ci * wc;
      const uint32_t key = ((v >> 10) << 22) | (uint32_t(ci * SDVLB_CELL + ((v >> 5) & 31)) << 11) |
                           uint32_t(cj * SDVLB_CELL + (v & 31));
      if (use_smem) s_keys[j] = key;
      else lk[j] = key;
    }
  }
  __syncthreads();

  // ---- level-wide retainBest (fast_detector.cc:146-148): warp 0 in shared memory; one thread on the global list when
  // the level holds more than SEL_SMEM_KEYS candidates
  SEL_MARK(4);
  int fin = fits ? total : 0;
  if (fits && total > nfeatures) {
    if (use_smem) {
      // one warp, in shared memory (a CTA-wide variant of the two-sided passes, select_warp.cuh, measured slower: 11.7
      // instead of 8.9 us on 440 keypoints -- two block barriers per 256 elements cost more than walking 32 at a time)
      if (warp == 0) {
        const int f2 = sdvlb_sel::warp_retain_best<22, uint16_t>(s_keys, total, nfeatures, s_pos);
        if (lane == 0) s_final = f2;
      }
      __syncthreads();
      fin = s_final;
    } else {
      if (tid == 0) s_final = sdvlb_sel::retain_best<22>(lk, total, nfeatures);
      __syncthreads();
      fin = s_final;
    }
  }
  if (use_smem)
    for (int i = tid; i < fin; i += SEL_THREADS) lk[i] = s_keys[i];
  if (tid == 0) level_cnt[frame * SDVLB_MAX_LEVELS + level] = fin;

  // ---- last CTA of this frame concatenates the levels (fast_detector.cc:150-151,170-173)
  SEL_MARK(5);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(&tickets[frame * SDVLB_TICKET_STRIDE], 1);
  __syncthreads();
#ifdef SDVLB_SELECT_DEBUG
  if (false)
    printf("sel LAST of lvl %d (part %d): waterfill %llu cells %llu wait %llu gather %llu retain %llu ns, total %d nfeat %d (start %llu end %llu)\n",
           level, int(blockIdx.x), t_dbg[1] - t_dbg[0], t_dbg[2] - t_dbg[1], t_dbg[3] - t_dbg[2], t_dbg[4] - t_dbg[3],
           t_dbg[5] - t_dbg[4], total, nfeatures, t_dbg[0] % 1000000ull, t_dbg[5] % 1000000ull);
#endif
  if (s_ticket != A.n_fast_levels - 1) return;
  __threadfence();
  const FrameDev& fr = B.f[blockIdx.y];
  int4* const mirror = reinterpret_cast<int4*>(fr.host_mirror);   // header at [0], records from [1]
  // The same pass feeds the spatial index of the matcher (a counting sort of the corner indices by 32-px cell of their
  // level-0 position; the level-0 FAST grid has at most max_cells cells, so nleft / nsel are free to hold the counts):
  // a corner's cell is counted while its record is written and kept in shared memory for the scatter.
  const int gw = A.g.wcells[0], gcells = gw * A.g.hcells[0];
  int32_t* const start = fr.grid;
  int32_t* const cursor = fr.grid + gcells + 1;
  int32_t* const item = cursor + gcells;
  int* const s_cnt = nleft;    // count, then running cursor
  uint16_t* const s_cell = reinterpret_cast<uint16_t*>(work);
  constexpr int kCellSmem = SEL_SMEM_KEYS * 4;   // corners whose cell fits in the work area
  for (int i = tid; i <= gcells; i += SEL_THREADS) s_cnt[i] = 0;
  __syncthreads();
  int base = 0;
  for (int l = 0; l < A.n_fast_levels; l++) {
    const int n = *reinterpret_cast<volatile int32_t*>(&level_cnt[frame * SDVLB_MAX_LEVELS + l]);
    const uint32_t* src = level_kp + size_t(frame) * A.level_kp_total + A.level_kp_off[l];
    if (base + n > A.corner_cap) {
      if (tid == 0) {
        atomicExch(&tickets[frame * SDVLB_TICKET_STRIDE + SDVLB_TICKET_OVERFLOW], 1);
        if (!mirror) *reinterpret_cast<volatile int32_t*>(A.overflow_flag) = 1;
      }
      break;
    }
    for (int i = tid; i < n; i += SEL_THREADS) {
      const uint32_t key = __ldcg(src + i);
      const int x = int(key & 2047), y = int((key >> 11) & 2047);
      const int4 rec = make_int4(x, y, l, int32_t(key >> 22));
      fr.corners[base + i] = rec;
      if (mirror && base + i < fr.mirror_cap) mirror[1 + base + i] = rec;
      const int cell = ((y << l) >> 5) * gw + ((x << l) >> 5);
      if (base + i < kCellSmem) s_cell[base + i] = uint16_t(cell);
      atomicAdd(&s_cnt[cell], 1);
    }
    base += n;
  }
  if (tid == 0) {
    *fr.n_corners = base;
    const int ovf = atomicExch(&tickets[frame * SDVLB_TICKET_STRIDE + SDVLB_TICKET_OVERFLOW], 0);   // also re-arms it
    if (mirror) mirror[0] = make_int4(base, ovf, 0, 0);
    tickets[frame * SDVLB_TICKET_STRIDE] = 0;
  }
  __syncthreads();
  block_exclusive_scan(s_cnt, gcells, s_tmp);
  for (int i = tid; i <= gcells; i += SEL_THREADS) {
    start[i] = s_cnt[i];
    if (i < gcells) cursor[i] = s_cnt[i];   // kept for the layout's sake (readers use start[] and item[])
  }
  __syncthreads();
  // (the order of a cell's items is arbitrary: SearchPoint keeps the minimum of (score, corner index), which is what
  // the reference's in-order scan yields)
  for (int i = tid; i < base; i += SEL_THREADS) {
    int cell;
    if (i < kCellSmem) cell = s_cell[i];
    else { const int4 c = fr.corners[i]; cell = ((c.y << c.z) >> 5) * gw + ((c.x << c.z) >> 5); }
    item[atomicAdd(&s_cnt[cell], 1)] = i;
  }
#ifdef SDVLB_SELECT_DEBUG
  SEL_MARK(6);
  if (tid == 0)
    printf("sel LAST of frame %d (lvl %d, part %d): waterfill %llu cells %llu wait %llu gather %llu retain %llu concat+index %llu ns, span of this CTA %llu ns (start %llu end %llu) [ticket %llu loads %llu scan %llu gather %llu]\n",
           int(blockIdx.y), level, int(blockIdx.x), t_dbg[1] - t_dbg[0], t_dbg[2] - t_dbg[1], t_dbg[3] - t_dbg[2], t_dbg[4] - t_dbg[3], t_dbg[5] - t_dbg[4],
           t_dbg[6] - t_dbg[5], t_dbg[6] - t_dbg[0], t_dbg[0] % 1000000ull, t_dbg[6] % 1000000ull, t_dbg[7] - t_dbg[3],
           t_dbg[8] - t_dbg[7], t_dbg[9] - t_dbg[8], t_dbg[4] - t_dbg[9]);
#endif
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
  int parts = 0;
  for (int l = 0; l < SDVLB_MAX_LEVELS + 1; l++) {
    A.part_off[l] = parts;
    if (l < p.max_fast_levels) parts += (g.wcells[l] * g.hcells[l] + SEL_PART_CELLS - 1) / SEL_PART_CELLS;
  }
  plan->nfeatures = nfeatures;
}

cudaError_t sdvlb_launch_fast_cells(const FrameBatch& B, const FastPlan& plan, uint32_t* cell_kp, int32_t* cell_cnt,
                                    cudaStream_t stream) {
  const FastArgs& A = plan.args;
  dim3 g1((A.g.total_cells + DET_WARPS - 1) / DET_WARPS, B.n);
  // (Fewer FAST CTAs per SM, to leave room for the tracking kernels' CTAs, was tried with dynamic-memory padding: 5 -> 4
  // -> 3 CTAs per SM makes this kernel 76 -> 84 -> 99 us and the whole step slower, 225 k -> 220 k -> 212 k frames/s.)
  SDVLB_PREPARE(fast_cells_kernel, 0);
  fast_cells_kernel<<<g1, DET_THREADS, 0, stream>>>(B, A, cell_kp, cell_cnt);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_fast_select(const FrameBatch& B, const FastPlan& plan, uint32_t* cell_kp, int32_t* cell_cnt,
                                     int32_t* cell_kept, uint32_t* level_kp, int32_t* level_cnt, int32_t* tickets,
                                     cudaStream_t stream) {
  const FastArgs& A = plan.args;
  dim3 g2(A.part_off[A.n_fast_levels], B.n);
  // work area: the per-warp cell buffers (keys + positions) or, in the last CTA of a level, the level list + positions
  const size_t work = std::max(size_t(SEL_THREADS / 32) * SDVLB_CELL_CAP * (sizeof(uint32_t) + 2),
                               size_t(SEL_SMEM_KEYS) * (sizeof(uint32_t) + 2 * sizeof(uint16_t)));
  const size_t dyn = size_t(2 * (plan.max_cells_level + 1)) * sizeof(int) + work;
  if (dyn > 200 * 1024) return cudaErrorInvalidValue;
  SDVLB_PREPARE(fast_select_kernel, dyn);
  fast_select_kernel<<<g2, SEL_THREADS, dyn, stream>>>(B, A, cell_kp, cell_cnt, cell_kept, level_kp, level_cnt, tickets,
                                                      plan.max_cells_level);
  return cudaGetLastError();
}
