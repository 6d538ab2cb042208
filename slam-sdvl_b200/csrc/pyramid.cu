// pyramid.cu — K1: Gaussian 5x5 /2 image pyramid, bit-exact with cv::pyrDown on CV_8UC1 as called by
// Frame::CreatePyramid (frame.cc:114-120): dst(y,x) = (sum_{i,j} k_i k_j src(r(2y+i-2), r(2x+j-2)) + 128) >> 8,
// k = [1 4 6 4 1], r = BORDER_REFLECT_101, dst size = (cols/2, rows/2).  Pure integer arithmetic.
//
// The 25-tap sum is evaluated with the 4-way byte dot product (IDP.4A): for a source row r with vertical weight k_r and
// three consecutive aligned source words w0 w1 w2 (w1 starting at column 2x, x even),
//   dst(x)   += dp4a(w0, k_r*(0,0,1,4)) + dp4a(w1, k_r*(6,4,1,0))
//   dst(x+1) += dp4a(w1, k_r*(1,4,6,4)) + dp4a(w2, k_r*(1,0,0,0))
// i.e. 10 instructions per output pixel and no byte unpacking (all weights k_r*k_j <= 36 fit a byte).
//
//   upload_kernel   : level 0 of a whole frame batch from device-visible memory (pinned host memory over PCIe, or
//                     device memory) in one launch.
//   pyr_down_roll_kernel : one level -> next level for a frame batch in separable form with a rolling window of
//                     horizontal sums (round 2): a thread owns 8 output columns and walks down 4 or 8 output rows.
//   pyr_down_kernel : the direct form (10 IDP.4A per output, 8 outputs of one row per thread); rows that are not
//                     word-aligned, shared-memory sources (pyr_tail_kernel) and the SDVLB_PYR_DIRECT comparison knob.
//   pyr_tail_kernel : all remaining levels once a level and its successor fit in shared memory, one CTA per frame
//                     (levels 2..4 of a 752x480 frame from level 1: one launch instead of three).
#include <cstdlib>

#include "common.cuh"

namespace {

// BORDER_REFLECT_101; t is first clamped into the range where one reflection suffices (taps that need more are only
// ever requested for outputs that are discarded, the clamp just keeps their addresses valid).
__device__ __forceinline__ int reflect101(int t, int n) {
  t = max(-(n - 1), min(2 * n - 2, t));
  if (t < 0) return -t;
  if (t >= n) return 2 * n - 2 - t;
  return t;
}

// k_r * (0,0,1,4), k_r * (6,4,1,0), k_r * (1,4,6,4), k_r * (1,0,0,0) as little-endian byte vectors
__device__ __forceinline__ void pair_acc(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t kr, uint32_t& e, uint32_t& o) {
  e = __dp4a(w0, 0x04010000u * kr, e);
  e = __dp4a(w1, 0x00010406u * kr, e);
  o = __dp4a(w1, 0x04060401u * kr, o);
  o = __dp4a(w2, 0x00000001u * kr, o);
}

// 8 consecutive outputs x0 .. x0+7 (x0 % 8 == 0) of destination row y from a source whose rows are word-aligned
// (sw % 4 == 0, stride % 4 == 0, base 4-byte aligned).  Works on global or shared memory.  Returns the outputs packed
// in two words.  Borders cost two register patches, no extra memory access: the only out-of-row taps that reach a kept
// output are columns -2, -1 (reflected to 2, 1, which sit in the first word) and column sw (reflected to sw - 2, which
// sits in the last in-row word); words beyond the row are loaded from a clamped in-row address and ignored.
template <bool kGlobal>
__device__ __forceinline__ uint2 down8(const uint8_t* __restrict__ src, int stride, int sw, int sh, int y, int x0) {
  uint32_t e[4] = {0, 0, 0, 0}, o[4] = {0, 0, 0, 0};
  const int c0 = 2 * x0 - 4;                 // column of w[0]
  const int kb = (sw - c0) >> 2;             // index of the first word that lies beyond the row (>= 6: none)
#pragma unroll
  for (int r = 0; r < 5; r++) {
    const uint32_t kr = (r == 0 || r == 4) ? 1u : (r == 2 ? 6u : 4u);
    const uint8_t* __restrict__ p = src + size_t(reflect101(2 * y - 2 + r, sh)) * stride;
    uint32_t w[6];   // w[k] = columns c0 + 4k .. + 3
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int c = min(max(c0 + 4 * k, 0), sw - 4);
      const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(p + c);
      w[k] = kGlobal ? __ldg(q) : *q;
    }
    if (x0 == 0) w[0] = __byte_perm(w[1], 0, 0x1200);          // columns -2, -1 <- 2, 1
#pragma unroll
    for (int k = 1; k < 6; k++)
      if (k == kb) w[k] = (w[k - 1] >> 16) & 255u;             // column sw <- sw - 2
#pragma unroll
    for (int j = 0; j < 4; j++) pair_acc(w[j], w[j + 1], w[j + 2], kr, e[j], o[j]);
  }
  uint2 out;
  out.x = ((e[0] + 128) >> 8) | (((o[0] + 128) >> 8) << 8) | (((e[1] + 128) >> 8) << 16) | (((o[1] + 128) >> 8) << 24);
  out.y = ((e[2] + 128) >> 8) | (((o[2] + 128) >> 8) << 8) | (((e[3] + 128) >> 8) << 16) | (((o[3] + 128) >> 8) << 24);
  return out;
}

// Stores up to 8 packed outputs at drow + x0 (bound: x0 + k < dw); word stores when the row allows it.
__device__ __forceinline__ void store8(uint8_t* __restrict__ drow, int x0, int dw, uint2 v, bool words_ok) {
  if (words_ok && x0 + 8 <= dw) {
    *reinterpret_cast<uint32_t*>(drow + x0) = v.x;
    *reinterpret_cast<uint32_t*>(drow + x0 + 4) = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (x0 + k < dw) drow[x0 + k] = uint8_t(((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 255u);
  }
}

// Generic output (any alignment): byte taps with reflection.
__device__ __forceinline__ uint8_t down1(const uint8_t* __restrict__ src, int stride, int sw, int sh, int y, int x) {
  int acc = 0;
#pragma unroll
  for (int r = 0; r < 5; r++) {
    const int kr = (r == 0 || r == 4) ? 1 : (r == 2 ? 6 : 4);
    const uint8_t* __restrict__ p = src + size_t(reflect101(2 * y - 2 + r, sh)) * stride;
    const int h = p[reflect101(2 * x - 2, sw)] + 4 * p[reflect101(2 * x - 1, sw)] + 6 * p[reflect101(2 * x, sw)] +
                  4 * p[reflect101(2 * x + 1, sw)] + p[reflect101(2 * x + 2, sw)];
    acc += kr * h;
  }
  return uint8_t((acc + 128) >> 8);
}

constexpr int PD_THREADS = 256;
#ifndef SDVLB_PYR_ROWS_DEFAULT
#define SDVLB_PYR_ROWS_DEFAULT 0   // 0: by level size; 4 / 8: forced (experiments)
#endif

__global__ void __launch_bounds__(PD_THREADS) pyr_down_kernel(const __grid_constant__ FrameBatch B, int src_off, int sw,
                                                              int sh, int dst_off, int dw, int dh) {
  uint8_t* const pyr = B.f[blockIdx.y].pyr;
  const uint8_t* __restrict__ src = pyr + src_off;
  uint8_t* __restrict__ dst = pyr + dst_off;
  const int gpr = (dw + 7) >> 3;                       // groups of 8 outputs per row
  const int task = blockIdx.x * PD_THREADS + threadIdx.x;
  if (task >= gpr * dh) return;
  const int y = task / gpr, x0 = (task - y * gpr) * 8;
  if ((sw & 3) == 0) {                                 // rows are word-aligned (level offsets are 256-byte aligned)
    const uint2 v = down8<true>(src, sw, sw, sh, y, x0);
    store8(dst + size_t(y) * dw, x0, dw, v, (dw & 3) == 0);
  } else {
    for (int k = 0; k < 8; k++)
      if (x0 + k < dw) dst[size_t(y) * dw + x0 + k] = down1(src, sw, sw, sh, y, x0 + k);
  }
}

// ---- separable form with a rolling window (round 2) -------------------------------------------------------------------
// The 25-tap sum is separable in exact integer arithmetic: h(s, x) = sum_j k_j src(s, 2x+j-2) <= 16 * 255 per source row
// s, dst(y, x) = (h(2y-2) + 4 h(2y-1) + 6 h(2y) + 4 h(2y+1) + h(2y+2) + 128) >> 8 <= 65408 >> 8.  A thread owns 8
// adjacent output columns and walks down PR output rows: every step loads TWO new source rows (one 16-byte load plus
// the word on either side), forms their 8 horizontal sums with IDP.4A (2 per sum) packed as 16-bit pairs, and combines
// the five live rows with plain 32-bit adds on the pairs (a lane never exceeds 16 bits, so no carry crosses) -- ~16
// instead of ~70 instructions per output of the direct form, which recomputes every horizontal sum 2.5 times and
// clamps the address of each of its 30 loads.
// Borders cost nothing per row: a word outside the row is not loaded (it reads as 0) and the reflected taps are folded
// into the IDP.4A coefficients of the one thread they concern -- columns -2, -1 <- 2, 1 turn (6,4,1,0) on the first
// word of the row into (6,8,2,0) (coef_e0), column sw <- sw - 2 turns (1,4,6,4) on the last word into (1,4,7,4) for the
// last output of the row (coef_o[j]).
template <int ALIGN>
struct RawRow { uint32_t w[6]; };   // w[k] = columns cm - 4 + 4k .. + 3

template <int ALIGN>
__device__ __forceinline__ RawRow<ALIGN> load_row8(const uint8_t* __restrict__ p, int cm, int sw) {
  RawRow<ALIGN> r;
  r.w[0] = cm > 0 ? __ldg(reinterpret_cast<const uint32_t*>(p + cm - 4)) : 0u;
  if (ALIGN == 16) {   // sw % 16 == 0: the four middle words always lie inside the row
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + cm));
    r.w[1] = v.x; r.w[2] = v.y; r.w[3] = v.z; r.w[4] = v.w;
  } else if (ALIGN == 8) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p + cm));
    uint2 b = make_uint2(0u, 0u);
    if (cm + 16 <= sw) b = __ldg(reinterpret_cast<const uint2*>(p + cm + 8));
    r.w[1] = a.x; r.w[2] = a.y; r.w[3] = b.x; r.w[4] = b.y;
  } else {
#pragma unroll
    for (int k = 1; k < 5; k++)
      r.w[k] = cm + 4 * k <= sw ? __ldg(reinterpret_cast<const uint32_t*>(p + cm + 4 * k - 4)) : 0u;
  }
  r.w[5] = cm + 20 <= sw ? __ldg(reinterpret_cast<const uint32_t*>(p + cm + 16)) : 0u;
  return r;
}

template <int ALIGN>
__device__ __forceinline__ void hsum8(const RawRow<ALIGN>& r, uint32_t coef_e0, const uint32_t coef_o[4], uint32_t h[4]) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint32_t e = __dp4a(r.w[j + 1], j == 0 ? coef_e0 : 0x00010406u, __dp4a(r.w[j], 0x04010000u, 0u));
    const uint32_t o = __dp4a(r.w[j + 2], 0x00000001u, __dp4a(r.w[j + 1], coef_o[j], 0u));
    h[j] = e + (o << 16);
  }
}

template <int ALIGN, int PR>   // PR output rows per thread: 2 * PR + 3 source rows are summed for them
__global__ void __launch_bounds__(PD_THREADS) pyr_down_roll_kernel(const __grid_constant__ FrameBatch B, int src_off, int sw,
                                                                   int sh, int dst_off, int dw, int dh) {
  uint8_t* const pyr = B.f[blockIdx.y].pyr;
  const uint8_t* __restrict__ src = pyr + src_off;
  uint8_t* __restrict__ dst = pyr + dst_off;
  const int gpr = (dw + 7) >> 3;                       // groups of 8 outputs per row
  const int task = blockIdx.x * PD_THREADS + threadIdx.x;
  const int rb = task / gpr;
  const int y0 = rb * PR, x0 = (task - rb * gpr) * 8;
  if (y0 >= dh) return;
  const int cm = 2 * x0;                               // column of the first middle word
  const bool words_ok = (dw & 3) == 0;
  const uint32_t coef_e0 = cm == 0 ? 0x00020806u : 0x00010406u;
  uint32_t coef_o[4];
#pragma unroll
  for (int j = 0; j < 4; j++) coef_o[j] = x0 + 2 * j + 1 == dw - 1 ? 0x04070401u : 0x04060401u;
  // rows beyond the image (a last, partial block of rows) are clamped: loaded, summed, not stored
  auto row = [&](int s) { return src + size_t(reflect101(min(s, sh), sh)) * sw; };
  uint32_t ha[4], hb[4], hc[4], hd[4], he[4];
  {
    const RawRow<ALIGN> ra = load_row8<ALIGN>(row(2 * y0 - 2), cm, sw), rb_ = load_row8<ALIGN>(row(2 * y0 - 1), cm, sw),
                        rc = load_row8<ALIGN>(row(2 * y0), cm, sw);
    hsum8<ALIGN>(ra, coef_e0, coef_o, ha);
    hsum8<ALIGN>(rb_, coef_e0, coef_o, hb);
    hsum8<ALIGN>(rc, coef_e0, coef_o, hc);
  }
  RawRow<ALIGN> nd = load_row8<ALIGN>(row(2 * y0 + 1), cm, sw), ne = load_row8<ALIGN>(row(2 * y0 + 2), cm, sw);
#pragma unroll
  for (int r = 0; r < PR; r++) {
    const int y = y0 + r;
    const RawRow<ALIGN> cd = nd, ce = ne;
    if (r + 1 < PR) {   // the next step's rows are in flight while this step is summed
      nd = load_row8<ALIGN>(row(2 * y + 3), cm, sw);
      ne = load_row8<ALIGN>(row(2 * y + 4), cm, sw);
    }
    hsum8<ALIGN>(cd, coef_e0, coef_o, hd);
    hsum8<ALIGN>(ce, coef_e0, coef_o, he);
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = ha[j] + he[j] + 0x00800080u + 4u * (hb[j] + hd[j]) + 6u * hc[j];
    uint2 out;
    out.x = __byte_perm(v[0], v[1], 0x7531);           // bits 8..15 of every 16-bit lane
    out.y = __byte_perm(v[2], v[3], 0x7531);
    if (y < dh) store8(dst + size_t(y) * dw, x0, dw, out, words_ok);
#pragma unroll
    for (int j = 0; j < 4; j++) { ha[j] = hc[j]; hb[j] = hd[j]; hc[j] = he[j]; }
  }
}

// All levels after `first` for one frame per CTA: level `first` is staged in shared memory (row stride padded to a
// multiple of 4), every further level is computed from its predecessor in shared memory and written to global memory.
constexpr int PT_THREADS = 1024;
constexpr size_t kTailSmemMax = 64 * 1024;   // small on purpose: one CTA per frame only pays off for the tiny levels

__device__ __host__ inline size_t level_smem(int w, int h) { return (size_t((w + 3) & ~3) * h + 15) & ~size_t(15); }

__global__ void __launch_bounds__(PT_THREADS) pyr_tail_kernel(const __grid_constant__ FrameBatch B,
                                                              const __grid_constant__ PyrGeom G, int first) {
  extern __shared__ __align__(16) uint8_t s_buf[];
  uint8_t* const pyr = B.f[blockIdx.x].pyr;
  const int tid = threadIdx.x;
  int sw = G.w[first], sh = G.h[first];
  int sst = (sw + 3) & ~3;
  uint8_t* s_a = s_buf;                                  // source level
  uint8_t* s_b = s_buf + level_smem(sw, sh);             // its successor (levels shrink, so ping-pong always fits)
  {
    const uint8_t* __restrict__ g = pyr + G.off[first];
    if ((sw & 3) == 0) {
      const int n4 = (sw * sh) >> 2;
      for (int i = tid; i < n4; i += PT_THREADS)
        reinterpret_cast<uint32_t*>(s_a)[i] = __ldg(reinterpret_cast<const uint32_t*>(g) + i);
    } else {
      for (int i = tid; i < sw * sh; i += PT_THREADS) {
        const int r = i / sw;
        s_a[r * sst + (i - r * sw)] = __ldg(g + i);
      }
    }
  }
  __syncthreads();
  for (int l = first + 1; l < G.levels; l++) {
    const int dw = G.w[l], dh = G.h[l];
    const int dst_st = (dw + 3) & ~3;
    uint8_t* __restrict__ gdst = pyr + G.off[l];
    const bool last = (l + 1 == G.levels);
    if ((sw & 3) == 0) {
      const int gpr = (dw + 7) >> 3;
      for (int task = tid; task < gpr * dh; task += PT_THREADS) {
        const int y = task / gpr, x0 = (task - y * gpr) * 8;
        const uint2 v = down8<false>(s_a, sst, sw, sh, y, x0);
        store8(gdst + size_t(y) * dw, x0, dw, v, (dw & 3) == 0);
        if (!last) store8(s_b + size_t(y) * dst_st, x0, dst_st, v, true);   // padding columns are never used as taps
      }
    } else {   // odd-sized source (94 -> 47 of a 752-wide frame): byte taps, ONE output per thread -- these levels are
               // tiny, so the step lasts as long as its longest thread (8 outputs per thread: 180 busy threads of 1024)
      for (int i = tid; i < dw * dh; i += PT_THREADS) {
        const int y = i / dw, x = i - y * dw;
        const uint8_t px = down1(s_a, sst, sw, sh, y, x);
        gdst[size_t(y) * dw + x] = px;
        if (!last) s_b[size_t(y) * dst_st + x] = px;
      }
    }
    __syncthreads();
    uint8_t* t = s_a; s_a = s_b; s_b = t;
    sw = dw; sh = dh; sst = dst_st;
  }
}

// Level 0 of every frame of the batch from device-visible memory (pinned host memory read over PCIe, or device
// memory): one launch instead of one cudaMemcpyAsync per frame.  16-byte accesses when both sides allow it.
__global__ void __launch_bounds__(256) upload_kernel(const __grid_constant__ FrameBatch B,
                                                     const __grid_constant__ ImageBatch I, int bytes) {
  const uint8_t* __restrict__ src = I.src[blockIdx.y];
  uint8_t* __restrict__ dst = B.f[blockIdx.y].pyr;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int n16 = bytes >> 4;
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src);
    uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst);
    int i = tid;
    for (; i + 3 * nth < n16; i += 4 * nth) {   // four PCIe reads in flight per thread
      const uint4 a = __ldcs(s4 + i), b = __ldcs(s4 + i + nth), c = __ldcs(s4 + i + 2 * nth), d = __ldcs(s4 + i + 3 * nth);
      d4[i] = a; d4[i + nth] = b; d4[i + 2 * nth] = c; d4[i + 3 * nth] = d;
    }
    for (; i < n16; i += nth) d4[i] = __ldcs(s4 + i);
    for (int i = (n16 << 4) + tid; i < bytes; i += nth) dst[i] = src[i];
  } else {
    for (int i = tid; i < bytes; i += nth) dst[i] = src[i];
  }
}

// ---- level-0 upload through the bulk-copy engine (TMA, UBLKCP) ------------------------------------------------------
// One thread per CTA drives a ring of shared-memory stages: cp.async.bulk global -> shared (completion on an mbarrier)
// then cp.async.bulk shared -> global.  The reads of pinned host memory take a PCIe round trip each; issued by the LSU
// (upload_kernel) they sit in the SM's load pipeline and delay the loads of every compute CTA resident on the same SM,
// issued by the bulk-copy engine they do not touch it.  PCIe wants ~100 KB in flight in total, so a few CTAs suffice.
constexpr int UB_STAGES = 4;
constexpr int UB_CHUNK = 8192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(32) upload_bulk_kernel(const __grid_constant__ FrameBatch B,
                                                         const __grid_constant__ ImageBatch I, int bytes,
                                                         unsigned long long min_ns) {
  __shared__ __align__(128) uint8_t stage[UB_STAGES][UB_CHUNK];
  __shared__ __align__(8) uint64_t full[UB_STAGES];
  if (threadIdx.x != 0) return;
  unsigned long long t_start = 0;
  if (min_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
  // the chunks of all frames of the batch form one list, dealt round-robin to the CTAs of the (small) grid: what is in
  // flight on PCIe is gridDim.x * UB_STAGES * UB_CHUNK bytes whatever the batch size
  const int per_frame = (bytes + UB_CHUNK - 1) / UB_CHUNK;
  const int n_chunks = per_frame * B.n;
  const int mine = (n_chunks - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  if (mine <= 0) return;
  for (int s = 0; s < UB_STAGES; s++)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  auto locate = [&](int k, int& fr, int& off, int& sz) {
    const int id = int(blockIdx.x) + k * int(gridDim.x);
    fr = id / per_frame;
    off = (id - fr * per_frame) * UB_CHUNK;
    sz = min(UB_CHUNK, bytes - off);   // multiple of 16 (checked by the launcher)
  };
  auto issue = [&](int k) {
    const int s = k % UB_STAGES;
    int fr, off, sz;
    locate(k, fr, off, sz);
    const uint32_t bar = smem_u32(&full[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(sz) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(stage[s])), "l"(I.src[fr] + off), "r"(sz), "r"(bar) : "memory");
  };
  for (int k = 0; k < min(mine, UB_STAGES); k++) issue(k);
  for (int k = 0; k < mine; k++) {
    const int s = k % UB_STAGES;
    const uint32_t bar = smem_u32(&full[s]), parity = (k / UB_STAGES) & 1;
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    int fr, off, sz;
    locate(k, fr, off, sz);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(B.f[fr].pyr + off), "r"(smem_u32(stage[s])), "r"(sz) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (k + UB_STAGES < mine) {   // the stage is free again once the store has read it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue(k + UB_STAGES);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (min_ns) {   // experiment knob (SDVLB_UPLOAD_MIN_NS_PER_FRAME): emulate the duration of a PCIe transfer
    unsigned long long t;
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t_start < min_ns);
  }
}

// Camera::UndistortImage = cv::undistort (camera.cc:100-105): per destination pixel the fixed-point map of
// cv::initUndistortRectifyMap (m1type CV_16SC2: source pixel + 5-bit fractions) and cv::remap's fixed-point bilinear
// kernel (weights with 15 fractional bits, taps outside the image are 0).  Four destination pixels per thread, one
// 32-bit store; the taps are byte gathers from the raw image (L1/L2: neighbouring pixels share their sources).
__global__ void __launch_bounds__(256) undistort_kernel(const __grid_constant__ FrameBatch B,
                                                        const __grid_constant__ ImageBatch R,
                                                        const __grid_constant__ UndistortArgs U) {
  const uint8_t* __restrict__ src = R.src[blockIdx.y];
  uint8_t* __restrict__ dst = B.f[blockIdx.y].pyr;
  const int w = U.w, h = U.h;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;   // group of four pixels; rows are multiples of 4 wide or padded
  const int gpr = (w + 3) >> 2;
  if (q >= gpr * h) return;
  const int i = q / gpr, j0 = (q - i * gpr) * 4;
  const double ir0 = 1.0 / U.fx, ir2 = -U.u0 / U.fx, ir4 = 1.0 / U.fy, ir5 = -U.v0 / U.fy;
  const double y = i * ir4 + ir5;
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int j = j0 + k;
    const double x = j * ir0 + ir2;
    const double x2 = x * x, y2 = y * y;
    const double r2 = x2 + y2, _2xy = 2 * x * y;
    const double kr = 1 + ((U.k3 * r2 + U.k2) * r2 + U.k1) * r2;   // the rational part k4..k6 is not used by SDVL
    const double xd = x * kr + U.p1 * _2xy + U.p2 * (r2 + 2 * x2);
    const double yd = y * kr + U.p1 * (r2 + 2 * y2) + U.p2 * _2xy;
    const double u = U.fx * xd + U.u0, v = U.fy * yd + U.v0;
    const int iu = __double2int_rn(u * 32), iv = __double2int_rn(v * 32);   // saturate_cast<int>(x) = cvRound(x)
    const int sx = int(short(iu >> 5)), sy = int(short(iv >> 5));          // the map stores them as short
    const int fx = iu & 31, fy = iv & 31;
    int acc = 0;
    if (sx >= 0 && sy >= 0 && sx + 1 < w && sy + 1 < h) {
      const uint8_t* p = src + size_t(sy) * w + sx;
      acc = int(__ldg(p)) * ((32 - fx) * (32 - fy)) + int(__ldg(p + 1)) * (fx * (32 - fy)) +
            int(__ldg(p + w)) * ((32 - fx) * fy) + int(__ldg(p + w + 1)) * (fx * fy);
    } else {
      auto tap = [&](int yy, int xx) -> int {
        return (xx >= 0 && xx < w && yy >= 0 && yy < h) ? int(__ldg(src + size_t(yy) * w + xx)) : 0;
      };
      acc = tap(sy, sx) * ((32 - fx) * (32 - fy)) + tap(sy, sx + 1) * (fx * (32 - fy)) +
            tap(sy + 1, sx) * ((32 - fx) * fy) + tap(sy + 1, sx + 1) * (fx * fy);
    }
    // weights carry a factor 32 (2^15 scale): (acc * 32 + 2^14) >> 15 == (acc + 2^9) >> 10
    const uint32_t px = uint32_t((acc + (1 << 9)) >> 10);
    if (j < w) packed |= px << (8 * k);
  }
  uint8_t* o = dst + size_t(i) * w + j0;
  if (j0 + 3 < w && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) *reinterpret_cast<uint32_t*>(o) = packed;
  else
    for (int k = 0; k < 4 && j0 + k < w; k++) o[k] = uint8_t(packed >> (8 * k));
}

// The level the tail kernel starts from: the first one that fits in shared memory together with its successor
// (-1: none, every level is produced by pyr_down_kernel).
int tail_source_level(const PyrGeom& g, size_t* smem_bytes) {
  for (int l = 1; l + 1 < g.levels; l++) {
    const size_t need = level_smem(g.w[l], g.h[l]) + level_smem(g.w[l + 1], g.h[l + 1]);
    if (need <= kTailSmemMax) {
      if (smem_bytes) *smem_bytes = need;
      return l;
    }
  }
  return -1;
}

}  // namespace

cudaError_t sdvlb_launch_upload(const FrameBatch& B, const ImageBatch& I, int bytes, bool src_on_device, cudaStream_t stream) {
  static const bool emulate_pcie = getenv("SDVLB_UPLOAD_MIN_NS_PER_FRAME") != nullptr;   // experiment knob, see below
  if (src_on_device && !emulate_pcie) {   // HBM -> HBM: no PCIe depth to manage, a plain wide copy (8 CTAs per frame)
    SDVLB_PREPARE(upload_kernel, 0);
    upload_kernel<<<dim3(8, B.n), 256, 0, stream>>>(B, I, bytes);
    return cudaGetLastError();
  }
  // PCIe wants its bandwidth-delay product in flight (~100 KB) and not much more: every other PCIe read -- above all
  // the GPU front end fetching the launch commands of the compute kernels -- queues behind what the upload has
  // outstanding.  The bulk kernel keeps SDVLB_UPLOAD_CTAS (default 2) x 32 KB in flight whatever the batch size
  // (measured on B200, 64 sequences in 8 groups: 2 CTAs 114 k frames/s, 4: 112 k, 8: 103 k, 16: 98 k).
  // SDVLB_UPLOAD=ldg selects the load/store kernel (per-frame CTAs; kept for comparison and for unaligned images).
  static int ctas = 0, bulk = 1;
  if (ctas == 0) {
    const char* m = getenv("SDVLB_UPLOAD");
    bulk = !(m && m[0] == 'l');
    const char* e = getenv("SDVLB_UPLOAD_CTAS");
    ctas = e ? atoi(e) : (bulk ? 2 : 1);
    if (ctas < 1 || ctas > 148) ctas = bulk ? 2 : 1;
  }
  bool aligned = (bytes & 15) == 0;
  for (int i = 0; i < B.n && aligned; i++)
    aligned = ((reinterpret_cast<uintptr_t>(I.src[i]) | reinterpret_cast<uintptr_t>(B.f[i].pyr)) & 15) == 0;
  if (bulk && aligned) {
    static long long min_ns = -1;
    if (min_ns < 0) { const char* e = getenv("SDVLB_UPLOAD_MIN_NS_PER_FRAME"); min_ns = e ? atoll(e) : 0; }
    upload_bulk_kernel<<<ctas, 32, 0, stream>>>(B, I, bytes, (unsigned long long)(min_ns * B.n));
  } else {
    dim3 grid(bulk ? 1 : ctas, B.n);
    SDVLB_PREPARE(upload_kernel, 0);
    upload_kernel<<<grid, 256, 0, stream>>>(B, I, bytes);
  }
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_undistort(const FrameBatch& B, const ImageBatch& raw, const UndistortArgs& U, cudaStream_t stream) {
  const int groups = ((U.w + 3) >> 2) * U.h;
  dim3 grid((groups + 255) / 256, B.n);
  SDVLB_PREPARE(undistort_kernel, 0);
  undistort_kernel<<<grid, 256, 0, stream>>>(B, raw, U);
  return cudaGetLastError();
}

int sdvlb_pyramid_launches(const PyrGeom& g) {
  const int src = tail_source_level(g, nullptr);
  return src > 0 ? src + 1 : g.levels - 1;
}

// Builds levels 1..L-1 for the frames of the batch (level 0 already resident).
cudaError_t sdvlb_launch_pyramid(const FrameBatch& B, const PyrGeom& g, cudaStream_t stream) {
  size_t smem = 0;
  const int tail_src = tail_source_level(g, &smem);
  const int direct_to = tail_src > 0 ? tail_src : g.levels - 1;   // levels 1..direct_to by pyr_down_kernel
  static const bool direct = getenv("SDVLB_PYR_DIRECT") != nullptr;   // comparison knob: the round-1 direct-form kernel
  static const int rows_per_thread = getenv("SDVLB_PYR_ROWS") ? atoi(getenv("SDVLB_PYR_ROWS")) : SDVLB_PYR_ROWS_DEFAULT;
  for (int l = 1; l <= direct_to; l++) {
    const int sw = g.w[l - 1];
    if ((sw & 3) == 0 && !direct) {   // word-aligned rows (level offsets are 256-byte aligned): rolling separable form
#define SDVLB_ROLL(AL, R)                                                                                            \
  do {                                                                                                               \
    const int tasks = ((g.w[l] + 7) >> 3) * ((g.h[l] + R - 1) / R);                                                  \
    dim3 grid((tasks + PD_THREADS - 1) / PD_THREADS, B.n);                                                           \
    SDVLB_PREPARE((pyr_down_roll_kernel<AL, R>), 0);                                                                 \
    pyr_down_roll_kernel<AL, R><<<grid, PD_THREADS, 0, stream>>>(B, g.off[l - 1], sw, g.h[l - 1], g.off[l], g.w[l], g.h[l]); \
  } while (0)
      // 8 rows per thread sum 19 source rows for 8 output rows, 4 rows 11 for 4 (+16 % work) but twice the threads: the
      // small levels are latency-bound and take the latter (A/B on B200: 33.3 / 32.0 us per 64 frames with 8 / 4 everywhere)
      const int rpt = rows_per_thread ? rows_per_thread : (g.h[l] >= 200 ? 8 : 4);
      if ((sw & 15) == 0) { if (rpt == 8) SDVLB_ROLL(16, 8); else SDVLB_ROLL(16, 4); }
      else if ((sw & 7) == 0) { if (rpt == 8) SDVLB_ROLL(8, 8); else SDVLB_ROLL(8, 4); }
      else { if (rpt == 8) SDVLB_ROLL(4, 8); else SDVLB_ROLL(4, 4); }
#undef SDVLB_ROLL
      continue;
    }
    const int tasks = ((g.w[l] + 7) >> 3) * g.h[l];
    dim3 grid((tasks + PD_THREADS - 1) / PD_THREADS, B.n);
    SDVLB_PREPARE(pyr_down_kernel, 0);
    pyr_down_kernel<<<grid, PD_THREADS, 0, stream>>>(B, g.off[l - 1], g.w[l - 1], g.h[l - 1], g.off[l], g.w[l], g.h[l]);
  }
  if (tail_src > 0) {
    SDVLB_PREPARE(pyr_tail_kernel, smem);
    pyr_tail_kernel<<<B.n, PT_THREADS, smem, stream>>>(B, g, tail_src);
  }
  return cudaGetLastError();
}
