// pyramid.cu — K1: Gaussian 5x5 /2 image pyramid, bit-exact with cv::pyrDown on CV_8UC1 as called by
// Frame::CreatePyramid (frame.cc:114-120): dst(y,x) = (sum_{i,j} k_i k_j src(r(2y+i-2), r(2x+j-2)) + 128) >> 8,
// k = [1 4 6 4 1], r = BORDER_REFLECT_101, dst size = (cols/2, rows/2).  Pure integer arithmetic.
//
// One launch per destination level, batched over frames (grid.z).  A CTA produces a 64x16 tile: it stages the
// (2*64+8) x (2*16+3) source window in shared memory with aligned 32-bit loads (byte loads at borders / unaligned
// rows), filters rows into a u16 buffer, then columns, and stores packed 32-bit words.
#include "common.cuh"

namespace {

constexpr int TW = 64, TH = 16;
constexpr int SROWS = 2 * TH + 3;     // 35 source rows
constexpr int SCOLS = 2 * TW + 8;     // 136 bytes: source x in [2*tx0-4, 2*tx0+132)
constexpr int THREADS = 256;

// BORDER_REFLECT_101; t is first clamped into the range where one reflection suffices (taps that need more are
// only ever requested for outputs that are discarded, the clamp just keeps their addresses valid).
__device__ __forceinline__ int reflect101(int t, int n) {
  t = max(-(n - 1), min(2 * n - 2, t));
  if (t < 0) return -t;
  if (t >= n) return 2 * n - 2 - t;
  return t;
}

__global__ void __launch_bounds__(THREADS) pyr_down_kernel(const __grid_constant__ FrameBatch B, int src_off, int sw,
                                                           int sh, int dst_off, int dw, int dh) {
  __shared__ __align__(16) uint8_t s_src[SROWS][SCOLS];
  __shared__ uint16_t s_h[SROWS][TW];

  uint8_t* const pyr = B.f[blockIdx.z].pyr;
  const uint8_t* __restrict__ src = pyr + src_off;
  uint8_t* __restrict__ dst = pyr + dst_off;
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
  const int sx0 = 2 * tx0 - 4;          // source x of s_src[.][0]
  const int sy0 = 2 * ty0 - 2;          // source y of s_src[0][.]
  const int tid = threadIdx.x;

  // ---- stage source window
  const bool row_words_ok = (sx0 >= 0) && (sx0 + SCOLS <= sw) && ((sw & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
  if (row_words_ok) {
    constexpr int WPR = SCOLS / 4;  // 34 words per row
    for (int i = tid; i < SROWS * WPR; i += THREADS) {
      const int r = i / WPR, c = i - r * WPR;
      const int sy = reflect101(sy0 + r, sh);
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + size_t(sy) * sw + sx0) + c);
      *reinterpret_cast<uint32_t*>(&s_src[r][4 * c]) = v;
    }
  } else {
    for (int i = tid; i < SROWS * SCOLS; i += THREADS) {
      const int r = i / SCOLS, c = i - r * SCOLS;
      const int sy = reflect101(sy0 + r, sh);
      const int sx = reflect101(sx0 + c, sw);
      s_src[r][c] = __ldg(src + size_t(sy) * sw + sx);
    }
  }
  __syncthreads();

  // ---- horizontal pass: s_h[r][x] = sum_j k_j * src(r, 2(tx0+x)+j-2)  -> s_src column 2x+j+2
  for (int i = tid; i < SROWS * TW; i += THREADS) {
    const int r = i / TW, x = i - r * TW;
    const uint8_t* p = &s_src[r][2 * x + 2];
    s_h[r][x] = uint16_t(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
  }
  __syncthreads();

  // ---- vertical pass + store: each thread 4 consecutive x of one row
  const int qy = tid / (TW / 4), qx = (tid - qy * (TW / 4)) * 4;
  const int oy = ty0 + qy, ox = tx0 + qx;
  if (oy < dh && ox < dw) {
    uint32_t packed = 0;
    uint8_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int x = qx + k;
      const int acc = s_h[2 * qy][x] + 4 * s_h[2 * qy + 1][x] + 6 * s_h[2 * qy + 2][x] + 4 * s_h[2 * qy + 3][x] +
                      s_h[2 * qy + 4][x];
      o[k] = uint8_t((acc + 128) >> 8);
      packed |= uint32_t(o[k]) << (8 * k);
    }
    uint8_t* d = dst + size_t(oy) * dw + ox;
    if (ox + 3 < dw && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(d) = packed;
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (ox + k < dw) d[k] = o[k];
    }
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
    for (int i = tid; i < n16; i += nth) d4[i] = __ldcs(s4 + i);
    for (int i = (n16 << 4) + tid; i < bytes; i += nth) dst[i] = src[i];
  } else {
    for (int i = tid; i < bytes; i += nth) dst[i] = src[i];
  }
}

}  // namespace

cudaError_t sdvlb_launch_upload(const FrameBatch& B, const ImageBatch& I, int bytes, cudaStream_t stream) {
  dim3 grid(24, B.n);
  upload_kernel<<<grid, 256, 0, stream>>>(B, I, bytes);
  return cudaGetLastError();
}

int sdvlb_pyramid_launches(const PyrGeom& g) { return g.levels - 1; }

// Builds levels 1..L-1 for the frames of the batch (level 0 already resident). One launch per level.
cudaError_t sdvlb_launch_pyramid(const FrameBatch& B, const PyrGeom& g, cudaStream_t stream) {
  for (int l = 1; l < g.levels; l++) {
    dim3 grid((g.w[l] + TW - 1) / TW, (g.h[l] + TH - 1) / TH, B.n);
    pyr_down_kernel<<<grid, THREADS, 0, stream>>>(B, g.off[l - 1], g.w[l - 1], g.h[l - 1], g.off[l], g.w[l], g.h[l]);
  }
  return cudaGetLastError();
}
