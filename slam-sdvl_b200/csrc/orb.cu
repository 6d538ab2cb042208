// orb.cu — ORB descriptor mode on the device: ORBDetector::GetOrientation and
// GetDescriptor (extra/orb_detector.cc:350-437) for a list of positions, for every corner of the frames of a build
// batch (Frame::descriptors_, filled lazily by the reference, matcher.cc:265-269; here once, when the frame is built),
// and for the init features of the map points a sequence receives (Feature::descriptor_).  Device functions: orb.cuh.
#include "orb.cuh"
#include "seq.cuh"

using namespace sdvlb_orb;

namespace {

constexpr int ORB_WARPS = 8;

// positions: n x (x, y, level) in level coordinates
__global__ void __launch_bounds__(ORB_WARPS * 32) orb_positions_kernel(const uint8_t* __restrict__ pyr,
                                                                        const __grid_constant__ PyrGeom G, int levels,
                                                                        const int32_t* __restrict__ xyl, int n,
                                                                        uint8_t* __restrict__ desc, float* __restrict__ angle) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * ORB_WARPS + (threadIdx.x >> 5);
  if (i >= n) return;
  const int x = __ldg(xyl + 3 * i), y = __ldg(xyl + 3 * i + 1), l = __ldg(xyl + 3 * i + 2);
  uint32_t byte = 0;
  float deg = -1.f;
  uint32_t pw[8];
  orb_load_pattern(pw);
  if (l >= 0 && l < levels && x >= kOrbLimit && x < G.w[l] - kOrbLimit && y >= kOrbLimit && y < G.h[l] - kOrbLimit)
    deg = orb_describe(pyr + G.off[l], G.w[l], x, y, &byte, pw);
  desc[size_t(i) * 32 + lane] = uint8_t(byte);
  if (angle && lane == 0) angle[i] = deg;
}

// Every corner of every frame of a build batch: Frame::descriptors_ (filled lazily by the reference, matcher.cc:265-269,
// frame.cc:148-161; a descriptor is a function of the frame and the corner, so computing all of them when the frame is
// built gives the same bytes).  grid (ORB_FRAME_CTAS, frames); a CTA takes chunks of ORB_CHUNK corners of its frame
// through three phases: moments (a warp per corner), orientation + rotation (a THREAD per corner: the double-precision
// cos / sin dominate the instruction count and are the same for all 32 lanes of a warp-per-corner design), rotated
// tests (a warp per corner).
constexpr int ORB_FRAME_CTAS = 32;
constexpr int ORB_CHUNK = 32;
__global__ void __launch_bounds__(ORB_WARPS * 32) orb_frames_kernel(const __grid_constant__ FrameBatch B,
                                                                     const __grid_constant__ PyrGeom G, int levels, int cap) {
  __shared__ int s_m10[ORB_CHUNK], s_m01[ORB_CHUNK];
  __shared__ float s_a[ORB_CHUNK], s_b[ORB_CHUNK];
  const FrameDev& f = B.f[blockIdx.y];
  if (!f.desc) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = min(*f.n_corners, cap);
  uint32_t pw[8];
  orb_load_pattern(pw);
  for (int c0 = blockIdx.x * ORB_CHUNK; c0 < n; c0 += gridDim.x * ORB_CHUNK) {
    const int m = min(ORB_CHUNK, n - c0);
    for (int k = warp; k < m; k += ORB_WARPS) {
      const int4 c = __ldg(f.corners + c0 + k);
      int m10 = 0, m01 = 0;
      const bool ok = c.z >= 0 && c.z < levels && c.x >= kOrbLimit && c.x < G.w[c.z] - kOrbLimit && c.y >= kOrbLimit &&
                      c.y < G.h[c.z] - kOrbLimit;
      if (ok) orb_moments(f.pyr + G.off[c.z] + size_t(c.y) * G.w[c.z] + c.x, G.w[c.z], m10, m01);
      if (lane == 0) { s_m10[k] = m10; s_m01[k] = m01; }
    }
    __syncthreads();
    if (int(threadIdx.x) < m) {
      float a, b;
      orb_angle(s_m10[threadIdx.x], s_m01[threadIdx.x], a, b);
      s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    }
    __syncthreads();
    for (int k = warp; k < m; k += ORB_WARPS) {
      const int4 c = __ldg(f.corners + c0 + k);
      const bool ok = c.z >= 0 && c.z < levels && c.x >= kOrbLimit && c.x < G.w[c.z] - kOrbLimit && c.y >= kOrbLimit &&
                      c.y < G.h[c.z] - kOrbLimit;
      uint32_t byte = 0;
      if (ok) byte = orb_sample(f.pyr + G.off[c.z] + size_t(c.y) * G.w[c.z] + c.x, G.w[c.z], s_a[k], s_b[k], pw);
      reinterpret_cast<uint8_t*>(f.desc)[size_t(c0 + k) * 32 + lane] = uint8_t(byte);
    }
    __syncthreads();
  }
}

// Feature::descriptor_ of the init features of the map points carried by the SEQC_ADD_POINTS commands of a
// submission (blockIdx.y = command): computed from the keyframe the points come from, before the align kernel appends
// the points -- and their descriptors -- to the sequence's feature list.
__global__ void __launch_bounds__(ORB_WARPS * 32) seq_orb_points_kernel(const SeqCmd* __restrict__ cmds,
                                                                         const __grid_constant__ PyrGeom G) {
  const SeqCmd& C = cmds[blockIdx.y];
  if (C.kind != SEQC_ADD_POINTS || !C.desc) return;
  const int lane = threadIdx.x & 31;
  for (int k = blockIdx.x * ORB_WARPS + (threadIdx.x >> 5); k < C.n; k += gridDim.x * ORB_WARPS) {
    const double px0 = C.pts[k].ref_px[0], px1 = C.pts[k].ref_px[1];
    const int level = C.pts[k].ref_level;
    const uint32_t w = orb_feature_word(C.kf_pyr, G, level, px0, px1);
    if (lane < 8) C.desc[size_t(k) * 8 + lane] = w;
  }
}

}  // namespace

cudaError_t sdvlb_launch_orb_positions(const uint8_t* pyr, const PyrGeom& g, int levels, const int32_t* d_xyl, int n,
                                       uint8_t* d_desc, float* d_angle, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  SDVLB_PREPARE(orb_positions_kernel, 0);
  orb_positions_kernel<<<(n + ORB_WARPS - 1) / ORB_WARPS, ORB_WARPS * 32, 0, stream>>>(pyr, g, levels, d_xyl, n, d_desc, d_angle);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_orb_frames(const FrameBatch& B, const PyrGeom& g, int levels, int cap, cudaStream_t stream) {
  if (B.n <= 0 || cap <= 0) return cudaSuccess;
  SDVLB_PREPARE(orb_frames_kernel, 0);
  orb_frames_kernel<<<dim3(ORB_FRAME_CTAS, B.n), ORB_WARPS * 32, 0, stream>>>(B, g, levels, cap);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_seq_orb_points(const SeqCmd* d_cmds, int n_cmds, int max_points, const PyrGeom& g, cudaStream_t stream) {
  if (n_cmds <= 0 || max_points <= 0) return cudaSuccess;
  SDVLB_PREPARE(seq_orb_points_kernel, 0);
  const int ctas = (max_points + ORB_WARPS - 1) / ORB_WARPS;
  seq_orb_points_kernel<<<dim3(ctas < 16 ? ctas : 16, n_cmds), ORB_WARPS * 32, 0, stream>>>(d_cmds, g);
  return cudaGetLastError();
}
