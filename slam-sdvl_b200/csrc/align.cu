// align.cu — K3: ImageAlign::ComputePose (image_align.cc:46-267) as ONE persistent kernel per call: all pyramid
// levels and all Gauss-Newton iterations run on the device (one CTA per alignment = per sequence), including the
// 6x6 LDLT solve, the SE3 update T <- T*Exp(-x) and the reference's accept/rollback logic.  No host round trip, no
// atomics: partial sums go through warp shuffles and one fixed-order pass over the warps' totals.
//
// Two entry kernels share the body (align_core): image_align_kernel (class API: features marshalled by the host,
// AlignJobDev) and seq_align_kernel (resident sequences: the CTA first applies the mapping thread's commands, sets the
// motion-model prior -- SDVL::SetMotionModel, sdvl.cc:278-281 -- and reads its features straight from the sequence's
// feature list; what used to be a separate "prep" launch).
//
// Arithmetic mirrors the reference's type ledger: fp32 pixel weights / patch / gradients / residual,
// fp64 geometry, Jacobian, H, b, pose.  The reference's per-pixel J = (dx*Jp0 + dy*Jp1)*fx/2^l is linear in (dx,dy),
// so per feature b_f = -s*(Jp0*sum(dx*res) + Jp1*sum(dy*res)) and H_f = s^2*(A Jp0Jp0' + B(Jp0Jp1'+Jp1Jp0') + C Jp1Jp1')
// with A,B,C = sum(dx^2, dx*dy, dy^2): identical in exact arithmetic, and within 1e-15 relative in fp64.
//
// Inverse compositional: H_f does not change inside a level, and H only depends on WHICH features project into the
// current image.  So PrecomputePatches also forms H_all = sum_f H_f once per level; an iteration reduces b, chi2 and
// three counters only (8 values instead of 29) and H is H_all minus the H_f of the features that fell outside the
// current image (a handful near the borders): the warp that owns such a feature re-forms its H_f from the cached
// Jacobian rows and gradient moments, lane k computing entry k, so the correction needs no shuffle and no extra pass.
// As long as the SET of outside features does not change, H does not either and its LDL^T factor is reused: most
// iterations are two triangular substitutions.
//
// Per-feature caches (xyz, Jacobian rows, reference patch and its gradients) live in SHARED memory for the first
// `cap` features (256 or 512, chosen by the launcher); features beyond that use the same layout in global scratch.
// Quirks kept: sticky visible_fts_/patch_cache_ across levels with J zeroed per level, stop_/chi2_ never reset,
// fx used for both Jacobian rows, chi2 compared as float(chi2)/float(n_meas).
#include <cstdlib>

#include "seq.cuh"

namespace {

constexpr int AL_THREADS = 256;
constexpr int AL_WARPS = AL_THREADS / 32;
constexpr int ND = 20;                        // cached doubles per feature: xyz[3], j0[6], j1[6] (times fx / 2^level), px[2],
                                              // gradient moments sum(dx^2, dx dy, dy^2) of the level
constexpr int PF = SDVLB_ALIGN_SC_FLOATS;     // cached floats per feature: patch[16], dx[16], dy[16], 4 pad.  A row
                                              // stride of 52 words spreads the 16-byte loads of 8 neighbouring
                                              // threads over all 32 banks (48 would make them collide 4-way)
constexpr int NRED = 7;                       // b[6], chi2
constexpr int AL_PASSES_MAX = 16;             // features per alignment <= 16 * 256 for the factor-reuse bookkeeping

// Eight consecutive pixels starting at p, from aligned 32-bit words: a row of the 5x5 / 7x7 footprints costs two or
// three loads instead of five or seven byte loads (the load/store unit is what the residual phase waits for).
// Reads up to 3 bytes before p and up to 11 after it, inside the frame's pyramid allocation (256-B slack).
__device__ __forceinline__ uint64_t load_pixels8(const uint8_t* __restrict__ p, bool third) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  const unsigned sh = unsigned(a & 3) * 8;
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
  const uint32_t w2 = third ? __ldg(w + 2) : 0u;
  const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
  return (uint64_t(hi) << 32) | lo;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void store_Rt12(const DSE3& T, double* __restrict__ Rt) {
  double R[9];
  se3_rot(T, R);
#pragma unroll
  for (int i = 0; i < 9; i++) Rt[i] = R[i];
  Rt[9] = T.tx; Rt[10] = T.ty; Rt[11] = T.tz;
}

// Unpivoted LDL^T of a symmetric 6x6 matrix given by its upper triangle (row-major, 21 values), in registers.
// fac[0..14] = L (strictly lower, row-major: L10, L20, L21, L30 ...), fac[15..20] = 1 / D.  Returns false when a pivot
// is not safely positive (the caller then uses Eigen's pivoted algorithm, ldlt_solve6).  Same operations as
// ldlt_solve6_spd (common.cuh).
__device__ __forceinline__ bool ldlt_factor6(const double* __restrict__ up, double* __restrict__ fac) {
  double A[6][6];
  {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int q = r; q < 6; q++) { A[r][q] = up[k]; A[q][r] = up[k]; k++; }
  }
  double L[6][6], W[6][6], Dinv[6];
  double dmax = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) dmax = fmax(dmax, A[i][i]);
  const double tiny = dmax * 1e-13;
  bool ok = dmax > 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double d = A[k][k];
#pragma unroll
    for (int j = 0; j < k; j++) d -= L[k][j] * W[k][j];
    ok = ok && (d > tiny);
    const double inv = pivot_rcp(d);
    Dinv[k] = inv;
#pragma unroll
    for (int i = k + 1; i < 6; i++) {
      double s = A[i][k];
#pragma unroll
      for (int j = 0; j < k; j++) s -= L[i][j] * W[k][j];
      W[i][k] = s;
      L[i][k] = s * inv;
    }
  }
  int k = 0;
#pragma unroll
  for (int i = 1; i < 6; i++)
#pragma unroll
    for (int j = 0; j < i; j++) fac[k++] = L[i][j];
#pragma unroll
  for (int i = 0; i < 6; i++) fac[15 + i] = Dinv[i];
  return ok;
}

__device__ __forceinline__ void ldlt_apply6(const double* __restrict__ fac, const double b[6], double x[6]) {
  double L[6][6];
  {
    int k = 0;
#pragma unroll
    for (int i = 1; i < 6; i++)
#pragma unroll
      for (int j = 0; j < i; j++) L[i][j] = fac[k++];
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double s = b[i];
#pragma unroll
    for (int j = 0; j < i; j++) s -= L[i][j] * y[j];
    y[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] = y[i] * fac[15 + i];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
#pragma unroll
    for (int j = i + 1; j < 6; j++) s -= L[j][i] * x[j];
    x[i] = s;
  }
}

// What align_core needs to know about one alignment.
struct AlignView {
  const uint8_t* ref_pyr;
  const uint8_t* cur_pyr;
  int n;
  int fast;
  double* out_pose;        // 7 doubles
  double* cur_pose;        // 7 doubles: the current frame's pose slot (FrameDev::pose)
  int32_t* out_info;       // [0] n_meas of the last ComputeResiduals, [1] iterations run
  double* out_error;       // GetError()
  int32_t* out_cycles;     // optional: [4] SM cycles in PrecomputePatches / residuals / reduction / solve + update
  sdvlb_gn_iter* trace;    // optional
  int trace_cap;
  const double* forced_T;  // teacher forcing (parity tests)
  const int32_t* forced_iters;
  double* g_d;             // [ND][n - cap] overflow cache
  float* g_f;              // [n - cap][PF]
  int32_t* g_flags;        // [n] flags of the overflow features (indexed by feature)
};

// Shared-memory plan of the kernel (dynamic): caches of the first `cap` features, then the small reduction areas.
struct AlignSmem {
  double* d;               // [ND][cap]
  float* f;                // [cap][PF]
  uint8_t* flag;           // [cap]  bit0 visible (sticky), bit1 has a Jacobian at this level, bit2 inside the current
                           //        image in the iteration being evaluated, bit3 the feature observes a live point
  double* wred;            // [AL_WARPS][8] per-warp partials of an iteration
  double* hred;            // [AL_WARPS][21] per-warp partials of H_all
  double* hcor;            // [AL_WARPS][21] per-warp sums of the H_f of the features outside the current image
  unsigned* outside;       // [AL_PASSES_MAX][AL_WARPS] ballots of those features at the last factorisation
  int* wredi;              // [AL_WARPS][4]
  double* Hall;            // [21] sum of H_f over the features that have a Jacobian at this level
  double* Hcur;            // [21] H of the iteration being solved
  double* fac;             // [21] LDL^T factor
  double* red;             // [8]  b[6], chi2
  double* T;               // [7]  current relative pose
  double* Rt;              // [12] its rotation + translation
  int* ctrl;               // [0] continue, [1] level break, [2] n_meas, [3] n_inv, [4] n_vis_j
};
__host__ __device__ inline size_t align_smem_bytes(int cap) {
  return size_t(cap) * (ND * 8 + PF * 4 + 1) + 16 + (AL_WARPS * (8 + 21 + 21) + 21 * 3 + 8 + 7 + 12) * 8 +
         (AL_WARPS * 4 + 8 + AL_PASSES_MAX * AL_WARPS) * 4;
}
__device__ __forceinline__ void align_carve(unsigned char* mem, int cap, AlignSmem& s) {
  s.d = reinterpret_cast<double*>(mem);
  s.f = reinterpret_cast<float*>(s.d + size_t(ND) * cap);
  s.flag = reinterpret_cast<uint8_t*>(s.f + size_t(cap) * PF);
  // (offset arithmetic on `mem`, which is 16-byte aligned -- a pointer -> integer -> pointer round trip would hide from
  // the compiler that everything below lives in shared memory, and every access would become a generic LD / ST)
  const size_t off = (size_t(cap) * (ND * 8 + PF * 4 + 1) + 15) & ~size_t(15);
  s.wred = reinterpret_cast<double*>(mem + off);
  s.hred = s.wred + AL_WARPS * 8;
  s.hcor = s.hred + AL_WARPS * 21;
  s.Hall = s.hcor + AL_WARPS * 21;
  s.Hcur = s.Hall + 21;
  s.fac = s.Hcur + 21;
  s.red = s.fac + 21;
  s.T = s.red + 8;
  s.Rt = s.T + 7;
  s.wredi = reinterpret_cast<int*>(s.Rt + 12);
  s.ctrl = s.wredi + AL_WARPS * 4;
  s.outside = reinterpret_cast<unsigned*>(s.ctrl + 8);
}

// Features marshalled by the host (class API, AlignJobDev::feats).
struct JobFeatures {
  const sdvlb_align_feat* __restrict__ feats;
  __device__ __forceinline__ void load(int f, double& px0, double& px1, double& X, double& Y, double& Z, bool& valid) const {
    const sdvlb_align_feat ft = feats[f];
    px0 = ft.px[0]; px1 = ft.px[1];
    X = ft.v[0] * ft.depth; Y = ft.v[1] * ft.depth; Z = ft.v[2] * ft.depth;   // xyz_ref = v * depth (image_align.cc:160,235)
    valid = ft.valid != 0;
  }
};
// Features of a resident sequence's last frame; depth = |point - camera centre of that frame| (image_align.cc:159,234).
struct SeqFeatures {
  const SeqFeat* __restrict__ list;
  double C[3];
  __device__ __forceinline__ void load(int f, double& px0, double& px1, double& X, double& Y, double& Z, bool& valid) const {
    const SeqFeat& ft = list[f];
    px0 = ft.px[0]; px1 = ft.px[1];
    valid = (ft.flags & SEQF_HAS_POINT) != 0;
    const double dx = ft.pos[0] - C[0], dy = ft.pos[1] - C[1], dz = ft.pos[2] - C[2];
    const double depth = valid ? sqrt(dx * dx + dy * dy + dz * dz) : 1.0;
    X = ft.v[0] * depth; Y = ft.v[1] * depth; Z = ft.v[2] * depth;
  }
};

// PrecomputePatches for one feature (image_align.cc:208-267).  d / ds: the feature's double cache and its stride,
// fl: its float row.  Returns the new flags; adds the feature's H_f to Hacc and stores it at gH[k * n].
// Entry (r, q) of a feature's H_f = sum over its 16 pixels of J J^T, from the gradient moments (sa, sb, sc) and the
// scaled Jacobian rows.  One definition with explicit roundings: PrecomputePatches (H_all) and the correction for
// features outside the image must produce the same bits.
__device__ __forceinline__ double hf_entry(double sa, double sb, double sc, double j0r, double j0q, double j1r, double j1q) {
  const double aa = __dmul_rn(j0r, j0q);
  const double ab = __fma_rn(j0r, j1q, __dmul_rn(j1r, j0q));
  const double bb = __dmul_rn(j1r, j1q);
  return __fma_rn(sa, aa, __fma_rn(sb, ab, __dmul_rn(sc, bb)));
}
__constant__ int c_hr[21] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5};
__constant__ int c_hq[21] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};

// PrecomputePatches for one feature (image_align.cc:208-267).  d / ds: the feature's double cache and its stride,
// fl: its float row.  Returns the new flags; adds the feature's H_f to Hacc.
__device__ __forceinline__ int precompute_feature(int old_flags, const uint8_t* __restrict__ img1, int W, int Hh, float scale,
                                                  int border, double fs, double* __restrict__ d, int ds,
                                                  float* __restrict__ fl, double Hacc[21]) {
  int flags = old_flags & 9;   // J zeroed per level (image_align.cc:69), visibility sticky
  const bool valid = (old_flags & 8) != 0;
  const float u_ref = float(d[15 * ds] * double(scale));
  const float v_ref = float(d[16 * ds] * double(scale));
  const bool in_img = u_ref >= 0.f && v_ref >= 0.f && u_ref < float(W) && v_ref < float(Hh);   // guards the int cast
  const int ui = in_img ? int(floorf(u_ref)) : -1, vi = in_img ? int(floorf(v_ref)) : -1;
  if (!(valid && !(ui - border < 0 || vi - border < 0 || ui + border >= W || vi + border >= Hh))) return flags;
  flags = 11;
  const float su = u_ref - float(ui), sv = v_ref - float(vi);
  const float wtl = float((1.0 - su) * (1.0 - sv));
  const float wtr = float(su * (1.0 - sv));
  const float wbl = float((1.0 - su) * sv);
  const float wbr = float(su * sv);
  // 7x7 footprint: rows vi-3..vi+3, cols ui-3..ui+3
  float px[7][7];
#pragma unroll
  for (int r = 0; r < 7; r++) {
    const uint64_t v = load_pixels8(img1 + size_t(vi - 3 + r) * W + (ui - 3), true);
#pragma unroll
    for (int c = 0; c < 7; c++) px[r][c] = float(unsigned(v >> (8 * c)) & 0xffu);
  }
  double sa = 0, sb = 0, sc = 0;
#pragma unroll
  for (int y = 0; y < 4; y++) {
    float pv[4], gx[4], gy[4];
#pragma unroll
    for (int x = 0; x < 4; x++) {
      // p = &img[(vi+y-2)][ui-2+x]  -> px[y+1][x+1]
      const int r = y + 1, c = x + 1;
      const float val = wtl * px[r][c] + wtr * px[r][c + 1] + wbl * px[r + 1][c] + wbr * px[r + 1][c + 1];
      const float dx = 0.5f * ((wtl * px[r][c + 1] + wtr * px[r][c + 2] + wbl * px[r + 1][c + 1] + wbr * px[r + 1][c + 2]) -
                               (wtl * px[r][c - 1] + wtr * px[r][c] + wbl * px[r + 1][c - 1] + wbr * px[r + 1][c]));
      const float dy = 0.5f * ((wtl * px[r + 1][c] + wtr * px[r + 1][c + 1] + wbl * px[r + 2][c] + wbr * px[r + 2][c + 1]) -
                               (wtl * px[r - 1][c] + wtr * px[r - 1][c + 1] + wbl * px[r][c] + wbr * px[r][c + 1]));
      pv[x] = val; gx[x] = dx; gy[x] = dy;
      sa += double(dx) * double(dx);
      sb += double(dx) * double(dy);
      sc += double(dy) * double(dy);
    }
    reinterpret_cast<float4*>(fl)[y] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    reinterpret_cast<float4*>(fl)[4 + y] = make_float4(gx[0], gx[1], gx[2], gx[3]);
    reinterpret_cast<float4*>(fl)[8 + y] = make_float4(gy[0], gy[1], gy[2], gy[3]);
  }
  double j0[6], j1[6];
  jacobian3d_to_plane(d[0], d[ds], d[2 * ds], j0, j1);
#pragma unroll
  for (int r = 0; r < 6; r++) {
    j0[r] *= fs; j1[r] *= fs;
    d[(3 + r) * ds] = j0[r];
    d[(9 + r) * ds] = j1[r];
  }
  d[17 * ds] = sa; d[18 * ds] = sb; d[19 * ds] = sc;
  // sum over the 16 pixels of J J^T: constant for the whole level (inverse compositional)
  int k = 0;
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int q = r; q < 6; q++) {
      Hacc[k] += hf_entry(sa, sb, sc, j0[r], j0[q], j1[r], j1[q]);
      k++;
    }
  return flags;
}

// ComputeResiduals for one feature that is marked visible (image_align.cc:127-206).  Returns the new flags (bit2 = the
// projection fell inside the current image); adds to b / chi2 / the counters.
__device__ __forceinline__ int residual_feature(int flags, const double* __restrict__ Rt, const sdvlb_camera& cam,
                                                const uint8_t* __restrict__ img2, int W, int Hh, float scale, int border,
                                                const double* __restrict__ d, int ds, const float* __restrict__ fl,
                                                double b[6], float& chi2_f, int& n_meas, int& n_inv, int& n_visj) {
  flags &= 11;
  const double X = d[0], Y = d[ds], Z = d[2 * ds];
  const double xc = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9];
  const double yc = Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10];
  const double zc = Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11];
  const double pu = cam.u0 + cam.fx * xc / zc;   // Camera::Project (camera.cc:69-72)
  const double pv = cam.v0 + cam.fy * yc / zc;
  const float u_cur = float(pu * double(scale));
  const float v_cur = float(pv * double(scale));
  bool inside = u_cur >= 0.f && v_cur >= 0.f && u_cur < float(W) && v_cur < float(Hh);   // NaN / inf / out of range
  int ui = 0, vi = 0;
  if (inside) {
    ui = int(floorf(u_cur)); vi = int(floorf(v_cur));
    inside = !(ui < 0 || vi < 0 || ui - border < 0 || vi - border < 0 || ui + border >= W || vi + border >= Hh);
  }
  if (!inside) {
    if (flags & 2) n_inv++;
    return flags;
  }
  const float su = u_cur - float(ui), sv = v_cur - float(vi);
  const float wtl = float((1.0 - su) * (1.0 - sv));
  const float wtr = float(su * (1.0 - sv));
  const float wbl = float((1.0 - su) * sv);
  const float wbr = float(su * sv);
  float px[5][5];
#pragma unroll
  for (int r = 0; r < 5; r++) {
    const uint64_t v = load_pixels8(img2 + size_t(vi - 2 + r) * W + (ui - 2), false);
#pragma unroll
    for (int c = 0; c < 5; c++) px[r][c] = float(unsigned(v >> (8 * c)) & 0xffu);
  }
  double sdx = 0, sdy = 0;
  float chi = 0.0f;
  const float4* __restrict__ row = reinterpret_cast<const float4*>(fl);
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const float4 pt = row[y], gx = row[4 + y], gy = row[8 + y];
    const float ptv[4] = {pt.x, pt.y, pt.z, pt.w};
    const float gxv[4] = {gx.x, gx.y, gx.z, gx.w};
    const float gyv[4] = {gy.x, gy.y, gy.z, gy.w};
#pragma unroll
    for (int x = 0; x < 4; x++) {
      const float ic = wtl * px[y][x] + wtr * px[y][x + 1] + wbl * px[y + 1][x] + wbr * px[y + 1][x + 1];
      const float res = ic - ptv[x];
      chi += res * res;
      sdx += double(res) * double(gxv[x]);
      sdy += double(res) * double(gyv[x]);
    }
  }
  chi2_f += chi;
  n_meas += 16;
  if (flags & 2) {
#pragma unroll
    for (int r = 0; r < 6; r++) b[r] -= d[(3 + r) * ds] * sdx + d[(9 + r) * ds] * sdy;
    n_visj++;
    flags |= 4;
  }
  return flags;
}

template <typename Src>
__device__ void align_core(const AlignView& J, const Src& src, const PyrGeom& G, const DevParams& dp, int cap,
                           const AlignSmem& S, const double* T_ref, const double* T_cur) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = J.n;
  const int n_over = n > cap ? n - cap : 0;
  const sdvlb_params& P = dp.p;
  const sdvlb_camera& cam = dp.cam;

  // thread-0 state (image_align.cc:35-41)
  double chi2_ = 1e10, error_ = 1e10;
  bool stop_ = false;
  int n_meas_last = 0;
  int trace_n = 0;
  int forced_k = 0;
  DSE3 T, T_bk;

  if (n == 0) {   // image_align.cc:55-58: nothing to track, frame2 keeps its pose
    if (tid == 0) {
      for (int i = 0; i < 7; i++) { J.out_pose[i] = T_cur[i]; J.cur_pose[i] = T_cur[i]; }
      J.out_info[0] = 0; J.out_info[1] = 0;
      *J.out_error = 1e10;
    }
    return;
  }
  if (tid == 0) {
    const DSE3 T1 = se3_load(T_ref), T2 = se3_load(T_cur);
    T = se3_mul(T2, se3_inverse(T1));   // image_align.cc:66
    se3_store(T, S.T);
    store_Rt12(T, S.Rt);
  }
  // A feature is always handled by the same thread (f = tid, tid + 256, ...), in every phase: its caches are private
  // to that thread and need no barrier; only the totals and the H correction cross threads.
  for (int f = tid; f < n; f += AL_THREADS) {
    double px0, px1, X, Y, Z;
    bool valid;
    src.load(f, px0, px1, X, Y, Z, valid);
    if (f < cap) {
      S.flag[f] = valid ? 8 : 0;
      double* d = S.d + f;
      d[0] = X; d[cap] = Y; d[2 * cap] = Z; d[15 * cap] = px0; d[16 * cap] = px1;
    } else {
      J.g_flags[f] = valid ? 8 : 0;
      double* d = J.g_d + (f - cap);
      d[0] = X; d[n_over] = Y; d[2 * n_over] = Z; d[15 * n_over] = px0; d[16 * n_over] = px1;
    }
  }
  __syncthreads();

  long long cyc[4] = {0, 0, 0, 0};   // thread 0: PrecomputePatches, residuals, reduction, solve + update (SM cycles)
  long long t_mark = clock64();
  bool level_break = false;   // uniform (derived from shared state)
  for (int level = P.max_align_level; level >= P.min_align_level && !level_break; level--) {
    const int W = G.w[level], Hh = G.h[level];
    const uint8_t* __restrict__ img1 = J.ref_pyr + G.off[level];
    const uint8_t* __restrict__ img2 = J.cur_pyr + G.off[level];
    const float scale = 1.0f / float(1 << level);
    const int border = P.align_patch_size / 2 + 1;   // 3
    const int n_forced = J.forced_T ? J.forced_iters[level] : -1;
    if (n_forced == 0) continue;
    if (tid == 0) T_bk = T;
    int fac_state = 0;   // thread 0: 1 = S.fac is the LDL^T factor of H for the set of outside features in S.outside
    if (lane == 0)   // every warp keeps its own entries: no barrier needed
      for (int k = 0; k < AL_PASSES_MAX; k++) S.outside[k * AL_WARPS + warp] = 0u;

    const int max_its = J.forced_T ? n_forced : P.max_img_align_its;
    for (int it = 0; it < max_its; it++) {
      if (J.forced_T) {   // parity tests only: the schedule dictates the pose of every iteration
        if (tid == 0) { T = se3_load(J.forced_T + 7 * forced_k); se3_store(T, S.T); store_Rt12(T, S.Rt); forced_k++; }
        __syncthreads();
      }
      // ---- PrecomputePatches(level) on the first iteration (image_align.cc:208-267) + partials of H_all
      if (it == 0) {
        const double fs = cam.fx / double(1 << level);
        double Hacc[21];
#pragma unroll
        for (int k = 0; k < 21; k++) Hacc[k] = 0.0;
        for (int f = tid; f < n; f += AL_THREADS) {
          if (f < cap) {
            S.flag[f] = uint8_t(precompute_feature(S.flag[f], img1, W, Hh, scale, border, fs, S.d + f, cap,
                                                   S.f + size_t(f) * PF, Hacc));
          } else {
            J.g_flags[f] = precompute_feature(J.g_flags[f], img1, W, Hh, scale, border, fs, J.g_d + (f - cap), n_over,
                                              J.g_f + size_t(f - cap) * PF, Hacc);
          }
        }
#pragma unroll
        for (int k = 0; k < 21; k++) {
          const double s = warp_sum_d(Hacc[k]);
          if (lane == 0) S.hred[warp * 21 + k] = s;
        }
        if (tid == 0) { const long long t = clock64(); cyc[0] += t - t_mark; t_mark = t; }
      }

      // ---- ComputeResiduals (image_align.cc:127-206)
      double b[6] = {0, 0, 0, 0, 0, 0};
      float chi2_f = 0.0f;
      int n_meas = 0, n_inv = 0, n_visj = 0;
      double hc = 0.0;          // lane k < 21: entry k of the sum of H_f over this warp's features outside the image
      int set_changed = 0;      // lane 0: the set of those features differs from the one H was last factorised for
      for (int f0 = 0, pass = 0; f0 < n; f0 += AL_THREADS, pass++) {   // uniform trip count: every warp votes
        const int f = f0 + tid;
        bool outside = false;
        if (f < n) {
          const int flags = f < cap ? int(S.flag[f]) : J.g_flags[f];
          if (flags & 1) {
            const int inv_before = n_inv;
            int nf;
            if (f < cap) {
              nf = residual_feature(flags, S.Rt, cam, img2, W, Hh, scale, border, S.d + f, cap, S.f + size_t(f) * PF, b,
                                    chi2_f, n_meas, n_inv, n_visj);
              S.flag[f] = uint8_t(nf);
            } else {
              nf = residual_feature(flags, S.Rt, cam, img2, W, Hh, scale, border, J.g_d + (f - cap), n_over,
                                    J.g_f + size_t(f - cap) * PF, b, chi2_f, n_meas, n_inv, n_visj);
              J.g_flags[f] = nf;
            }
            outside = n_inv != inv_before;   // has a Jacobian at this level but projects outside the current image
          }
        }
        unsigned bal = __ballot_sync(0xffffffffu, outside);
        if (lane == 0 && pass < AL_PASSES_MAX) {
          if (S.outside[pass * AL_WARPS + warp] != bal) { S.outside[pass * AL_WARPS + warp] = bal; set_changed = 1; }
        }
        if (pass >= AL_PASSES_MAX && bal) set_changed = 1;   // beyond the bookkeeping: never reuse a factor
        // the owning warp re-forms H_f of each such feature, lane k computing entry k (fixed order: ascending feature)
        __syncwarp();   // the caches of a feature were written by its own lane
        while (bal) {
          const int src = __ffs(bal) - 1;
          bal &= bal - 1;
          const int fo = f0 + warp * 32 + src;
          if (lane < 21) {
            const double* __restrict__ d = fo < cap ? S.d + fo : J.g_d + (fo - cap);
            const int ds = fo < cap ? cap : n_over;
            const int r = c_hr[lane], q = c_hq[lane];
            hc += hf_entry(d[17 * ds], d[18 * ds], d[19 * ds], d[(3 + r) * ds], d[(3 + q) * ds], d[(9 + r) * ds], d[(9 + q) * ds]);
          }
        }
      }
      {
        double v[NRED];
#pragma unroll
        for (int r = 0; r < 6; r++) v[r] = b[r];
        v[6] = double(chi2_f);
#pragma unroll
        for (int k = 0; k < NRED; k++) {
          const double s = warp_sum_d(v[k]);
          if (lane == 0) S.wred[warp * 8 + k] = s;
        }
        const int m = int(__reduce_add_sync(0xffffffffu, unsigned(n_meas)));
        const int ni = int(__reduce_add_sync(0xffffffffu, unsigned(n_inv)));
        const int nv = int(__reduce_add_sync(0xffffffffu, unsigned(n_visj)));
        if (lane == 0) {
          S.wredi[warp * 4] = m; S.wredi[warp * 4 + 1] = ni; S.wredi[warp * 4 + 2] = nv; S.wredi[warp * 4 + 3] = set_changed;
        }
        if (lane < 21) S.hcor[warp * 21 + lane] = hc;
      }
      __syncthreads();
      if (tid == 0) { const long long t = clock64(); cyc[1] += t - t_mark; t_mark = t; }

      // ---- warp 0: totals in fixed order, H of this iteration, then the Optimize step on thread 0
      if (warp == 0) {
        if (it == 0 && lane < 21) {
          double s = 0;
#pragma unroll
          for (int w = 0; w < AL_WARPS; w++) s += S.hred[w * 21 + lane];
          S.Hall[lane] = s;
        }
        if (lane < NRED) {
          double tot = 0;
#pragma unroll
          for (int w = 0; w < AL_WARPS; w++) tot += S.wred[w * 8 + lane];
          S.red[lane] = tot;
        }
        int ti = 0;
        if (lane < 4) {
#pragma unroll
          for (int w = 0; w < AL_WARPS; w++) ti += S.wredi[w * 4 + lane];
        }
        const int nm = __shfl_sync(0xffffffffu, ti, 0);
        const int ninv = __shfl_sync(0xffffffffu, ti, 1);
        const int nvisj = __shfl_sync(0xffffffffu, ti, 2);
        const bool set_changed_any = __shfl_sync(0xffffffffu, ti, 3) != 0;
        // H: every feature with a Jacobian projects inside the image -> H_all; none does -> exactly 0 (as the
        // reference's empty sum); otherwise H_all minus the H_f of the ones outside
        if (ninv > 0 && lane < 21) {
          double h = 0.0;
          if (nvisj > 0) {
            double c = 0.0;
#pragma unroll
            for (int w = 0; w < AL_WARPS; w++) c += S.hcor[w * 21 + lane];
            h = S.Hall[lane] - c;
          }
          S.Hcur[lane] = h;
        }
        __syncwarp();

        if (lane == 0) {
          { const long long t = clock64(); cyc[2] += t - t_mark; t_mark = t; }
          const double* __restrict__ Hup = ninv > 0 ? S.Hcur : S.Hall;
          double bb[6], x[6];
#pragma unroll
          for (int r = 0; r < 6; r++) bb[r] = S.red[r];
          n_meas_last = nm;
          const double new_chi2 = double(float(S.red[6]) / float(nm));   // float / size_t -> float (image_align.cc:205)
          if (nm == 0) stop_ = true;
          // H.ldlt().solve(Jres) (image_align.cc:102): register-only LDL^T when H is safely positive definite (the
          // normal case; the factor of H_all serves every iteration of the level), Eigen's pivoted algorithm otherwise
          // (singular / empty systems, NaN propagation)
          bool solved = false;
          if (fac_state == 1 && !set_changed_any) {   // same H as the last factorisation: substitutions only
            ldlt_apply6(S.fac, bb, x);
            solved = true;
          } else {
            double fac[21];
            fac_state = 0;
            if (ldlt_factor6(Hup, fac)) {
              ldlt_apply6(fac, bb, x);
              solved = true;
#pragma unroll
              for (int k = 0; k < 21; k++) S.fac[k] = fac[k];
              fac_state = 1;
            }
          }
          if (!solved) {
            double H[36];
            int k = 0;
            for (int r = 0; r < 6; r++)
              for (int q = r; q < 6; q++) { H[r * 6 + q] = Hup[k]; H[q * 6 + r] = Hup[k]; k++; }
            ldlt_solve6(H, bb, x);
          }
          bool nan = false;
          if (isnan(x[0])) { stop_ = true; nan = true; }
          int flags = (nan ? 2 : 0) | (nm == 0 ? 4 : 0);
          int cont = 1;
          const bool reject = (it > 0 && new_chi2 > chi2_) || stop_;
          const DSE3 T_in = T;
          if (J.forced_T) {
            if (reject) flags |= 1;     // what the reference would have decided; control stays with the schedule
          } else if (reject) {
            T = T_bk;
            flags |= 1;
            cont = 0;
          } else {
            T_bk = T;
            double mx[6];
            for (int r = 0; r < 6; r++) mx[r] = -x[r];
            T = se3_mul(T, se3_exp(mx));   // image_align.cc:116
            chi2_ = new_chi2;
            double e = -1;
            for (int r = 0; r < 6; r++) e = fmax(e, fabs(x[r]));   // AbsMax (utils.cc:28-42)
            error_ = e;
            if (error_ <= 1e-10) cont = 0;
          }
          se3_store(T, S.T);
          store_Rt12(T, S.Rt);           // the pose the next iteration (or level) evaluates its residuals at
          S.ctrl[0] = cont;
          if (J.trace && trace_n < J.trace_cap) {
            sdvlb_gn_iter& rec = J.trace[trace_n];
            rec.level = level; rec.iter = it; rec.n_meas = nm; rec.flags = flags;
            se3_store(T_in, rec.T_in);
            int k = 0;
            for (int r = 0; r < 6; r++)
              for (int q = r; q < 6; q++) { rec.H[r * 6 + q] = Hup[k]; rec.H[q * 6 + r] = Hup[k]; k++; }
            for (int i = 0; i < 6; i++) { rec.b[i] = bb[i]; rec.x[i] = x[i]; }
            rec.chi2 = new_chi2;
          }
          trace_n++;
          { const long long t = clock64(); cyc[3] += t - t_mark; t_mark = t; }
        }
      }
      __syncthreads();
      if (!S.ctrl[0]) break;
    }
    // image_align.cc:73-76 (relocalisation only)
    if (tid == 0) {
      int lb = 0;
      if (!J.forced_T && J.fast && error_ > 0.01) { error_ = 1e10; lb = 1; }
      S.ctrl[1] = lb;
    }
    __syncthreads();
    level_break = S.ctrl[1] != 0;
  }

  // Programmatic dependent launch: the next kernel of the chain may be brought in; it waits for this grid to complete.
  // The trigger stays at the END on purpose: raised when the last pyramid level starts (to hide the ~9 us between this
  // kernel's end and the first SearchPoint CTA) it costs throughput -- 197 k instead of 223 k frames/s, A/B on one box
  // -- because the ~850 SearchPoint CTAs of a group then sit in griddepcontrol.wait on SM slots that the frame-build
  // kernels of the other groups need.
  sdvlb_launch_dependents();
  if (tid == 0) {
    const DSE3 out = se3_mul(T, se3_load(T_ref));   // frame2_->SetPose(current_se3 * frame1_->GetPose())
    se3_store(out, J.out_pose);
    se3_store(out, J.cur_pose);
    J.out_info[0] = n_meas_last;
    J.out_info[1] = trace_n;
    *J.out_error = error_;
    if (J.out_cycles)
      for (int i = 0; i < 4; i++) J.out_cycles[i] = int(cyc[i]);
  }
}

struct AlignArgs {
  PyrGeom g;
  DevParams dp;
  int cap;
};

__global__ void __launch_bounds__(AL_THREADS, 1) image_align_kernel(const AlignJobDev* __restrict__ jobs,
                                                                    const __grid_constant__ AlignArgs A) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ double s_Tref[7], s_Tcur[7];
  AlignSmem S;
  align_carve(s_dyn, A.cap, S);
  const AlignJobDev& Jd = jobs[blockIdx.x];
  if (Jd.n < 0) return;
  if (threadIdx.x < 7) { s_Tref[threadIdx.x] = Jd.T_ref[threadIdx.x]; s_Tcur[threadIdx.x] = Jd.T_cur[threadIdx.x]; }
  __syncthreads();
  AlignView J;
  J.ref_pyr = Jd.ref.pyr; J.cur_pyr = Jd.cur.pyr;
  J.n = Jd.n; J.fast = Jd.fast;
  J.out_pose = Jd.out_pose; J.cur_pose = Jd.cur.pose; J.out_info = Jd.out_info; J.out_error = Jd.out_error;
  J.out_cycles = Jd.out_cycles;
  J.trace = Jd.trace; J.trace_cap = Jd.trace_cap; J.forced_T = Jd.forced_T; J.forced_iters = Jd.forced_iters;
  J.g_d = Jd.sc_d;
  J.g_f = Jd.sc_f;
  J.g_flags = Jd.sc_flags;
  JobFeatures src{Jd.feats};
  align_core(J, src, A.g, A.dp, A.cap, S, s_Tref, s_Tcur);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ resident sequences
// One CTA per sequence of the step: the mapping thread's commands, the motion-model prior and ImageAlign.
namespace {

__global__ void __launch_bounds__(AL_THREADS, 1) seq_align_kernel(const __grid_constant__ SeqStepArgs A, int cap) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ double s_Tref[7], s_Tcur[7], s_C[3];
  __shared__ int s_base;
  AlignSmem Sm;
  align_carve(s_dyn, cap, Sm);
  SeqState* S = A.seq[blockIdx.x];
  const int tid = threadIdx.x;
  sdvlb_grid_dependency_wait();   // queued behind the previous step's post kernel: the sequence state is its output
  if (A.cmd_range[blockIdx.x].y > 0) {   // the mapping thread's commands for this sequence (keyframes), in order
    seq_apply_commands(A.cmds, A.cmd_range[blockIdx.x], A.dp, &s_base);
    __threadfence_block();
    __syncthreads();
  }
  const int has_last = S->has_last;
  const int n = has_last ? S->n_list : 0;
  if (tid == 0) {
    const DSE3 T_last = se3_load(S->T_last);
    const DSE3 Twc = se3_inverse(T_last);   // Frame::GetWorldPosition (frame.h)
    s_C[0] = Twc.tx; s_C[1] = Twc.ty; s_C[2] = Twc.tz;
    const DSE3 prior = se3_mul(se3_exp(S->vel), T_last);   // SDVL::SetMotionModel (sdvl.cc:278-281)
    se3_store(T_last, s_Tref);
    se3_store(prior, s_Tcur);
    S->align_info[0] = 0; S->align_info[1] = 0;
  }
  __syncthreads();
  if (!has_last || S->hold) return;   // no track yet (no reset) / on hold: the post kernel reports which
  AlignView J;
  J.ref_pyr = S->last.pyr; J.cur_pyr = A.cur[blockIdx.x].pyr;
  J.n = n; J.fast = 0;
  J.out_pose = S->align_pose; J.cur_pose = A.cur[blockIdx.x].pose; J.out_info = S->align_info;
  J.out_error = &S->align_error; J.out_cycles = S->align_cycles;
  J.trace = nullptr; J.trace_cap = 0; J.forced_T = nullptr; J.forced_iters = nullptr;
  // the sequence's own scratch, carved by ITS capacity (sequences of one submission may differ)
  const size_t nn = size_t(S->max_feats);
  double* sc_d = reinterpret_cast<double*>(S->align_scratch);
  J.g_d = sc_d;
  J.g_f = reinterpret_cast<float*>(S->align_scratch + nn * SDVLB_ALIGN_SC_DOUBLES * 8);
  J.g_flags = reinterpret_cast<int32_t*>(S->align_scratch + nn * (SDVLB_ALIGN_SC_DOUBLES * 8 + SDVLB_ALIGN_SC_FLOATS * 4));
  SeqFeatures src;
  src.list = S->list[S->cur];
  src.C[0] = s_C[0]; src.C[1] = s_C[1]; src.C[2] = s_C[2];
  align_core(J, src, A.g, A.dp, cap, Sm, s_Tref, s_Tcur);
}

// Features cached in shared memory: 256 (85 KB) covers the tracking configurations of the reference (at most
// max(max_matches, points per keyframe) features per frame); 512 (170 KB) is used when the caller's bound is larger.
// More features than that still work (global scratch), more slowly.
int align_cap_for(int n_bound) {
  static const int forced = [] { const char* e = getenv("SDVLB_ALIGN_CAP"); return e ? atoi(e) : 0; }();   // experiment knob
  if (forced == 256 || forced == 512) return forced;
  return n_bound <= 256 ? 256 : 512;
}

}  // namespace

size_t sdvlb_align_scratch_bytes(int n) { return SDVLB_ALIGN_SC_BYTES(n) + 256; }

cudaError_t sdvlb_launch_align(const void* d_jobs, int n_jobs, int n_bound, const PyrGeom& g, const DevParams& dp,
                               cudaStream_t stream) {
  AlignArgs A;
  A.g = g;
  A.dp = dp;
  A.cap = align_cap_for(n_bound);
  const size_t dyn = align_smem_bytes(A.cap);
  SDVLB_PREPARE(image_align_kernel, dyn);
  image_align_kernel<<<n_jobs, AL_THREADS, dyn, stream>>>(static_cast<const AlignJobDev*>(d_jobs), A);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_seq_align(const SeqStepArgs& A, int n_bound, cudaStream_t stream) {
  const int cap = align_cap_for(n_bound);
  const size_t dyn = align_smem_bytes(cap);
  SDVLB_PREPARE(seq_align_kernel, dyn);
  return sdvlb_launch_dependent(seq_align_kernel, dim3(A.n), dim3(AL_THREADS), dyn, stream, A, cap);
}

size_t sdvlb_align_job_size() { return sizeof(AlignJobDev); }
