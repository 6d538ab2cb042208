// align.cu — K3: ImageAlign::ComputePose (image_align.cc:46-267) as ONE persistent kernel per call: all pyramid
// levels and all Gauss-Newton iterations run on the device (one CTA per alignment job = per sequence), including the
// 6x6 LDLT solve, the SE3 update T <- T*Exp(-x) and the reference's accept/rollback logic.  No host round trip, no
// atomics: per-feature partials go through shared memory and a fixed-order tree.
//
// Arithmetic mirrors the reference's type ledger: fp32 pixel weights / patch / gradients / residual,
// fp64 geometry, Jacobian, H, b, pose.  The reference's per-pixel J = (dx*Jp0 + dy*Jp1)*fx/2^l is linear in (dx,dy),
// so per feature b_f = -s*(Jp0*sum(dx*res) + Jp1*sum(dy*res)) and H_f = s^2*(A Jp0Jp0' + B(Jp0Jp1'+Jp1Jp0') + C Jp1Jp1')
// with A,B,C = sum(dx^2, dx*dy, dy^2): identical in exact arithmetic, and within 1e-15 relative in fp64.
// Quirks kept: sticky visible_fts_/patch_cache_ across levels with J zeroed per level, stop_/chi2_ never reset,
// fx used for both Jacobian rows, chi2 compared as float(chi2)/float(n_meas).
#include "common.cuh"

namespace {

constexpr int AL_THREADS = 256;
constexpr int NV = 30;   // reduced values: H(21) b(6) chi2 n_meas + pad

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

struct AlignArgs {
  PyrGeom g;
  DevParams dp;
};

__global__ void __launch_bounds__(AL_THREADS) image_align_kernel(const AlignJobDev* __restrict__ jobs,
                                                                 const __grid_constant__ AlignArgs A) {
  extern __shared__ double s_dyn_part[];                 // NV x AL_THREADS partials (60 KB, opt-in dynamic)
  double (*s_part)[AL_THREADS] = reinterpret_cast<double (*)[AL_THREADS]>(s_dyn_part);
  __shared__ double s_red[NV];
  __shared__ double s_T[7];        // current relative pose
  __shared__ double s_Rt[12];      // rotation + translation of s_T
  __shared__ int s_ctrl[4];        // [0] continue flag

  const AlignJobDev& J = jobs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = J.n;
  const sdvlb_params& P = A.dp.p;
  const sdvlb_camera& cam = A.dp.cam;

  float* __restrict__ c_patch = J.sc_f;
  float* __restrict__ c_dx = J.sc_f + size_t(n) * 16;
  float* __restrict__ c_dy = J.sc_f + size_t(n) * 32;
  double* __restrict__ c_xyz = J.sc_d;
  double* __restrict__ c_j0 = J.sc_d + size_t(n) * 3;
  double* __restrict__ c_j1 = J.sc_d + size_t(n) * 9;
  double* __restrict__ c_H = J.sc_d + size_t(n) * 15;   // per-feature J J^T sum of the level, [k * n + f], k < 21
  int32_t* __restrict__ c_flags = J.sc_flags;

  // thread-0 state (image_align.cc:35-41)
  double chi2_ = 1e10, error_ = 1e10;
  bool stop_ = false;
  int n_meas_last = 0;
  int trace_n = 0;
  int forced_k = 0;
  DSE3 T, T_bk;

  if (n < 0) return;   // resident sequence without a reference frame yet: nothing to align
  if (n == 0) {   // image_align.cc:55-58: nothing to track, frame2 keeps its pose
    if (tid == 0) {
      for (int i = 0; i < 7; i++) { J.out_pose[i] = J.T_cur[i]; J.cur.pose[i] = J.T_cur[i]; }
      J.out_info[0] = 0; J.out_info[1] = 0;
      *J.out_error = 1e10;
    }
    return;
  }
  if (tid == 0) {
    const DSE3 T1 = se3_load(J.T_ref), T2 = se3_load(J.T_cur);
    T = se3_mul(T2, se3_inverse(T1));   // image_align.cc:66
    se3_store(T, s_T);
  }
  for (int f = tid; f < n; f += AL_THREADS) {
    c_flags[f] = 0;
    const sdvlb_align_feat ft = J.feats[f];
    c_xyz[3 * f + 0] = ft.v[0] * ft.depth;   // xyz_ref = v * depth (image_align.cc:160,235)
    c_xyz[3 * f + 1] = ft.v[1] * ft.depth;
    c_xyz[3 * f + 2] = ft.v[2] * ft.depth;
  }
  __syncthreads();

  long long cyc[4] = {0, 0, 0, 0};   // thread 0: PrecomputePatches, residuals, reduction, solve + update (SM cycles)
  long long t_mark = clock64();
  bool level_break = false;   // uniform (derived from shared state)
  for (int level = P.max_align_level; level >= P.min_align_level && !level_break; level--) {
    const int W = A.g.w[level], Hh = A.g.h[level];
    const uint8_t* __restrict__ img1 = J.ref.pyr + A.g.off[level];
    const uint8_t* __restrict__ img2 = J.cur.pyr + A.g.off[level];
    const float scale = 1.0f / float(1 << level);
    const int border = P.align_patch_size / 2 + 1;   // 3
    const int n_forced = J.forced_T ? J.forced_iters[level] : -1;
    if (n_forced == 0) continue;
    if (tid == 0) T_bk = T;

    const int max_its = J.forced_T ? n_forced : P.max_img_align_its;
    for (int it = 0; it < max_its; it++) {
      if (J.forced_T) {
        if (tid == 0) { T = se3_load(J.forced_T + 7 * forced_k); se3_store(T, s_T); forced_k++; }
      }
      if (tid == 0) {
        double R[9];
        se3_rot(T, R);
        for (int i = 0; i < 9; i++) s_Rt[i] = R[i];
        s_Rt[9] = T.tx; s_Rt[10] = T.ty; s_Rt[11] = T.tz;
      }
      // ---- PrecomputePatches(level) on the first iteration (image_align.cc:208-267)
      if (it == 0) {
        const double fs = cam.fx / double(1 << level);
        for (int f = tid; f < n; f += AL_THREADS) {
          const sdvlb_align_feat ft = J.feats[f];
          int flags = c_flags[f] & 1;   // J zeroed per level (image_align.cc:69), visibility sticky
          const float u_ref = float(ft.px[0] * double(scale));
          const float v_ref = float(ft.px[1] * double(scale));
          const bool in_img = u_ref >= 0.f && v_ref >= 0.f && u_ref < float(W) && v_ref < float(Hh);   // guards the int cast
          const int ui = in_img ? int(floorf(u_ref)) : -1, vi = in_img ? int(floorf(v_ref)) : -1;
          if (ft.valid && !(ui - border < 0 || vi - border < 0 || ui + border >= W || vi + border >= Hh)) {
            flags = 3;
            double j0[6], j1[6];
            jacobian3d_to_plane(c_xyz[3 * f], c_xyz[3 * f + 1], c_xyz[3 * f + 2], j0, j1);
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
            for (int y = 0; y < 4; y++)
#pragma unroll
              for (int x = 0; x < 4; x++) {
                // p = &img[(vi+y-2)][ui-2+x]  -> px[y+1][x+1]
                const int r = y + 1, c = x + 1;
                const float val = wtl * px[r][c] + wtr * px[r][c + 1] + wbl * px[r + 1][c] + wbr * px[r + 1][c + 1];
                const float dx = 0.5f * ((wtl * px[r][c + 1] + wtr * px[r][c + 2] + wbl * px[r + 1][c + 1] + wbr * px[r + 1][c + 2]) -
                                         (wtl * px[r][c - 1] + wtr * px[r][c] + wbl * px[r + 1][c - 1] + wbr * px[r + 1][c]));
                const float dy = 0.5f * ((wtl * px[r + 1][c] + wtr * px[r + 1][c + 1] + wbl * px[r + 2][c] + wbr * px[r + 2][c + 1]) -
                                         (wtl * px[r - 1][c] + wtr * px[r - 1][c + 1] + wbl * px[r][c] + wbr * px[r][c + 1]));
                c_patch[f * 16 + y * 4 + x] = val;
                c_dx[f * 16 + y * 4 + x] = dx;
                c_dy[f * 16 + y * 4 + x] = dy;
                sa += double(dx) * double(dx);
                sb += double(dx) * double(dy);
                sc += double(dy) * double(dy);
              }
#pragma unroll
            for (int r = 0; r < 6; r++) { j0[r] *= fs; j1[r] *= fs; c_j0[6 * f + r] = j0[r]; c_j1[6 * f + r] = j1[r]; }
            // sum over the 16 pixels of J J^T: constant for the whole level (inverse compositional), so it is formed
            // once here instead of in every Gauss-Newton iteration
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
              for (int q = r; q < 6; q++) {
                c_H[size_t(k) * n + f] = sa * j0[r] * j0[q] + sb * (j0[r] * j1[q] + j1[r] * j0[q]) + sc * j1[r] * j1[q];
                k++;
              }
          }
          c_flags[f] = flags;
        }
      }
      __syncthreads();
      if (tid == 0) { const long long t = clock64(); cyc[0] += t - t_mark; t_mark = t; }

      // ---- ComputeResiduals (image_align.cc:127-206)
      double acc[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) acc[i] = 0.0;
      float chi2_f = 0.0f;
      int n_meas = 0;
      for (int f = tid; f < n; f += AL_THREADS) {
        const int flags = c_flags[f];
        if (!(flags & 1)) continue;
        const double X = c_xyz[3 * f], Y = c_xyz[3 * f + 1], Z = c_xyz[3 * f + 2];
        const double xc = s_Rt[0] * X + s_Rt[1] * Y + s_Rt[2] * Z + s_Rt[9];
        const double yc = s_Rt[3] * X + s_Rt[4] * Y + s_Rt[5] * Z + s_Rt[10];
        const double zc = s_Rt[6] * X + s_Rt[7] * Y + s_Rt[8] * Z + s_Rt[11];
        const double pu = cam.u0 + cam.fx * xc / zc;   // Camera::Project (camera.cc:69-72)
        const double pv = cam.v0 + cam.fy * yc / zc;
        const float u_cur = float(pu * double(scale));
        const float v_cur = float(pv * double(scale));
        if (!(u_cur >= 0.f && v_cur >= 0.f && u_cur < float(W) && v_cur < float(Hh))) continue;   // NaN/inf/out of range
        const int ui = int(floorf(u_cur)), vi = int(floorf(v_cur));
        if (ui < 0 || vi < 0 || ui - border < 0 || vi - border < 0 || ui + border >= W || vi + border >= Hh) continue;
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
        const float4* pp = reinterpret_cast<const float4*>(c_patch + f * 16);
        const float4* pdx = reinterpret_cast<const float4*>(c_dx + f * 16);
        const float4* pdy = reinterpret_cast<const float4*>(c_dy + f * 16);
#pragma unroll
        for (int y = 0; y < 4; y++) {
          const float4 pt = pp[y], gx = pdx[y], gy = pdy[y];
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
          double j0[6], j1[6];
#pragma unroll
          for (int r = 0; r < 6; r++) { j0[r] = c_j0[6 * f + r]; j1[r] = c_j1[6 * f + r]; }
#pragma unroll
          for (int k = 0; k < 21; k++) acc[k] += c_H[size_t(k) * n + f];
#pragma unroll
          for (int r = 0; r < 6; r++) acc[21 + r] -= j0[r] * sdx + j1[r] * sdy;
        }
      }
      acc[27] = double(chi2_f);
      acc[28] = double(n_meas);
#pragma unroll
      for (int i = 0; i < NV; i++) s_part[i][tid] = acc[i];
      __syncthreads();
      if (tid == 0) { const long long t = clock64(); cyc[1] += t - t_mark; t_mark = t; }
      // fixed-order tree: warp w reduces rows w, w+8, ...
      for (int v = warp; v < 29; v += AL_THREADS / 32) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < AL_THREADS / 32; k++) s += s_part[v][lane + 32 * k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_red[v] = s;
      }
      __syncthreads();

      // ---- Optimize step on thread 0 (image_align.cc:91-124)
      if (tid == 0) {
        { const long long t = clock64(); cyc[2] += t - t_mark; t_mark = t; }
        double Hm[6][6], b[6], x[6];
        {
          int k = 0;
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int q = r; q < 6; q++) { Hm[r][q] = s_red[k]; Hm[q][r] = s_red[k]; k++; }
        }
#pragma unroll
        for (int r = 0; r < 6; r++) b[r] = s_red[21 + r];
        const int nm = int(s_red[28]);
        n_meas_last = nm;
        const double new_chi2 = double(float(s_red[27]) / float(nm));   // float / size_t -> float (image_align.cc:205)
        if (nm == 0) stop_ = true;
        // H.ldlt().solve(Jres) (image_align.cc:102): register-only LDL^T when H is safely positive definite (the normal
        // case), Eigen's pivoted algorithm otherwise (singular / empty systems, NaN propagation)
        if (!ldlt_solve6_spd(Hm, b, x)) {
          double H[36];
          for (int r = 0; r < 6; r++)
            for (int q = 0; q < 6; q++) H[r * 6 + q] = Hm[r][q];
          ldlt_solve6(H, b, x);
        }
        bool nan = false;
        if (isnan(x[0])) { stop_ = true; nan = true; }
        int flags = (nan ? 2 : 0) | (nm == 0 ? 4 : 0);
        int cont = 1;
        const bool reject = (it > 0 && new_chi2 > chi2_) || stop_;
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
        if (J.trace && trace_n < J.trace_cap) {
          sdvlb_gn_iter& rec = J.trace[trace_n];
          rec.level = level; rec.iter = it; rec.n_meas = nm; rec.flags = flags;
          for (int i = 0; i < 7; i++) rec.T_in[i] = s_T[i];
          for (int r = 0; r < 6; r++)
            for (int q = 0; q < 6; q++) rec.H[r * 6 + q] = Hm[r][q];
          for (int i = 0; i < 6; i++) { rec.b[i] = b[i]; rec.x[i] = x[i]; }
          rec.chi2 = new_chi2;
        }
        trace_n++;
        se3_store(T, s_T);
        s_ctrl[0] = cont;
        { const long long t = clock64(); cyc[3] += t - t_mark; t_mark = t; }
      }
      __syncthreads();
      if (!s_ctrl[0]) break;
    }
    // image_align.cc:73-76 (relocalisation only)
    if (tid == 0) {
      int lb = 0;
      if (!J.forced_T && J.fast && error_ > 0.01) { error_ = 1e10; lb = 1; }
      s_ctrl[1] = lb;
    }
    __syncthreads();
    level_break = s_ctrl[1] != 0;
  }

  if (tid == 0) {
    const DSE3 out = se3_mul(T, se3_load(J.T_ref));   // frame2_->SetPose(current_se3 * frame1_->GetPose())
    se3_store(out, J.out_pose);
    se3_store(out, J.cur.pose);
    J.out_info[0] = n_meas_last;
    J.out_info[1] = trace_n;
    *J.out_error = error_;
    if (J.out_cycles)
      for (int i = 0; i < 4; i++) J.out_cycles[i] = int(cyc[i]);
  }
}

}  // namespace

size_t sdvlb_align_scratch_floats(int n) { return size_t(n) * 48; }
size_t sdvlb_align_scratch_doubles(int n) { return size_t(n) * SDVLB_ALIGN_SC_DOUBLES; }

cudaError_t sdvlb_launch_align(const void* d_jobs, int n_jobs, const PyrGeom& g, const DevParams& dp,
                               cudaStream_t stream) {
  AlignArgs A;
  A.g = g;
  A.dp = dp;
  const size_t dyn = sizeof(double) * NV * AL_THREADS;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(image_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  sdvlb_common_carveout(image_align_kernel);
  image_align_kernel<<<n_jobs, AL_THREADS, dyn, stream>>>(static_cast<const AlignJobDev*>(d_jobs), A);
  return cudaGetLastError();
}

size_t sdvlb_align_job_size() { return sizeof(AlignJobDev); }
