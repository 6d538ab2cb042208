// seq.cu — resident sequences: everything SDVL::ProcessFrame does after ImageAlign and SearchPoint, on the device.
//
//   seq_apply_kernel : host commands for sequences that are not part of the step being submitted (the others get
//                      theirs applied by their own seq_align_kernel CTA, align.cu)
//   [seq_align_kernel, search_seq_kernel: align.cu / search.cu]
//   seq_post_kernel  : FeatureAlign::ProjectPoint binning + SelectPoints (feature_align.cc:88-150,323-339),
//                      SelectInliers (RANSAC, :152-216), OptimizePose / RescueOutliers / RemoveOutliers (:73-82,
//                      :218-256), ConvergePose (:341-421), GetMotionModel (sdvl.cc:266-276), CalcTrackingQuality
//                      (sdvl.cc:240-264) and Map::NeedKeyframe (map.cc:170-188); writes the sequence's next feature list
//                      and its pinned host result
//   pose_call_kernel : SelectInliers / OptimizePose alone (sdvlb_select_inliers, sdvlb_optimize_pose)
//
// One CTA per sequence.  SelectPoints is order-exact without walking the reference's per-cell std::list: the reference
// visits the 32-px cells in cell_order_ order, inside a cell the points by descending Point::Score() (stable), stops a
// cell at its first match and everything at max_matches matches.  Here every candidate gets its rank inside its cell
// (score desc, candidate index asc; candidates of a cell are chained through a shared-memory list, so a rank costs a
// walk over the handful of candidates of that cell), a cell's match is its lowest-ranked found candidate, a prefix sum
// over the cells in cell_order_ gives both the max_matches cut-off and each match's index in fs_found.  RANSAC
// evaluates hypotheses in batches (8, 24, then the rest of max_ransac_its), one thread per hypothesis (5-point
// Gauss-Newton in registers) with the supporter counts taken by all threads, and after each batch one thread replays
// the reference's adaptive iteration count over the results: the same hypothesis wins and rand() advances by exactly
// the number of draws the reference makes (with mostly-inlier matches the loop ends inside the first batch).  The next
// frame's random_shuffle of cell_order_ (the only other rand() consumer) runs on a fifth warp while the other four
// refine the pose.
#include <climits>

#include "seq.cuh"

namespace {

// Main threads of the FeatureAlign kernel.  192 instead of the original 128: a frame has up to max_matches = 150
// observations, which 128 threads took in two passes in every Gauss-Newton iteration of OptimizePose (and the rank /
// SelectPoints loops over ~200 candidates in two); A/B on one box, three alternating runs each: kernel 95.8 -> 85.1 us
// per 64 sequences, tracked frames/s 225 k -> 237 k (160 threads 88 us, 256 threads 89 us).
#ifndef SDVLB_PO_MAIN
#define SDVLB_PO_MAIN 192
#endif
constexpr int PO_MAIN = SDVLB_PO_MAIN;        // threads doing the FeatureAlign work (named barrier 1)
constexpr int PO_THREADS = PO_MAIN + 32;      // + the shuffle warp
constexpr int NVP = 28;           // A (21, upper triangle) + b (6) + chi2
constexpr int RANSAC_MAX_PTS = 8;
constexpr int HYP_MAX = 256;      // max_ransac_its supported

constexpr double KMADNorm = 1.4826;            // feature_align.h
constexpr double KTukeyC = 4.6851 * 4.6851;

__device__ __forceinline__ void main_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PO_MAIN) : "memory"); }

__device__ __forceinline__ void store_Rt(const DSE3& T, double* Rt) {
  double R[9];
  se3_rot(T, R);
#pragma unroll
  for (int i = 0; i < 9; i++) Rt[i] = R[i];
  Rt[9] = T.tx; Rt[10] = T.ty; Rt[11] = T.tz;
}

// error = (SimpleProject(v) - SimpleProject(se3 * pos)) * 2^-level (feature_align.cc:269-273)
__device__ __forceinline__ void reproj(const double* __restrict__ Rt, double a0, double a1, double X, double Y, double Z,
                                       double s, double& x, double& y, double& z, double& ex, double& ey) {
  x = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9];
  y = Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10];
  z = Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11];
  ex = (a0 - x / z) * s;
  ey = (a1 - y / z) * s;
}

// One observation's contribution to A, b, chi2 (feature_align.cc:389-398).  fp64 issue rate is what bounds these
// kernels, so the reference's expressions are regrouped to the fewest fp64 instructions (same mathematics, rounding
// differs in the last bits): one reciprocal of z serves the projection and the Jacobian, the Tukey argument
// (|e|/scale)^2 is e.e * 1/scale^2 (no square root, no division), 2^-level multiplies the weights instead of the
// twelve Jacobian entries, and the structural zeros of Jacobian3DToPlane (J0[1] = J1[0] = 0) are not multiplied out.
// inv_scale2 = 1 / scale^2.
__device__ __forceinline__ void accumulate_obs(const double* __restrict__ Rt, const PoseProblem& P, int i, double inv_scale2,
                                               double acc[NVP]) {
  const double s = P.o_scale[i];
  const double X = P.o_pos[3 * i], Y = P.o_pos[3 * i + 1], Z = P.o_pos[3 * i + 2];
  const double x = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9];
  const double y = Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10];
  const double z = Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11];
  const double zi = pivot_rcp(z);
  const double bx = x * zi, by = y * zi;
  const double ex = (P.o_a[2 * i] - bx) * s, ey = (P.o_a[2 * i + 1] - by) * s;   // error * 2^-level
  // Jacobian3DToPlane (extra/utils.cc:99-118) without the level factor
  const double a = -zi;                 // J0[0] = J1[1]
  const double j02 = bx * zi, j12 = by * zi;
  const double j03 = bx * by;           // x y / z^2 = -J1[4]
  const double j04 = -(1.0 + bx * bx);
  const double j05 = by;
  const double j13 = 1.0 + by * by;
  const double j15 = -bx;
  const double J0[6] = {a, 0.0, j02, j03, j04, j05};
  const double J1[6] = {0.0, a, j12, j13, -j03, j15};
  const double e2 = ex * ex + ey * ey;
  const double x_square = e2 * inv_scale2;
  double w = 0.0;
  if (x_square <= KTukeyC) {            // GetTukeyValue (feature_align.cc:423-431)
    const double tmp = 1.0 - x_square * (1.0 / KTukeyC);
    w = tmp * tmp;
  }
  const double ws = w * s, wss = ws * s;   // J rows carry 2^-level each
  // rows 0 and 1 of A: only one of the two Jacobian rows is non-zero there
  acc[0] += (a * a) * wss;              // (0,0)
  // (0,1) = 0
  acc[2] += (a * j02) * wss;  acc[3] += (a * j03) * wss;  acc[4] += (a * j04) * wss;  acc[5] += (a * j05) * wss;
  acc[6] += (a * a) * wss;              // (1,1)
  acc[7] += (a * j12) * wss;  acc[8] += (a * j13) * wss;  acc[9] += (a * -j03) * wss; acc[10] += (a * j15) * wss;
  int k = 11;
#pragma unroll
  for (int r = 2; r < 6; r++)
#pragma unroll
    for (int q = r; q < 6; q++) { acc[k] += (J0[r] * J0[q] + J1[r] * J1[q]) * wss; k++; }
#pragma unroll
  for (int r = 0; r < 6; r++) acc[21 + r] -= (J0[r] * ex + J1[r] * ey) * ws;
  acc[27] += e2 * w;
}

// error.norm() <= thr for error = (a - SimpleProject(Rt * pos)) * s (feature_align.cc:269-274), without the two
// divisions and the square root: |a z - (x, y)|^2 s^2 <= thr^2 z^2.  Both sides carry a relative rounding error below
// 1e-14; inside a 1e-11 band around equality the reference's own expression decides, so the outcome is always the
// reference's.
__device__ __forceinline__ bool within_threshold(const double* __restrict__ Rt, const PoseProblem& P, int i, double thr) {
  const double s = P.o_scale[i];
  const double X = P.o_pos[3 * i], Y = P.o_pos[3 * i + 1], Z = P.o_pos[3 * i + 2];
  const double a0 = P.o_a[2 * i], a1 = P.o_a[2 * i + 1];
  const double x = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9];
  const double y = Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10];
  const double z = Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11];
  const double dx = a0 * z - x, dy = a1 * z - y;
  const double lhs = (dx * dx + dy * dy) * (s * s);
  const double tz = thr * z;
  const double rhs = tz * tz;
  if (fabs(lhs - rhs) > 1e-11 * rhs) return lhs < rhs;
  const double ex = (a0 - x / z) * s, ey = (a1 - y / z) * s;
  return sqrt(ex * ex + ey * ey) <= thr;
}

// A.ldlt().solve(b): register LDL^T when A is safely positive definite, Eigen's pivoted algorithm otherwise
__device__ __forceinline__ void solve6(const double acc[NVP], double x[6]) {
  double Hm[6][6], b[6];
  int k = 0;
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int q = r; q < 6; q++) { Hm[r][q] = acc[k]; Hm[q][r] = acc[k]; k++; }
#pragma unroll
  for (int r = 0; r < 6; r++) b[r] = acc[21 + r];
  if (!ldlt_solve6_spd(Hm, b, x)) {
    double H[36];
    for (int r = 0; r < 6; r++)
      for (int q = 0; q < 6; q++) H[r * 6 + q] = Hm[r][q];
    ldlt_solve6(H, b, x);
  }
}

struct PoseShared {
  double red[NVP];
  double T[7];        // *se3 of ConvergePose
  double Rt[12];
  double scale;
  int ctrl[4];        // [0] continue, [1] number of selected observations
};

// FeatureAlign::ConvergePose (feature_align.cc:341-421) over the observations whose flag == sel, by the PO_MAIN main
// threads.  The result is left in sh.T; returns false (uniformly) when no observation is selected.
__device__ __forceinline__ bool converge_pose_cta(const PoseProblem& P, int sel, const double* T_init, const DevParams& dp,
                                  PoseShared& sh, double (*part)[PO_MAIN]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  DSE3 T, last_T;
  double chi2 = 0.0;
  if (tid == 0) {
    T = se3_load(T_init);
    last_T = T;
    se3_store(T, sh.T);
    store_Rt(T, sh.Rt);
    sh.ctrl[1] = 0;
    sh.scale = 0.0;
  }
  main_sync();
  int cnt = 0;
  for (int i = tid; i < P.n; i += PO_MAIN) {
    double e = -1.0;
    if (P.o_flag[i] == sel) {
      double x, y, z, ex, ey;
      reproj(sh.Rt, P.o_a[2 * i], P.o_a[2 * i + 1], P.o_pos[3 * i], P.o_pos[3 * i + 1], P.o_pos[3 * i + 2], P.o_scale[i],
             x, y, z, ex, ey);
      e = sqrt(ex * ex + ey * ey);
      cnt++;
    }
    P.o_err[i] = e;
  }
  if (cnt) atomicAdd(&sh.ctrl[1], cnt);
  main_sync();
  const int m = sh.ctrl[1];
  if (m == 0) return false;
  // GetMedianVector (utils.cc:215-220): nth_element at size/2 == the value of rank size/2
  const int mid = m / 2;
  for (int i = tid; i < P.n; i += PO_MAIN) {
    const double e = P.o_err[i];
    if (!(e >= 0.0)) continue;
    int rank = 0;
    for (int j = 0; j < P.n; j++) {
      const double ej = P.o_err[j];
      rank += (ej >= 0.0 && (ej < e || (ej == e && j < i))) ? 1 : 0;
    }
    if (rank == mid) sh.scale = KMADNorm * e;
  }
  main_sync();

  const int max_its = dp.p.max_optim_pose_its;
  for (int it = 0; it < max_its; it++) {
    const double scale = it >= 5 ? 0.85 / dp.cam.fx : sh.scale;   // "force estimator after 5th iteration"
    const double inv_scale2 = 1.0 / (scale * scale);
    double acc[NVP];
#pragma unroll
    for (int k = 0; k < NVP; k++) acc[k] = 0.0;
    for (int i = tid; i < P.n; i += PO_MAIN)
      if (P.o_flag[i] == sel) accumulate_obs(sh.Rt, P, i, inv_scale2, acc);
#pragma unroll
    for (int k = 0; k < NVP; k++) part[k][tid] = acc[k];
    main_sync();
    for (int v = warp; v < NVP; v += PO_MAIN / 32) {   // fixed-order tree
      double s = 0;
#pragma unroll
      for (int k = 0; k < PO_MAIN / 32; k++) s += part[v][lane + 32 * k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) sh.red[v] = s;
    }
    main_sync();
    if (tid == 0) {
      double dT[6], a[NVP];
#pragma unroll
      for (int k = 0; k < NVP; k++) a[k] = sh.red[k];
      solve6(a, dT);
      const double new_chi2 = a[27];
      int cont = 1;
      if ((it > 0 && new_chi2 > chi2) || isnan(dT[0])) {
        T = last_T;   // roll-back
        cont = 0;
      } else {
        const DSE3 T_new = se3_mul(se3_exp(dT), T);
        last_T = T;
        T = T_new;
        chi2 = new_chi2;
        double amax = -1;
#pragma unroll
        for (int r = 0; r < 6; r++) amax = fmax(amax, fabs(dT[r]));
        if (amax <= 1e-10) cont = 0;
      }
      se3_store(T, sh.T);
      store_Rt(T, sh.Rt);
      sh.ctrl[0] = cont;
    }
    main_sync();
    if (!sh.ctrl[0]) break;
  }
  return true;
}

// ConvergePose for one RANSAC hypothesis (np <= 8 observations (base + k) % size), entirely by one thread.
// NP > 0: np == NP known at compile time, so the per-observation chains (two divisions, a square root, the Tukey
// weight) of one Gauss-Newton iteration are interleaved by the scheduler instead of running back to back.
template <int NP>
__device__ __forceinline__ bool converge_pose_small(const PoseProblem& P, int base, int np_rt, int size, const DSE3& T0,
                                                    const DevParams& dp, DSE3* out) {
  const int np = NP > 0 ? NP : np_rt;
  DSE3 T = T0, last_T = T0;
  double chi2 = 0.0;
  double Rt[12];
  store_Rt(T, Rt);
  if (np <= 0) return false;
  double e[RANSAC_MAX_PTS];
#pragma unroll
  for (int k = 0; k < RANSAC_MAX_PTS; k++) {
    e[k] = -1.0;
    if (k < np) {
      const int i = (base + k) % size;
      double x, y, z, ex, ey;
      reproj(Rt, P.o_a[2 * i], P.o_a[2 * i + 1], P.o_pos[3 * i], P.o_pos[3 * i + 1], P.o_pos[3 * i + 2], P.o_scale[i],
             x, y, z, ex, ey);
      e[k] = sqrt(ex * ex + ey * ey);
    }
  }
  double med = 0.0;
  const int mid = np / 2;
#pragma unroll
  for (int k = 0; k < RANSAC_MAX_PTS; k++) {
    int rank = 0;
#pragma unroll
    for (int j = 0; j < RANSAC_MAX_PTS; j++) rank += (j < np && (e[j] < e[k] || (e[j] == e[k] && j < k))) ? 1 : 0;
    if (k < np && rank == mid) med = e[k];
  }
  const double scale0 = KMADNorm * med;
  const int max_its = dp.p.max_optim_pose_its;
  for (int it = 0; it < max_its; it++) {
    const double scale = it >= 5 ? 0.85 / dp.cam.fx : scale0;
    const double inv_scale2 = 1.0 / (scale * scale);
    double acc[NVP];
#pragma unroll
    for (int k = 0; k < NVP; k++) acc[k] = 0.0;
    if (NP > 0) {
#pragma unroll
      for (int k = 0; k < (NP > 0 ? NP : 1); k++) accumulate_obs(Rt, P, (base + k) % size, inv_scale2, acc);
    } else {
      for (int k = 0; k < np; k++) accumulate_obs(Rt, P, (base + k) % size, inv_scale2, acc);
    }
    double dT[6];
    solve6(acc, dT);
    const double new_chi2 = acc[27];
    if ((it > 0 && new_chi2 > chi2) || isnan(dT[0])) {
      T = last_T;
      break;
    }
    const DSE3 T_new = se3_mul(se3_exp(dT), T);
    last_T = T;
    T = T_new;
    chi2 = new_chi2;
    store_Rt(T, Rt);
    double amax = -1;
#pragma unroll
    for (int r = 0; r < 6; r++) amax = fmax(amax, fabs(dT[r]));
    if (amax <= 1e-10) break;
  }
  *out = T;
  return true;
}

struct RansacShared {
  sdvlb_rand backup;         // the stream before the speculative draws
  int best, it, nits, best_supporters, more;
};
// The per-hypothesis arrays, carved from dynamic shared memory.  Every thread derives the pointers itself (registers):
// read back from a structure in memory they would be opaque to the compiler and every access through them a generic
// LD / ST instead of LDS / STS (the same holds for the observation arrays and the per-cell arrays below).
struct RansacArrays {
  double (*hRt)[12];         // [R] pose of every hypothesis as rotation + translation
  int* hsup;                 // [R] supporters
  int* hconv;                // [R] ConvergePose returned true
  int* rnd;                  // [R] speculative draws
};
__host__ __device__ inline size_t ransac_bytes(int R) { return size_t(R) * (12 * sizeof(double) + 3 * sizeof(int)); }
__device__ __forceinline__ RansacArrays ransac_carve(unsigned char* mem, int R) {   // mem 8-byte aligned
  RansacArrays ra;
  ra.hRt = reinterpret_cast<double (*)[12]>(mem);
  ra.hsup = reinterpret_cast<int*>(mem + size_t(R) * 12 * sizeof(double));
  ra.hconv = ra.hsup + R;
  ra.rnd = ra.hconv + R;
  return ra;
}

// FeatureAlign::SelectInliers (feature_align.cc:152-216) over P (= fs_found): flags every observation INLIER/OUTLIER.
// T_frame: frame->GetPose().  *rng advances by the reference's number of rand() calls.  Main threads only.
//
// Hypotheses h0..h1 of a batch are spread over the four warps (hypothesis h0 + k on warp k % 4): a warp waits for its
// slowest Gauss-Newton loop, so the first, small batch -- the one that usually settles the loop -- has two per warp.
__device__ __forceinline__ void select_inliers_cta(const PoseProblem& P, const double* T_frame, const DevParams& dp,
                                                   sdvlb_rand* rng, RansacShared& rs, const RansacArrays ra,
                                                   long long* stamps = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int size = P.n;
  if (size == 0) return;
  const int R = min(dp.p.max_ransac_its, HYP_MAX);
  const int np = min(min(dp.p.max_ransac_points, size), RANSAC_MAX_PTS);
  const double thr = dp.p.inlier_error_threshold / dp.cam.fx;
  if (tid < 34) rs.backup.r[tid] = rng->r[tid];
  if (tid == 0) {
    rs.backup.n = rng->n;
    rs.best = -1; rs.it = 0; rs.nits = dp.p.max_ransac_its; rs.best_supporters = 0; rs.more = 1;
  }
  main_sync();
  int h0 = 0;
  while (h0 < R) {
    const int h1 = min(R, h0 == 0 ? 8 : (h0 == 8 ? 32 : R));
    if (tid == 0)   // speculative draws of the batch; the stream is rewound to the reference's count below
      for (int h = h0; h < h1; h++) ra.rnd[h] = rand_next(rng);
    main_sync();
    // consecutive hypotheses on different warps -- of the first four only, one per SM sub-partition: a hypothesis is a
    // serial fp64 chain, and two warps of them on one scheduler take turns (14.3 -> 17.7 us with six warps)
    for (int k = lane * 4 + warp; warp < 4 && h0 + k < h1; k += 128) {
      const int h = h0 + k;
      {
        DSE3 T;
        const bool ok = np == 5 ? converge_pose_small<5>(P, ra.rnd[h] % size, np, size, se3_load(T_frame), dp, &T)
                                : converge_pose_small<0>(P, ra.rnd[h] % size, np, size, se3_load(T_frame), dp, &T);
        ra.hconv[h] = ok ? 1 : 0;
        ra.hsup[h] = 0;
        double Rt[12];
        store_Rt(T, Rt);
#pragma unroll
        for (int q = 0; q < 12; q++) ra.hRt[h][q] = Rt[q];
      }
    }
    main_sync();
    if (stamps && h0 == 0) stamps[0] = clock64();
    // CheckReprojectionError of every hypothesis of the batch against every match (feature_align.cc:258-283): thread t
    // tests matches t, t + 128, ... against all of them; counts by ballot, one shared-memory add per warp and hypothesis
    for (int i0 = 0; i0 < size; i0 += PO_MAIN) {
      const int i = i0 + tid;
      for (int h = h0; h < h1; h++) {
        const bool in = i < size && within_threshold(ra.hRt[h], P, i, thr);
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0 && bal) atomicAdd(&ra.hsup[h], __popc(bal));   // integer adds: order does not matter
      }
    }
    main_sync();
    if (tid == 0) {   // the reference's loop, replayed over the hypotheses computed so far
      const double sprob = 0.99;
      int nits = rs.nits, it = rs.it, best_supporters = rs.best_supporters, best = rs.best;
      while (it < nits && it < h1) {
        if (ra.hconv[it] && ra.hsup[it] > best_supporters) {
          best = it;
          best_supporters = ra.hsup[it];
          const double epsilon = 1.0 - (double(best_supporters) / double(size));
          double tmp = 1.0 - epsilon;
          for (int k = 1; k < np; k++) tmp *= tmp;
          if (tmp < 1e-5) nits = dp.p.max_ransac_its;
          else nits = min(dp.p.max_ransac_its, int(log(1.0 - sprob) / log(1.0 - tmp)));
        }
        it++;
      }
      rs.nits = nits; rs.it = it; rs.best_supporters = best_supporters; rs.best = best;
      rs.more = (it < nits && h1 < R) ? 1 : 0;
    }
    main_sync();
    h0 = h1;
    if (!rs.more) break;
  }
  if (stamps) stamps[1] = clock64();
  if (tid == 0) {   // rewind: the reference drew rs.it numbers only
    for (int k = 0; k < 34; k++) rng->r[k] = rs.backup.r[k];
    rng->n = rs.backup.n;
    for (int k = 0; k < rs.it; k++) rand_next(rng);
  }
  // "Get low innovation inliers" with best_se3 (identity when no hypothesis ever had a supporter)
  double Rt[12];
  if (rs.best >= 0) {
#pragma unroll
    for (int k = 0; k < 12; k++) Rt[k] = ra.hRt[rs.best][k];
  } else {
#pragma unroll
    for (int k = 0; k < 12; k++) Rt[k] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0;
  }
  for (int i = tid; i < size; i += PO_MAIN) {
    double x, y, z, ex, ey;
    reproj(Rt, P.o_a[2 * i], P.o_a[2 * i + 1], P.o_pos[3 * i], P.o_pos[3 * i + 1], P.o_pos[3 * i + 2], P.o_scale[i], x, y,
           z, ex, ey);
    P.o_flag[i] = sqrt(ex * ex + ey * ey) <= thr ? SDVLB_OBS_INLIER : SDVLB_OBS_OUTLIER;
  }
  main_sync();
}

// Re-partition after a pose change: observations flagged `from` whose error at sh.Rt exceeds / is within thr.
// Returns (uniformly) how many changed list.
__device__ __forceinline__ int recheck_cta(const PoseProblem& P, const double* Rt, int from, bool move_if_within, double thr, int to,
                           PoseShared& sh) {
  const int tid = threadIdx.x;
  if (tid == 0) sh.ctrl[2] = 0;
  main_sync();
  int moved = 0;
  for (int i = tid; i < P.n; i += PO_MAIN) {
    if (P.o_flag[i] != from) continue;
    double x, y, z, ex, ey;
    reproj(Rt, P.o_a[2 * i], P.o_a[2 * i + 1], P.o_pos[3 * i], P.o_pos[3 * i + 1], P.o_pos[3 * i + 2], P.o_scale[i], x, y,
           z, ex, ey);
    const bool within = sqrt(ex * ex + ey * ey) <= thr;
    if (within == move_if_within) { P.o_flag[i] = to; moved++; }
  }
  if (moved) atomicAdd(&sh.ctrl[2], moved);
  main_sync();
  return sh.ctrl[2];
}

// FeatureAlign::OptimizePose(frame) without RemoveOutliers (feature_align.cc:73-82,218-243).  T_frame (shared, 7
// doubles) holds frame->GetPose() on entry and the refined pose on return.
__device__ __forceinline__ void optimize_pose_cta(const PoseProblem& P, double* T_frame, const DevParams& dp, PoseShared& sh,
                                  double (*part)[PO_MAIN]) {
  const int tid = threadIdx.x;
  const double thr = dp.p.inlier_error_threshold / dp.cam.fx;
  for (int round = 0; round < 2; round++) {
    if (converge_pose_cta(P, SDVLB_OBS_INLIER, T_frame, dp, sh, part)) {
      if (tid < 7) T_frame[tid] = sh.T[tid];   // frame->SetPose(se3)
      main_sync();
      recheck_cta(P, sh.Rt, SDVLB_OBS_INLIER, false, thr, SDVLB_OBS_OUTLIER, sh);
    }
    if (round == 1) break;
    // RescueOutliers at the frame pose
    if (tid == 0) store_Rt(se3_load(T_frame), sh.Rt);
    main_sync();
    if (recheck_cta(P, sh.Rt, SDVLB_OBS_OUTLIER, true, 2 * thr, SDVLB_OBS_INLIER, sh) == 0) break;
  }
}

// ------------------------------------------------------------------------------------------------ commands
// One CTA per sequence that has commands (sequences that are not part of the step being submitted).
__global__ void __launch_bounds__(128) seq_apply_kernel(const SeqCmd* __restrict__ cmds, const int2* __restrict__ ranges,
                                                        const __grid_constant__ DevParams dp) {
  __shared__ int s_base;
  seq_apply_commands(cmds, ranges[blockIdx.x], dp, &s_base);
}

// ------------------------------------------------------------------------------------------------ post
// Completion of a submission without a separate 1-thread kernel: every CTA checks in after its results are fenced to
// the host; the last one publishes the submission's sequence number in the pinned completion word.
__device__ __forceinline__ void signal_done(const SeqStepArgs& A) {
  if (!A.h_flag) return;
  // release at GPU scope; the last CTA's system fence below is cumulative over everything that happened before the
  // arrivals it has observed, i.e. over the result writes of every CTA of the submission
  __threadfence();
  const unsigned arrived = atomicAdd(A.d_done, 1u) + 1u;
  if (arrived == unsigned(A.n)) {
    *A.d_done = 0u;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(A.h_flag) = A.seq_no;
  }
}

// The observation arrays of the pose refinement (a, pos, scale, err: 7 doubles, and a flag per observation) live in
// shared memory when the problem is small enough (always, for FeatureAlign's max_matches <= 256): every Gauss-Newton
// iteration of every RANSAC hypothesis re-reads them.
constexpr int OBS_SMEM_MAX = 256;
__host__ __device__ inline size_t obs_smem_bytes(int cap) { return size_t(cap) * (7 * sizeof(double) + sizeof(int32_t)); }
__device__ __forceinline__ void obs_carve(unsigned char* mem, int cap, PoseProblem& P) {   // mem 8-byte aligned
  double* d = reinterpret_cast<double*>(mem);
  P.o_a = d; P.o_pos = d + 2 * size_t(cap); P.o_scale = d + 5 * size_t(cap); P.o_err = d + 6 * size_t(cap);
  P.o_flag = reinterpret_cast<int32_t*>(d + 7 * size_t(cap));
}

struct PostShared {
  PoseShared ps;
  RansacShared rs;
  double T_frame[7];
  sdvlb_rand rng;
  int scan[PO_MAIN];
  int attempts, n_found, n_inl, n_outl;
  int adopt;
  int kf_live[SDVLB_SEQ_KF_CAP];
};

// kObsSmem: max_matches <= OBS_SMEM_MAX, the observation arrays are carved from shared memory (a compile-time fact, so
// that the accesses are LDS / STS; chosen at run time the pointers would be generic).
template <bool kObsSmem>
__global__ void __launch_bounds__(PO_THREADS, 1) seq_post_kernel(const __grid_constant__ SeqStepArgs A) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const long long t_entry = clock64();
  PostShared& sh = *reinterpret_cast<PostShared*>(s_raw);
  double (*part)[PO_MAIN] = reinterpret_cast<double (*)[PO_MAIN]>(s_raw + ((sizeof(PostShared) + 15) / 16) * 16);
  SeqState* S = A.seq[blockIdx.x];
  const int tid = threadIdx.x;
  const int n_cells = A.g.wcells[0] * A.g.hcells[0];
  const int gw = A.g.wcells[0];
  sdvlb_grid_dependency_wait();   // programmatic dependent launch: everything above overlaps the search kernel's tail
  SeqResultHost* Rz = S->result[A.slot];
  if (!S->has_last || S->hold) {   // uniform: no track yet (no reset) / waiting for the caller: the frame is not consumed
    if (tid == 0) {
      Rz->status = S->has_last ? SDVLB_SEQ_HELD : SDVLB_SEQ_IDLE;
      Rz->error = S->overflow;
      Rz->need_keyframe = 0;
      Rz->lost_frames = S->lost_frames;
      signal_done(A);
    }
    return;
  }
  // small shared memory on purpose: this CTA must fit beside the build stream's kernels.  Every thread derives the
  // array pointers itself (see RansacArrays)
  unsigned char* const mem0 = reinterpret_cast<unsigned char*>(&part[0][0]) + sizeof(double) * NVP * PO_MAIN;
  const RansacArrays ra = ransac_carve(mem0, A.dp.p.max_ransac_its);
  int* const p_win = reinterpret_cast<int*>(mem0 + (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16);
  int* const p_slot = p_win + n_cells;     // [n_cells] index of its match in fs_found, -1: not visited
  int* const p_order = p_slot + n_cells;   // [n_cells] cell_order_
  int* const p_head = p_order + n_cells;   // [n_cells] first candidate of the cell's chain, -1: empty
                                           // p_win: [n_cells] per cell: rank of its match (INT_MAX: none)

  // ---- load the persistent FeatureAlign state (all PO_THREADS threads)
  for (int i = tid; i < n_cells; i += PO_THREADS) {
    p_order[i] = S->cell_order[i]; p_win[i] = INT_MAX; p_slot[i] = -1; p_head[i] = -1;
  }
  if (tid < 34) sh.rng.r[tid] = S->rng.r[tid];
  if (tid == 0) {
    sh.rng.n = S->rng.n;
    sh.attempts = 0; sh.n_found = 0; sh.n_inl = 0; sh.n_outl = 0;
    for (int i = 0; i < 7; i++) sh.T_frame[i] = A.cur[blockIdx.x].pose[i];   // ImageAlign's result
  }
  if (tid < SDVLB_SEQ_KF_CAP) sh.kf_live[tid] = 0;
  __syncthreads();

  if (tid >= PO_MAIN) {
    // ---- shuffle warp: std::random_shuffle(cell_order_) for the NEXT frame (feature_align.cc:103), once SelectInliers
    // has taken its draws.  Waits on barrier 2 (all PO_THREADS threads).
    asm volatile("bar.sync 2, %0;" ::"n"(PO_THREADS) : "memory");
    const int lane = tid - PO_MAIN;
    if (n_cells >= 64) {
      // The n_cells - 1 draws in one go.  rand() is r[n] = r[n-3] + r[n-31] (mod 2^32): the 30 values of a round only
      // need values of earlier rounds through the lag-31 term, so a round is three interleaved prefix sums (stride 3)
      // of known terms -- four shuffle steps instead of 30 dependent updates.  seq[] = the last 34 values in
      // chronological order, then the new ones; it lives in win/slot, idle since SelectPoints.
      uint32_t* seq = reinterpret_cast<uint32_t*>(p_win);
      const int nd = n_cells - 1, n0 = sh.rng.n;
      for (int k = lane; k < 34; k += 32) seq[k] = sh.rng.r[(n0 - 34 + k) % 34];
      __syncwarp();
      for (int m0 = 0; m0 < nd; m0 += 30) {
        uint32_t t = lane < 30 ? seq[34 + m0 + lane - 31] : 0u;
#pragma unroll
        for (int o = 3; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, t, o);
          if (lane >= o) t += v;
        }
        if (lane < 30) seq[34 + m0 + lane] = t + seq[34 + m0 + (lane % 3) - 3];
        __syncwarp();
      }
      // state after exactly nd draws, then value -> swap partner j = rand() % (i + 1)
      const int n1 = n0 + nd;
      for (int k = lane; k < 34; k += 32) S->rng.r[(n1 - 34 + k) % 34] = seq[nd + k];
      if (lane == 0) S->rng.n = n1 >= 34 * 1000000 ? n1 - 34 * 999999 : n1;
      __syncwarp();
      for (int i = 1 + lane; i < n_cells; i += 32) seq[34 + i - 1] = (seq[34 + i - 1] >> 1) % uint32_t(i + 1);
      __syncwarp();
      if (lane == 0) {
#pragma unroll 4
        for (int i = 1; i < n_cells; ++i) {
          const int j = int(seq[34 + i - 1]);
          const int a = p_order[i], b = p_order[j];
          p_order[i] = b; p_order[j] = a;
        }
      }
      __syncwarp();
      for (int i = lane; i < n_cells; i += 32) S->cell_order[i] = p_order[i];
      if (lane == 0) Rz->phase_cycles[7] = int(clock64() - t_entry);
      return;
    }
    if (tid == PO_MAIN) {
      for (int i = 1; i < n_cells; ++i) {
        const int j = rand_next(&sh.rng) % (i + 1);
        const int a = p_order[i], b = p_order[j];
        p_order[i] = b; p_order[j] = a;
      }
    }
    __syncwarp();
    for (int i = lane; i < n_cells; i += 32) S->cell_order[i] = p_order[i];
    for (int i = lane; i < 34; i += 32) S->rng.r[i] = sh.rng.r[i];
    if (lane == 0) { S->rng.n = sh.rng.n; Rz->phase_cycles[7] = int(clock64() - t_entry); }
    return;
  }

  // =============================== main threads (barrier 1) ===============================
  long long t_phase[8];
  t_phase[0] = clock64();
  SeqFeat* __restrict__ L = S->list[S->cur];
  SeqFeat* __restrict__ NL = S->list[S->cur ^ 1];
  // Candidates = the features of the last frame that observe a point, in list order (FeatureAlign::ProjectPoints,
  // feature_align.cc:296-321); a candidate's index is its feature index.
  const int nc = S->n_list;
  const sdvlb_match* __restrict__ M = S->matches;
  // cell / score / chain link of every candidate: in shared memory (the GN partials area is idle here) unless there
  // are too many
  constexpr int kSmemCands = NVP * PO_MAIN * 2 / 3;   // three int arrays in NVP * PO_MAIN doubles
  const bool in_smem = nc <= kSmemCands;
  int32_t* __restrict__ c_cell = in_smem ? reinterpret_cast<int32_t*>(&part[0][0]) : S->c_cell;
  int32_t* __restrict__ c_score = in_smem ? reinterpret_cast<int32_t*>(&part[0][0]) + kSmemCands : S->c_score;
  int32_t* __restrict__ c_next = in_smem ? reinterpret_cast<int32_t*>(&part[0][0]) + 2 * kSmemCands : S->c_next;
  int32_t* __restrict__ c_rank = S->c_rank;

  // ---- ProjectPoint bookkeeping (feature_align.cc:323-339): cell of every seen point, chained per cell
  for (int i = tid; i < nc; i += PO_MAIN) {
    const SeqFeat& ft = L[i];
    int cell = -1;
    if (ft.flags & SEQF_HAS_POINT) {
      const sdvlb_match m = M[i];
      if (m.status != SDVLB_MATCH_UNSEEN) cell = int(m.proj[1] / SDVLB_CELL) * gw + int(m.proj[0] / SDVLB_CELL);
      if (cell >= n_cells) cell = -1;
    }
    c_cell[i] = cell;
    c_score[i] = ft.n_successful;
    if (cell >= 0) c_next[i] = atomicExch(&p_head[cell], i);   // chain order is arbitrary; ranks do not depend on it
  }
  main_sync();
  // ---- rank inside the cell: Score() descending, stable (std::list::sort, feature_align.cc:111)
  for (int i = tid; i < nc; i += PO_MAIN) {
    const int cell = c_cell[i];
    if (cell < 0) { c_rank[i] = -1; continue; }
    const int sc = c_score[i];
    int rank = 0;
    for (int j = p_head[cell]; j >= 0; j = c_next[j]) {
      const int sj = c_score[j];
      rank += (sj > sc || (sj == sc && j < i)) ? 1 : 0;
    }
    c_rank[i] = rank;
    if (M[i].status == SDVLB_MATCH_FOUND) atomicMin(&p_win[cell], rank);
  }
  main_sync();
  t_phase[1] = clock64();
  // ---- cells in cell_order_: exclusive prefix sum of "has a match" -> max_matches cut-off and fs_found index
  {
    const int per = (n_cells + PO_MAIN - 1) / PO_MAIN;
    const int q0 = tid * per, q1 = min(n_cells, q0 + per);
    int local = 0;
    for (int q = q0; q < q1; q++) local += p_win[p_order[q]] != INT_MAX ? 1 : 0;
    // exclusive scan over the PO_MAIN per-thread counts: shuffles inside a warp, then the warp totals
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) sh.scan[tid >> 5] = incl;
    main_sync();
    int pre = incl - local;
    int total = 0;
#pragma unroll
    for (int w = 0; w < PO_MAIN / 32; w++) {
      const int v = sh.scan[w];
      if (w < (tid >> 5)) pre += v;
      total += v;
    }
    if (tid == 0) sh.n_found = min(total, A.dp.p.max_matches);
    main_sync();
    const int max_matches = A.dp.p.max_matches;
    for (int q = q0; q < q1; q++) {
      const int c = p_order[q];
      const bool has = p_win[c] != INT_MAX;
      // the loop `for (i < size && matches_ < max_matches_)` visits this cell iff fewer than max_matches so far
      if (pre < max_matches) p_slot[c] = has ? pre : -2;   // -2: visited, no match
      pre += has ? 1 : 0;
    }
  }
  main_sync();
  const int n_found = sh.n_found;
  PoseProblem P;
  const int obs_cap = min(OBS_SMEM_MAX, (A.dp.p.max_matches + 7) & ~7);
  if constexpr (kObsSmem) {
    obs_carve(mem0 + (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16 + ((size_t(4 * n_cells) * sizeof(int) + 15) / 16) * 16,
              obs_cap, P);
  } else {
    P.o_a = S->o_a; P.o_pos = S->o_pos; P.o_scale = S->o_scale; P.o_err = S->o_err; P.o_flag = S->o_flag;
  }
  P.n = n_found;
  // ---- SelectPoints side effects (feature_align.cc:112-147): attempts, Promote / Unpromote, new features
  {
    int my_attempts = 0;
    for (int i = tid; i < nc; i += PO_MAIN) {
      const int cell = c_cell[i];
      if (cell < 0) continue;
      const int slot = p_slot[cell];
      if (slot == -1) continue;                       // cell never visited
      const int rank = c_rank[i], win = p_win[cell];
      if (rank > win) continue;                       // the cell was left at its first match
      my_attempts++;
      if (rank == win && slot >= 0) {                 // found: Promote, new Feature(frame, px, level)
        const sdvlb_match m = M[i];
        SeqFeat f = L[i];
        f.px[0] = m.px[0]; f.px[1] = m.px[1];
        cam_unproject_unit(A.dp.cam, m.px[0], m.px[1], f.v);
        f.level = m.level;
        f.n_successful += 1; f.n_failed = 0;          // Point::Promote (point.cc:102-106)
        f.status = SEQP_FOUND;
        NL[slot] = f;
        if (S->fdesc[0]) {   // ORB mode: the point keeps its init feature, hence that feature's descriptor
          const uint4* src = reinterpret_cast<const uint4*>(S->fdesc[S->cur] + size_t(i) * 8);
          uint4* dst = reinterpret_cast<uint4*>(S->fdesc[S->cur ^ 1] + size_t(slot) * 8);
          dst[0] = src[0]; dst[1] = src[1];
        }
        P.o_a[2 * slot] = f.v[0] / f.v[2];            // Camera::SimpleProject(feature->GetVector())
        P.o_a[2 * slot + 1] = f.v[1] / f.v[2];
        P.o_pos[3 * slot] = f.pos[0]; P.o_pos[3 * slot + 1] = f.pos[1]; P.o_pos[3 * slot + 2] = f.pos[2];
        P.o_scale[slot] = 1.0 / double(1 << m.level);
        P.o_flag[slot] = 0;
      }
      // not found: Point::Unpromote / Map::DeletePoint (point.cc:108-115); the point leaves the track either way,
      // because the next frame's candidates are this frame's features
    }
    if (my_attempts) atomicAdd(&sh.attempts, my_attempts);
  }
  main_sync();

  // ---- SDVL::CalcTrackingQuality (sdvl.cc:240-264), when the policy asks for it
  const sdvlb_seq_policy pol = S->policy;
  int quality = SDVLB_TRACKING_GOOD;
  int lost_frames = S->lost_frames;
  if (pol.tracking_quality) {
    const int attempts = sh.attempts;
    const double ratio = attempts == 0 ? 0.0 : double(n_found) / double(attempts);
    if (ratio > 0.2) { quality = SDVLB_TRACKING_GOOD; lost_frames = 0; }
    else if (n_found < A.dp.p.min_matches) { quality = SDVLB_TRACKING_BAD; lost_frames += 1; }
    else { quality = SDVLB_TRACKING_INSUFFICIENT; lost_frames = 0; }
  }
  const bool adopt = quality != SDVLB_TRACKING_BAD;   // sdvl.cc:99,119: last_frame_ advances unless tracking is bad
  if (!adopt) {
    // The frame is dropped but the points were visited: Promote / Unpromote stay with them (they live in the list of
    // the frame that remains the reference); a point that failed too often is deleted (point.cc:102-115, Map::DeletePoint)
    for (int i = tid; i < nc; i += PO_MAIN) {
      const int cell = c_cell[i];
      if (cell < 0) continue;
      const int slot = p_slot[cell];
      if (slot == -1) continue;
      const int rank = c_rank[i], win = p_win[cell];
      if (rank > win) continue;
      SeqFeat& f = L[i];
      if (rank == win && slot >= 0) { f.n_successful += 1; f.n_failed = 0; }
      else {
        f.n_failed += 1; f.n_unpromoted += 1;
        if (f.n_failed > A.dp.p.max_failed) f.flags &= ~SEQF_HAS_POINT;
      }
    }
  }

  // ---- SelectInliers (RANSAC)
  t_phase[2] = clock64();
  t_phase[6] = t_phase[7] = t_phase[2];
  select_inliers_cta(P, sh.T_frame, A.dp, &sh.rng, sh.rs, ra, t_phase + 6);
  t_phase[3] = clock64();
  asm volatile("bar.sync 2, %0;" ::"n"(PO_THREADS) : "memory");   // releases the shuffle warp: rand() is its from here on

  // ---- OptimizePose + RescueOutliers + OptimizePose (feature_align.cc:73-82)
  optimize_pose_cta(P, sh.T_frame, A.dp, sh.ps, part);
  main_sync();
  t_phase[4] = clock64();

  // ---- RemoveOutliers (feature_align.cc:245-256), result, GetMotionModel, state update
  sdvlb_launch_dependents();   // a queued next step may bring its align CTAs in; they wait for this grid to finish
  sdvlb_seq_feat* hfe = reinterpret_cast<sdvlb_seq_feat*>(reinterpret_cast<unsigned char*>(Rz) + sizeof(SeqResultHost));
  {
    int inl = 0, outl = 0;
    for (int i = tid; i < n_found; i += PO_MAIN) {
      const bool is_in = P.o_flag[i] == SDVLB_OBS_INLIER;
      SeqFeat& f = NL[i];
      if (is_in) { inl++; atomicAdd(&sh.kf_live[f.kf], 1); }
      else { outl++; f.flags &= ~SEQF_HAS_POINT; f.status = SEQP_NOT_FOUND; }
      sdvlb_seq_feat o;
      o.px[0] = f.px[0]; o.px[1] = f.px[1];
      o.user_id = f.user_id;
      o.level = f.level;
      o.flags = is_in ? SDVLB_FEAT_HAS_POINT : 0;
      hfe[i] = o;
    }
    if (inl) atomicAdd(&sh.n_inl, inl);
    if (outl) atomicAdd(&sh.n_outl, outl);
  }
  main_sync();
  if (!adopt) {   // the reference frame stays: its keyframes are the live ones
    if (tid < SDVLB_SEQ_KF_CAP) sh.kf_live[tid] = 0;
    main_sync();
    for (int i = tid; i < nc; i += PO_MAIN)
      if (L[i].flags & SEQF_HAS_POINT) atomicAdd(&sh.kf_live[L[i].kf], 1);
    main_sync();
  }
  if (tid < SDVLB_SEQ_KF_CAP) Rz->kf_live[tid] = sh.kf_live[tid];
  if (tid == 0) {
    const DSE3 T_new = se3_load(sh.T_frame);
    const DSE3 mov = se3_mul(T_new, se3_inverse(se3_load(S->T_last)));   // SDVL::GetMotionModel, sdvl.cc:266-276
    double vel[6];
    se3_log(mov, vel);
    for (int i = 0; i < 6; i++) S->vel[i] = 0.9 * (0.5 * vel[i] + 0.5 * S->vel[i]);
    for (int i = 0; i < 7; i++) { Rz->pose[i] = sh.T_frame[i]; A.cur[blockIdx.x].pose[i] = sh.T_frame[i]; }
    int need_kf = 0;
    if (adopt) {
      for (int i = 0; i < 7; i++) S->T_last[i] = sh.T_frame[i];
      S->last = A.cur[blockIdx.x];
      S->n_list = n_found;
      S->cur ^= 1;
      S->frame_id += 1;
      // Map::NeedKeyframe (map.cc:170-188), asked only when tracking is good (sdvl.cc:100)
      if (pol.keyframe_rule && quality == SDVLB_TRACKING_GOOD) {
        const int npoints = sh.n_inl;   // frame->GetNumPoints()
        const bool enough_its = (S->frame_id - S->last_kf_frame) >= pol.min_keyframe_its;
        const bool lost_many = double(npoints) < double(S->last_matches) * pol.lost_ratio;
        const bool lost_some = double(npoints) < double(S->last_matches) * 0.9;
        S->last_matches = max(S->last_matches, npoints);
        if ((enough_its && lost_some) || lost_many) { S->last_matches = npoints; need_kf = 1; }
      }
    }
    S->lost_frames = lost_frames;
    if (need_kf || (pol.tracking_quality && lost_frames >= 3)) S->hold = 1;   // sdvl.cc:74-91: relocalisation is the caller's
    Rz->stats[0] = S->align_info[0];
    Rz->stats[1] = n_found;
    Rz->stats[2] = sh.attempts;
    Rz->stats[3] = sh.n_inl;
    Rz->stats[4] = sh.n_outl;
    Rz->stats[5] = sh.n_inl;        // Frame::GetNumPoints(): features that still have a point
    Rz->stats[6] = S->align_info[1];
    Rz->stats[7] = n_found;
    Rz->error = S->overflow | (nc > A.max_feats ? 2 : 0);   // 2: more features than the submission was sized for
    Rz->status = SDVLB_SEQ_TRACKED;
    Rz->quality = quality;
    Rz->need_keyframe = need_kf;
    Rz->lost_frames = lost_frames;
    // latency breakdown of this kernel in SM cycles
    t_phase[5] = clock64();
    Rz->phase_cycles[0] = int(t_phase[1] - t_phase[0]);
    Rz->phase_cycles[1] = int(t_phase[2] - t_phase[1]);
    Rz->phase_cycles[2] = int(t_phase[6] - t_phase[2]);   // RANSAC: draws + hypotheses of the first batch
    Rz->phase_cycles[3] = int(t_phase[0] - t_entry);      // kernel entry -> state loaded
    Rz->phase_cycles[4] = int(t_phase[3] - t_phase[6]);   //         supporters, replay, further batches, inlier flags
    Rz->phase_cycles[5] = int(t_phase[4] - t_phase[3]);
    Rz->phase_cycles[6] = int(t_phase[5] - t_phase[4]);
    for (int i = 0; i < 4; i++) Rz->align_cycles[i] = S->align_cycles[i];
  }
  // Everything the host reads has been written (to pinned memory): publish the submission's completion.  One system
  // fence, by thread 0 AFTER the CTA barrier (signal_done): fences are cumulative, so it orders the writes of all the
  // threads the barrier has synchronised with before the completion word -- a fence per thread before the barrier made
  // every warp wait for a PCIe round trip of its own (8 % of the kernel's stall samples).
  main_sync();
  if (tid == 0) signal_done(A);
}

// ------------------------------------------------------------------------------------------------ standalone calls
struct PoseCallShared {
  PoseShared ps;
  RansacShared rs;
  double T_frame[7];
  sdvlb_rand rng;
};

__global__ void __launch_bounds__(PO_MAIN, 1) pose_call_kernel(const __grid_constant__ PoseCallArgs A) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  PoseCallShared& sh = *reinterpret_cast<PoseCallShared*>(s_raw);
  double (*part)[PO_MAIN] = reinterpret_cast<double (*)[PO_MAIN]>(s_raw + ((sizeof(PoseCallShared) + 15) / 16) * 16);
  const int tid = threadIdx.x;
  const RansacArrays ra =
      ransac_carve(reinterpret_cast<unsigned char*>(&part[0][0]) + sizeof(double) * NVP * PO_MAIN, A.dp.p.max_ransac_its);
  PoseProblem P;
  P.n = A.n;
  if (A.n <= OBS_SMEM_MAX) {
    obs_carve(reinterpret_cast<unsigned char*>(&part[0][0]) + sizeof(double) * NVP * PO_MAIN +
                  (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16,
              (A.n + 7) & ~7, P);
  } else {
    P.o_a = A.scratch; P.o_pos = A.scratch + 2 * size_t(A.n); P.o_scale = A.scratch + 5 * size_t(A.n);
    P.o_err = A.scratch + 6 * size_t(A.n);
    P.o_flag = A.iscratch;
  }
  for (int i = tid; i < A.n; i += PO_MAIN) {
    const sdvlb_pose_obs o = A.obs[i];
    P.o_a[2 * i] = o.v[0] / o.v[2]; P.o_a[2 * i + 1] = o.v[1] / o.v[2];
    P.o_pos[3 * i] = o.pos[0]; P.o_pos[3 * i + 1] = o.pos[1]; P.o_pos[3 * i + 2] = o.pos[2];
    P.o_scale[i] = 1.0 / double(1 << o.level);
    P.o_flag[i] = o.flags;
  }
  if (tid < 7) sh.T_frame[tid] = A.T[tid];
  if (A.mode == 0) {
    if (tid < 34) sh.rng.r[tid] = A.rng->r[tid];
    if (tid == 0) sh.rng.n = A.rng->n;
  }
  main_sync();
  if (A.mode == 0) {
    select_inliers_cta(P, sh.T_frame, A.dp, &sh.rng, sh.rs, ra);
    if (tid < 34) A.rng->r[tid] = sh.rng.r[tid];
    if (tid == 0) A.rng->n = sh.rng.n;
  } else {
    optimize_pose_cta(P, sh.T_frame, A.dp, sh.ps, part);
    main_sync();
    if (tid < 7) A.T[tid] = sh.T_frame[tid];
  }
  main_sync();
  for (int i = tid; i < A.n; i += PO_MAIN) A.obs[i].flags = P.o_flag[i];
}

}  // namespace

cudaError_t sdvlb_launch_seq_apply(const SeqCmd* d_cmds, const int2* d_ranges, int n_ranges, const DevParams& dp,
                                   cudaStream_t stream) {
  if (n_ranges <= 0) return cudaSuccess;
  SDVLB_PREPARE(seq_apply_kernel, 0);
  seq_apply_kernel<<<n_ranges, 128, 0, stream>>>(d_cmds, d_ranges, dp);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_seq_post(const SeqStepArgs& A, cudaStream_t stream) {
  const int n_cells = A.g.wcells[0] * A.g.hcells[0];
  const int obs_cap = A.dp.p.max_matches <= OBS_SMEM_MAX ? ((A.dp.p.max_matches + 7) & ~7) : 0;
  const size_t dyn = ((sizeof(PostShared) + 15) / 16) * 16 + sizeof(double) * NVP * PO_MAIN +
                     (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16 + ((size_t(4 * n_cells) * sizeof(int) + 15) / 16) * 16 +
                     obs_smem_bytes(obs_cap);
  if (obs_cap > 0) {
    SDVLB_PREPARE(seq_post_kernel<true>, dyn);
    return sdvlb_launch_dependent(seq_post_kernel<true>, dim3(A.n), dim3(PO_THREADS), dyn, stream, A);
  }
  SDVLB_PREPARE(seq_post_kernel<false>, dyn);
  return sdvlb_launch_dependent(seq_post_kernel<false>, dim3(A.n), dim3(PO_THREADS), dyn, stream, A);
}

cudaError_t sdvlb_launch_pose_call(const PoseCallArgs& A, cudaStream_t stream) {
  const size_t dyn = ((sizeof(PoseCallShared) + 15) / 16) * 16 + sizeof(double) * NVP * PO_MAIN +
                     (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16 +
                     (A.n <= OBS_SMEM_MAX ? obs_smem_bytes((A.n + 7) & ~7) : 0);
  SDVLB_PREPARE(pose_call_kernel, dyn);
  pose_call_kernel<<<1, PO_MAIN, dyn, stream>>>(A);
  return cudaGetLastError();
}
