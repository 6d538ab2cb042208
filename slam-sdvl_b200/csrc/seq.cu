// seq.cu — resident sequences: everything SDVL::ProcessFrame does around ImageAlign and SearchPoint, on the device.
//
//   seq_apply_kernel : host commands (restart a track, append the mapping thread's new points to the current frame)
//   seq_prep_kernel  : SetMotionModel (sdvl.cc:278-281), ImageAlign feature marshalling (image_align.cc:154-160,
//                      229-235), FeatureAlign::ProjectPoints candidate list (feature_align.cc:296-321)
//   [image_align_kernel, search_points_kernel: align.cu / search.cu]
//   seq_post_kernel  : FeatureAlign::ProjectPoint binning + SelectPoints (feature_align.cc:88-150,323-339),
//                      SelectInliers (RANSAC, :152-216), OptimizePose / RescueOutliers / RemoveOutliers (:73-82,
//                      :218-256), ConvergePose (:341-421), GetMotionModel (sdvl.cc:266-276); writes the sequence's
//                      next feature list and its pinned host result
//   pose_call_kernel : SelectInliers / OptimizePose alone (sdvlb_select_inliers, sdvlb_optimize_pose)
//
// One CTA per sequence.  SelectPoints is order-exact without walking the reference's per-cell std::list: the reference
// visits the 32-px cells in cell_order_ order, inside a cell the points by descending Point::Score() (stable), stops a
// cell at its first match and everything at max_matches matches.  Here every candidate gets its rank inside its cell
// (score desc, candidate index asc), a cell's match is its lowest-ranked found candidate, a prefix sum over the cells in
// cell_order_ gives both the max_matches cut-off and each match's index in fs_found.  RANSAC evaluates all
// max_ransac_its hypotheses at once (one thread each, 5-point Gauss-Newton in registers), then one thread replays the
// reference's adaptive iteration count over the results, so the same hypothesis wins and rand() advances by exactly the
// number of draws the reference makes.  The next frame's random_shuffle of cell_order_ (the only other rand() consumer)
// runs on a ninth warp while the other eight refine the pose.
#include <climits>

#include "seq.cuh"

namespace {

constexpr int PO_MAIN = 128;      // threads doing the FeatureAlign work (named barrier 1)
constexpr int PO_THREADS = 160;   // + the shuffle warp
constexpr int NVP = 28;           // A (21, upper triangle) + b (6) + chi2
constexpr int RANSAC_MAX_PTS = 8;
constexpr int HYP_MAX = 256;      // max_ransac_its supported

constexpr double KMADNorm = 1.4826;            // feature_align.h
constexpr double KTukeyC = 4.6851 * 4.6851;

__device__ __forceinline__ void main_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

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
__device__ bool converge_pose_cta(const PoseProblem& P, int sel, const double* T_init, const DevParams& dp,
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
  double (*hRt)[12];         // [R] pose of every hypothesis as rotation + translation   } carved from dynamic shared
  int* hsup;                 // [R] supporters                                            } memory by ransac_carve()
  int* hconv;                // [R] ConvergePose returned true
  int* rnd;                  // [R] speculative draws
  int best, draws;
};
__host__ __device__ inline size_t ransac_bytes(int R) { return size_t(R) * (12 * sizeof(double) + 3 * sizeof(int)); }
__device__ inline void ransac_carve(RansacShared& rs, unsigned char* mem, int R) {   // mem 8-byte aligned
  rs.hRt = reinterpret_cast<double (*)[12]>(mem);
  rs.hsup = reinterpret_cast<int*>(mem + size_t(R) * 12 * sizeof(double));
  rs.hconv = rs.hsup + R;
  rs.rnd = rs.hconv + R;
}

// FeatureAlign::SelectInliers (feature_align.cc:152-216) over P (= fs_found): flags every observation INLIER/OUTLIER.
// T_frame: frame->GetPose().  *rng advances by the reference's number of rand() calls.  Main threads only.
__device__ void select_inliers_cta(const PoseProblem& P, const double* T_frame, const DevParams& dp, sdvlb_rand* rng,
                                   RansacShared& rs, long long* stamps = nullptr) {
  const int tid = threadIdx.x;
  const int size = P.n;
  if (size == 0) return;
  const int R = min(dp.p.max_ransac_its, HYP_MAX);
  const int np = min(min(dp.p.max_ransac_points, size), RANSAC_MAX_PTS);
  const double thr = dp.p.inlier_error_threshold / dp.cam.fx;
  if (tid < 34) rs.backup.r[tid] = rng->r[tid];
  if (tid == 0) rs.backup.n = rng->n;
  main_sync();
  if (tid == 0)   // speculative draws for every hypothesis; the stream is rewound to the reference's count below
    for (int h = 0; h < R; h++) rs.rnd[h] = rand_next(rng);
  main_sync();
  for (int h = tid; h < R; h += PO_MAIN) {
    DSE3 T;
    const bool ok = np == 5 ? converge_pose_small<5>(P, rs.rnd[h] % size, np, size, se3_load(T_frame), dp, &T)
                            : converge_pose_small<0>(P, rs.rnd[h] % size, np, size, se3_load(T_frame), dp, &T);
    rs.hconv[h] = ok ? 1 : 0;
    double Rt[12];
    store_Rt(T, Rt);
#pragma unroll
    for (int k = 0; k < 12; k++) rs.hRt[h][k] = Rt[k];
    // CheckReprojectionError of this hypothesis against every match (feature_align.cc:258-283); the observation
    // loads are warp-uniform
    int sup = 0;
    for (int i = 0; i < size; i++) sup += within_threshold(Rt, P, i, thr) ? 1 : 0;
    rs.hsup[h] = sup;
  }
  main_sync();
  if (stamps) stamps[0] = clock64();
  if (stamps) stamps[1] = clock64();
  if (tid == 0) {   // the reference's loop, replayed over the precomputed hypotheses
    const double sprob = 0.99;
    int nits = dp.p.max_ransac_its, it = 0, best_supporters = 0, best = -1;
    while (it < nits && it < R) {
      if (rs.hconv[it] && rs.hsup[it] > best_supporters) {
        best = it;
        best_supporters = rs.hsup[it];
        const double epsilon = 1.0 - (double(best_supporters) / double(size));
        double tmp = 1.0 - epsilon;
        for (int k = 1; k < np; k++) tmp *= tmp;
        if (tmp < 1e-5) nits = dp.p.max_ransac_its;
        else nits = min(dp.p.max_ransac_its, int(log(1.0 - sprob) / log(1.0 - tmp)));
      }
      it++;
    }
    rs.best = best;
    rs.draws = it;
    if (it < R) {   // rewind: the reference drew `it` numbers only
      for (int k = 0; k < 34; k++) rng->r[k] = rs.backup.r[k];
      rng->n = rs.backup.n;
      for (int k = 0; k < it; k++) rand_next(rng);
    }
  }
  main_sync();
  // "Get low innovation inliers" with best_se3 (identity when no hypothesis ever had a supporter)
  double Rt[12];
  if (rs.best >= 0) {
#pragma unroll
    for (int k = 0; k < 12; k++) Rt[k] = rs.hRt[rs.best][k];
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
__device__ int recheck_cta(const PoseProblem& P, const double* Rt, int from, bool move_if_within, double thr, int to,
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
__device__ void optimize_pose_cta(const PoseProblem& P, double* T_frame, const DevParams& dp, PoseShared& sh,
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
// The commands of one sequence (contiguous, applied in order) by the threads of one CTA.
__device__ void apply_commands(const SeqCmd* __restrict__ cmds, int2 range, const DevParams& dp, int* s_base) {
  const int tid = threadIdx.x;
  for (int ci = range.x; ci < range.x + range.y; ci++) {
    const SeqCmd& C = cmds[ci];
    SeqState* S = C.seq;
    if (C.kind == 0) {
      if (tid == 0) {
        for (int i = 0; i < 7; i++) S->T_last[i] = C.T[i];
        for (int i = 0; i < 6; i++) S->vel[i] = 0.0;
        S->last = C.frame;
        S->has_last = 1;
        S->n_list = 0;
        S->n_cands = 0;
        for (int i = 0; i < 7; i++) C.frame.pose[i] = C.T[i];
      }
      __syncthreads();
      continue;
    }
    // append points to the current list (keyframe seeding / mapping thread output)
    if (tid == 0) {
      *s_base = S->n_list;
      SeqKf& K = S->kf[C.kf_slot];
      K.pyr = C.kf_pyr;
      for (int i = 0; i < 7; i++) K.T[i] = C.T[i];
    }
    __syncthreads();
    const int base = *s_base;
    SeqFeat* L = S->list[S->cur];
    for (int k = tid; k < C.n; k += blockDim.x) {
      if (base + k >= S->max_feats) break;
      const sdvlb_seq_point p = C.pts[k];
      SeqFeat f;
      f.px[0] = p.cur_px[0]; f.px[1] = p.cur_px[1];
      cam_unproject_unit(dp.cam, p.cur_px[0], p.cur_px[1], f.v);
      f.pos[0] = p.pos[0]; f.pos[1] = p.pos[1]; f.pos[2] = p.pos[2];
      f.ref_px[0] = p.ref_px[0]; f.ref_px[1] = p.ref_px[1];
      cam_unproject_unit(dp.cam, p.ref_px[0], p.ref_px[1], f.ref_v);
      f.idepth = p.idepth; f.idepth_std = p.idepth_std;
      f.user_id = p.user_id;
      f.level = p.cur_level; f.ref_level = p.ref_level;
      f.kf = C.kf_slot;
      f.flags = SEQF_HAS_POINT | ((p.flags & SDVLB_CAND_FIXED) ? SEQF_FIXED : 0);
      f.n_successful = p.n_successful; f.n_failed = p.n_failed;
      f.status = SEQP_FOUND;
      f.n_unpromoted = 0;
      L[base + k] = f;
    }
    __syncthreads();
    if (tid == 0) S->n_list = min(base + C.n, S->max_feats);
    __syncthreads();
  }
}

// One CTA per sequence that has commands (sequences that are not part of the step being submitted).
__global__ void __launch_bounds__(128) seq_apply_kernel(const SeqCmd* __restrict__ cmds, const int2* __restrict__ ranges,
                                                        const __grid_constant__ DevParams dp) {
  __shared__ int s_base;
  apply_commands(cmds, ranges[blockIdx.x], dp, &s_base);
}

// ------------------------------------------------------------------------------------------------ prep
__global__ void __launch_bounds__(128) seq_prep_kernel(const __grid_constant__ SeqStepArgs A) {
  SeqState* S = A.seq[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double s_C[3];
  __shared__ int s_warp_cnt[4];
  __shared__ int s_base;
  if (A.cmd_range[blockIdx.x].y > 0) {   // the mapping thread's commands for this sequence (keyframes), in order
    apply_commands(A.cmds, A.cmd_range[blockIdx.x], A.dp, &s_base);
    __threadfence_block();
    __syncthreads();
  }
  AlignJobDev& J = A.jobs[blockIdx.x];
  const int n = S->has_last ? S->n_list : 0;
  if (tid == 0) {
    A.frames[blockIdx.x] = A.cur[blockIdx.x];
    s_base = 0;
    const DSE3 T_last = se3_load(S->T_last);
    const DSE3 Twc = se3_inverse(T_last);   // Frame::GetWorldPosition (frame.h)
    s_C[0] = Twc.tx; s_C[1] = Twc.ty; s_C[2] = Twc.tz;
    const DSE3 prior = se3_mul(se3_exp(S->vel), T_last);   // SDVL::SetMotionModel (sdvl.cc:278-281)
    J.ref = S->last;
    J.cur = A.cur[blockIdx.x];
    J.feats = S->afeat;
    J.n = S->has_last ? n : -1;
    J.fast = 0;
    for (int i = 0; i < 7; i++) J.T_ref[i] = S->T_last[i];
    se3_store(prior, J.T_cur);
    J.out_pose = S->align_pose;
    J.out_info = S->align_info;
    J.out_error = &S->align_error;
    J.out_cycles = S->align_cycles;
    J.trace = nullptr; J.trace_cap = 0; J.forced_n = 0; J.forced_T = nullptr; J.forced_iters = nullptr;
    const size_t nn = size_t(A.max_feats);
    uint8_t* sc = S->align_scratch;
    J.sc_d = reinterpret_cast<double*>(sc);
    J.sc_f = reinterpret_cast<float*>(sc + nn * SDVLB_ALIGN_SC_DOUBLES * 8);
    J.sc_flags = reinterpret_cast<int32_t*>(sc + nn * SDVLB_ALIGN_SC_DOUBLES * 8 + nn * 48 * 4);
    S->align_info[0] = 0; S->align_info[1] = 0;
    if (n == 0) {   // ImageAlign::ComputePose returns at once (image_align.cc:55-58); frame2 keeps the prior
      se3_store(prior, S->align_pose);
      se3_store(prior, A.cur[blockIdx.x].pose);
    }
  }
  __syncthreads();
  const SeqFeat* __restrict__ L = S->list[S->cur];
  // ImageAlign features: every feature of the last frame; candidates: the ones that still observe a point
  for (int b = 0; b < n; b += blockDim.x) {
    const int f = b + tid;
    bool valid = false;
    SeqFeat ft;
    if (f < n) {
      ft = L[f];
      valid = (ft.flags & SEQF_HAS_POINT) != 0;
      sdvlb_align_feat a;
      a.px[0] = ft.px[0]; a.px[1] = ft.px[1];
      a.v[0] = ft.v[0]; a.v[1] = ft.v[1]; a.v[2] = ft.v[2];
      const double dx = ft.pos[0] - s_C[0], dy = ft.pos[1] - s_C[1], dz = ft.pos[2] - s_C[2];
      a.depth = valid ? sqrt(dx * dx + dy * dy + dz * dz) : 1.0;   // image_align.cc:159,234
      a.valid = valid ? 1 : 0;
      a.pad_ = 0;
      S->afeat[f] = a;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; w++) off += s_warp_cnt[w];
    if (valid) {
      const int ci = off + __popc(bal & ((1u << lane) - 1));
      SearchCandDev c;
      const SeqKf& K = S->kf[ft.kf];
      c.ref_pyr = K.pyr;
#pragma unroll
      for (int i = 0; i < 7; i++) c.ref_T[i] = K.T[i];
      c.ref_px[0] = ft.ref_px[0]; c.ref_px[1] = ft.ref_px[1];
      c.ref_v[0] = ft.ref_v[0]; c.ref_v[1] = ft.ref_v[1]; c.ref_v[2] = ft.ref_v[2];
      c.idepth = ft.idepth; c.idepth_std = ft.idepth_std;
      c.px[0] = 0.0; c.px[1] = 0.0;
      c.pos[0] = ft.pos[0]; c.pos[1] = ft.pos[1]; c.pos[2] = ft.pos[2];
      c.ref_level = ft.ref_level;
      c.flags = SDVLB_CAND_PROJECT | ((ft.flags & SEQF_FIXED) ? SDVLB_CAND_FIXED : 0);
      c.cur_index = blockIdx.x;
      c.pad_ = 0;
      S->cands[ci] = c;
      S->cand_feat[ci] = f;
    }
    __syncthreads();
    if (tid == 0) { for (int w = 0; w < 4; w++) s_base += s_warp_cnt[w]; }
    __syncthreads();
  }
  if (tid == 0) S->n_cands = s_base;
}

// ------------------------------------------------------------------------------------------------ post
// Completion of a submission without a separate 1-thread kernel: every CTA checks in after its results are fenced to
// the host; the last one publishes the submission's sequence number in the pinned completion word.
__device__ __forceinline__ void signal_done(const SeqStepArgs& A) {
  if (!A.h_flag) return;
  __threadfence_system();
  const unsigned arrived = atomicAdd(A.d_done, 1u) + 1u;
  if (arrived == unsigned(A.n)) {
    *A.d_done = 0u;
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(A.h_flag) = A.seq_no;
  }
}

struct PostShared {
  PoseShared ps;
  RansacShared rs;
  double T_frame[7];
  sdvlb_rand rng;
  int* win;                    // [n_cells] per cell: rank of its match (INT_MAX: none)          } dynamic shared
  int* slot;                   // [n_cells] index of its match in fs_found, -1: not visited     } memory
  int* order;                  // [n_cells] cell_order_
  int scan[PO_MAIN];
  int attempts, n_found, n_inl, n_outl;
  int kf_live[SDVLB_SEQ_KF_CAP];
};

__global__ void __launch_bounds__(PO_THREADS, 1) seq_post_kernel(const __grid_constant__ SeqStepArgs A) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const long long t_entry = clock64();
  PostShared& sh = *reinterpret_cast<PostShared*>(s_raw);
  double (*part)[PO_MAIN] = reinterpret_cast<double (*)[PO_MAIN]>(s_raw + ((sizeof(PostShared) + 15) / 16) * 16);
  SeqState* S = A.seq[blockIdx.x];
  const int tid = threadIdx.x;
  const int n_cells = A.g.wcells[0] * A.g.hcells[0];
  const int gw = A.g.wcells[0];
  if (!S->has_last) {         // uniform: nothing was tracked (no reset yet)
    if (tid == 0) signal_done(A);
    return;
  }
  if (tid == 0) {             // small shared memory on purpose: this CTA must fit beside the build stream's kernels
    unsigned char* mem = reinterpret_cast<unsigned char*>(&part[0][0]) + sizeof(double) * NVP * PO_MAIN;
    ransac_carve(sh.rs, mem, A.dp.p.max_ransac_its);
    mem += (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16;
    sh.win = reinterpret_cast<int*>(mem);
    sh.slot = sh.win + n_cells;
    sh.order = sh.slot + n_cells;
  }
  __syncthreads();

  // ---- load the persistent FeatureAlign state (all PO_THREADS threads)
  for (int i = tid; i < n_cells; i += PO_THREADS) { sh.order[i] = S->cell_order[i]; sh.win[i] = INT_MAX; sh.slot[i] = -1; }
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
    asm volatile("bar.sync 2, 160;" ::: "memory");
    const int lane = tid - PO_MAIN;
    if (n_cells >= 64) {
      // The n_cells - 1 draws in one go.  rand() is r[n] = r[n-3] + r[n-31] (mod 2^32): the 30 values of a round only
      // need values of earlier rounds through the lag-31 term, so a round is three interleaved prefix sums (stride 3)
      // of known terms -- four shuffle steps instead of 30 dependent updates.  seq[] = the last 34 values in
      // chronological order, then the new ones; it lives in win/slot, idle since SelectPoints.
      uint32_t* seq = reinterpret_cast<uint32_t*>(sh.win);
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
          const int a = sh.order[i], b = sh.order[j];
          sh.order[i] = b; sh.order[j] = a;
        }
      }
      __syncwarp();
      for (int i = lane; i < n_cells; i += 32) S->cell_order[i] = sh.order[i];
      if (lane == 0) S->result->phase_cycles[7] = int(clock64() - t_entry);
      return;
    }
    if (tid == PO_MAIN) {
      for (int i = 1; i < n_cells; ++i) {
        const int j = rand_next(&sh.rng) % (i + 1);
        const int a = sh.order[i], b = sh.order[j];
        sh.order[i] = b; sh.order[j] = a;
      }
    }
    __syncwarp();
    for (int i = lane; i < n_cells; i += 32) S->cell_order[i] = sh.order[i];
    for (int i = lane; i < 34; i += 32) S->rng.r[i] = sh.rng.r[i];
    if (lane == 0) { S->rng.n = sh.rng.n; S->result->phase_cycles[7] = int(clock64() - t_entry); }
    return;
  }

  // =============================== main threads (barrier 1) ===============================
  long long t_phase[8];
  t_phase[0] = clock64();
  const SeqFeat* __restrict__ L = S->list[S->cur];
  SeqFeat* __restrict__ NL = S->list[S->cur ^ 1];
  const int nc = S->n_cands;
  const sdvlb_match* __restrict__ M = S->matches;
  // cell / score of every candidate: in shared memory (the GN partials area is idle here) unless there are too many
  constexpr int kSmemCands = NVP * PO_MAIN;   // two int arrays in NVP * PO_MAIN doubles
  int32_t* __restrict__ c_cell = nc <= kSmemCands ? reinterpret_cast<int32_t*>(&part[0][0]) : S->c_cell;
  int32_t* __restrict__ c_score = nc <= kSmemCands ? reinterpret_cast<int32_t*>(&part[0][0]) + kSmemCands : S->c_score;
  int32_t* __restrict__ c_rank = S->c_rank;

  // ---- ProjectPoint bookkeeping (feature_align.cc:323-339): cell of every seen point
  for (int i = tid; i < nc; i += PO_MAIN) {
    const sdvlb_match m = M[i];
    int cell = -1;
    if (m.status != SDVLB_MATCH_UNSEEN) cell = int(m.proj[1] / SDVLB_CELL) * gw + int(m.proj[0] / SDVLB_CELL);
    if (cell >= n_cells) cell = -1;
    c_cell[i] = cell;
    c_score[i] = L[S->cand_feat[i]].n_successful;
  }
  main_sync();
  // ---- rank inside the cell: Score() descending, stable (std::list::sort, feature_align.cc:111)
  for (int i = tid; i < nc; i += PO_MAIN) {
    const int cell = c_cell[i];
    if (cell < 0) { c_rank[i] = -1; continue; }
    const int sc = c_score[i];
    int rank = 0;
    for (int j = 0; j < nc; j++) {
      const int cj = c_cell[j], sj = c_score[j];
      rank += (cj == cell && (sj > sc || (sj == sc && j < i))) ? 1 : 0;
    }
    c_rank[i] = rank;
    if (M[i].status == SDVLB_MATCH_FOUND) atomicMin(&sh.win[cell], rank);
  }
  main_sync();
  t_phase[1] = clock64();
  // ---- cells in cell_order_: exclusive prefix sum of "has a match" -> max_matches cut-off and fs_found index
  {
    const int per = (n_cells + PO_MAIN - 1) / PO_MAIN;
    const int q0 = tid * per, q1 = min(n_cells, q0 + per);
    int local = 0;
    for (int q = q0; q < q1; q++) local += sh.win[sh.order[q]] != INT_MAX ? 1 : 0;
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
      const int c = sh.order[q];
      const bool has = sh.win[c] != INT_MAX;
      // the loop `for (i < size && matches_ < max_matches_)` visits this cell iff fewer than max_matches so far
      if (pre < max_matches) sh.slot[c] = has ? pre : -2;   // -2: visited, no match
      pre += has ? 1 : 0;
    }
  }
  main_sync();
  const int n_found = sh.n_found;
  PoseProblem P;
  P.o_a = S->o_a; P.o_pos = S->o_pos; P.o_scale = S->o_scale; P.o_err = S->o_err; P.o_flag = S->o_flag;
  P.n = n_found;
  // ---- SelectPoints side effects (feature_align.cc:112-147): attempts, Promote / Unpromote, new features
  {
    int my_attempts = 0;
    for (int i = tid; i < nc; i += PO_MAIN) {
      const int cell = c_cell[i];
      if (cell < 0) continue;
      const int slot = sh.slot[cell];
      if (slot == -1) continue;                       // cell never visited
      const int rank = c_rank[i], win = sh.win[cell];
      if (rank > win) continue;                       // the cell was left at its first match
      my_attempts++;
      if (rank == win && slot >= 0) {                 // found: Promote, new Feature(frame, px, level)
        const sdvlb_match m = M[i];
        SeqFeat f = L[S->cand_feat[i]];
        f.px[0] = m.px[0]; f.px[1] = m.px[1];
        cam_unproject_unit(A.dp.cam, m.px[0], m.px[1], f.v);
        f.level = m.level;
        f.n_successful += 1; f.n_failed = 0;          // Point::Promote (point.cc:102-106)
        f.status = SEQP_FOUND;
        NL[slot] = f;
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

  // ---- SelectInliers (RANSAC)
  t_phase[2] = clock64();
  t_phase[6] = t_phase[7] = t_phase[2];
  select_inliers_cta(P, sh.T_frame, A.dp, &sh.rng, sh.rs, t_phase + 6);
  t_phase[3] = clock64();
  asm volatile("bar.sync 2, 160;" ::: "memory");   // releases the shuffle warp: rand() is its from here on

  // ---- OptimizePose + RescueOutliers + OptimizePose (feature_align.cc:73-82)
  optimize_pose_cta(P, sh.T_frame, A.dp, sh.ps, part);
  main_sync();
  t_phase[4] = clock64();

  // ---- RemoveOutliers (feature_align.cc:245-256), result, GetMotionModel, state update
  SeqResultHost* Rz = S->result;
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
  if (tid < SDVLB_SEQ_KF_CAP) Rz->kf_live[tid] = sh.kf_live[tid];
  if (tid == 0) {
    const DSE3 T_new = se3_load(sh.T_frame);
    const DSE3 mov = se3_mul(T_new, se3_inverse(se3_load(S->T_last)));   // sdvl.cc:266-276
    double vel[6];
    se3_log(mov, vel);
    for (int i = 0; i < 6; i++) S->vel[i] = 0.9 * (0.5 * vel[i] + 0.5 * S->vel[i]);
    for (int i = 0; i < 7; i++) { S->T_last[i] = sh.T_frame[i]; Rz->pose[i] = sh.T_frame[i]; A.cur[blockIdx.x].pose[i] = sh.T_frame[i]; }
    S->last = A.cur[blockIdx.x];
    S->n_list = n_found;
    S->cur ^= 1;
    S->frame_id += 1;
    Rz->stats[0] = S->align_info[0];
    Rz->stats[1] = n_found;
    Rz->stats[2] = sh.attempts;
    Rz->stats[3] = sh.n_inl;
    Rz->stats[4] = sh.n_outl;
    Rz->stats[5] = sh.n_inl;        // Frame::GetNumPoints(): features that still have a point
    Rz->stats[6] = S->align_info[1];
    Rz->stats[7] = n_found;
    Rz->error = 0;
    // latency breakdown of this kernel in SM cycles
    t_phase[5] = clock64();
    Rz->phase_cycles[0] = int(t_phase[1] - t_phase[0]);
    Rz->phase_cycles[1] = int(t_phase[2] - t_phase[1]);
    Rz->phase_cycles[2] = int(t_phase[6] - t_phase[2]);   // RANSAC: draws + hypotheses
    Rz->phase_cycles[3] = int(t_phase[0] - t_entry);      // kernel entry -> state loaded
    Rz->phase_cycles[4] = int(t_phase[3] - t_phase[6]);   //         supporters, replay + final inlier flags
    Rz->phase_cycles[5] = int(t_phase[4] - t_phase[3]);
    Rz->phase_cycles[6] = int(t_phase[5] - t_phase[4]);
    for (int i = 0; i < 4; i++) Rz->align_cycles[i] = S->align_cycles[i];
  }
  // everything the host reads is in pinned memory by now: publish the submission's completion
  __threadfence_system();
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
  if (tid == 0)
    ransac_carve(sh.rs, reinterpret_cast<unsigned char*>(&part[0][0]) + sizeof(double) * NVP * PO_MAIN, A.dp.p.max_ransac_its);
  PoseProblem P;
  P.n = A.n;
  P.o_a = A.scratch; P.o_pos = A.scratch + 2 * size_t(A.n); P.o_scale = A.scratch + 5 * size_t(A.n);
  P.o_err = A.scratch + 6 * size_t(A.n);
  P.o_flag = A.iscratch;
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
    select_inliers_cta(P, sh.T_frame, A.dp, &sh.rng, sh.rs);
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

template <typename K>
cudaError_t opt_in_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
}

}  // namespace

cudaError_t sdvlb_launch_seq_apply(const SeqCmd* d_cmds, const int2* d_ranges, int n_ranges, const DevParams& dp,
                                   cudaStream_t stream) {
  if (n_ranges <= 0) return cudaSuccess;
  sdvlb_common_carveout(seq_apply_kernel);
  seq_apply_kernel<<<n_ranges, 128, 0, stream>>>(d_cmds, d_ranges, dp);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_seq_prep(const SeqStepArgs& A, cudaStream_t stream) {
  sdvlb_common_carveout(seq_prep_kernel);
  seq_prep_kernel<<<A.n, 128, 0, stream>>>(A);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_seq_post(const SeqStepArgs& A, cudaStream_t stream) {
  const int n_cells = A.g.wcells[0] * A.g.hcells[0];
  const size_t dyn = ((sizeof(PostShared) + 15) / 16) * 16 + sizeof(double) * NVP * PO_MAIN +
                     (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16 + size_t(3 * n_cells) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = opt_in_smem(seq_post_kernel, 160 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  sdvlb_common_carveout(seq_post_kernel);
  seq_post_kernel<<<A.n, PO_THREADS, dyn, stream>>>(A);
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_pose_call(const PoseCallArgs& A, cudaStream_t stream) {
  const size_t dyn = ((sizeof(PoseCallShared) + 15) / 16) * 16 + sizeof(double) * NVP * PO_MAIN +
                     (ransac_bytes(A.dp.p.max_ransac_its) + 15) / 16 * 16;
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = opt_in_smem(pose_call_kernel, dyn);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  sdvlb_common_carveout(pose_call_kernel);
  pose_call_kernel<<<1, PO_MAIN, dyn, stream>>>(A);
  return cudaGetLastError();
}
