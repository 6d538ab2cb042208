// search.cu — K4: Matcher::SearchPoint (matcher.cc:45-121, non-ORB) for a batch of map points, one warp per
// candidate: optional FeatureAlign::ProjectPoint (feature_align.cc:323-339), depth-range projection test, affine warp
// (WarpMatrixAffine :293-312), search level (:314-323), warped 10x10 reference patch (CreatePatch :325-357), corner
// gating (GetCornersInRange :123-230), integer ZMSSD (:447-476, bit-exact) and the 3-parameter inverse-compositional
// Lucas-Kanade refinement (AlignPatch :359-445, fp32 as in the reference).
// Geometry is evaluated redundantly by all 32 lanes in fp64 (uniform control flow); patch pixels, corner scan and LK
// pixels are spread over the lanes; the LK normal equations are reduced with xor-butterfly shuffles (every lane ends
// with the same bits).  Ties in ZMSSD resolve to the lowest corner index, as the reference's in-order scan does.
#include "common.cuh"
#include "orb.cuh"
#include "seq.cuh"

namespace {

constexpr int SE_WARPS = 4;
constexpr int SE_THREADS = SE_WARPS * 32;
// CTAs per SM the batch kernels are compiled for.  The kernels are latency-bound (30 % issue-active at 118 registers = 4
// CTAs per SM); capping the registers at 85 costs ~230 bytes of spills per thread and still wins: 64 -> 59 (5 CTAs) ->
// 55 (6) -> 57 us (8) per 64 sequences, +3 % frames/s for the whole step (A/B on one box, round 2).
constexpr int SE_MIN_CTAS = 6;

struct SearchArgs {
  PyrGeom g;
  DevParams dp;
  int n;
};

__device__ __forceinline__ void cam_project(const sdvlb_camera& c, double x, double y, double z, double& u, double& v) {
  u = c.u0 + c.fx * x / z;   // camera.cc:69-72
  v = c.v0 + c.fy * y / z;
}
__device__ __forceinline__ void cam_unproject(const sdvlb_camera& c, double u, double v, double& x, double& y,
                                              double& z) {  // camera.cc:74-79
  x = (u - c.u0) / c.fx;
  y = (v - c.v0) / c.fy;
  z = 1.0;
  const double n = sqrt(x * x + y * y + z * z);
  x /= n; y /= n; z /= n;
}
__device__ __forceinline__ bool finite2(double a, double b) { return isfinite(a) && isfinite(b); }

// Interpolate8U (extra/utils.cc:44-59), then (uint8_t) truncation (matcher.cc:348)
__device__ __forceinline__ uint8_t interp8u(const uint8_t* __restrict__ img, int stride, float u, float v) {
  const int x = int(floorf(u)), y = int(floorf(v));
  const float sx = u - float(x), sy = v - float(y);
  const float w00 = (1.0f - sx) * (1.0f - sy);
  const float w01 = (1.0f - sx) * sy;
  const float w10 = sx * (1.0f - sy);
  const float w11 = 1.0f - w00 - w01 - w10;
  const uint8_t* p = img + size_t(y) * stride + x;
  const float r = w00 * float(__ldg(p)) + w01 * float(__ldg(p + stride)) + w10 * float(__ldg(p + 1)) +
                  w11 * float(__ldg(p + stride + 1));
  return uint8_t(int(r));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One Matcher::SearchPoint by one warp.  s_bp / s_patch: this warp's 10x10 border patch and 8x8 patch in shared memory.
// kOrb = Config::UseORB() (matcher.cc:79,131-134,243-277): corners are gated with the ORB margin and scored by the
// Hamming distance between the feature's descriptor (qword: word (lane & 7) of its 8 words, in every lane) and the
// corner's (cur.desc, 8 words per corner, written when the frame was built, orb.cu) with threshold
// MIN_ORB_THRESHOLD = 100; everything else is the same code.
template <bool kOrb = false>
__device__ __forceinline__ void search_one(const SearchCandDev& C, const FrameDev& cur, sdvlb_match* __restrict__ out_m,
                                           const PyrGeom& G, const DevParams& dp, uint8_t* s_bp_w, uint8_t* s_patch_w,
                                           uint32_t qword = 0u) {
  const uint32_t* __restrict__ cur_desc = cur.desc;
  const int lane = threadIdx.x & 31;
  const sdvlb_params& P = dp.p;
  const sdvlb_camera& cam = dp.cam;
  const int ps = 8, half = 4;

  sdvlb_match m;
  m.px[0] = m.px[1] = 0.0;
  m.proj[0] = C.px[0]; m.proj[1] = C.px[1];
  m.level = 0; m.status = SDVLB_MATCH_NOT_FOUND; m.zmssd = -1; m.n_in_range = 0;

  const DSE3 T_cur = se3_load(cur.pose);
  const DSE3 T_ref = se3_load(C.ref_T);
  const DSE3 T_ref_w = se3_inverse(T_ref);              // ref_frame->GetWorldPose()
  const DSE3 pose = se3_mul(T_cur, T_ref_w);            // matcher.cc:55
  const bool fixed = (C.flags & SDVLB_CAND_FIXED) != 0;
  double px0 = C.px[0], px1 = C.px[1];
  bool alive = true;

  if (C.flags & SDVLB_CAND_PROJECT) {                   // feature_align.cc:323-339
    double x, y, z, u, v;
    se3_apply(T_cur, C.pos[0], C.pos[1], C.pos[2], x, y, z);
    bool ok = !(z < 0.0);
    if (ok) {
      cam_project(cam, x, y, z, u, v);
      ok = finite2(u, v) && fabs(u) < 1e9 && fabs(v) < 1e9;
      if (ok) {
        const int iu = int(u), iv = int(v);             // Vector2d::cast<int>() truncates
        const int mrg = P.patch_size;
        ok = iu >= mrg && iu < cam.width - mrg && iv >= mrg && iv < cam.height - mrg;   // camera.h:93-95
      }
    }
    if (!ok) { m.status = SDVLB_MATCH_UNSEEN; alive = false; }
    else { px0 = u; px1 = v; m.proj[0] = u; m.proj[1] = v; }
  }

  double pxa0 = 0, pxa1 = 0, pxb0 = 0, pxb1 = 0;
  if (alive) {                                          // matcher.cc:57-77
    const double zmin = 1.0 / (C.idepth + 2.0 * C.idepth_std);
    double wx, wy, wz, x, y, z;
    se3_apply(T_ref_w, C.ref_v[0] * zmin, C.ref_v[1] * zmin, C.ref_v[2] * zmin, wx, wy, wz);
    se3_apply(T_cur, wx, wy, wz, x, y, z);
    if (z < 0.0) alive = false;
    else cam_project(cam, x, y, z, pxa0, pxa1);
    if (alive && !fixed) {
      const double zmax = 1.0 / fmax(C.idepth - 2.0 * C.idepth_std, 0.00000001);
      se3_apply(T_ref_w, C.ref_v[0] * zmax, C.ref_v[1] * zmax, C.ref_v[2] * zmax, wx, wy, wz);
      se3_apply(T_cur, wx, wy, wz, x, y, z);
      if (z < 0.0) alive = false;
      else cam_project(cam, x, y, z, pxb0, pxb1);
    }
  }
  const int level = C.ref_level;
  if (alive) {                                          // matcher.cc:83, camera.h:96-98, feature.h:93-95
    const int lx = int(C.ref_px[0] / double(1 << level)), ly = int(C.ref_px[1] / double(1 << level));
    const int mrg = ps / 2 + 2;
    if (!(lx >= mrg && lx < cam.width / double(1 << level) - mrg && ly >= mrg &&
          ly < cam.height / double(1 << level) - mrg))
      alive = false;
  }

  int slevel = 0;
  double Ai00 = 0, Ai01 = 0, Ai10 = 0, Ai11 = 0;
  if (alive) {                                          // WarpMatrixAffine (matcher.cc:293-312)
    const double depth = 1.0 / C.idepth;
    const double p0 = C.ref_v[0] * depth, p1 = C.ref_v[1] * depth, p2 = C.ref_v[2] * depth;
    double dux, duy, duz, dvx, dvy, dvz;
    cam_unproject(cam, C.ref_px[0] + 5.0 * double(1 << level), C.ref_px[1] + 0.0 * double(1 << level), dux, duy, duz);
    cam_unproject(cam, C.ref_px[0] + 0.0 * double(1 << level), C.ref_px[1] + 5.0 * double(1 << level), dvx, dvy, dvz);
    const double su = p2 / duz, sv = p2 / dvz;
    dux *= su; duy *= su; duz *= su;
    dvx *= sv; dvy *= sv; dvz *= sv;
    double x, y, z, c0, c1, u0, u1, v0, v1;
    se3_apply(pose, p0, p1, p2, x, y, z);       cam_project(cam, x, y, z, c0, c1);
    se3_apply(pose, dux, duy, duz, x, y, z);    cam_project(cam, x, y, z, u0, u1);
    se3_apply(pose, dvx, dvy, dvz, x, y, z);    cam_project(cam, x, y, z, v0, v1);
    const double A00 = (u0 - c0) / 5, A10 = (u1 - c1) / 5, A01 = (v0 - c0) / 5, A11 = (v1 - c1) / 5;
    // GetSearchLevel (matcher.cc:314-323)
    double det = A00 * A11 - A01 * A10;
    const double det0 = det;
    const int maxl = P.max_fast_levels - 1;
    while (det > 3.0 && slevel < maxl) { slevel += 1; det *= 0.25; }
    // Matrix2d::inverse (matcher.cc:330)
    const double invdet = 1.0 / det0;
    Ai00 = A11 * invdet; Ai01 = -A01 * invdet; Ai10 = -A10 * invdet; Ai11 = A00 * invdet;
    if (isnan(Ai00)) alive = false;   // the reference keeps a stale patch here (matcher.cc:331-334); we report NOT_FOUND
  }

  if (alive) {                                          // CreatePatch (matcher.cc:325-357)
    const int W = G.w[level], Hh = G.h[level];
    const uint8_t* __restrict__ img = C.ref_pyr + G.off[level];
    const double pyrx = C.ref_px[0] / double(1 << level), pyry = C.ref_px[1] / double(1 << level);
    for (int i = lane; i < 100; i += 32) {
      const int y = i / 10, x = i - y * 10;
      double ppx = double(x - 5), ppy = double(y - 5);
      ppx *= double(1 << slevel); ppy *= double(1 << slevel);
      const double q0 = Ai00 * ppx + Ai01 * ppy + pyrx;
      const double q1 = Ai10 * ppx + Ai11 * ppy + pyry;
      uint8_t val = 0;
      if (!(q0 < 0 || q1 < 0 || q0 >= W - 1 || q1 >= Hh - 1) && finite2(q0, q1)) val = interp8u(img, W, float(q0), float(q1));
      s_bp_w[i] = val;
      if (y >= 1 && y < 9 && x >= 1 && x < 9) s_patch_w[(y - 1) * 8 + (x - 1)] = val;
    }
  }
  __syncwarp();

  // ---- GetCornersInRange + SearchFeatures (matcher.cc:123-291)
  unsigned long long best_key = ~0ull;
  int n_in_range = 0;
  const int threshold = kOrb ? 100 : ps * ps * 500;   // MIN_ORB_THRESHOLD / MAX_SSD_PER_PIXEL (matcher.h:36-37)
  if (alive) {
    double range = double(P.search_size);
    for (int i = 1; i <= slevel; i++) range *= 1.2;
    const double range2 = range * range;
    const int margin = kOrb ? 4 + 31 / 2 : 1 + ps / 2;   // matcher.cc:131-134
    // epipolar-capsule constants (matcher.cc:140-150)
    double nx = 0, ny = 0, normdist = 0, xdiff = 0, ydiff = 0, vline = 1;
    if (!fixed) {
      double ex = pxa0 - pxb0, ey = pxa1 - pxb1;
      const double en2 = ex * ex + ey * ey;
      if (en2 > 0) { const double en = sqrt(en2); ex /= en; ey /= en; }
      nx = ey; ny = -ex;
      normdist = pxa0 * nx + pxa1 * ny;
      xdiff = pxb0 - pxa0; ydiff = pxb1 - pxa1;
      vline = xdiff * xdiff + ydiff * ydiff;
    }
    // GetZMSSDScore (matcher.cc:447-457): template sums, two pixels per lane
    int sumA = 0, sumAA = 0;
    if constexpr (!kOrb) {
      const int ta0 = s_patch_w[lane], ta1 = s_patch_w[lane + 32];
      sumA = int(__reduce_add_sync(0xffffffffu, unsigned(ta0 + ta1)));
      sumAA = int(__reduce_add_sync(0xffffffffu, unsigned(ta0 * ta0 + ta1 * ta1)));
    }
    const int nc = *cur.n_corners;
    // GetCornersInRange scans every corner of the frame (matcher.cc:123-230); only corners within `range` of the
    // predicted position (fixed points) or of the epipolar segment can pass, so only the 32-px index cells that
    // overlap that region are visited.  The exact test below is the reference's.
    const int gw = G.wcells[0], gh = G.hcells[0];
    const int32_t* __restrict__ g_start = cur.grid;
    const int32_t* __restrict__ g_item = cur.grid + 2 * gw * gh + 1;
    double bx0, bx1, by0, by1;
    if (fixed) { bx0 = bx1 = px0; by0 = by1 = px1; }
    else { bx0 = fmin(pxa0, pxb0); bx1 = fmax(pxa0, pxb0); by0 = fmin(pxa1, pxb1); by1 = fmax(pxa1, pxb1); }
    const double pad = range + 1.0;
    // Degenerate inputs make the reference's test accept every corner (NaN comparisons are false; a zero-length
    // epipolar segment gives u = NaN): visit the whole index then.
    int cx0 = 0, cx1 = gw - 1, cy0 = 0, cy1 = gh - 1;
    const bool box_ok = isfinite(bx0) && isfinite(bx1) && isfinite(by0) && isfinite(by1) && (fixed || vline > 0.0);
    if (box_ok) {
      cx0 = int(fmax(0.0, floor((bx0 - pad) * (1.0 / 32.0))));
      cy0 = int(fmax(0.0, floor((by0 - pad) * (1.0 / 32.0))));
      cx1 = int(fmin(double(gw - 1), floor((bx1 + pad) * (1.0 / 32.0))));
      cy1 = int(fmin(double(gh - 1), floor((by1 + pad) * (1.0 / 32.0))));
    }
    for (int gy = cy0; gy <= cy1; gy++) {
      // cells of one grid row are contiguous in the item list
      const int i0 = __ldg(g_start + gy * gw + cx0), i1 = __ldg(g_start + gy * gw + cx1 + 1);
      for (int base = i0; base < i1; base += 32) {
        const int it = base + lane;
        bool in = false;
        int idx = 0, cx = 0, cy = 0, cl = 0;
        if (it < i1) {
          idx = __ldg(g_item + it);
          const int4 cc = __ldg(cur.corners + idx);
          cx = cc.x; cy = cc.y; cl = cc.z;
          in = idx < nc && abs(cl - level) <= 1 && !(cx - margin < 0 || cy - margin < 0) &&
               !(cy + margin >= G.h[cl] || cx + margin >= G.w[cl]);
          if (in) {
            const double posx = double(cx * (1 << cl)), posy = double(cy * (1 << cl));
            if (fixed) {
              const double dx = px0 - posx, dy = px1 - posy;
              in = !(dx * dx + dy * dy > range2);
            } else {
              const double dist = normdist - (posx * nx + posy * ny);
              if (fabs(dist) > range) in = false;
              else {
                const double u = ((posx - pxa0) * xdiff + (posy - pxa1) * ydiff) / vline;
                if (u > 1) { const double dx = posx - pxb0, dy = posy - pxb1; if (dx * dx + dy * dy > range2) in = false; }
                if (in && u < 0) { const double dx = posx - pxa0, dy = posy - pxa1; if (dx * dx + dy * dy > range2) in = false; }
              }
            }
          }
        }
        uint32_t todo = __ballot_sync(0xffffffffu, in);
        n_in_range += __popc(todo);
        // CompareZMSSDScore (matcher.cc:459-476) for each in-range corner, the 8x8 pixels spread over the warp
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int bcx = __shfl_sync(0xffffffffu, cx, src), bcy = __shfl_sync(0xffffffffu, cy, src);
          const int bcl = __shfl_sync(0xffffffffu, cl, src), bidx = __shfl_sync(0xffffffffu, idx, src);
          int score;
          if constexpr (kOrb) {   // ORBDetector::Distance (extra/orb_detector.cc:399-410): 8 words, one per lane
            const unsigned bits = lane < 8 ? unsigned(__popc(qword ^ __ldg(cur_desc + size_t(bidx) * 8 + lane))) : 0u;
            score = int(__reduce_add_sync(0xffffffffu, bits));
            (void)bcx; (void)bcy; (void)bcl;
          } else {
            const int Wc = G.w[bcl];
            const int py = lane >> 2, pxx = (lane & 3) * 2;
            const uint8_t* __restrict__ cp = cur.pyr + G.off[bcl] + size_t(bcy - half + py) * Wc + (bcx - half + pxx);
            const int b0 = __ldg(cp), b1 = __ldg(cp + 1);
            const int a0 = s_patch_w[py * 8 + pxx], a1 = s_patch_w[py * 8 + pxx + 1];
            const int sB = int(__reduce_add_sync(0xffffffffu, unsigned(b0 + b1)));
            const int sBB = int(__reduce_add_sync(0xffffffffu, unsigned(b0 * b0 + b1 * b1)));
            const int sAB = int(__reduce_add_sync(0xffffffffu, unsigned(a0 * b0 + a1 * b1)));
            score = sumAA - 2 * sAB + sBB - (sumA * sumA - 2 * sumA * sB + sB * sB) / 64;
          }
          const unsigned long long key = (static_cast<unsigned long long>(uint32_t(score)) << 32) | uint32_t(bidx);
          if (key < best_key) best_key = key;
        }
      }
    }
    m.n_in_range = n_in_range;
    int best_score = threshold + 1;
    if (best_key != ~0ull && int(best_key >> 32) < best_score) best_score = int(best_key >> 32);
    m.zmssd = n_in_range == 0 ? -1 : best_score;
    if (best_score >= threshold) alive = false;
  }

  // ---- AlignPatch (matcher.cc:359-445) at the search level
  if (alive) {
    const int bidx = int(best_key & 0xffffffffu);
    const int4 bc = __ldg(cur.corners + bidx);
    const int bl = bc.z;
    const double bx = double(bc.x * (1 << bl)), by = double(bc.y * (1 << bl));
    const int W = G.w[slevel], Hh = G.h[slevel];
    const uint8_t* __restrict__ img = cur.pyr + G.off[slevel];
    // template gradients; each lane owns pixels lane and lane+32
    float gdx[2], gdy[2], tp[2];
    float h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int i = lane + 32 * k, y = i >> 3, x = i & 7;
      const uint8_t* it = &s_bp_w[(y + 1) * 10 + x + 1];
      gdx[k] = float(0.5 * double(int(it[1]) - int(it[-1])));
      gdy[k] = float(0.5 * double(int(it[10]) - int(it[-10])));
      tp[k] = float(s_patch_w[i]);
      h00 += gdx[k] * gdx[k]; h01 += gdx[k] * gdy[k]; h02 += gdx[k];
      h11 += gdy[k] * gdy[k]; h12 += gdy[k];
    }
    // all terms are multiples of 0.25 below 2^24: any summation order is exact
    h00 = warp_sum(h00); h01 = warp_sum(h01); h02 = warp_sum(h02); h11 = warp_sum(h11); h12 = warp_sum(h12);
    const float Hm[3][3] = {{h00, h01, h02}, {h01, h11, h12}, {h02, h12, 64.0f}};
    float Hinv[3][3];
    {
      // Eigen Matrix3f::inverse(): cofactor(i,j) = m(i1,j1)*m(i2,j2) - m(i1,j2)*m(i2,j1); inv(i,j) = cof(j,i)/det
      float cof[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
          cof[i][j] = Hm[i1][j1] * Hm[i2][j2] - Hm[i1][j2] * Hm[i2][j1];
        }
      const float det = cof[0][0] * Hm[0][0] + cof[1][0] * Hm[1][0] + cof[2][0] * Hm[2][0];
      const float invdet = 1.0f / det;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Hinv[i][j] = cof[j][i] * invdet;
    }
    float mean_diff = 0.0f;
    float u = float(bx / double(1 << slevel));
    float v = float(by / double(1 << slevel));
    const float min_update_squared = float(0.03 * 0.03);
    bool converged = false, failed = false;
    for (int iter = 0; iter < P.max_align_its; iter++) {
      if (isnan(u) || isnan(v)) { failed = true; break; }
      if (!(u >= float(half) && v >= float(half) && u < float(W) && v < float(Hh))) break;   // also guards int conversion
      const int u_r = int(floorf(u)), v_r = int(floorf(v));
      if (u_r < half || v_r < half || u_r >= W - half || v_r >= Hh - half) break;
      const float sx = u - float(u_r), sy = v - float(v_r);
      const float wTL = float((1.0 - sx) * (1.0 - sy));
      const float wTR = float(sx * (1.0 - sy));
      const float wBL = float((1.0 - sx) * sy);
      const float wBR = float(sx * sy);
      float j0 = 0, j1 = 0, j2 = 0;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int i = lane + 32 * k, y = i >> 3, x = i & 7;
        const uint8_t* it = img + size_t(v_r + y - half) * W + (u_r - half + x);
        const float sp = wTL * float(__ldg(it)) + wTR * float(__ldg(it + 1)) + wBL * float(__ldg(it + W)) +
                         wBR * float(__ldg(it + W + 1));
        const float res = sp - tp[k] + mean_diff;
        j0 -= res * gdx[k];
        j1 -= res * gdy[k];
        j2 -= res;
      }
      j0 = warp_sum(j0); j1 = warp_sum(j1); j2 = warp_sum(j2);
      const float up0 = Hinv[0][0] * j0 + Hinv[0][1] * j1 + Hinv[0][2] * j2;
      const float up1 = Hinv[1][0] * j0 + Hinv[1][1] * j1 + Hinv[1][2] * j2;
      const float up2 = Hinv[2][0] * j0 + Hinv[2][1] * j1 + Hinv[2][2] * j2;
      u += up0; v += up1; mean_diff += up2;
      if (up0 * up0 + up1 * up1 < min_update_squared) { converged = true; break; }
    }
    if (converged && !failed) {
      m.status = SDVLB_MATCH_FOUND;
      m.px[0] = double(u) * double(1 << slevel);
      m.px[1] = double(v) * double(1 << slevel);
      m.level = slevel;
    }
  }
  if (lane == 0) *out_m = m;
}

// q_desc (kOrb): 8 words per candidate, feature->GetDescriptor() of the candidate's init feature.
template <bool kOrb>
__global__ void __launch_bounds__(SE_THREADS, SE_MIN_CTAS) search_points_kernel(const SearchCandDev* __restrict__ cands,
                                                                   const FrameDev* __restrict__ frames,
                                                                   sdvlb_match* __restrict__ out,
                                                                   const __grid_constant__ SearchArgs A,
                                                                   const uint32_t* __restrict__ q_desc) {
  __shared__ uint8_t s_bp[SE_WARPS][104];
  __shared__ uint8_t s_patch[SE_WARPS][64];
  const int warp = threadIdx.x >> 5;
  const int ci = blockIdx.x * SE_WARPS + warp;
  if (ci >= A.n) return;
  const SearchCandDev& C = cands[ci];
  const FrameDev cur = frames[C.cur_index];
  uint32_t qword = 0u;
  if constexpr (kOrb) qword = __ldg(q_desc + size_t(ci) * 8 + (threadIdx.x & 7));
  search_one<kOrb>(C, cur, out + ci, A.g, A.dp, s_bp[warp], s_patch[warp], qword);
}

// Resident sequences: blockIdx.y = sequence of the step, one warp per feature of the sequence's last frame.  A feature
// that observes a point is a candidate (FeatureAlign::ProjectPoints, feature_align.cc:296-321): lane 0 assembles the
// SearchPoint arguments from the feature and its keyframe slot in shared memory, the match goes to the sequence's own
// device state at the feature's index.
template <bool kOrb>
__global__ void __launch_bounds__(SE_THREADS, SE_MIN_CTAS) search_seq_kernel(const __grid_constant__ SeqStepArgs A) {
  __shared__ uint8_t s_bp[SE_WARPS][104];
  __shared__ uint8_t s_patch[SE_WARPS][64];
  __shared__ SearchCandDev s_cand[SE_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SeqState* S = A.seq[blockIdx.y];
  const int fi = blockIdx.x * SE_WARPS + warp;
  sdvlb_grid_dependency_wait();   // the align kernel's commands decide which features exist
  sdvlb_launch_dependents();      // the post kernel's CTAs (one per sequence) may take their places and wait
  if (!S->has_last || S->hold || fi >= S->n_list) return;
  const SeqFeat& ft = S->list[S->cur][fi];
  if (!(ft.flags & SEQF_HAS_POINT)) return;
  if (lane == 0) {
    SearchCandDev& c = s_cand[warp];
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
    c.cur_index = blockIdx.y;
    c.pad_ = 0;
  }
  __syncwarp();
  uint32_t qword = 0u;
  if constexpr (kOrb) {   // Feature::descriptor_ of the point's init feature, kept beside the feature list
    const uint32_t* fd = S->fdesc[S->cur];
    if (fd) qword = fd[size_t(fi) * 8 + (lane & 7)];
  }
  search_one<kOrb>(s_cand[warp], A.cur[blockIdx.y], S->matches + fi, A.g, A.dp, s_bp[warp], s_patch[warp], qword);
}

// ---- Map::UpdateCandidates (map.cc:397-498): one warp per depth-filter seed.  The geometry is a few dozen fp64
// operations, evaluated by every lane (uniform control flow into search_one); lane 0 writes the seed back.
struct SeedArgs {
  PyrGeom g;
  DevParams dp;
  sdvlb_seed_params sp;
  int n;
};

__device__ __forceinline__ double parallax_cos(const double a[3], const double b[3], const double p[3]) {   // utils.cc:207-213
  double v1[3] = {a[0] - p[0], a[1] - p[1], a[2] - p[2]}, v2[3] = {b[0] - p[0], b[1] - p[1], b[2] - p[2]};
  const double n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
  const double n2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  for (int i = 0; i < 3; i++) { v1[i] /= n1; v2[i] /= n2; }
  return v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
}

template <bool kOrb>
__global__ void __launch_bounds__(SE_THREADS) seed_update_kernel(sdvlb_seed* __restrict__ seeds, const FrameDev cur,
                                                                 const __grid_constant__ SeedArgs A) {
  __shared__ uint8_t s_bp[SE_WARPS][104];
  __shared__ uint8_t s_patch[SE_WARPS][64];
  __shared__ SearchCandDev s_cand[SE_WARPS];
  __shared__ sdvlb_match s_match[SE_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int si = blockIdx.x * SE_WARPS + warp;
  if (si >= A.n) return;
  sdvlb_seed& S = seeds[si];
  const sdvlb_camera& cam = A.dp.cam;
  const DSE3 T_cur = se3_load(cur.pose), T_ref = se3_load(S.ref_T);
  const DSE3 T_ref_w = se3_inverse(T_ref), T_cur_w = se3_inverse(T_cur);
  const double c_ref[3] = {T_ref_w.tx, T_ref_w.ty, T_ref_w.tz}, c_cur[3] = {T_cur_w.tx, T_cur_w.ty, T_cur_w.tz};
  const double v[3] = {S.ref_v[0], S.ref_v[1], S.ref_v[2]};
  double rho = S.rho, sigma2 = S.sigma2, a = S.a, b = S.b;
  int status;
  double depth = 0.0;

  // Point::GetPosition (point.cc:128-142) and Frame::IsPointVisible (frame.cc:104-112)
  double pos[3], rel[3];
  se3_apply(T_ref_w, v[0] / rho, v[1] / rho, v[2] / rho, pos[0], pos[1], pos[2]);
  se3_apply(T_cur, pos[0], pos[1], pos[2], rel[0], rel[1], rel[2]);
  const bool init = A.sp.mode == SDVLB_SEEDS_INIT;   // Map::InitCandidates: no visibility test, no filter update
  bool visible = init || !(rel[2] < 0.0);
  if (visible && !init) {
    double pu, pv;
    cam_project(cam, rel[0], rel[1], rel[2], pu, pv);
    visible = finite2(pu, pv) && fabs(pu) < 1e9 && fabs(pv) < 1e9;
    if (visible) {
      const int iu = int(pu), iv = int(pv);   // Vector2d::cast<int>() truncates
      visible = iu >= 0 && iu < cam.width && iv >= 0 && iv < cam.height;
    }
  }
  const double dx = c_cur[0] - c_ref[0], dy = c_cur[1] - c_ref[1], dz = c_cur[2] - c_ref[2];
  const double baseline = sqrt(dx * dx + dy * dy + dz * dz);   // Frame::DistanceTo (frame.h:129-131)
  if (!visible) {
    status = S.last_kf_id < A.sp.min_kf_id ? SDVLB_SEED_DELETE_OLD : SDVLB_SEED_NOT_VISIBLE;
  } else if (baseline / A.sp.depth_mean < 0.01) {
    status = SDVLB_SEED_SHORT_BASELINE;
  } else {
    // matcher.SearchPoint(frame, feature, point->GetInverseDepth(), point->GetStd(), false, &imgpos, &level)
    if (lane == 0) {
      SearchCandDev& C = s_cand[warp];
      C.ref_pyr = reinterpret_cast<const uint8_t*>(S.ref_frame);   // resolved to the device pyramid by the host
      for (int i = 0; i < 7; i++) C.ref_T[i] = S.ref_T[i];
      C.ref_px[0] = S.ref_px[0]; C.ref_px[1] = S.ref_px[1];
      for (int i = 0; i < 3; i++) { C.ref_v[i] = v[i]; C.pos[i] = 0.0; }
      C.idepth = rho; C.idepth_std = sqrt(sigma2);
      C.px[0] = C.px[1] = 0.0;
      C.ref_level = S.ref_level; C.flags = 0; C.cur_index = 0; C.pad_ = 0;
    }
    __syncwarp();
    uint32_t qword = 0u;
    // ORB mode: the candidate's init-feature descriptor, recomputed from its keyframe (the seed record has no room for it)
    if constexpr (kOrb)
      qword = sdvlb_orb::orb_feature_word(reinterpret_cast<const uint8_t*>(S.ref_frame), A.g, S.ref_level, S.ref_px[0], S.ref_px[1]);
    search_one<kOrb>(s_cand[warp], cur, &s_match[warp], A.g, A.dp, s_bp[warp], s_patch[warp], qword);
    __syncwarp();
    const sdvlb_match m = s_match[warp];
    if (m.status != SDVLB_MATCH_FOUND) {
      if (init) {
        status = SDVLB_SEED_NOT_FOUND;        // map.cc:320-321: next corner
      } else {
        const int nf = S.n_failed + 1;        // Point::Unpromote (point.cc:109-116)
        b += 1.0;
        status = nf > A.dp.p.max_failed ? SDVLB_SEED_DELETE_FAILED : SDVLB_SEED_NOT_FOUND;
        if (lane == 0) { S.n_failed = nf; S.b = b; }
      }
    } else {
      // GetDepthFromTriangulation(pose, feature->GetVector(), v3d, &depth) (utils.cc:193-205)
      const DSE3 pose = se3_mul(T_cur, T_ref_w);
      double R[9], rv[3], vc[3];
      se3_rot(pose, R);
      mat3_mul_vec(R, v[0], v[1], v[2], rv[0], rv[1], rv[2]);
      cam_unproject_unit(cam, m.px[0], m.px[1], vc);
      const double t[3] = {pose.tx, pose.ty, pose.tz};
      const double a00 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
      const double a01 = rv[0] * vc[0] + rv[1] * vc[1] + rv[2] * vc[2];
      const double a11 = vc[0] * vc[0] + vc[1] * vc[1] + vc[2] * vc[2];
      const double det = a00 * a11 - a01 * a01;
      if (lane == 0) { S.px[0] = m.px[0]; S.px[1] = m.px[1]; S.level = m.level; }
      if (det < 0.000001) {
        status = SDVLB_SEED_NO_DEPTH;
      } else {
        const double at0 = rv[0] * t[0] + rv[1] * t[1] + rv[2] * t[2];
        const double at1 = vc[0] * t[0] + vc[1] * t[1] + vc[2] * t[2];
        depth = fabs(-((a11 / det) * at0 + (-a01 / det) * at1));
        double p3d[3];
        se3_apply(T_ref_w, depth * v[0], depth * v[1], depth * v[2], p3d[0], p3d[1], p3d[2]);
        const double cos_alpha = parallax_cos(c_ref, c_cur, p3d);   // map.cc:467-472
        if (cos_alpha >= 0.999999) {
          status = SDVLB_SEED_NO_PARALLAX;
        } else if (depth < A.sp.map_scale * A.sp.scale_min_dist || depth < A.sp.depth_mean * A.sp.scale_min_dist) {
          status = SDVLB_SEED_TOO_CLOSE;
        } else if (init) {
          status = SDVLB_SEED_UPDATED;   // the caller runs Point::InitCandidate(feature, depth)
        } else {
          // Point::Update(frame, depth, px_error_angle) (point.cc:63-100)
          status = SDVLB_SEED_UPDATED;
          const double px_error_angle = atan(1.0 / (2.0 * cam.fx)) * 2.0;   // camera.h:104-107
          const DSE3 pu = se3_mul(T_ref, T_cur_w);
          const double PI = 3.14159265;
          const double tt[3] = {pu.tx, pu.ty, pu.tz};
          const double av[3] = {v[0] * depth - tt[0], v[1] * depth - tt[1], v[2] * depth - tt[2]};
          const double t_norm = sqrt(tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
          const double a_norm = sqrt(av[0] * av[0] + av[1] * av[1] + av[2] * av[2]);
          const double alpha = acos((v[0] * tt[0] + v[1] * tt[1] + v[2] * tt[2]) / t_norm);
          const double beta = acos(-(av[0] * tt[0] + av[1] * tt[1] + av[2] * tt[2]) / (t_norm * a_norm));
          const double beta_plus = beta + px_error_angle;
          const double gamma_plus = PI - alpha - beta_plus;
          const double depth_plus = t_norm * sin(beta_plus) / sin(gamma_plus);
          const double tau = depth_plus - depth;
          const double tau_inverse = 0.5 * (1.0 / fmax(0.0000001, depth - tau) - 1.0 / (depth + tau));
          const double tau2 = tau_inverse * tau_inverse;
          const double x = 1. / depth;
          const double norm_scale = sqrt(sigma2 + tau2);
          double ca = S.cos_alpha, ld = S.last_distance;
          if (!isnan(norm_scale)) {
            const double s2 = 1. / (1. / sigma2 + 1. / tau2);
            const double mm = s2 * (rho / sigma2 + x / tau2);
            double pdf = 0.0;
            if (norm_scale > 0) {   // PDFNormal (point.cc:200-216)
              double ex = x - rho;
              ex *= -ex;
              ex /= 2 * norm_scale * norm_scale;
              pdf = exp(ex) / (norm_scale * sqrt(2.0 * PI));
            }
            double C1 = a / (a + b) * pdf;
            double C2 = b / (a + b) * 1. / S.z_range;
            const double nc = C1 + C2;
            C1 /= nc;
            C2 /= nc;
            const double f = C1 * (a + 1.) / (a + b + 1.) + C2 * a / (a + b + 1.);
            const double e = C1 * (a + 1.) * (a + 2.) / ((a + b + 1.) * (a + b + 2.)) +
                             C2 * a * (a + 1.0) / ((a + b + 1.0) * (a + b + 2.0));
            const double rho_new = C1 * mm + C2 * rho;
            sigma2 = C1 * (s2 + mm * mm) + C2 * (sigma2 + rho * rho) - rho_new * rho_new;
            rho = rho_new;
            a = (e - f) / (f - e / f);
            b = a * (1.0 - f) / f;
            se3_apply(T_ref_w, v[0] / rho, v[1] / rho, v[2] / rho, pos[0], pos[1], pos[2]);
            ca = parallax_cos(c_ref, c_cur, pos);
            const double ex = c_cur[0] - pos[0], ey = c_cur[1] - pos[1], ez = c_cur[2] - pos[2];
            ld = sqrt(ex * ex + ey * ey + ez * ez);   // Frame::DistanceTo(point)
            // Point::HasConverged (point.cc:162-176)
            const double std_d = sqrt(sigma2) / (rho * rho);
            if (4 * std_d * ca / ld < 0.1) status = SDVLB_SEED_CONVERGED;
          }
          if (lane == 0) {
            S.rho = rho; S.sigma2 = sigma2; S.a = a; S.b = b; S.cos_alpha = ca; S.last_distance = ld;
            if (!isnan(norm_scale)) S.n_failed = 0;
            if (status == SDVLB_SEED_CONVERGED) { S.p3d[0] = pos[0]; S.p3d[1] = pos[1]; S.p3d[2] = pos[2]; }
          }
        }
      }
    }
  }
  if (lane == 0) { S.status = status; S.depth = depth; }
}

// Completion signal of a tracking submission: the stream reaches this 1-thread kernel after ImageAlign and SearchPoint
// have finished (their results are already in pinned host memory); it publishes the submission's sequence number in
// pinned host memory, so the host can poll a plain memory word instead of calling into the CUDA driver.
__global__ void signal_kernel(volatile uint32_t* flag, uint32_t seq) {
  __threadfence_system();
  *flag = seq;
}

}  // namespace

cudaError_t sdvlb_launch_signal(uint32_t* h_flag, uint32_t seq, cudaStream_t stream) {
  SDVLB_PREPARE(signal_kernel, 0);
  signal_kernel<<<1, 1, 0, stream>>>(h_flag, seq);
  return cudaGetLastError();
}

// d_qdesc != nullptr: Config::UseORB(), 8 descriptor words per candidate; the frames must carry descriptors
cudaError_t sdvlb_launch_search(const SearchCandDev* d_cands, int n, const FrameDev* d_frames, sdvlb_match* d_out,
                                const PyrGeom& g, const DevParams& dp, cudaStream_t stream, const uint32_t* d_qdesc) {
  if (n <= 0) return cudaSuccess;
  SearchArgs A;
  A.g = g;
  A.dp = dp;
  A.n = n;
  if (d_qdesc) {
    SDVLB_PREPARE(search_points_kernel<true>, 0);
    search_points_kernel<true><<<(n + SE_WARPS - 1) / SE_WARPS, SE_THREADS, 0, stream>>>(d_cands, d_frames, d_out, A, d_qdesc);
  } else {
    SDVLB_PREPARE(search_points_kernel<false>, 0);
    search_points_kernel<false><<<(n + SE_WARPS - 1) / SE_WARPS, SE_THREADS, 0, stream>>>(d_cands, d_frames, d_out, A, nullptr);
  }
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_seed_update(sdvlb_seed* d_seeds, int n, const FrameDev& cur, const PyrGeom& g,
                                     const DevParams& dp, const sdvlb_seed_params& sp, cudaStream_t stream, bool orb) {
  if (n <= 0) return cudaSuccess;
  SeedArgs A;
  A.g = g;
  A.dp = dp;
  A.sp = sp;
  A.n = n;
  if (orb) {
    SDVLB_PREPARE(seed_update_kernel<true>, 0);
    seed_update_kernel<true><<<(n + SE_WARPS - 1) / SE_WARPS, SE_THREADS, 0, stream>>>(d_seeds, cur, A);
  } else {
    SDVLB_PREPARE(seed_update_kernel<false>, 0);
    seed_update_kernel<false><<<(n + SE_WARPS - 1) / SE_WARPS, SE_THREADS, 0, stream>>>(d_seeds, cur, A);
  }
  return cudaGetLastError();
}

cudaError_t sdvlb_launch_search_seq(const SeqStepArgs& A, cudaStream_t stream) {
  const dim3 grid((A.max_feats + SE_WARPS - 1) / SE_WARPS, A.n);   // max_feats: the caller's bound on features per sequence
  if (A.use_orb) {
    SDVLB_PREPARE(search_seq_kernel<true>, 0);
    return sdvlb_launch_dependent(search_seq_kernel<true>, grid, dim3(SE_THREADS), 0, stream, A);
  }
  SDVLB_PREPARE(search_seq_kernel<false>, 0);
  return sdvlb_launch_dependent(search_seq_kernel<false>, grid, dim3(SE_THREADS), 0, stream, A);
}
