// synth.cc — seeded synthetic world for the SDVL front-end: a textured plane z=0 seen by a moving pinhole camera.
// TUM/EuRoC/ICL-NUIM are not available offline (BASELINE.json), so parity tests and the bench use this generator
// (SURVEY.md §8d): value-noise + rectangle texture, exact homography render (bilinear, round-to-nearest u8),
// smooth seeded trajectories with ground-truth world->camera poses {q0,q1,q2,q3,tx,ty,tz}.
// Pure host code, no dependency on the oracle or on the CUDA library.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {  // splitmix64
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }
  int range(int lo, int hi) { return lo + int(next() % uint64_t(hi - lo)); }
};

inline float lattice(uint32_t x, uint32_t y, uint32_t seed) {
  uint32_t h = x * 0x85EBCA6Bu ^ y * 0xC2B2AE35u ^ seed * 0x27D4EB2Fu;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return float(h & 0xFFFFFF) * (1.0f / 16777215.0f);
}

void quat_mul(const double a[4], const double b[4], double c[4]) {
  c[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  c[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  c[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  c[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

void quat_to_R(const double q[4], double R[3][3]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - w * z);     R[0][2] = 2 * (x * z + w * y);
  R[1][0] = 2 * (x * y + w * z);     R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - w * x);
  R[2][0] = 2 * (x * z - w * y);     R[2][1] = 2 * (y * z + w * x);     R[2][2] = 1 - 2 * (x * x + y * y);
}

}  // namespace

extern "C" {

// size x size u8 texture: 4 octaves of value noise plus `n_rects` random axis-aligned / rotated rectangles.
void synth_texture(uint8_t* tex, int size, uint32_t seed, int n_rects) {
  std::vector<float> f(size_t(size) * size);
  for (int y = 0; y < size; y++)
    for (int x = 0; x < size; x++) {
      float acc = 0, amp = 1.0f, norm = 0;
      for (int o = 0; o < 4; o++) {
        const int cell = 256 >> (2 * o);  // 256, 64, 16, 4 texels
        const int gx = x / cell, gy = y / cell;
        float fx = float(x % cell) / cell, fy = float(y % cell) / cell;
        fx = fx * fx * (3 - 2 * fx);
        fy = fy * fy * (3 - 2 * fy);
        const float a = lattice(gx, gy, seed + o), b = lattice(gx + 1, gy, seed + o);
        const float c = lattice(gx, gy + 1, seed + o), d = lattice(gx + 1, gy + 1, seed + o);
        acc += amp * ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy);
        norm += amp;
        amp *= 0.5f;
      }
      f[size_t(y) * size + x] = 40.0f + 175.0f * acc / norm;
    }
  Rng rng(0xC0FFEEull ^ (uint64_t(seed) << 20));
  for (int r = 0; r < n_rects; r++) {
    const int cx = rng.range(0, size), cy = rng.range(0, size);
    const int hw = rng.range(6, 90), hh = rng.range(6, 90);
    const float val = float(rng.range(0, 256));
    const bool rot = (rng.next() & 1) != 0;
    const double ang = rng.uni() * 3.14159265358979;
    const float ca = float(std::cos(ang)), sa = float(std::sin(ang));
    const int ext = rot ? int(std::ceil(std::sqrt(double(hw * hw + hh * hh)))) : (hw > hh ? hw : hh);
    for (int y = cy - ext; y <= cy + ext; y++) {
      if (y < 0 || y >= size) continue;
      for (int x = cx - ext; x <= cx + ext; x++) {
        if (x < 0 || x >= size) continue;
        const float dx = float(x - cx), dy = float(y - cy);
        float u = dx, v = dy;
        if (rot) { u = ca * dx + sa * dy; v = -sa * dx + ca * dy; }
        if (std::fabs(u) <= hw && std::fabs(v) <= hh) f[size_t(y) * size + x] = val;
      }
    }
  }
  // three 3x3 box-blur passes (~Gaussian sigma 1.4 texels) so bilinear sampling at ~2 texels/px does not alias
  std::vector<float> g(f.size());
  for (int pass = 0; pass < 3; pass++) {
    for (int y = 0; y < size; y++)
      for (int x = 0; x < size; x++) {
        float acc = 0;
        for (int j = -1; j <= 1; j++)
          for (int i = -1; i <= 1; i++) {
            int xx = x + i, yy = y + j;
            xx = xx < 0 ? 0 : (xx >= size ? size - 1 : xx);
            yy = yy < 0 ? 0 : (yy >= size ? size - 1 : yy);
            acc += f[size_t(yy) * size + xx];
          }
        g[size_t(y) * size + x] = acc * (1.0f / 9.0f);
      }
    f.swap(g);
  }
  for (size_t i = 0; i < f.size(); i++) {
    float v = f[i];
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    tex[i] = uint8_t(v + 0.5f);
  }
}

// Renders the plane z=0 (texture centred at the world origin, `texel_m` metres per texel) through a pinhole camera
// cam = {width,height,fx,fy,u0,v0} at world->camera pose T. Pixels whose ray misses the texture are 0.
static void render_rows(const uint8_t* tex, int size, double texel_m, const double* cam, const double* T, uint8_t* out,
                        int w, int h, int y0, int y1) {
  double R[3][3];
  quat_to_R(T, R);
  // camera centre in world: C = -R^T t
  const double C[3] = {-(R[0][0] * T[4] + R[1][0] * T[5] + R[2][0] * T[6]),
                       -(R[0][1] * T[4] + R[1][1] * T[5] + R[2][1] * T[6]),
                       -(R[0][2] * T[4] + R[1][2] * T[5] + R[2][2] * T[6])};
  const double fx = cam[2], fy = cam[3], u0 = cam[4], v0 = cam[5];
  const double half = 0.5 * size;
  for (int y = y0; y < y1; y++) {
    for (int x = 0; x < w; x++) {
      const double rc[3] = {(x - u0) / fx, (y - v0) / fy, 1.0};
      // ray in world = R^T rc
      const double d[3] = {R[0][0] * rc[0] + R[1][0] * rc[1] + R[2][0] * rc[2],
                           R[0][1] * rc[0] + R[1][1] * rc[1] + R[2][1] * rc[2],
                           R[0][2] * rc[0] + R[1][2] * rc[1] + R[2][2] * rc[2]};
      uint8_t val = 0;
      if (std::fabs(d[2]) > 1e-12) {
        const double s = -C[2] / d[2];
        if (s > 0) {
          const double X = C[0] + s * d[0], Y = C[1] + s * d[1];
          const double tu = X / texel_m + half, tv = Y / texel_m + half;
          if (tu >= 0 && tv >= 0 && tu < size - 1 && tv < size - 1) {
            const int iu = int(tu), iv = int(tv);
            const float fu = float(tu - iu), fv = float(tv - iv);
            const uint8_t* p = tex + size_t(iv) * size + iu;
            const float top = p[0] + fu * (float(p[1]) - float(p[0]));
            const float bot = p[size] + fu * (float(p[size + 1]) - float(p[size]));
            const float v = top + fv * (bot - top);
            val = uint8_t(v + 0.5f);
          }
        }
      }
      out[size_t(y) * w + x] = val;
    }
  }
}

void synth_render(const uint8_t* tex, int size, double texel_m, const double cam[6], const double T[7], uint8_t* out,
                  int w, int h) {
  render_rows(tex, size, texel_m, cam, T, out, w, h, 0, h);
}

// n frames, poses = n x 7, out = n x h x w; uses up to `threads` host threads.
void synth_render_batch(const uint8_t* tex, int size, double texel_m, const double cam[6], const double* poses, int n,
                        uint8_t* out, int w, int h, int threads) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([=]() {
      for (int i = t; i < n; i += threads)
        render_rows(tex, size, texel_m, cam, poses + 7 * i, out + size_t(i) * w * h, w, h, 0, h);
    });
  for (auto& th : pool) th.join();
}

// Smooth seeded trajectory over the plane. kind 0: EuRoC-shaped (<=0.5 m/s, <=15 deg/s), kind 1: TUM-shaped fast
// motion (<=1.5 m/s, <=60 deg/s). fps = frame rate. Camera looks down the world -Z axis from ~`height` metres.
void synth_trajectory(int kind, uint32_t seed, int n, double fps, double height, double* poses) {
  Rng rng(0x7A3Bull * (seed + 1) + 17);
  const double vmax = kind == 0 ? 0.5 : 1.5;                       // m/s
  const double wmax = (kind == 0 ? 15.0 : 60.0) * 3.14159265358979 / 180.0;  // rad/s
  double ph[9], fr[9];
  for (int i = 0; i < 9; i++) { ph[i] = rng.uni() * 6.283185307; fr[i] = 0.15 + 0.35 * rng.uni(); }  // Hz
  for (int k = 0; k < n; k++) {
    const double t = k / fps;
    // position: amplitude chosen so peak speed = A*2*pi*f <= vmax/sqrt(3)
    double C[3];
    for (int a = 0; a < 3; a++) {
      const double A = (vmax / 1.7320508) / (6.283185307 * fr[a]);
      C[a] = A * std::sin(6.283185307 * fr[a] * t + ph[a]) - A * std::sin(ph[a]);
    }
    C[2] = height + 0.3 * C[2];
    double ang[3];
    for (int a = 0; a < 3; a++) {
      const double A = (wmax / 1.7320508) / (6.283185307 * fr[3 + a]);
      ang[a] = A * std::sin(6.283185307 * fr[3 + a] * t + ph[3 + a]) - A * std::sin(ph[3 + a]);
      if (a < 2) ang[a] *= 0.5;  // keep the optical axis near the plane normal
    }
    // base orientation: camera x = world X, y = -world Y, z = -world Z  (rotation by pi about X): q = (0,1,0,0)
    const double qb[4] = {0, 1, 0, 0};
    const double hx = 0.5 * ang[0], hy = 0.5 * ang[1], hz = 0.5 * ang[2];
    const double qx[4] = {std::cos(hx), std::sin(hx), 0, 0};
    const double qy[4] = {std::cos(hy), 0, std::sin(hy), 0};
    const double qz[4] = {std::cos(hz), 0, 0, std::sin(hz)};
    double q1[4], q2[4], q[4];
    quat_mul(qz, qy, q1);
    quat_mul(q1, qx, q2);
    quat_mul(q2, qb, q);   // R_cw = Rz Ry Rx R_base
    double R[3][3];
    quat_to_R(q, R);
    double* P = poses + 7 * k;
    P[0] = q[0]; P[1] = q[1]; P[2] = q[2]; P[3] = q[3];
    P[4] = -(R[0][0] * C[0] + R[0][1] * C[1] + R[0][2] * C[2]);
    P[5] = -(R[1][0] * C[0] + R[1][1] * C[1] + R[1][2] * C[2]);
    P[6] = -(R[2][0] * C[0] + R[2][1] * C[1] + R[2][2] * C[2]);
  }
}

}  // extern "C"
