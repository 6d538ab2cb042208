// eigen_shim_sanity.cc (TEST INFRASTRUCTURE) -- how fast are the reference's hot Eigen expressions when they are
// compiled against the stand-in header oracle/ref_shim/Eigen/Dense (what oracle/_ref, the CPU baseline, uses) compared
// with the same arithmetic written as plain loops over arrays?  Real Eigen is not installed here, so this is the
// bound that can be measured: if the stand-in is not slower than straightforward loops, the CPU baseline is not
// inflated by it.  Expressions: image_align.cc:198-199 (H += J J^T w, Jres -= J res w; J 6x1) and
// feature_align.cc:396-397 (A += J^T J w, b -= J^T e w; J 2x6).
//
//   g++ -O3 -march=x86-64-v3 -std=c++14 -Ioracle/ref_shim oracle/eigen_shim_sanity.cc -o /tmp/eigen_shim_sanity && /tmp/eigen_shim_sanity
#include <Eigen/Dense>
#include <chrono>
#include <cstdio>

using Eigen::Matrix;
typedef Matrix<double, 6, 6> M6;
typedef Matrix<double, 6, 1> V6;
typedef Matrix<double, 2, 6> M26;
typedef Matrix<double, 2, 1> V2;

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main() {
  const int N = 20000000;
  volatile double sink = 0;
  // ---- image_align.cc:198-199
  {
    M6 H; V6 b; V6 J;
    for (int i = 0; i < 6; i++) J(i) = 0.1 * (i + 1);
    const double t0 = now();
    for (int n = 0; n < N; n++) {
      const float res = float(n & 7) * 0.25f;
      const double weight = 1.0;
      J(n % 6) += 1e-9;
      H.noalias() += J * J.transpose() * weight;
      b.noalias() -= J * res * weight;
    }
    const double t1 = now();
    sink += H(3, 4) + b(2);
    double Hh[36] = {0}, bh[6] = {0}, Jh[6];
    for (int i = 0; i < 6; i++) Jh[i] = 0.1 * (i + 1);
    const double t2 = now();
    for (int n = 0; n < N; n++) {
      const float res = float(n & 7) * 0.25f;
      const double weight = 1.0;
      Jh[n % 6] += 1e-9;
      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) Hh[c * 6 + r] += Jh[r] * Jh[c] * weight;
      for (int r = 0; r < 6; r++) bh[r] -= Jh[r] * res * weight;
    }
    const double t3 = now();
    sink += Hh[22] + bh[2];
    printf("image_align.cc:198-199  H += J J^T w, Jres -= J res w   stand-in %.2f ns   plain loops %.2f ns   per pixel\n",
           (t1 - t0) / N * 1e9, (t3 - t2) / N * 1e9);
  }
  // ---- feature_align.cc:396-397
  {
    M6 A; V6 b; M26 J; V2 e(0.01, -0.02);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 6; j++) J(i, j) = 0.1 * (i + 1) + 0.01 * j;
    const double t0 = now();
    for (int n = 0; n < N; n++) {
      const double weight = 0.5 + 0.0625 * (n & 7);
      J(n & 1, n % 6) += 1e-9;
      A.noalias() += J.transpose() * J * weight;
      b.noalias() -= J.transpose() * e * weight;
    }
    const double t1 = now();
    sink += A(3, 4) + b(2);
    double Ah[36] = {0}, bh[6] = {0}, Jh[2][6], eh[2] = {0.01, -0.02};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 6; j++) Jh[i][j] = 0.1 * (i + 1) + 0.01 * j;
    const double t2 = now();
    for (int n = 0; n < N; n++) {
      const double weight = 0.5 + 0.0625 * (n & 7);
      Jh[n & 1][n % 6] += 1e-9;
      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) Ah[c * 6 + r] += (Jh[0][r] * Jh[0][c] + Jh[1][r] * Jh[1][c]) * weight;
      for (int r = 0; r < 6; r++) bh[r] -= (Jh[0][r] * eh[0] + Jh[1][r] * eh[1]) * weight;
    }
    const double t3 = now();
    sink += Ah[22] + bh[2];
    printf("feature_align.cc:396-397  A += J^T J w, b -= J^T e w    stand-in %.2f ns   plain loops %.2f ns   per observation\n",
           (t1 - t0) / N * 1e9, (t3 - t2) / N * 1e9);
  }
  return sink == 12345.0;
}
