// Dependent-chain latency of the instructions the serial sections of the ImageAlign / FeatureAlign kernels are made of
// (one warp on one SM): DADD, DFMA, DMUL, FFMA, LDS, SHFL, clock64 itself.   nvcc -arch=sm_100a -O3 fp64_latency.cu && ./a.out
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = seed + lane; sm[lane + 32] = seed;
  __syncwarp();
  double a = seed, b = seed * 0.5;
  float fa = float(seed), fb = 0.5f;
  long long t0, t1;
  const int N = 512;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) a = a + b;
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) a = fma(a, b, b);
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) a = a * b;
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) fa = fmaf(fa, fb, fb);
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  int idx = lane;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) idx = int(sm[idx & 63]) & 63;
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0;
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  t0 = clock64();
  long long acc = 0;
#pragma unroll 1
  for (int i = 0; i < N; i++) acc += clock64();
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  double s, c;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) { sincos(a * 1e-3 + 0.3, &s, &c); a = s + c; }
  t1 = clock64(); if (lane == 0) cyc[7] = (t1 - t0) * 8;   // scaled to N = 512
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) a = sqrt(a + 2.0);
  t1 = clock64(); if (lane == 0) cyc[8] = (t1 - t0) * 8;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) a = 1.0 / (a + 2.0);
  t1 = clock64(); if (lane == 0) cyc[9] = (t1 - t0) * 8;
  out[lane] = a + fa + idx + double(acc);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256); cudaMallocManaged(&cyc, 128);
  for (int r = 0; r < 2; r++) { k<<<1, 32>>>(out, cyc, 1.0000001); cudaDeviceSynchronize(); }
  const char* names[10] = {"DADD", "DFMA", "DMUL", "FFMA", "LDS (dependent, double)", "SHFL + DADD", "clock64 + IADD", "sincos(double)", "sqrt(double)", "1/x (double)"};
  for (int i = 0; i < 10; i++) printf("%-26s %7.1f cycles per dependent op (loop overhead included)\n", names[i], cyc[i] / 512.0);
  return 0;
}
