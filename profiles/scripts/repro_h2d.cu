#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
int main() {
  unsigned char* d = nullptr; size_t stride = 360960; 
  printf("malloc %d\n", cudaMalloc((void**)&d, stride * 64 * 8));
  unsigned char* h = (unsigned char*)malloc(2 * stride);
  cudaStream_t s; int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
  printf("stream %d\n", cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
  printf("copy0 %d\n", cudaMemcpyAsync(d, h, stride, cudaMemcpyHostToDevice, s));
  printf("copy1 %d\n", cudaMemcpyAsync(d + 64 * stride, h + stride, stride, cudaMemcpyHostToDevice, s));
  printf("copy1b %d\n", cudaMemcpyAsync(d + 64 * stride, h + 528, stride, cudaMemcpyHostToDevice, s));
  printf("sync %d\n", cudaStreamSynchronize(s));
  return 0;
}
