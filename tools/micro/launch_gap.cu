// Microbenchmark (tooling): cost of back-to-back dependent launches on one stream, with and without
// shared-memory carveout changes between consecutive kernels.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void tiny(float* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f; }
__global__ void tiny_smem(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += s[1]; }
template <class F> float timeit(F f, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 20; ++i) f();
  cudaDeviceSynchronize(); cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1000.f / iters;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(tiny_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int grids[3] = {16, 148, 1184};
  for (int g : grids) {
    float t1 = timeit([&] { tiny<<<g, 256>>>(d); }, 2000);
    float t2 = timeit([&] { tiny_smem<<<g, 256, 200 * 1024>>>(d); }, 2000);
    float t3 = timeit([&] { tiny<<<g, 256>>>(d); tiny_smem<<<g, 256, 200 * 1024>>>(d); }, 1000) / 2;
    float t4 = timeit([&] { tiny_smem<<<g, 256, 64 * 1024>>>(d); tiny_smem<<<g, 256, 200 * 1024>>>(d); }, 1000) / 2;
    printf("grid %4d: small-only %.2f us | 200KB-smem-only %.2f us | alternating small/200KB %.2f us | alternating 64KB/200KB %.2f us per launch\n", g, t1, t2, t3, t4);
  }
  cudaFuncSetAttribute(tiny, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  for (int g : grids) {
    float t3 = timeit([&] { tiny<<<g, 256>>>(d); tiny_smem<<<g, 256, 200 * 1024>>>(d); }, 1000) / 2;
    printf("grid %4d: alternating small(carveout=100)/200KB %.2f us per launch\n", g, t3);
  }
  return 0;
}
