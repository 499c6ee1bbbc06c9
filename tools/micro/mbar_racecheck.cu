// Does compute-sanitizer racecheck model mbarrier hand-offs?  Warp 0 writes a shared tile, arrives on `full`;
// warp 1 waits on `full`, reads the tile, arrives on `empty`; warp 0 waits on `empty` before the next write.
// Correct by construction (release / acquire through the mbarrier); if racecheck reports hazards here, its
// reports on the warp-specialised kernels (hand-offs by mbarrier only) are the same false positives.
//   nvcc -arch=sm_100a -o mbar_racecheck mbar_racecheck.cu && compute-sanitizer --tool racecheck ./mbar_racecheck
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
}
__global__ void k(int* out, int iters) {
  __shared__ int tile[32];
  __shared__ uint64_t full, empty;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full, 1); mbar_init(&empty, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  int acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (warp == 0) {
      if (it > 0) mbar_wait(&empty, (it - 1) & 1);
      tile[lane] = it * 32 + lane;
      __syncwarp();
      if (lane == 0) mbar_arrive(&full);
    } else {
      mbar_wait(&full, it & 1);
      acc += tile[31 - lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty);
    }
  }
  if (warp == 1) out[lane] = acc;
}
int main() {
  int* d; cudaMalloc(&d, 128);
  k<<<1, 64>>>(d, 8);
  int h[32]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  int expect = 0; for (int it = 0; it < 8; ++it) expect += it * 32 + 31;
  printf("lane 0 sum %d (expected %d) %s\n", h[0], expect, cudaGetLastError() == cudaSuccess && h[0] == expect ? "ok" : "FAIL");
  return 0;
}
