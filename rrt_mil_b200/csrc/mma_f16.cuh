// Warp-level tensor-core helpers shared by the attention kernels (forward and backward):
// ldmatrix fragment loads and mma.sync m16n8k16 (f16 operands, fp32 accumulate).
//
// Fragment conventions (g = lane >> 2, t = lane & 3):
//   A (16x16, row):  a0 (row g,   k 2t..2t+1)  a1 (row g+8, k 2t..)  a2 (row g, k 2t+8..)  a3 (row g+8, k 2t+8..)
//   B (16x8,  col):  b0 (k 2t..2t+1, n g)      b1 (k 2t+8.., n g)
//   C (16x8):        c0,c1 (row g, n 2t,2t+1)  c2,c3 (row g+8, n 2t,2t+1)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rrt {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                                uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (16 rows x 16 k) of a row-major smem matrix [row][ld halves]: rows row0.., k-step ks
template <int LD>
__device__ __forceinline__ void load_a_rowmajor(uint32_t (&a)[4], const __half* m, int row0, int ks,
                                                int lane) {
  ldsm_x4(a, m + (size_t)(row0 + (lane & 15)) * LD + ks * 16 + (lane >> 4) * 8);
}
// B fragments of X . Y^T with Y row-major [n][ld]: n-tiles (n0..n0+7 -> b[0],b[1]; n0+8.. -> b[2],b[3])
template <int LD>
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], const __half* y, int n0, int ks, int lane) {
  ldsm_x4(b, y + (size_t)(n0 + (lane & 7) + (lane >> 4) * 8) * LD + ks * 16 + ((lane >> 3) & 1) * 8);
}
// B fragments of X . V with V row-major [k][ld]: the 16 k-rows k0.., n columns c0..c0+7 -> b[0],b[1];
// c0+8.. -> b[2],b[3]
template <int LD>
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], const __half* v, int k0, int c0, int lane) {
  ldsm_x4_trans(b, v + (size_t)(k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + c0 + (lane >> 4) * 8);
}

}  // namespace rrt
