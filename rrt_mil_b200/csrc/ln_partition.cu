// LayerNorm (+ zero padding + region partition) streaming kernels.  HBM-bound: one warp per
// token row, float4 loads/stores, two-pass statistics held in registers.
#include "kernels.cuh"

namespace rrt {

__device__ __forceinline__ void store4(float* row, int idx4, float4 v) {
  reinterpret_cast<float4*>(row)[idx4] = v;
}
__device__ __forceinline__ void store4(__half* row, int idx4, float4 v) {
  reinterpret_cast<uint2*>(row)[idx4] = pack_h4(v);
}

template <int V, typename OutT>  // D = 128 * V
__device__ __forceinline__ void ln_row(const float* __restrict__ xrow, const float* __restrict__ x0row,
                                       const float* __restrict__ gamma,
                                       const float* __restrict__ beta, OutT* __restrict__ orow,
                                       int lane) {
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = __ldg(reinterpret_cast<const float4*>(xrow) + lane + 32 * i);
    if (x0row != nullptr) {
      float4 u = __ldg(reinterpret_cast<const float4*>(x0row) + lane + 32 * i);
      v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
    }
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float inv_d = 1.f / (128.f * V);
  float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  float rstd = rsqrtf(warp_sum(q) * inv_d + kLnEps);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
    o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
    o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
    o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
    store4(orow, lane + 32 * i, o);
  }
}

// Two adjacent slots per warp, both rows' loads issued before either is reduced (2 x D x 4 bytes in flight per
// warp), so that the whole bag is ONE resident wave (N=9000, D=512: 576 CTAs at 4 per SM; one slot per warp was
// 1152 CTAs = 1.56 waves at 5 per SM, the second one 56 % full: 9.7 us for 27.8 MB).
template <int V>
__global__ void __launch_bounds__(256, 4) ln_partition_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              __half* __restrict__ z, Grid grid) {
  constexpr int D = 128 * V, U = V <= 4 ? 2 : 1;   // wide rows: one per warp (registers)
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int slot0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * U;
  if (slot0 >= grid.Np) return;
  float4 v[U][V];
  bool real[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int slot = slot0 + u;
    const int t = slot < grid.Np ? grid.slot_to_token(slot) : grid.L;
    real[u] = t < grid.L;
#pragma unroll
    for (int i = 0; i < V; ++i)
      v[u][i] = real[u] ? __ldg(reinterpret_cast<const float4*>(x + (size_t)t * D) + lane + 32 * i)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_d = 1.f / D;
  float mean[U], rstd[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
    mean[u] = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float a = v[u][i].x - mean[u], b = v[u][i].y - mean[u], c = v[u][i].z - mean[u], d = v[u][i].w - mean[u];
      q += (a * a + b * b) + (c * c + d * d);
    }
    rstd[u] = rsqrtf(warp_sum(q) * inv_d + kLnEps);
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (slot0 + u >= grid.Np) continue;
      uint2* zrow = reinterpret_cast<uint2*>(z + (size_t)(slot0 + u) * D);
      // pad token: exact zeros AFTER the norm (modules/rmsa.py:200)
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (real[u]) {
        o.x = (v[u][i].x - mean[u]) * rstd[u] * gm.x + bt.x;
        o.y = (v[u][i].y - mean[u]) * rstd[u] * gm.y + bt.y;
        o.z = (v[u][i].z - mean[u]) * rstd[u] * gm.z + bt.z;
        o.w = (v[u][i].w - mean[u]) * rstd[u] * gm.w + bt.w;
      }
      zrow[lane + 32 * i] = real[u] ? pack_h4(o) : make_uint2(0u, 0u);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x1,
                                                            const float* __restrict__ x0,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            float* __restrict__ out, int L) {
  const int D = 128 * V;
  int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= L) return;
  ln_row<V>(x1 + (size_t)t * D, x0 ? x0 + (size_t)t * D : nullptr, gamma, beta,
            out + (size_t)t * D, threadIdx.x & 31);
}

#define RRT_DISPATCH_V(D, ...)                         \
  switch ((D) / 128) {                                 \
    case 1: { constexpr int V = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int V = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int V = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int V = 4; __VA_ARGS__; break; } \
    case 6: { constexpr int V = 6; __VA_ARGS__; break; } \
    case 8: { constexpr int V = 8; __VA_ARGS__; break; } \
    default: return cudaErrorInvalidValue;             \
  }

cudaError_t launch_ln_partition(const float* x, const float* gamma, const float* beta, __half* z,
                                const Grid& grid, int D, cudaStream_t stream) {
  if (D % 128) return cudaErrorInvalidValue;
  const int wpb = 8, rows_per_cta = (D <= 512 ? 2 : 1) * wpb;   // U of the kernel
  int blocks = (grid.Np + rows_per_cta - 1) / rows_per_cta;
  RRT_DISPATCH_V(D, prefer_max_shared(ln_partition_kernel<V>);
                 return launch_chain_kernel(ln_partition_kernel<V>, dim3(blocks), dim3(wpb * 32), 0, stream, x,
                                            gamma, beta, z, grid));
  return cudaGetLastError();
}

cudaError_t launch_add_layernorm(const float* x1, const float* x0, const float* gamma,
                                 const float* beta, float* out, int L, int D,
                                 cudaStream_t stream) {
  if (D % 128) return cudaErrorInvalidValue;
  if (L == 0) return cudaSuccess;
  const int wpb = 8;
  int blocks = (L + wpb - 1) / wpb;
  RRT_DISPATCH_V(D, prefer_max_shared(add_layernorm_kernel<V>); add_layernorm_kernel<V><<<blocks, wpb * 32, 0, stream>>>(x1, x0, gamma, beta, out, L));
  return cudaGetLastError();
}

cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float* out,
                             int L, int D, cudaStream_t stream) {
  return launch_add_layernorm(x, nullptr, gamma, beta, out, L, D, stream);
}

}  // namespace rrt
