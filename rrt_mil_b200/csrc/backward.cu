// Streaming kernels of the backward pass (HBM / L2 bound, fp32 math):
//   amax_kernel            max|g| of an fp32 gradient (the stage's fp16 scale, backward.cuh)
//   grad_tile_kernel       fp32 token-order gradient -> scaled fp16 slot-order rows (+ zero pads)
//                          + its transpose [C, M64] (the K-major operand of the weight-gradient GEMM)
//                          + column sums (bias gradient); or fp16 in -> transpose (+ column sums)
//   ln_bwd_kernel          LayerNorm backward (+ residual gradient, + gamma/beta gradients, + amax)
//   wt_convert_kernel      fp32 weight [N,K] -> fp16 transpose [K,N] (the dgrad GEMM's B operand)
// Reference semantics: autograd of modules/rrt.py:117-125 (x + attn(norm(x))) and nn.LayerNorm.
#include "backward.cuh"
#include "kernels.cuh"

namespace rrt {
namespace {

__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, size_t n4,
                                                   uint32_t* __restrict__ amax) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_amax(amax, m);
}

// One CTA = 64 rows (slots) x 128 columns.  IN_F32: `in` is fp32 [L, C] in token order (row = the
// slot's token through `grid`, pads -> 0; grid.H == 0 means identity), values are multiplied by
// S(amax); the scaled fp16 rows go to `rows` [M, C] (nullable).  !IN_F32: `in` is fp16 [M, C].
// Both: transpose to outT [C, M64] (nullable), columns >= M zero; colsum[c] += sum over rows of the
// UNSCALED value (fp16 input: value * 1/S(amax) when amax != null).
// mask_src (IN_F32): mode 0 = keep (x mask_scale) where mask_src != 0 (ReLU read off its output); mode 1 =
// multiply by gelu'(mask_src), mask_src = the layer's pre-activations.
constexpr int kTileR = 64, kTileC = 128, kLdt = 130;
template <bool IN_F32>
__global__ void __launch_bounds__(256) grad_tile_kernel(const void* __restrict__ in_,
                                                        __half* __restrict__ rows,
                                                        __half* __restrict__ outT,
                                                        float* __restrict__ colsum,
                                                        const uint32_t* __restrict__ amax, Grid grid,
                                                        int M, int M64, int C, Dropout drop,
                                                        const float* __restrict__ mask_src, float mask_scale,
                                                        int mask_mode) {
  __shared__ __align__(16) __half tile[kTileR * kLdt];
  __shared__ float red[8][kTileC];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * kTileR, c0 = blockIdx.y * kTileC;
  const uint32_t ab = amax ? __ldg(amax) : 0u;
  const float S = (IN_F32 && amax) ? grad_scale(ab) : 1.f;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = warp * 8 + j, slot = s0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot < M) {
      if (IN_F32) {
        int t = grid.H > 0 ? grid.slot_to_token(slot) : slot;
        if (t < grid.L) {
          v = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(in_) + (size_t)t * C + c0) + lane);
          if (drop.on()) {
            const float4 m = dropout_scale4(drop, (unsigned long long)t * C + c0 + 4 * lane);
            v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
          }
          if (IN_F32 && mask_src) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src + (size_t)t * C + c0) + lane);
            if (mask_mode == 1) {  // mask_src = pre-activations of an nn.GELU
              v.x *= gelu_grad(m.x); v.y *= gelu_grad(m.y); v.z *= gelu_grad(m.z); v.w *= gelu_grad(m.w);
            } else {
            v.x = m.x != 0.f ? v.x * mask_scale : 0.f; v.y = m.y != 0.f ? v.y * mask_scale : 0.f;
            v.z = m.z != 0.f ? v.z * mask_scale : 0.f; v.w = m.w != 0.f ? v.w * mask_scale : 0.f;
            }
          }
        }
      } else {
        v = unpack_h4(__ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(in_) + (size_t)slot * C + c0) + lane));
      }
    }
    csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
    uint2 pk = IN_F32 ? pack_h4(make_float4(v.x * S, v.y * S, v.z * S, v.w * S)) : pack_h4(v);
    if (IN_F32 && rows && slot < M)
      *(reinterpret_cast<uint2*>(rows + (size_t)slot * C + c0) + lane) = pk;
    uint32_t* trow = reinterpret_cast<uint32_t*>(tile + r * kLdt);
    trow[2 * lane] = pk.x;
    trow[2 * lane + 1] = pk.y;
  }
  if (colsum) *reinterpret_cast<float4*>(&red[warp][4 * lane]) = csum;
  __syncthreads();
  if (colsum && tid < kTileC) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][tid];
    if (!IN_F32 && amax) s *= grad_inv_scale(ab);
    if (s != 0.f) atomicAdd(colsum + c0 + tid, s);
  }
  if (outT) {
    // item = (col, chunk of 8 rows): a warp writes 4 transposed rows x 128 contiguous bytes
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int item = it * 256 + tid, chunk = item & 7, col = item >> 3;
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t lo = *reinterpret_cast<const uint16_t*>(tile + (chunk * 8 + 2 * j) * kLdt + col);
        uint32_t hi = *reinterpret_cast<const uint16_t*>(tile + (chunk * 8 + 2 * j + 1) * kLdt + col);
        w[j] = lo | (hi << 16);
      }
      *reinterpret_cast<uint4*>(outT + (size_t)(c0 + col) * M64 + s0 + chunk * 8) =
          make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// LayerNorm backward, warp per token (grid-stride), D = 128 * V.
//   dz: gradient wrt the LayerNorm OUTPUT, either fp16 [Np, D] in slot order, scaled by S(amax_in)
//       (DZ_F16, the R-MSA blocks) or fp32 [L, D] in token order (final norm / CR-MSA norm)
//   dx[t] = (dres[t]) (+ dres2[t]) + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dz * gamma
//   dgamma += dz * xhat, dbeta += dz (fp32 atomics, one per column per CTA); amax_out: max|dx|
template <int V, bool DZ_F16>
__global__ void __launch_bounds__(256) ln_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ x_add, const float* __restrict__ gamma,
    const void* __restrict__ dz_, const uint32_t* __restrict__ amax_in, const float* __restrict__ dres,
    const float* __restrict__ dres2, float* __restrict__ dx, float* __restrict__ dgamma,
    float* __restrict__ dbeta, uint32_t* __restrict__ amax_out, Grid grid) {
  constexpr int D = 128 * V;
  __shared__ float red[8][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float inv_s = (DZ_F16 && amax_in) ? grad_inv_scale(__ldg(amax_in)) : 1.f;
  const float inv_d = 1.f / D;
  float4 gm[V], dg[V], db[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float amax = 0.f;
  for (int t = blockIdx.x * wpb + warp; t < grid.L; t += gridDim.x * wpb) {
    float4 v[V], d[V];
    float s = 0.f;
    const size_t zrow = DZ_F16 ? (size_t)(grid.H > 0 ? grid.token_to_slot(t) : t) : (size_t)t;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = __ldg(reinterpret_cast<const float4*>(x + (size_t)t * D) + lane + 32 * i);
      if (x_add) {
        float4 u = __ldg(reinterpret_cast<const float4*>(x_add + (size_t)t * D) + lane + 32 * i);
        v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      if (DZ_F16) {
        d[i] = unpack_h4(__ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(dz_) + zrow * D) + lane + 32 * i));
        d[i].x *= inv_s; d[i].y *= inv_s; d[i].z *= inv_s; d[i].w *= inv_s;
      } else {
        d[i] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(dz_) + zrow * D) + lane + 32 * i);
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + kLnEps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
      dg[i].x = fmaf(d[i].x, v[i].x, dg[i].x); dg[i].y = fmaf(d[i].y, v[i].y, dg[i].y);
      dg[i].z = fmaf(d[i].z, v[i].z, dg[i].z); dg[i].w = fmaf(d[i].w, v[i].w, dg[i].w);
      db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
      d[i].x *= gm[i].x; d[i].y *= gm[i].y; d[i].z *= gm[i].z; d[i].w *= gm[i].w;  // g
      m1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      m2 += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
    m1 = warp_sum(m1) * inv_d;
    m2 = warp_sum(m2) * inv_d;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o;
      o.x = rstd * (d[i].x - m1 - v[i].x * m2); o.y = rstd * (d[i].y - m1 - v[i].y * m2);
      o.z = rstd * (d[i].z - m1 - v[i].z * m2); o.w = rstd * (d[i].w - m1 - v[i].w * m2);
      if (dres) {
        float4 u = __ldg(reinterpret_cast<const float4*>(dres + (size_t)t * D) + lane + 32 * i);
        o.x += u.x; o.y += u.y; o.z += u.z; o.w += u.w;
      }
      if (dres2) {
        float4 u = __ldg(reinterpret_cast<const float4*>(dres2 + (size_t)t * D) + lane + 32 * i);
        o.x += u.x; o.y += u.y; o.z += u.z; o.w += u.w;
      }
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
      reinterpret_cast<float4*>(dx + (size_t)t * D)[lane + 32 * i] = o;
    }
  }
  if (amax_out) {
    amax = warp_max(amax);
    if (lane == 0 && amax > 0.f) atomic_amax(amax_out, amax);
  }
  // CTA reduction of the per-lane column partials, then one atomic per column
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dgamma : dbeta;
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < V; ++i)
      *reinterpret_cast<float4*>(&red[warp][4 * (lane + 32 * i)]) = pass == 0 ? dg[i] : db[i];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < wpb; ++w) s += red[w][c];
      atomicAdd(dst + c, s);
    }
    __syncthreads();
  }
}

// wT[k, n] = (half) w[n, k]
__global__ void __launch_bounds__(256) wt_convert_kernel(const float* __restrict__ w,
                                                         __half* __restrict__ wT, int N, int K) {
  __shared__ float t[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    t[r][tx] = (n0 + r < N && k0 + tx < K) ? __ldg(w + (size_t)(n0 + r) * K + k0 + tx) : 0.f;
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (k0 + r < K && n0 + tx < N) wT[(size_t)(k0 + r) * N + n0 + tx] = __float2half_rn(t[tx][r]);
}

}  // namespace

cudaError_t launch_amax(const float* x, size_t n, uint32_t* amax, cudaStream_t stream) {
  if (n % 4 || (reinterpret_cast<uintptr_t>(x) & 15)) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  amax_kernel<<<blocks, 256, 0, stream>>>(x, n4, amax);
  return cudaGetLastError();
}

cudaError_t launch_grad_partition(const float* g, const Grid& grid, int M, int C, const uint32_t* amax,
                                  __half* rows, __half* rowsT, float* colsum, cudaStream_t stream,
                                  const Dropout& drop, const float* mask_src, float mask_scale, int mask_mode) {
  if (C % kTileC || M <= 0) return cudaErrorInvalidValue;
  const int M64 = (M + 63) / 64 * 64;
  dim3 gr(M64 / kTileR, C / kTileC);
  grad_tile_kernel<true><<<gr, 256, 0, stream>>>(g, rows, rowsT, colsum, amax, grid, M, M64, C, drop, mask_src,
                                                 mask_scale, mask_mode);
  return cudaGetLastError();
}

namespace {
template <bool MASK_ONLY>
__global__ void __launch_bounds__(256) dropout_kernel(float* __restrict__ x, size_t n4, Dropout drop) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 m = dropout_scale4(drop, 4ull * i);
    float4 v = MASK_ONLY ? make_float4(1.f, 1.f, 1.f, 1.f) : reinterpret_cast<float4*>(x)[i];
    reinterpret_cast<float4*>(x)[i] = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
  }
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) add2_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                   const float* __restrict__ b, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(a) + i);
    if (b) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(b) + i);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}
}  // namespace

// out = a (+ b); n % 4 == 0
cudaError_t launch_add2(float* out, const float* a, const float* b, size_t n, cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  add2_kernel<<<blocks, 256, 0, stream>>>(out, a, b, n / 4);
  return cudaGetLastError();
}

cudaError_t launch_dropout_inplace(float* x, size_t n, const Dropout& drop, cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  if (n == 0 || !drop.on()) return cudaSuccess;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dropout_kernel<false><<<blocks, 256, 0, stream>>>(x, n / 4, drop);
  return cudaGetLastError();
}

cudaError_t launch_dropout_mask(float* out, size_t n, const Dropout& drop, cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  Dropout d = drop;
  if (!d.on()) { d.thresh = 0; d.scale = 1.f; }
  dropout_kernel<true><<<blocks, 256, 0, stream>>>(out, n / 4, d);
  return cudaGetLastError();
}

cudaError_t launch_transpose_f16(const __half* in, int M, int C, __half* outT, float* colsum,
                                 const uint32_t* amax, cudaStream_t stream) {
  if (C % kTileC || M <= 0) return cudaErrorInvalidValue;
  const int M64 = (M + 63) / 64 * 64;
  dim3 gr(M64 / kTileR, C / kTileC);
  Grid none{};
  grad_tile_kernel<false><<<gr, 256, 0, stream>>>(in, nullptr, outT, colsum, amax, none, M, M64, C, Dropout{},
                                                  nullptr, 1.f, 0);
  return cudaGetLastError();
}

#define RRT_BWD_DISPATCH_V(D, ...)                       \
  switch ((D) / 128) {                                   \
    case 1: { constexpr int V = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int V = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int V = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int V = 4; __VA_ARGS__; break; } \
    case 6: { constexpr int V = 6; __VA_ARGS__; break; } \
    case 8: { constexpr int V = 8; __VA_ARGS__; break; } \
    default: return cudaErrorInvalidValue;               \
  }

cudaError_t launch_ln_backward(const float* x, const float* x_add, const float* gamma, const void* dz,
                               bool dz_f16, const uint32_t* amax_in, const float* dres,
                               const float* dres2, float* dx, float* dgamma, float* dbeta,
                               uint32_t* amax_out, const Grid& grid, int D, cudaStream_t stream) {
  if (D % 128) return cudaErrorInvalidValue;
  if (grid.L == 0) return cudaSuccess;
  int blocks = (grid.L + 7) / 8;
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (dz_f16) {
    RRT_BWD_DISPATCH_V(D, ln_bwd_kernel<V, true><<<blocks, 256, 0, stream>>>(
                              x, x_add, gamma, dz, amax_in, dres, dres2, dx, dgamma, dbeta, amax_out, grid));
  } else {
    RRT_BWD_DISPATCH_V(D, ln_bwd_kernel<V, false><<<blocks, 256, 0, stream>>>(
                              x, x_add, gamma, dz, amax_in, dres, dres2, dx, dgamma, dbeta, amax_out, grid));
  }
  return cudaGetLastError();
}

cudaError_t launch_wt_convert(const float* w, __half* wT, int N, int K, cudaStream_t stream) {
  dim3 gr((K + 31) / 32, (N + 31) / 32);
  wt_convert_kernel<<<gr, 256, 0, stream>>>(w, wT, N, K);
  return cudaGetLastError();
}

}  // namespace rrt
