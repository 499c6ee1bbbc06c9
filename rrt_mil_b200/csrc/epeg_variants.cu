// EPEG ablation variants of InnerAttention (modules/rmsa.py:72-87, 104-129).  None of them
// is on in a shipped configuration (defaults: epeg_2d=False, epeg_type='attn'), so these are plain CUDA-core
// kernels written for correctness and bounded cost, not for the roofline:
//
//   epeg_type = 'value_bf' | 'value_af'   depthwise conv (k,1) or k x k over V folded into the region's rs x rs
//       grid.  The reference reshapes v [B_, h, N, d] -> permute(0,3,1,2) -> [B_, C, rs, rs]: the conv's channel
//       c' is (d, h)-major (c' = d*heads + h) while its result is read back (h', d')-major (c' = h'*hd + d'), so
//       column c' of the result comes from source column (c' % heads)*hd + c' / heads of V.
//         pe[slot, c'] = b[c'] + sum_{a,b} w[c', a, b] * V[slot(rho, pr + a - pad, pc + b - padw), src(c')]
//       value_bf: V[:, c'] += pe[:, c'] before the attention;  value_af: o[:, c'] += pe[:, c'] after it.
//   epeg_2d with epeg_type = 'attn'       k x k conv on the [P, P] logit map of every (region, head):
//         logits'[i, j] = S[i, j] + sum_{a,b} w[h, a, b] * S[i + a - pad, j + b - pad]     (zero padded)
//       (the conv bias is constant over the map and vanishes in the softmax).  One CTA per (region, head) keeps
//       S in shared memory (fp32), so regions are limited to kEpeg2dMaxP tokens.
#include "kernels.cuh"

namespace rrt {
namespace {

__global__ void __launch_bounds__(256) epeg_value_pe_kernel(const __half* __restrict__ qkv, const float* __restrict__ w,
                                                            const float* __restrict__ bias, __half* __restrict__ pe,
                                                            Grid g, int D, int heads, int k, int kw) {
  const int hd = D / heads, P = g.P, rs = g.rs;
  const size_t total = (size_t)g.Np * D;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int slot = (int)(idx / D), cs = (int)(idx - (size_t)slot * D);   // source column of V
    const int cp = (cs % hd) * heads + cs / hd;                               // conv channel c' = d*heads + h
    const int rho = slot / P, p = slot - rho * P, pr = p / rs, pc = p - pr * rs;
    float acc = bias ? __ldg(bias + cp) : 0.f;
    const float* wc = w + (size_t)cp * k * kw;
    for (int a = 0; a < k; ++a) {
      const int r = pr + a - k / 2;
      if (r < 0 || r >= rs) continue;
      for (int b = 0; b < kw; ++b) {
        const int c = pc + b - kw / 2;
        if (c < 0 || c >= rs) continue;
        acc = fmaf(__ldg(wc + a * kw + b),
                   __half2float(qkv[(size_t)(rho * P + r * rs + c) * 3 * D + 2 * D + cs]), acc);
      }
    }
    pe[(size_t)slot * D + cp] = __float2half_rn(acc);
  }
}

// dst[row * ld + col0 + c] += pe[row * D + c]
__global__ void __launch_bounds__(256) epeg_value_add_kernel(__half* __restrict__ dst, int ld, int col0,
                                                             const __half* __restrict__ pe, int rows, int D) {
  const size_t total = (size_t)rows * (D / 2);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / (D / 2)), c2 = (int)(idx - (size_t)row * (D / 2));
    __half2* d = reinterpret_cast<__half2*>(dst + (size_t)row * ld + col0) + c2;
    const float2 a = __half22float2(*d), b = __half22float2(reinterpret_cast<const __half2*>(pe + (size_t)row * D)[c2]);
    *d = __floats2half2_rn(a.x + b.x, a.y + b.y);
  }
}

constexpr int kEpeg2dMaxP = 160;

// grid (heads, R), 256 threads.  smem: S [P][P+1] fp32 | q, k, v [P][HD] f16 | taps [k*k] | prow [8][P]
template <int HD>
__global__ void __launch_bounds__(256) rmsa_attn_epeg2d_kernel(const __half* __restrict__ qkv, const float* __restrict__ taps,
                                                               __half* __restrict__ o, Grid g, int D, int k,
                                                               float scale) {
  extern __shared__ __align__(16) float smem_f[];
  const int P = g.P, PS = P + 1, h = blockIdx.x, rho = blockIdx.y;
  float* S = smem_f;                                         // [P][PS]
  float* tw = S + (((size_t)P * PS + 3) & ~(size_t)3);      // [k*k]
  float* prow = tw + ((k * k + 3) & ~3);                     // [8][P]
  __half* qs = reinterpret_cast<__half*>(prow + 8 * ((P + 3) & ~3));
  __half* ks = qs + (size_t)P * HD;
  __half* vs = ks + (size_t)P * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < k * k; i += 256) tw[i] = __ldg(taps + (size_t)h * k * k + i);
  for (int i = tid; i < P * (HD / 8); i += 256) {
    const int r = i / (HD / 8), c8 = i - r * (HD / 8);
    const uint4* src = reinterpret_cast<const uint4*>(qkv + (size_t)(rho * P + r) * 3 * D + h * HD) + c8;
    reinterpret_cast<uint4*>(qs + (size_t)r * HD)[c8] = __ldg(src);
    reinterpret_cast<uint4*>(ks + (size_t)r * HD)[c8] = __ldg(src + D / 8);
    reinterpret_cast<uint4*>(vs + (size_t)r * HD)[c8] = __ldg(src + 2 * D / 8);
  }
  __syncthreads();
  // S = scale * q k^T
  for (int e = tid; e < P * P; e += 256) {
    const int i = e / P, j = e - i * P;
    const __half2* qi = reinterpret_cast<const __half2*>(qs + (size_t)i * HD);
    const __half2* kj = reinterpret_cast<const __half2*>(ks + (size_t)j * HD);
    float acc = 0.f;
#pragma unroll 8
    for (int d2 = 0; d2 < HD / 2; ++d2) {
      const float2 a = __half22float2(qi[d2]), b = __half22float2(kj[d2]);
      acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
    }
    S[(size_t)i * PS + j] = acc * scale;
  }
  __syncthreads();
  const int pad = k / 2;
  float* pr = prow + warp * ((P + 3) & ~3);
  for (int i = warp; i < P; i += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < P; j += 32) {
      float acc = S[(size_t)i * PS + j];
      for (int a = 0; a < k; ++a) {
        const int ii = i + a - pad;
        if (ii < 0 || ii >= P) continue;
        const float* srow = S + (size_t)ii * PS;
        const float* wrow = tw + a * k;
        const int b0 = max(0, pad - j), b1 = min(k, P + pad - j);
        for (int b = b0; b < b1; ++b) acc = fmaf(wrow[b], srow[j + b - pad], acc);
      }
      pr[j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < P; j += 32) {
      const float e = __expf(pr[j] - mx);
      pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    for (int d = lane; d < HD; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < P; ++j) acc = fmaf(pr[j], __half2float(vs[(size_t)j * HD + d]), acc);
      o[(size_t)(rho * P + i) * D + h * HD + d] = __float2half_rn(acc * inv);
    }
    __syncwarp();
  }
}

size_t epeg2d_smem_bytes(int P, int HD, int k) {
  return ((((size_t)P * (P + 1) + 3) & ~(size_t)3) + ((k * k + 3) & ~3) + 8 * (size_t)((P + 3) & ~3)) * 4 +
         3 * (size_t)P * HD * 2;
}
}  // namespace

// pe [Np, D] f16 scratch; w [D, k, kw], bias [D] or null
cudaError_t launch_epeg_value_pe(const __half* qkv, const float* w, const float* bias, __half* pe, const Grid& grid,
                                 int D, int heads, int k, int kw, cudaStream_t stream) {
  if (!w || D % heads || k < 1 || (k & 1) == 0 || (kw != 1 && kw != k) || grid.rs * grid.rs != grid.P)
    return cudaErrorInvalidValue;
  size_t total = (size_t)grid.Np * D;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  epeg_value_pe_kernel<<<blocks, 256, 0, stream>>>(qkv, w, bias, pe, grid, D, heads, k, kw);
  return cudaGetLastError();
}

cudaError_t launch_epeg_value_add(__half* dst, int ld, int col0, const __half* pe, int rows, int D,
                                  cudaStream_t stream) {
  if (D % 2 || ld % 2 || col0 % 2) return cudaErrorInvalidValue;
  size_t total = (size_t)rows * (D / 2);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  epeg_value_add_kernel<<<blocks, 256, 0, stream>>>(dst, ld, col0, pe, rows, D);
  return cudaGetLastError();
}

bool rmsa_attention_epeg2d_supported(const Grid& grid, int D, int heads, int k) {
  const int hd = heads > 0 ? D / heads : 0;
  return (hd == 32 || hd == 64 || hd == 128) && D == hd * heads && grid.P <= kEpeg2dMaxP && k >= 1 && (k & 1) &&
         epeg2d_smem_bytes(grid.P, hd, k) <= 227 * 1024;
}

// taps [heads, k, k]
cudaError_t launch_rmsa_attention_epeg2d(const __half* qkv, const float* taps, __half* o, const Grid& grid, int D,
                                         int heads, int k, cudaStream_t stream) {
  if (!rmsa_attention_epeg2d_supported(grid, D, heads, k) || !taps) return cudaErrorInvalidValue;
  const int hd = D / heads;
  const size_t smem = epeg2d_smem_bytes(grid.P, hd, k);
  const float scale = 1.f / sqrtf((float)hd);
  dim3 gr(heads, grid.R);
#define RRT_E2D(HDV)                                                                                          \
  {                                                                                                           \
    static DeviceOnce configured;                                                                             \
    if (configured.needed()) {                                                                                \
      cudaError_t e = cudaFuncSetAttribute(rmsa_attn_epeg2d_kernel<HDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           227 * 1024);                                                       \
      if (e != cudaSuccess) return e;                                                                         \
    }                                                                                                         \
    rmsa_attn_epeg2d_kernel<HDV><<<gr, 256, smem, stream>>>(qkv, taps, o, grid, D, k, scale);                 \
  }
  if (hd == 32) RRT_E2D(32) else if (hd == 64) RRT_E2D(64) else RRT_E2D(128)
#undef RRT_E2D
  return cudaGetLastError();
}

}  // namespace rrt
