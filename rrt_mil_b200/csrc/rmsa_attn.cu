// R-MSA attention core, one CTA per (query chunk, head, region):
//   Q' = scale * (Q + dwconv1d_P(Q; taps_h))      -- EPEG moved from the logit map onto Q
//   O  = softmax(Q' K^T) V                          -- flash-style, online softmax over KV tiles
// (modules/rmsa.py:103-122; SURVEY.md 0.2-1 for the EPEG identity.)
// Generic in P (any region size, KV tiled by 64) and head_dim in {32,64,128}: the fallback for
// regions of more than 256 tokens (rmsa_attn_f16.cu covers the rest).  fp16 q/k/v in, fp16 O out;
// tensor math on mma.sync m16n8k8 tf32 with fp32 accumulation and fp32 softmax state.
#include "kernels.cuh"

namespace rrt {
namespace {

constexpr int BKV = 64;

template <int HD>
__global__ void __launch_bounds__(256) rmsa_attn_kernel(const __half* __restrict__ qkv,
                                                        const float* __restrict__ taps,
                                                        __half* __restrict__ o, Grid grid, int D,
                                                        int epeg_k, float qscale) {
  constexpr int LDS = HD + 4;
  constexpr int KS = HD / 8;  // k-steps over head_dim; also n-tiles of the output
  extern __shared__ __align__(16) float smem[];
  const int W = blockDim.x >> 5;
  const int pad = taps ? epeg_k / 2 : 0;
  const int qrows = 16 * W + 2 * pad;
  float* Ks = smem;               // [BKV][LDS]
  float* Vs = Ks + BKV * LDS;     // [BKV][LDS]
  float* Qs = Vs + BKV * LDS;     // [qrows][LDS]
  float* Ts = Qs + qrows * LDS;   // [epeg_k]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int rho = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 16 * W;
  const int P = grid.P;
  const size_t ld = 3 * (size_t)D;
  const __half* base = qkv + (size_t)rho * P * ld + h * HD;

  // ---- stage Q rows (with the conv halo) and the taps of this head
  for (int i = tid; i < qrows * (HD / 4); i += blockDim.x) {
    int r = i / (HD / 4), c = (i - r * (HD / 4)) * 4;
    int p = q0 - pad + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p >= 0 && p < P) v = unpack_h4(__ldg(reinterpret_cast<const uint2*>(base + (size_t)p * ld + c)));
    *reinterpret_cast<float4*>(Qs + r * LDS + c) = v;
  }
  if (taps)
    for (int i = tid; i < epeg_k; i += blockDim.x) Ts[i] = __ldg(taps + h * epeg_k + i);
  __syncthreads();

  // ---- Q' fragments (A operand of S = Q' K^T), pre-scaled by scale*log2(e)
  uint32_t qf[KS][4];
  {
    const int r0 = 16 * warp + g;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int r = r0 + (e & 1) * 8, c = ks * 8 + t + (e >> 1) * 4;
        float v = Qs[(r + pad) * LDS + c];
        if (taps) {
          float cv = 0.f;
          for (int j = 0; j < epeg_k; ++j) cv = fmaf(Ts[j], Qs[(r + j) * LDS + c], cv);
          v += cv;
        }
        qf[ks][e] = tf32_bits(v * qscale);
      }
    }
  }

  float oacc[KS][4];
#pragma unroll
  for (int i = 0; i < KS; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int kt0 = 0; kt0 < P; kt0 += BKV) {
    __syncthreads();  // previous tile fully consumed
    for (int i = tid; i < BKV * (HD / 4); i += blockDim.x) {
      int r = i / (HD / 4), c = (i - r * (HD / 4)) * 4;
      int p = kt0 + r;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (p < P) {
        const __half* src = base + (size_t)p * ld + c;
        kv = unpack_h4(__ldg(reinterpret_cast<const uint2*>(src + D)));
        vv = unpack_h4(__ldg(reinterpret_cast<const uint2*>(src + 2 * D)));
      }
      *reinterpret_cast<float4*>(Ks + r * LDS + c) = kv;
      *reinterpret_cast<float4*>(Vs + r * LDS + c) = vv;
    }
    __syncthreads();

    float s[BKV / 8][4];
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nt = 0; nt < BKV / 8; ++nt) {
        const float* p = Ks + (nt * 8 + g) * LDS + ks * 8 + t;
        uint32_t b[2] = {tf32_bits(p[0]), tf32_bits(p[4])};
        mma_tf32_16x8x8(s[nt], qf[ks], b);
      }
    }
    // mask the keys past the end of the region (tile padding only; zero-pad TOKENS are real keys)
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int key = kt0 + nt * 8 + 2 * t + (e & 1);
        if (key >= P) s[nt][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    }
    float corr[2] = {fast_exp2(m_run[0] - mx[0]), fast_exp2(m_run[1] - mx[1])};
    m_run[0] = mx[0];
    m_run[1] = mx[1];
    l_run[0] *= corr[0];
    l_run[1] *= corr[1];
#pragma unroll
    for (int i = 0; i < KS; ++i) {
      oacc[i][0] *= corr[0]; oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1]; oacc[i][3] *= corr[1];
    }
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pv = fast_exp2(s[nt][e] - mx[e >> 1]);
        l_run[e >> 1] += pv;
        s[nt][e] = pv;
      }
    }
    // O += P V.  The k index of each 8-key step is permuted (slot t <-> key 2t, slot t+4 <-> key
    // 2t+1) so that the S accumulator fragment IS the A fragment; V rows are read to match.
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      uint32_t a[4] = {tf32_bits(s[j][0]), tf32_bits(s[j][2]), tf32_bits(s[j][1]), tf32_bits(s[j][3])};
      const float* vrow = Vs + (j * 8 + 2 * t) * LDS + g;
#pragma unroll
      for (int nd = 0; nd < KS; ++nd) {
        uint32_t b[2] = {tf32_bits(vrow[nd * 8]), tf32_bits(vrow[LDS + nd * 8])};
        mma_tf32_16x8x8(oacc[nd], a, b);
      }
    }
  }

#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    int q = q0 + 16 * warp + g + hh * 8;
    if (q >= P) continue;
    float inv = 1.f / l_run[hh];
    __half* orow = o + ((size_t)rho * P + q) * D + h * HD + 2 * t;
#pragma unroll
    for (int nd = 0; nd < KS; ++nd)
      *reinterpret_cast<uint32_t*>(orow + nd * 8) =
          pack_h2(oacc[nd][hh * 2] * inv, oacc[nd][hh * 2 + 1] * inv);
  }
}

template <int HD>
cudaError_t launch(const __half* qkv, const float* taps, __half* o, const Grid& grid, int D,
                   int heads, int epeg_k, cudaStream_t stream) {
  int nb = (grid.P + 15) / 16;
  int chunks = (nb + 7) / 8;
  int W = (nb + chunks - 1) / chunks;
  int pad = taps ? epeg_k / 2 : 0;
  size_t smem = ((size_t)(2 * BKV + 16 * W + 2 * pad) * (HD + 4) + (taps ? epeg_k : 0)) * sizeof(float);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static DeviceOnce configured;  // per instantiation: the attribute call is slow
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rmsa_attn_kernel<HD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  const float kLog2e = 1.4426950408889634f;
  float qscale = kLog2e / sqrtf((float)HD);
  dim3 g(chunks, heads, grid.R);
  rmsa_attn_kernel<HD><<<g, 32 * W, smem, stream>>>(qkv, taps, o, grid, D, epeg_k, qscale);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_rmsa_attention(const __half* qkv, const float* taps, __half* o, const Grid& grid,
                                  int D, int heads, int epeg_k, cudaStream_t stream) {
  if (heads <= 0 || D % heads) return cudaErrorInvalidValue;
  if (grid.R > 65535) return cudaErrorInvalidValue;
  switch (D / heads) {
    case 32: return launch<32>(qkv, taps, o, grid, D, heads, epeg_k, stream);
    case 64: return launch<64>(qkv, taps, o, grid, D, heads, epeg_k, stream);
    case 128: return launch<128>(qkv, taps, o, grid, D, heads, epeg_k, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace rrt
