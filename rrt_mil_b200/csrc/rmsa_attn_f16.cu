// R-MSA attention core for regions of up to 256 tokens (every configuration the reference ships:
// P = 144 at N = 9000 / region_num 8, P = 196 at N = 50000 / region_num 16), one CTA per
// (region, head), the whole region resident in shared memory:
//   load   Q (with the EPEG halo), K, V rows of this head: fp32 global -> fp16 smem
//   conv   Q' = scale*log2e * (Q + dwconv1d_P(Q; taps_h))   fp32 math, fp16 result in smem
//   core   one warp per 16 query rows: S = Q' K^T (ldmatrix + mma.sync m16n8k16, fp32 accum),
//          online softmax in registers over KV tiles of 48 or 64 keys, O += P V
// fp16 operands carry the same 10-bit mantissa as tf32; softmax state and accumulators are fp32.
// (modules/rmsa.py:103-122; SURVEY.md 0.2-1 for the EPEG-on-Q identity.)
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace rrt {
namespace {

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
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// HD: head dim; NT: 8-key n-tiles per KV tile (6 -> 48 keys, 8 -> 64 keys)
template <int HD, int NT>
__global__ void __launch_bounds__(512) rmsa_attn_f16_kernel(const float* __restrict__ qkv,
                                                            const float* __restrict__ taps,
                                                            float* __restrict__ o, Grid grid, int D,
                                                            int epeg_k, float qscale, int n_kv_tiles,
                                                            bool round_out) {
  constexpr int LDH = HD + 8;   // halves per smem row: 16-byte row skew keeps ldmatrix conflict-free
  constexpr int KS = HD / 16;   // k16 steps over head_dim
  constexpr int ND = HD / 8;    // 8-wide n-tiles of the output
  constexpr int KV = 8 * NT;    // keys per KV tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = blockDim.x >> 5;
  const int P = grid.P;
  const int pad = taps ? epeg_k / 2 : 0;
  const int prow = 16 * W;            // query rows incl. padding of the last warp
  const int pk = n_kv_tiles * KV;     // key rows incl. padding of the last tile
  __half* Qraw = reinterpret_cast<__half*>(smem_raw);   // [prow + 2*pad][LDH]
  __half* Qp = Qraw + (size_t)(prow + 2 * pad) * LDH;   // [prow][LDH]
  __half* Ks = Qp + (size_t)prow * LDH;                 // [pk][LDH]
  __half* Vs = Ks + (size_t)pk * LDH;                   // [pk][LDH]
  float* Ts = reinterpret_cast<float*>(Vs + (size_t)pk * LDH);  // [epeg_k]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int rho = blockIdx.x;  // region
  const int h = blockIdx.y;
  const size_t ld = 3 * (size_t)D;
  const float* base = qkv + (size_t)rho * P * ld + h * HD;

  // ---- stage Q (halo), K, V as fp16 ---------------------------------------------------------
  constexpr int C4 = HD / 4;
  for (int i = tid; i < (prow + 2 * pad) * C4; i += blockDim.x) {
    int r = i / C4, c = (i - r * C4) * 4;
    int p = r - pad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p >= 0 && p < P) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * ld + c));
    uint2 u = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
    *reinterpret_cast<uint2*>(Qraw + (size_t)r * LDH + c) = u;
  }
  for (int i = tid; i < pk * C4; i += blockDim.x) {
    int r = i / C4, c = (i - r * C4) * 4;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (r < P) {
      kv = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * ld + D + c));
      vv = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * ld + 2 * D + c));
    }
    *reinterpret_cast<uint2*>(Ks + (size_t)r * LDH + c) = make_uint2(pack_h2(kv.x, kv.y), pack_h2(kv.z, kv.w));
    *reinterpret_cast<uint2*>(Vs + (size_t)r * LDH + c) = make_uint2(pack_h2(vv.x, vv.y), pack_h2(vv.z, vv.w));
  }
  if (taps)
    for (int i = tid; i < epeg_k; i += blockDim.x) Ts[i] = __ldg(taps + h * epeg_k + i);
  __syncthreads();

  // ---- EPEG on Q: warp w produces exactly the 16 rows it consumes -----------------------------
  {
    const int r0 = 16 * warp;
    for (int cp = lane; cp < HD / 2; cp += 32) {
      for (int r = 0; r < 16; ++r) {
        float2 acc = __half22float2(*reinterpret_cast<const __half2*>(Qraw + (size_t)(r0 + r + pad) * LDH + 2 * cp));
        if (taps) {
          float2 cv = make_float2(0.f, 0.f);
          for (int j = 0; j < epeg_k; ++j) {
            float2 q = __half22float2(*reinterpret_cast<const __half2*>(Qraw + (size_t)(r0 + r + j) * LDH + 2 * cp));
            float wj = Ts[j];
            cv.x = fmaf(wj, q.x, cv.x);
            cv.y = fmaf(wj, q.y, cv.y);
          }
          acc.x += cv.x;
          acc.y += cv.y;
        }
        *reinterpret_cast<__half2*>(Qp + (size_t)(r0 + r) * LDH + 2 * cp) =
            __floats2half2_rn(acc.x * qscale, acc.y * qscale);
      }
    }
  }
  __syncwarp();

  // ---- attention core ---------------------------------------------------------------------------
  uint32_t qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
    ldsm_x4(qa[ks], Qp + (size_t)(16 * warp + (lane & 15)) * LDH + ks * 16 + (lane >> 4) * 8);

  float oacc[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int kt = 0; kt < n_kv_tiles; ++kt) {
    const int kt0 = kt * KV;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b[4];
        ldsm_x4(b, Ks + (size_t)(kt0 + np * 16 + (lane & 7) + (lane >> 4) * 8) * LDH + ks * 16 +
                       ((lane >> 3) & 1) * 8);
        mma_f16_16x8x16(s[2 * np], qa[ks], b[0], b[1]);
        mma_f16_16x8x16(s[2 * np + 1], qa[ks], b[2], b[3]);
      }
    }
    if (kt0 + KV > P) {  // tile padding keys (zero-pad TOKENS are real keys and stay unmasked)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kt0 + nt * 8 + 2 * t + (e & 1) >= P) s[nt][e] = -INFINITY;
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    }
    const float corr[2] = {exp2f(m_run[0] - mx[0]), exp2f(m_run[1] - mx[1])};
    m_run[0] = mx[0];
    m_run[1] = mx[1];
    l_run[0] *= corr[0];
    l_run[1] *= corr[1];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      oacc[i][0] *= corr[0]; oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1]; oacc[i][3] *= corr[1];
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pv = exp2f(s[nt][e] - mx[e >> 1]);
        l_run[e >> 1] += pv;
        s[nt][e] = pv;
      }
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {  // 16 keys per step: two S n-tiles form one A fragment
      uint32_t pa[4] = {pack_h2(s[2 * j][0], s[2 * j][1]), pack_h2(s[2 * j][2], s[2 * j][3]),
                        pack_h2(s[2 * j + 1][0], s[2 * j + 1][1]),
                        pack_h2(s[2 * j + 1][2], s[2 * j + 1][3])};
#pragma unroll
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t b[4];
        ldsm_x4_trans(b, Vs + (size_t)(kt0 + j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDH +
                             np * 16 + (lane >> 4) * 8);
        mma_f16_16x8x16(oacc[2 * np], pa, b[0], b[1]);
        mma_f16_16x8x16(oacc[2 * np + 1], pa, b[2], b[3]);
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
    int q = 16 * warp + g + hh * 8;
    if (q >= P) continue;
    float inv = 1.f / l_run[hh];
    float* orow = o + ((size_t)rho * P + q) * D + h * HD + 2 * t;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      float a = oacc[nd][hh * 2] * inv, b = oacc[nd][hh * 2 + 1] * inv;
      if (round_out) { a = to_tf32(a); b = to_tf32(b); }
      *reinterpret_cast<float2*>(orow + nd * 8) = make_float2(a, b);
    }
  }
}

template <int HD, int NT>
cudaError_t launch(const float* qkv, const float* taps, float* o, const Grid& grid, int D, int heads,
                   int epeg_k, bool round_out, cudaStream_t stream) {
  const int W = (grid.P + 15) / 16;
  const int KV = 8 * NT;
  const int tiles = (grid.P + KV - 1) / KV;
  const int pad = taps ? epeg_k / 2 : 0;
  size_t smem = ((size_t)(16 * W + 2 * pad) + 16 * W + 2 * (size_t)tiles * KV) * (HD + 8) * sizeof(__half) +
                (taps ? epeg_k : 0) * sizeof(float) + 16;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(rmsa_attn_f16_kernel<HD, NT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const float kLog2e = 1.4426950408889634f;
  float qscale = kLog2e / sqrtf((float)HD);
  dim3 g(grid.R, heads);
  rmsa_attn_f16_kernel<HD, NT><<<g, 32 * W, smem, stream>>>(qkv, taps, o, grid, D, epeg_k, qscale,
                                                            tiles, round_out);
  return cudaGetLastError();
}

template <int HD>
cudaError_t launch_hd(const float* qkv, const float* taps, float* o, const Grid& grid, int D,
                      int heads, int epeg_k, bool round_out, cudaStream_t stream) {
  // KV tile of 48 or 64 keys, whichever pads the region's keys less
  int pad48 = (grid.P + 47) / 48 * 48, pad64 = (grid.P + 63) / 64 * 64;
  if (pad48 < pad64) return launch<HD, 6>(qkv, taps, o, grid, D, heads, epeg_k, round_out, stream);
  return launch<HD, 8>(qkv, taps, o, grid, D, heads, epeg_k, round_out, stream);
}
}  // namespace

bool rmsa_attention_f16_supported(const Grid& grid, int D, int heads) {
  int hd = heads > 0 ? D / heads : 0;
  return grid.P <= 256 && (hd == 32 || hd == 64 || hd == 128) && heads <= 65535;
}

cudaError_t launch_rmsa_attention_f16(const float* qkv, const float* taps, float* o,
                                      const Grid& grid, int D, int heads, int epeg_k,
                                      bool round_out, cudaStream_t stream) {
  if (!rmsa_attention_f16_supported(grid, D, heads)) return cudaErrorInvalidValue;
  switch (D / heads) {
    case 32: return launch_hd<32>(qkv, taps, o, grid, D, heads, epeg_k, round_out, stream);
    case 64: return launch_hd<64>(qkv, taps, o, grid, D, heads, epeg_k, round_out, stream);
    default: return launch_hd<128>(qkv, taps, o, grid, D, heads, epeg_k, round_out, stream);
  }
}

}  // namespace rrt
