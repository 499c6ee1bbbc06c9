// Backward of the R-MSA attention core (and of the landmark MHA: P = 64, no EPEG), one CTA per
// (region, head), the whole region resident in shared memory as in the forward kernel
// (rmsa_attn_f16.cu).  Autograd of modules/rmsa.py:103-122 through the EPEG-on-Q identity:
//   Qe = Q + dwconv1d_P(Q; taps_h),  S = scale * Qe K^T,  A = softmax(S),  O = A V
//   dV = A^T dO,  dA = dO V^T,  dS = A * (dA - rowsum(dO * O)),  dQe = scale * dS K,  dK = scale * dS^T Qe
//   dQ = dQe + dwconv1d_P^T(dQe; taps_h),   dtaps_h[d] = sum_{i,c} dQe[i,c] * Q[i + d - k/2, c]
// (the conv bias is constant along the softmax axis: its gradient is exactly zero.)
//
// Phases (warps own 16-row blocks; the softmax is recomputed, nothing but q/k/v/o is saved):
//   A   rows = queries: Q' = scale*log2e*Qe (Toeplitz MMA as in the forward) -> smem, row max / 1/sum,
//       D_i = rowsum(dO * O)
//   2   rows = keys:    S^T = K Q'^T, dP^T = V dO^T  ->  dV += P^T dO,  dK += dS^T Q'
//   1B  rows = queries: S = Q' K^T, dP = dO V^T      ->  dQe += dS K
//   C   EPEG transpose on dQe (smem), tap gradients, stores
// All gradients are in the caller's scaled fp16 domain (backward.cuh); dtaps is unscaled (fp32).
#include "backward.cuh"
#include "kernels.cuh"
#include "mma_f16.cuh"

namespace rrt {
namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// HD: head dim (32 | 64); RB: 16-row blocks per warp
template <int HD, int RB>
__global__ void __launch_bounds__(RB == 1 ? 320 : 256) rmsa_attn_bwd_kernel(
    const __half* __restrict__ qkv, const __half* __restrict__ o, const __half* __restrict__ dO,
    const float* __restrict__ taps, __half* __restrict__ dqkv, float* __restrict__ dtaps,
    const uint32_t* __restrict__ amax, int P, int D, int epeg_k, float scale, int PR, int q_rows) {
  constexpr int LDH = HD + 8;
  constexpr int KS = HD / 16;
  constexpr int ND = HD / 8;
  constexpr int C8 = HD / 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pad = taps ? epeg_k / 2 : 0;
  __half* Qs = reinterpret_cast<__half*>(smem_raw);  // [q_rows][LDH]: row r holds Q[r - pad]
  __half* Ks = Qs + (size_t)q_rows * LDH;             // [PR][LDH]
  __half* Vs = Ks + (size_t)PR * LDH;                 // [PR][LDH]
  __half* Gs = Vs + (size_t)PR * LDH;                 // [PR][LDH]  dO
  __half* Q2 = Gs + (size_t)PR * LDH;                 // [PR][LDH]  Q' (later dQe)
  float* m2s = reinterpret_cast<float*>(Q2 + (size_t)PR * LDH);  // [PR] row max (log2 domain)
  float* lis = m2s + PR;                              // [PR] 1 / row sum
  float* Drs = lis + PR;                              // [PR] rowsum(dO * O)
  float* Ts = Drs + PR;                               // [64] taps
  float* red = Ts + 64;                               // [64] tap gradient partials

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int W = blockDim.x >> 5;
  const int rho = blockIdx.x, h = blockIdx.y;
  const size_t ld = 3 * (size_t)D;
  const __half* base = qkv + (size_t)rho * P * ld + h * HD;
  const __half* gbase = dO + (size_t)rho * P * D + h * HD;
  const __half* obase = o + (size_t)rho * P * D + h * HD;
  const float qscale = scale * kLog2e;

  // ---- stage Q (halo), K, V, dO --------------------------------------------------------------
  for (int i = tid; i < q_rows * C8; i += blockDim.x) {
    int r = i / C8, c = (i - r * C8) * 8;
    int p = r - pad;
    bool ok = p >= 0 && p < P;
    cp_async16(Qs + (size_t)r * LDH + c, base + (size_t)(ok ? p : 0) * ld + c, ok);
  }
  for (int i = tid; i < PR * C8; i += blockDim.x) {
    int r = i / C8, c = (i - r * C8) * 8;
    bool ok = r < P;
    const __half* src = base + (size_t)(ok ? r : 0) * ld + c;
    cp_async16(Ks + (size_t)r * LDH + c, src + D, ok);
    cp_async16(Vs + (size_t)r * LDH + c, src + 2 * D, ok);
    cp_async16(Gs + (size_t)r * LDH + c, gbase + (size_t)(ok ? r : 0) * D + c, ok);
  }
  cp_async_commit();
  if (tid < 64) {
    Ts[tid] = (taps && tid < epeg_k) ? __ldg(taps + h * epeg_k + tid) : 0.f;
    red[tid] = 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- phase A ---------------------------------------------------------------------------------
#pragma unroll 1
  for (int rb = 0; rb < RB; ++rb) {
    const int i0 = 16 * (warp + rb * W);
    uint32_t qa[KS][4];
    if (taps) {
      float qacc[ND][4];
#pragma unroll
      for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) qacc[i][e] = 0.f;
      const int nkc = (16 + epeg_k - 1 + 15) / 16;
      for (int kc = 0; kc < nkc; ++kc) {
        uint32_t ca[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ro = g + (e & 1) * 8;
          const int co = 16 * kc + 2 * t + (e >> 1) * 8;
          float v[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            int d = co + u - ro;
            float x = (d >= 0 && d < epeg_k) ? Ts[d] : 0.f;
            v[u] = x + (d == pad ? 1.f : 0.f);
          }
          ca[e] = pack_h2(v[0], v[1]);
        }
#pragma unroll
        for (int np = 0; np < ND / 2; ++np) {
          uint32_t b[4];
          load_b_kn<LDH>(b, Qs, i0 + 16 * kc, np * 16, lane);
          mma_f16_16x8x16(qacc[2 * np], ca, b[0], b[1]);
          mma_f16_16x8x16(qacc[2 * np + 1], ca, b[2], b[3]);
        }
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qa[ks][0] = pack_h2(qacc[2 * ks][0] * qscale, qacc[2 * ks][1] * qscale);
        qa[ks][1] = pack_h2(qacc[2 * ks][2] * qscale, qacc[2 * ks][3] * qscale);
        qa[ks][2] = pack_h2(qacc[2 * ks + 1][0] * qscale, qacc[2 * ks + 1][1] * qscale);
        qa[ks][3] = pack_h2(qacc[2 * ks + 1][2] * qscale, qacc[2 * ks + 1][3] * qscale);
      }
    } else {
      const __half2 sc = __float2half2_rn(qscale);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        load_a_rowmajor<LDH>(qa[ks], Qs, i0, ks, lane);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __half2 v = __hmul2(*reinterpret_cast<__half2*>(&qa[ks][e]), sc);
          qa[ks][e] = *reinterpret_cast<uint32_t*>(&v);
        }
      }
    }
    // Q' rows to shared memory (B operand of dK = dS^T Q', A operand of phase 1B)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      *reinterpret_cast<uint32_t*>(Q2 + (size_t)(i0 + g) * LDH + ks * 16 + 2 * t) = qa[ks][0];
      *reinterpret_cast<uint32_t*>(Q2 + (size_t)(i0 + g + 8) * LDH + ks * 16 + 2 * t) = qa[ks][1];
      *reinterpret_cast<uint32_t*>(Q2 + (size_t)(i0 + g) * LDH + ks * 16 + 8 + 2 * t) = qa[ks][2];
      *reinterpret_cast<uint32_t*>(Q2 + (size_t)(i0 + g + 8) * LDH + ks * 16 + 8 + 2 * t) = qa[ks][3];
    }
    // softmax statistics of the 16 rows over all keys
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    for (int k0 = 0; k0 < P; k0 += 32) {
      float s[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b[4];
          load_b_nk<LDH>(b, Ks, k0 + np * 16, ks, lane);
          mma_f16_16x8x16(s[2 * np], qa[ks], b[0], b[1]);
          mma_f16_16x8x16(s[2 * np + 1], qa[ks], b[2], b[3]);
        }
      float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (k0 + nt * 8 + 2 * t + (e & 1) >= P) s[nt][e] = -INFINITY;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        l_run[hh] *= fast_exp2(m_run[hh] - mx[hh]);
        m_run[hh] = mx[hh];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) l_run[e >> 1] += fast_exp2(s[nt][e] - mx[e >> 1]);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
      l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
      const int row = i0 + g + hh * 8;
      float dsum = 0.f;
      if (row < P) {
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
          float2 ov = unpack_h2(__ldg(reinterpret_cast<const uint32_t*>(obase + (size_t)row * D + nd * 8 + 2 * t)));
          float2 gv = unpack_h2(*reinterpret_cast<const uint32_t*>(Gs + (size_t)row * LDH + nd * 8 + 2 * t));
          dsum = fmaf(ov.x, gv.x, dsum);
          dsum = fmaf(ov.y, gv.y, dsum);
        }
      }
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
      dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
      if (t == 0) {
        m2s[row] = m_run[hh];
        lis[row] = 1.f / l_run[hh];
        Drs[row] = dsum;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: key rows -> dK, dV -----------------------------------------------------------------
#pragma unroll 1
  for (int rb = 0; rb < RB; ++rb) {
    const int j0 = 16 * (warp + rb * W);
    if (j0 >= P) continue;
    uint32_t ka[KS][4], va[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      load_a_rowmajor<LDH>(ka[ks], Ks, j0, ks, lane);
      load_a_rowmajor<LDH>(va[ks], Vs, j0, ks, lane);
    }
    float dk[ND][4], dv[ND][4];
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dk[i][e] = 0.f; dv[i][e] = 0.f; }
    for (int q0 = 0; q0 < P; q0 += 32) {
      float st[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { st[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b[4];
          load_b_nk<LDH>(b, Q2, q0 + np * 16, ks, lane);
          mma_f16_16x8x16(st[2 * np], ka[ks], b[0], b[1]);
          mma_f16_16x8x16(st[2 * np + 1], ka[ks], b[2], b[3]);
          load_b_nk<LDH>(b, Gs, q0 + np * 16, ks, lane);
          mma_f16_16x8x16(dp[2 * np], va[ks], b[0], b[1]);
          mma_f16_16x8x16(dp[2 * np + 1], va[ks], b[2], b[3]);
        }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = q0 + nt * 8 + 2 * t + (e & 1);  // query index
          float pv = 0.f, dsv = 0.f;
          if (col < P) {
            pv = fast_exp2(st[nt][e] - m2s[col]) * lis[col];
            dsv = pv * (dp[nt][e] - Drs[col]);
          }
          st[nt][e] = pv;
          dp[nt][e] = dsv;
        }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t pa[4] = {pack_h2(st[2 * j][0], st[2 * j][1]), pack_h2(st[2 * j][2], st[2 * j][3]),
                          pack_h2(st[2 * j + 1][0], st[2 * j + 1][1]),
                          pack_h2(st[2 * j + 1][2], st[2 * j + 1][3])};
        uint32_t da[4] = {pack_h2(dp[2 * j][0], dp[2 * j][1]), pack_h2(dp[2 * j][2], dp[2 * j][3]),
                          pack_h2(dp[2 * j + 1][0], dp[2 * j + 1][1]),
                          pack_h2(dp[2 * j + 1][2], dp[2 * j + 1][3])};
#pragma unroll
        for (int np = 0; np < ND / 2; ++np) {
          uint32_t b[4];
          load_b_kn<LDH>(b, Gs, q0 + j * 16, np * 16, lane);
          mma_f16_16x8x16(dv[2 * np], pa, b[0], b[1]);
          mma_f16_16x8x16(dv[2 * np + 1], pa, b[2], b[3]);
          load_b_kn<LDH>(b, Q2, q0 + j * 16, np * 16, lane);
          mma_f16_16x8x16(dk[2 * np], da, b[0], b[1]);
          mma_f16_16x8x16(dk[2 * np + 1], da, b[2], b[3]);
        }
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int key = j0 + g + hh * 8;
      if (key >= P) continue;
      __half* krow = dqkv + ((size_t)rho * P + key) * ld + D + h * HD + 2 * t;
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        // Q' carries scale*log2e: dK = dS^T Q' / log2e
        *reinterpret_cast<uint32_t*>(krow + nd * 8) =
            pack_h2(dk[nd][hh * 2] * kLn2, dk[nd][hh * 2 + 1] * kLn2);
        *reinterpret_cast<uint32_t*>(krow + D + nd * 8) = pack_h2(dv[nd][hh * 2], dv[nd][hh * 2 + 1]);
      }
    }
  }

  // ---- phase 1B: query rows -> dQe ----------------------------------------------------------------
  float dq[RB][ND][4];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    const int i0 = 16 * (warp + rb * W);
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[rb][i][e] = 0.f;
    if (i0 >= P) continue;
    uint32_t qa[KS][4], ga[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      load_a_rowmajor<LDH>(qa[ks], Q2, i0, ks, lane);
      load_a_rowmajor<LDH>(ga[ks], Gs, i0, ks, lane);
    }
    const float mr[2] = {m2s[i0 + g], m2s[i0 + g + 8]};
    const float lr[2] = {lis[i0 + g], lis[i0 + g + 8]};
    const float dr[2] = {Drs[i0 + g], Drs[i0 + g + 8]};
    for (int k0 = 0; k0 < P; k0 += 32) {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b[4];
          load_b_nk<LDH>(b, Ks, k0 + np * 16, ks, lane);
          mma_f16_16x8x16(s[2 * np], qa[ks], b[0], b[1]);
          mma_f16_16x8x16(s[2 * np + 1], qa[ks], b[2], b[3]);
          load_b_nk<LDH>(b, Vs, k0 + np * 16, ks, lane);
          mma_f16_16x8x16(dp[2 * np], ga[ks], b[0], b[1]);
          mma_f16_16x8x16(dp[2 * np + 1], ga[ks], b[2], b[3]);
        }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = k0 + nt * 8 + 2 * t + (e & 1);  // key index
          float dsv = 0.f;
          if (col < P) dsv = fast_exp2(s[nt][e] - mr[e >> 1]) * lr[e >> 1] * (dp[nt][e] - dr[e >> 1]);
          dp[nt][e] = dsv;
        }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t da[4] = {pack_h2(dp[2 * j][0], dp[2 * j][1]), pack_h2(dp[2 * j][2], dp[2 * j][3]),
                          pack_h2(dp[2 * j + 1][0], dp[2 * j + 1][1]),
                          pack_h2(dp[2 * j + 1][2], dp[2 * j + 1][3])};
#pragma unroll
        for (int np = 0; np < ND / 2; ++np) {
          uint32_t b[4];
          load_b_kn<LDH>(b, Ks, k0 + j * 16, np * 16, lane);
          mma_f16_16x8x16(dq[rb][2 * np], da, b[0], b[1]);
          mma_f16_16x8x16(dq[rb][2 * np + 1], da, b[2], b[3]);
        }
      }
    }
  }

  if (!taps) {  // no EPEG: dQ = dQe
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int i0 = 16 * (warp + rb * W);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = i0 + g + hh * 8;
        if (row >= P) continue;
        __half* qrow = dqkv + ((size_t)rho * P + row) * ld + h * HD + 2 * t;
#pragma unroll
        for (int nd = 0; nd < ND; ++nd)
          *reinterpret_cast<uint32_t*>(qrow + nd * 8) =
              pack_h2(dq[rb][nd][hh * 2] * scale, dq[rb][nd][hh * 2 + 1] * scale);
      }
    }
    return;
  }

  // ---- phase C: EPEG transpose ----------------------------------------------------------------------
  __syncthreads();  // every warp is done with Q' / K / V / dO
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    const int i0 = 16 * (warp + rb * W);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = i0 + g + hh * 8;
      const bool ok = row < P;
#pragma unroll
      for (int nd = 0; nd < ND; ++nd)
        *reinterpret_cast<uint32_t*>(Q2 + (size_t)row * LDH + nd * 8 + 2 * t) =
            ok ? pack_h2(dq[rb][nd][hh * 2] * scale, dq[rb][nd][hh * 2 + 1] * scale) : 0u;
    }
  }
  __syncthreads();
  for (int item = tid; item < P * C8; item += blockDim.x) {
    const int m = item / C8, c = (item - m * C8) * 8;
    float acc[8];
    {
      uint4 u = *reinterpret_cast<const uint4*>(Q2 + (size_t)m * LDH + c);
      float2 a = unpack_h2(u.x), b = unpack_h2(u.y), cc = unpack_h2(u.z), d = unpack_h2(u.w);
      acc[0] = a.x; acc[1] = a.y; acc[2] = b.x; acc[3] = b.y;
      acc[4] = cc.x; acc[5] = cc.y; acc[6] = d.x; acc[7] = d.y;
    }
    for (int d = 0; d < epeg_k; ++d) {
      const int src = m - d + pad;  // Qe[src] read Q[src + d - pad] = Q[m]
      if (src < 0 || src >= P) continue;
      const float w = Ts[d];
      uint4 u = *reinterpret_cast<const uint4*>(Q2 + (size_t)src * LDH + c);
      float2 a = unpack_h2(u.x), b = unpack_h2(u.y), cc = unpack_h2(u.z), e = unpack_h2(u.w);
      acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
      acc[2] = fmaf(w, b.x, acc[2]); acc[3] = fmaf(w, b.y, acc[3]);
      acc[4] = fmaf(w, cc.x, acc[4]); acc[5] = fmaf(w, cc.y, acc[5]);
      acc[6] = fmaf(w, e.x, acc[6]); acc[7] = fmaf(w, e.y, acc[7]);
    }
    *reinterpret_cast<uint4*>(dqkv + ((size_t)rho * P + m) * ld + h * HD + c) =
        make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]),
                   pack_h2(acc[6], acc[7]));
  }
  if (dtaps) {
    for (int d = 0; d < epeg_k; ++d) {
      float part = 0.f;
      for (int item = tid; item < P * C8; item += blockDim.x) {
        const int i = item / C8, c = (item - i * C8) * 8;
        uint4 u = *reinterpret_cast<const uint4*>(Q2 + (size_t)i * LDH + c);
        uint4 v = *reinterpret_cast<const uint4*>(Qs + (size_t)(i + d) * LDH + c);  // Q[i + d - pad]
        float2 a0 = unpack_h2(u.x), a1 = unpack_h2(u.y), a2 = unpack_h2(u.z), a3 = unpack_h2(u.w);
        float2 b0 = unpack_h2(v.x), b1 = unpack_h2(v.y), b2 = unpack_h2(v.z), b3 = unpack_h2(v.w);
        part += (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) +
                (a2.x * b2.x + a2.y * b2.y) + (a3.x * b3.x + a3.y * b3.y);
      }
      part = warp_sum(part);
      if (lane == 0) atomicAdd(&red[d], part);
    }
    __syncthreads();
    if (tid < epeg_k) {
      const float inv = amax ? grad_inv_scale(__ldg(amax)) : 1.f;
      atomicAdd(dtaps + h * epeg_k + tid, red[tid] * inv);
    }
  }
}

template <int HD, int RB>
cudaError_t launch(const __half* qkv, const __half* o, const __half* dO, const float* taps,
                   __half* dqkv, float* dtaps, const uint32_t* amax, int R, int P, int D, int heads,
                   int epeg_k, cudaStream_t stream) {
  const int PR = (P + 31) / 32 * 32, nblk = PR / 16;
  if (nblk % RB) return cudaErrorInvalidValue;
  const int W = nblk / RB;
  const int pad = taps ? epeg_k / 2 : 0;
  const int nkc = taps ? (16 + epeg_k - 1 + 15) / 16 : 1;
  int q_rows = PR - 16 + 16 * nkc;                  // every warp's Toeplitz band exists
  if (q_rows < PR + 2 * pad) q_rows = PR + 2 * pad;  // and the tap-gradient reads Q[i + d - pad]
  size_t smem = ((size_t)q_rows + 4 * (size_t)PR) * (HD + 8) * sizeof(__half) +
                (3 * (size_t)PR + 128) * sizeof(float) + 16;
  if (smem > 227 * 1024 || W * 32 > (RB == 1 ? 320 : 256)) return cudaErrorInvalidValue;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rmsa_attn_bwd_kernel<HD, RB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  dim3 g(R, heads);
  rmsa_attn_bwd_kernel<HD, RB><<<g, 32 * W, smem, stream>>>(qkv, o, dO, taps, dqkv, dtaps, amax, P, D,
                                                            epeg_k, 1.f / sqrtf((float)HD), PR, q_rows);
  return cudaGetLastError();
}

}  // namespace

bool rmsa_attention_bwd_supported(int P, int D, int heads, int epeg_k) {
  int hd = heads > 0 ? D / heads : 0;
  return P >= 1 && P <= 256 && (hd == 32 || hd == 64) && heads <= 65535 && epeg_k <= 63;
}

cudaError_t launch_rmsa_attention_bwd(const __half* qkv, const __half* o, const __half* dO,
                                      const float* taps, __half* dqkv, float* dtaps,
                                      const uint32_t* amax, int R, int P, int D, int heads, int epeg_k,
                                      cudaStream_t stream) {
  if (!rmsa_attention_bwd_supported(P, D, heads, epeg_k)) return cudaErrorInvalidValue;
  const int nblk = (P + 31) / 32 * 2;
  if (D / heads == 32)
    return nblk <= 10 ? launch<32, 1>(qkv, o, dO, taps, dqkv, dtaps, amax, R, P, D, heads, epeg_k, stream)
                      : launch<32, 2>(qkv, o, dO, taps, dqkv, dtaps, amax, R, P, D, heads, epeg_k, stream);
  return nblk <= 10 ? launch<64, 1>(qkv, o, dO, taps, dqkv, dtaps, amax, R, P, D, heads, epeg_k, stream)
                    : launch<64, 2>(qkv, o, dO, taps, dqkv, dtaps, amax, R, P, D, heads, epeg_k, stream);
}

}  // namespace rrt
