// CR-MSA (modules/rmsa.py:290-337) without the [R,k,P,D] tensors the reference materialises:
//   stats/logits  : per slot LayerNorm statistics + logits = LN(x1) . phi          (streaming)
//   combine       : per region softmax_P / min / max of the logits, landmarks = cw . LN(x1)
//   landmark attn : MHA core over the 64 regions' landmarks (batch = k)
//   dispatch      : out = LN_final(x1 + (dmm*dw)^T . landmarks' (+ x0))             (streaming)
// The streaming kernels are HBM/L2-bound: one warp per token row, float4 accesses.
#include "kernels.cuh"

namespace rrt {
namespace {
__device__ long long* g_attn_trace_dev = nullptr;  // debug stamps (shared with tools/attn_trace.py)

// ------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) crmsa_stats_logits_kernel(
    const float* __restrict__ x1, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ phi, float2* __restrict__ stats, float* __restrict__ logits,
    Grid grid, int k) {
  constexpr int D = 128 * V;
  extern __shared__ __align__(16) float phi_t[];  // [k][D]: phi transposed so lanes read float4 runs
  if (phi) {
    for (int i = threadIdx.x; i < D * k; i += blockDim.x) {
      int c = i / k, n = i - c * k;
      phi_t[n * D + c] = __ldg(phi + i);
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float4 gm[V], bt[V];
  if (phi) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      bt[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    }
  }
  for (int slot = blockIdx.x * wpb + (threadIdx.x >> 5); slot < grid.Np; slot += gridDim.x * wpb) {
    int tok = grid.slot_to_token(slot);
    if (tok >= grid.L) {  // zero pad token: LN output forced to 0 -> logits 0
      if (lane == 0) stats[slot] = make_float2(0.f, 0.f);
      if (phi && lane < k) logits[(size_t)slot * k + lane] = 0.f;
      continue;
    }
    const float* xrow = x1 + (size_t)tok * D;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = __ldg(reinterpret_cast<const float4*>(xrow) + lane + 32 * i);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    float rstd = rsqrtf(warp_sum(q) * (1.f / D) + kLnEps);
    if (lane == 0) stats[slot] = make_float2(mean, rstd);
    if (!phi) continue;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x = (v[i].x - mean) * rstd * gm[i].x + bt[i].x;
      v[i].y = (v[i].y - mean) * rstd * gm[i].y + bt[i].y;
      v[i].z = (v[i].z - mean) * rstd * gm[i].z + bt[i].z;
      v[i].w = (v[i].w - mean) * rstd * gm[i].w + bt[i].w;
    }
    float mine = 0.f;  // lane n keeps logit n
    for (int n = 0; n < k; ++n) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float4 ph = *reinterpret_cast<const float4*>(phi_t + n * D + 4 * (lane + 32 * i));
        d = fmaf(v[i].x, ph.x, d); d = fmaf(v[i].y, ph.y, d);
        d = fmaf(v[i].z, ph.z, d); d = fmaf(v[i].w, ph.w, d);
      }
      d = warp_sum(d);
      if (lane == n) mine = d;
    }
    if (lane < k) logits[(size_t)slot * k + lane] = mine;
  }
}

__global__ void __launch_bounds__(256) crmsa_mlp_logits_kernel(const float* __restrict__ hidden,
                                                               const float* __restrict__ w2,
                                                               float* __restrict__ logits, int Np,
                                                               int Dh, int k) {
  int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (slot >= Np) return;
  const float* hrow = hidden + (size_t)slot * Dh;
  for (int n = 0; n < k; ++n) {
    float d = 0.f;
    for (int c = lane; c < Dh; c += 32) d = fmaf(__ldg(hrow + c), __ldg(w2 + (size_t)n * Dh + c), d);
    d = warp_sum(d);
    if (lane == 0) logits[(size_t)slot * k + n] = d;
  }
}

// ------------------------------------------------------------------------------------------
// grid (D/128, R), 256 threads.  smem: cw[P*k] | part[8][k][128]
template <int KMAX>
__global__ void __launch_bounds__(256) crmsa_combine_kernel(
    const float* __restrict__ x1, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float2* __restrict__ stats, const float* __restrict__ logits,
    __half* __restrict__ landmarks, float2* __restrict__ rstat, Grid grid, int D, int k) {
  extern __shared__ __align__(16) float smem[];
  const int P = grid.P, rho = blockIdx.y, chunk = blockIdx.x;
  float* cw = smem;                                   // [P][k]
  float* part = smem + (((size_t)P * k + 3) & ~(size_t)3);  // [8][k][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const float* lg = logits + (size_t)rho * P * k;
  for (int i = tid; i < P * k; i += blockDim.x) cw[i] = __ldg(lg + i);
  __syncthreads();
  // per landmark n: min / max / sum of exp over the P tokens of the region, then normalise
  for (int n = warp; n < k; n += 8) {
    float mx = -INFINITY, mn = INFINITY;
    for (int p = lane; p < P; p += 32) {
      float v = cw[p * k + n];
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) sum += __expf(cw[p * k + n] - mx);
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int p = lane; p < P; p += 32) cw[p * k + n] = __expf(cw[p * k + n] - mx) * inv;
    if (chunk == 0 && lane == 0) rstat[(size_t)rho * k + n] = make_float2(mn, mx);
  }
  __syncthreads();

  const int c0 = chunk * 128 + lane * 4;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c0));
  const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + c0));
  float4 acc[KMAX];
#pragma unroll
  for (int n = 0; n < KMAX; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);

  constexpr int U = KMAX <= 4 ? 8 : 4;  // rows in flight per warp
  for (int p0 = warp; p0 < P; p0 += 8 * U) {
    float4 xv[U];
    float2 st[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int p = p0 + 8 * u;
      st[u] = make_float2(0.f, 0.f);
      xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < P) {
        int slot = rho * P + p;
        st[u] = __ldg(stats + slot);
        if (st[u].y != 0.f) {
          int tok = grid.slot_to_token(slot);
          xv[u] = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)tok * D + c0));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int p = p0 + 8 * u;
      if (p >= P || st[u].y == 0.f) continue;  // pad rows are exact zeros after the norm
      float4 z;
      z.x = (xv[u].x - st[u].x) * st[u].y * gm.x + bt.x;
      z.y = (xv[u].y - st[u].x) * st[u].y * gm.y + bt.y;
      z.z = (xv[u].z - st[u].x) * st[u].y * gm.z + bt.z;
      z.w = (xv[u].w - st[u].x) * st[u].y * gm.w + bt.w;
#pragma unroll
      for (int n = 0; n < KMAX; ++n) {
        if (n < k) {
          float wgt = cw[p * k + n];
          acc[n].x = fmaf(wgt, z.x, acc[n].x);
          acc[n].y = fmaf(wgt, z.y, acc[n].y);
          acc[n].z = fmaf(wgt, z.z, acc[n].z);
          acc[n].w = fmaf(wgt, z.w, acc[n].w);
        }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < KMAX; ++n)
    if (n < k) *reinterpret_cast<float4*>(part + ((size_t)(warp * k + n)) * 128 + lane * 4) = acc[n];
  __syncthreads();
  for (int i = tid; i < k * 128; i += blockDim.x) {
    int n = i >> 7, c = i & 127;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[((size_t)(w * k + n)) * 128 + c];
    landmarks[((size_t)n * grid.R + rho) * D + chunk * 128 + c] = __float2half_rn(s);
  }
}

// ------------------------------------------------------------------------------------------
// CR-MSA front end, split in two all-SM kernels (the profile of the one-CTA-per-region kernel below:
// 6 M warp-instructions, FFMA only 20 % of them, 64 SMs busy -> instruction-bound).
//
// (1) crmsa_rowstats_kernel: warp per padded slot, grid-stride.  LayerNorm statistics and the k
//     logits from ONE batched butterfly:  G[c,n] = gamma[c]*phi[c,n], A[n] = sum_c G[c,n],
//     B[n] = sum_c beta[c]*phi[c,n]  ->  logits[n] = rstd*(x.G[:,n] - mean*A[n]) + B[n].
//     Writes logits [Np,k] and stats [Np] = (mean, rstd); rstd = 0 marks a zero pad slot.
// (2) crmsa_combine2_kernel: CTA per (region, 128-column chunk).  With w'[p,n] = cw[p,n]*rstd_p,
//     S1[n] = sum_p w'[p,n]*mean_p, S0[n] = sum_{real p} cw[p,n]:
//       landmarks[n,c] = gamma[c]*(sum_p w'[p,n]*x1[p,c] - S1[n]) + beta[c]*S0[n]
//     so the inner loop is one 16-byte load and k FMAs per 4 elements -- no per-element LayerNorm.
template <int V, int KMAX>
__global__ void __launch_bounds__(256) crmsa_rowstats_kernel(
    const float* __restrict__ x1, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ phi, float2* __restrict__ stats, float* __restrict__ logits, Grid grid,
    int k) {
  constexpr int D = 128 * V;
  extern __shared__ __align__(16) float smem[];
  float* Gt = smem;            // [KMAX][D]
  float* AB = Gt + KMAX * D;   // [2][KMAX]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wpb = blockDim.x >> 5;
  pdl_launch_dependents();
  if (phi) {  // weights only: runs ahead of the predecessor's completion
    for (int n = warp; n < k; n += wpb) {
      float a = 0.f, b = 0.f;
      for (int c = lane; c < D; c += 32) {
        float ph = __ldg(phi + (size_t)c * k + n);
        float gph = __ldg(gamma + c) * ph;
        Gt[n * D + c] = gph;
        a += gph;
        b = fmaf(__ldg(beta + c), ph, b);
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (lane == 0) { AB[n] = a; AB[KMAX + n] = b; }
    }
    __syncthreads();
  }
  pdl_wait();
  constexpr int U = 2;  // rows in flight per warp
  const int stride = gridDim.x * wpb * U;
  for (int s0 = (blockIdx.x * wpb + warp) * U; s0 < grid.Np; s0 += stride) {
    float4 v[U][V];
    int tk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int slot = s0 + u;
      int t = slot < grid.Np ? grid.slot_to_token(slot) : grid.L;
      tk[u] = t < grid.L ? t : -1;
#pragma unroll
      for (int i = 0; i < V; ++i)
        v[u][i] = tk[u] >= 0
                      ? __ldg(reinterpret_cast<const float4*>(x1 + (size_t)tk[u] * D) + lane + 32 * i)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float red[U][1 + KMAX];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float sacc = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) sacc += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
      red[u][0] = sacc;
#pragma unroll
      for (int n = 0; n < KMAX; ++n) red[u][1 + n] = 0.f;
    }
    if (phi) {
#pragma unroll
      for (int n = 0; n < KMAX; ++n) {
        if (n < k) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float4 gq = *reinterpret_cast<const float4*>(Gt + n * D + 4 * (lane + 32 * i));
#pragma unroll
            for (int u = 0; u < U; ++u) {
              float d = red[u][1 + n];
              d = fmaf(v[u][i].x, gq.x, d); d = fmaf(v[u][i].y, gq.y, d);
              d = fmaf(v[u][i].z, gq.z, d); d = fmaf(v[u][i].w, gq.w, d);
              red[u][1 + n] = d;
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < 1 + KMAX; ++j) red[u][j] += __shfl_xor_sync(0xffffffffu, red[u][j], o);
    float q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float mean = red[u][0] * (1.f / D);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float a = v[u][i].x - mean, b = v[u][i].y - mean, c = v[u][i].z - mean, d = v[u][i].w - mean;
        acc += (a * a + b * b) + (c * c + d * d);
      }
      q[u] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int slot = s0 + u;
      if (slot >= grid.Np) continue;
      float mean = red[u][0] * (1.f / D);
      float rstd = tk[u] >= 0 ? rsqrtf(q[u] * (1.f / D) + kLnEps) : 0.f;
      if (lane == 0) stats[slot] = make_float2(tk[u] >= 0 ? mean : 0.f, rstd);
      if (phi) {
        float mine = 0.f;
#pragma unroll
        for (int n = 0; n < KMAX; ++n)
          if (lane == n) mine = tk[u] >= 0 ? rstd * (red[u][1 + n] - mean * AB[n]) + AB[KMAX + n] : 0.f;
        if (lane < k) logits[(size_t)slot * k + lane] = mine;
      }
    }
  }
}

// grid (D/128, R), 256 threads.  smem: wq[P][KMAX] | tok[P] | part[8][KMAX][128] | s01[2][KMAX]
// FROM_PARTS: the statistics and logits do not exist yet; the projection GEMM that wrote x1 left, per token and
// 128-column part, [sum x, sum x^2, sum_c x_c gamma_c phi[c,n]] (GemmEpilogue::rs_part).  Every CTA of a region
// finishes mean / rstd / logits of its rows from those records (chunk 0 also publishes them in stats_out /
// logits_out for the dispatch kernel), so x1 is read once by the CR-MSA front instead of twice.
template <int KMAX, bool FROM_PARTS>
__global__ void __launch_bounds__(256) crmsa_combine2_kernel(
    const float* __restrict__ x1, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float2* __restrict__ stats, const float* __restrict__ logits, __half* __restrict__ landmarks,
    float2* __restrict__ rstat, Grid grid, int D, int k, const float* __restrict__ rs_part, int rs_parts,
    const float* __restrict__ phi, float2* __restrict__ stats_out, float* __restrict__ logits_out) {
  extern __shared__ __align__(16) float smem[];
  const int P = grid.P, rho = blockIdx.y, chunk = blockIdx.x;
  float* wq = smem;                                          // [P][KMAX]: logits -> cw*rstd
  int* tok = reinterpret_cast<int*>(wq + (size_t)P * KMAX);  // [P]
  float* part = reinterpret_cast<float*>(tok + ((P + 3) & ~3));  // [8][KMAX][128]
  float* s01 = part + 8 * KMAX * 128;                        // [2][KMAX]: S0, S1
  float2* sst = reinterpret_cast<float2*>(s01 + 2 * KMAX);   // [P] (mean, rstd); rstd = 0 marks a pad slot
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  if (FROM_PARTS) {
    // weights only (runs ahead of the predecessor): A[n] = sum_c gamma_c phi[c,n], B[n] = sum_c beta_c phi[c,n],
    // parked in s01 until the softmax below overwrites it
    for (int n = warp; n < k; n += 8) {
      float a = 0.f, b = 0.f;
      for (int c = lane; c < D; c += 32) {
        const float ph = __ldg(phi + (size_t)c * k + n);
        a = fmaf(__ldg(gamma + c), ph, a);
        b = fmaf(__ldg(beta + c), ph, b);
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (lane == 0) { s01[n] = a; s01[KMAX + n] = b; }
    }
    __syncthreads();
  }
  pdl_wait();

  for (int p = tid; p < P; p += 256) {
    int slot = rho * P + p;
    if (FROM_PARTS) {
      const int t = grid.slot_to_token(slot);
      float2 st = make_float2(0.f, 0.f);
      float lg[KMAX];
#pragma unroll
      for (int n = 0; n < KMAX; ++n) lg[n] = 0.f;
      if (t < grid.L) {
        float sum = 0.f, sq = 0.f, dot[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < rs_parts; ++q) {
          const float4* rec = reinterpret_cast<const float4*>(rs_part + ((size_t)t * rs_parts + q) * 8);
          const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1);
          sum += r0.x; sq += r0.y;
          dot[0] += r0.z; dot[1] += r0.w; dot[2] += r1.x; dot[3] += r1.y;
        }
        const float mean = sum / D;
        const float var = fmaxf(sq / D - mean * mean, 0.f);
        st = make_float2(mean, rsqrtf(var + kLnEps));
#pragma unroll
        for (int n = 0; n < KMAX; ++n)
          if (n < k && n < 4) lg[n] = st.y * (dot[n] - mean * s01[n]) + s01[KMAX + n];
      }
      sst[p] = st;
      tok[p] = t < grid.L ? t : -1;
#pragma unroll
      for (int n = 0; n < KMAX; ++n) wq[p * KMAX + n] = lg[n];
      if (chunk == 0) {
        stats_out[slot] = st;
        for (int n = 0; n < k; ++n) logits_out[(size_t)slot * k + n] = lg[n];
      }
    } else {
      float2 st = __ldg(stats + slot);
      sst[p] = st;
      tok[p] = st.y != 0.f ? grid.slot_to_token(slot) : -1;
#pragma unroll
      for (int n = 0; n < KMAX; ++n) wq[p * KMAX + n] = n < k ? __ldg(logits + (size_t)slot * k + n) : 0.f;
    }
  }
  __syncthreads();
  // per landmark: softmax over the region, min, max; then w' = cw * rstd, S0, S1
  for (int n = warp; n < k; n += 8) {
    float mx = -INFINITY, mn = INFINITY;
    for (int p = lane; p < P; p += 32) {
      float v = wq[p * KMAX + n];
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) sum += __expf(wq[p * KMAX + n] - mx);
    float inv = 1.f / warp_sum(sum);
    float s0 = 0.f, s1 = 0.f;
    for (int p = lane; p < P; p += 32) {
      float cw = __expf(wq[p * KMAX + n] - mx) * inv;
      float2 st = sst[p];
      float w = cw * st.y;  // rstd = 0 for pad rows: they contribute nothing
      wq[p * KMAX + n] = w;
      if (st.y != 0.f) s0 += cw;
      s1 = fmaf(w, st.x, s1);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
      s01[n] = s0;
      s01[KMAX + n] = s1;
      if (chunk == 0) rstat[(size_t)rho * k + n] = make_float2(mn, mx);
    }
  }
  __syncthreads();

  const int c0 = chunk * 128 + lane * 4;
  float4 acc[KMAX];
#pragma unroll
  for (int n = 0; n < KMAX; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = 6;  // rows in flight per warp
  for (int p0 = warp; p0 < P; p0 += 8 * U) {
    float4 xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int p = p0 + 8 * u;
      int t = p < P ? tok[p] : -1;
      xv[u] = t >= 0 ? __ldg(reinterpret_cast<const float4*>(x1 + (size_t)t * D + c0))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int p = p0 + 8 * u;
      if (p >= P) break;
      const float4 wv = *reinterpret_cast<const float4*>(wq + p * KMAX);  // KMAX = 4: one LDS.128
#pragma unroll
      for (int n = 0; n < KMAX; ++n) {
        float wgt = KMAX == 4 ? (n == 0 ? wv.x : n == 1 ? wv.y : n == 2 ? wv.z : wv.w) : wq[p * KMAX + n];
        acc[n].x = fmaf(wgt, xv[u].x, acc[n].x); acc[n].y = fmaf(wgt, xv[u].y, acc[n].y);
        acc[n].z = fmaf(wgt, xv[u].z, acc[n].z); acc[n].w = fmaf(wgt, xv[u].w, acc[n].w);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < KMAX; ++n)
    *reinterpret_cast<float4*>(part + ((size_t)(warp * KMAX + n)) * 128 + lane * 4) = acc[n];
  __syncthreads();
  for (int i = tid; i < k * 128; i += 256) {
    int n = i >> 7, c = i & 127;
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sacc += part[((size_t)(w * KMAX + n)) * 128 + c];
    int gc = chunk * 128 + c;
    float val = __ldg(gamma + gc) * (sacc - s01[KMAX + n]) + __ldg(beta + gc) * s01[n];
    landmarks[((size_t)n * grid.R + rho) * D + gc] = __float2half_rn(val);
  }
}

// ------------------------------------------------------------------------------------------
// Fused CR-MSA front end, one CTA per region (16 warps), two passes over the region's rows (which
// are L2-resident: the projection GEMM has just written them):
//   pass 1  warp per row, 3 rows in flight: LayerNorm statistics and the k logits from ONE batched
//           butterfly: with G[c,n] = gamma[c]*phi[c,n], A[n] = sum_c G[c,n], B[n] = sum_c beta[c]*phi[c,n]
//             logits[n] = rstd * (sum_c x[c]*G[c,n] - mean*A[n]) + B[n]      (== LN(x) . phi[:, n])
//           so sum(x) and the k dot products reduce together; sum((x-mean)^2) is the second round.
//           (phi == null: the logits come from the crmsa_mlp path and are only read)
//   mid     warp per landmark: softmax over the P tokens, min / max  -> cw[p, n], rstat[rho, n]
//   pass 2  warps tiled (row group x 128-column group): landmarks[n, :] += cw[p, n] * LN(x1)[p, :]
// Replaces the separate stats/logits and combine kernels: one launch, every load batched.
__device__ __forceinline__ void lstamp(int slot) {
  if (g_attn_trace_dev && blockIdx.x < 64 && threadIdx.x == 0) {
    long long t;
    if (slot == 0 || slot == 5) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));   // ns, comparable across SMs
      g_attn_trace_dev[blockIdx.x * 8 + (slot == 0 ? 6 : 7)] = t;
    }
    g_attn_trace_dev[blockIdx.x * 8 + slot] = clock64();
  }
}
template <int V, int KMAX>
__global__ void __launch_bounds__(512) crmsa_landmarks_kernel(
    const float* __restrict__ x1, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ phi, float* __restrict__ logits, __half* __restrict__ landmarks,
    float2* __restrict__ rstat, Grid grid, int k) {
  constexpr int D = 128 * V;
  constexpr int NW = 16;           // warps
  constexpr int CG = V;            // 128-column groups
  constexpr int RG = NW / CG;      // row groups of pass 2
  extern __shared__ __align__(16) float smem[];
  const int P = grid.P, rho = blockIdx.x;
  float* gam = smem;                       // [D]
  float* bet = gam + D;                    // [D]
  float* Gt = bet + D;                     // [k][D]  gamma[c] * phi[c, n]
  float* lg = Gt + (size_t)k * D;          // [P][k]  logits, then combine weights
  float2* st = reinterpret_cast<float2*>(lg + (((size_t)P * k + 3) & ~(size_t)3));  // [P] mean, rstd
  int* tok = reinterpret_cast<int*>(st + ((P + 1) & ~1));                            // [P]
  float* AB = reinterpret_cast<float*>(tok + ((P + 3) & ~3));                        // [2][KMAX]
  float* part = AB + 2 * 16;                                                         // [RG][k][D]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  lstamp(0);
  for (int i = tid; i < D; i += 512) { gam[i] = __ldg(gamma + i); bet[i] = __ldg(beta + i); }
  for (int p = tid; p < P; p += 512) {
    int t = grid.slot_to_token(rho * P + p);
    tok[p] = t < grid.L ? t : -1;
  }
  if (!phi)
    for (int i = tid; i < P * k; i += 512) lg[i] = __ldg(logits + (size_t)rho * P * k + i);
  __syncthreads();
  if (phi) {
    for (int n = warp; n < k; n += NW) {  // one warp per landmark column of phi
      float a = 0.f, b = 0.f;
      for (int c = lane; c < D; c += 32) {
        float ph = __ldg(phi + (size_t)c * k + n);
        float gph = gam[c] * ph;
        Gt[n * D + c] = gph;
        a += gph;
        b = fmaf(bet[c], ph, b);
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (lane == 0) { AB[n] = a; AB[16 + n] = b; }
    }
    __syncthreads();
  }

  lstamp(1);
  // ---- pass 1: statistics (+ logits) -----------------------------------------------------------
  constexpr int U1 = 3;  // rows in flight per warp
  for (int p0 = warp; p0 < P; p0 += NW * U1) {
    float4 v[U1][V];
    int tk[U1];
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      int p = p0 + u * NW;
      tk[u] = p < P ? tok[p] : -1;
#pragma unroll
      for (int i = 0; i < V; ++i)
        v[u][i] = tk[u] >= 0
                      ? __ldg(reinterpret_cast<const float4*>(x1 + (size_t)tk[u] * D) + lane + 32 * i)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // round 1: sum(x) and the k dot products with G, all rows of the batch in one butterfly
    float red[U1][1 + KMAX];
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
      red[u][0] = s;
#pragma unroll
      for (int n = 0; n < KMAX; ++n) red[u][1 + n] = 0.f;
    }
    if (phi) {
#pragma unroll
      for (int n = 0; n < KMAX; ++n) {
        if (n < k) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float4 gq = *reinterpret_cast<const float4*>(Gt + n * D + 4 * (lane + 32 * i));
#pragma unroll
            for (int u = 0; u < U1; ++u) {
              float d = red[u][1 + n];
              d = fmaf(v[u][i].x, gq.x, d); d = fmaf(v[u][i].y, gq.y, d);
              d = fmaf(v[u][i].z, gq.z, d); d = fmaf(v[u][i].w, gq.w, d);
              red[u][1 + n] = d;
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U1; ++u)
#pragma unroll
        for (int j = 0; j < 1 + KMAX; ++j)
          if (j == 0 || (phi && j - 1 < k)) red[u][j] += __shfl_xor_sync(0xffffffffu, red[u][j], o);
    // round 2: centred second moment
    float q[U1];
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      float mean = red[u][0] * (1.f / D);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float a = v[u][i].x - mean, b = v[u][i].y - mean, c = v[u][i].z - mean, d = v[u][i].w - mean;
        acc += (a * a + b * b) + (c * c + d * d);
      }
      q[u] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U1; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
#pragma unroll
    for (int u = 0; u < U1; ++u) {
      int p = p0 + u * NW;
      if (p >= P) continue;
      if (tk[u] < 0) {  // zero pad token: LN output forced to 0 -> logits 0
        if (lane == 0) st[p] = make_float2(0.f, 0.f);
        if (phi && lane < k) {
          lg[p * k + lane] = 0.f;
          logits[((size_t)rho * P + p) * k + lane] = 0.f;
        }
        continue;
      }
      float mean = red[u][0] * (1.f / D);
      float rstd = rsqrtf(q[u] * (1.f / D) + kLnEps);
      if (lane == 0) st[p] = make_float2(mean, rstd);
      if (phi) {
        float mine = 0.f;  // lane n keeps logit n
#pragma unroll
        for (int n = 0; n < KMAX; ++n)
          if (lane == n) mine = rstd * (red[u][1 + n] - mean * AB[n]) + AB[16 + n];
        if (lane < k) {
          lg[p * k + lane] = mine;
          logits[((size_t)rho * P + p) * k + lane] = mine;
        }
      }
    }
  }
  __syncthreads();
  lstamp(2);

  // ---- per landmark: softmax over the region's tokens, min, max --------------------------------
  for (int n = warp; n < k; n += NW) {
    float mx = -INFINITY, mn = INFINITY;
    for (int p = lane; p < P; p += 32) {
      float v = lg[p * k + n];
      mx = fmaxf(mx, v);
      mn = fminf(mn, v);
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) {
      float e = __expf(lg[p * k + n] - mx);
      lg[p * k + n] = e;
      sum += e;
    }
    float inv = 1.f / warp_sum(sum);
    for (int p = lane; p < P; p += 32) lg[p * k + n] *= inv;
    if (lane == 0) rstat[(size_t)rho * k + n] = make_float2(mn, mx);
  }
  __syncthreads();

  lstamp(3);
  // ---- pass 2: landmarks = cw^T . LN(x1) ---------------------------------------------------------
  {
    const int cg = warp % CG, rg = warp / CG;
    const int c0 = cg * 128 + lane * 4;
    const float4 gm = *reinterpret_cast<const float4*>(gam + c0);
    const float4 bt = *reinterpret_cast<const float4*>(bet + c0);
    float4 acc[KMAX];
#pragma unroll
    for (int n = 0; n < KMAX; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int U2 = 8;
    for (int p0 = rg; p0 < P; p0 += RG * U2) {
      float4 xv[U2];
      int pt[U2];
#pragma unroll
      for (int u = 0; u < U2; ++u) {
        int p = p0 + u * RG;
        pt[u] = p < P ? tok[p] : -1;
        xv[u] = pt[u] >= 0 ? __ldg(reinterpret_cast<const float4*>(x1 + (size_t)pt[u] * D + c0))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U2; ++u) {
        if (pt[u] < 0) continue;  // pad rows are exact zeros after the norm
        int p = p0 + u * RG;
        float2 ms = st[p];
        float4 z;
        z.x = (xv[u].x - ms.x) * ms.y * gm.x + bt.x;
        z.y = (xv[u].y - ms.x) * ms.y * gm.y + bt.y;
        z.z = (xv[u].z - ms.x) * ms.y * gm.z + bt.z;
        z.w = (xv[u].w - ms.x) * ms.y * gm.w + bt.w;
#pragma unroll
        for (int n = 0; n < KMAX; ++n) {
          if (n < k) {
            float wgt = lg[p * k + n];
            acc[n].x = fmaf(wgt, z.x, acc[n].x); acc[n].y = fmaf(wgt, z.y, acc[n].y);
            acc[n].z = fmaf(wgt, z.z, acc[n].z); acc[n].w = fmaf(wgt, z.w, acc[n].w);
          }
        }
      }
    }
#pragma unroll
    for (int n = 0; n < KMAX; ++n)
      if (n < k) *reinterpret_cast<float4*>(part + ((size_t)(rg * k + n)) * D + c0) = acc[n];
  }
  __syncthreads();
  lstamp(4);
  for (int i = tid; i < k * D; i += 512) {
    int n = i / D, c = i - n * D;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < RG; ++r) s += part[((size_t)(r * k + n)) * D + c];
    landmarks[((size_t)n * grid.R + rho) * D + c] = __float2half_rn(s);
  }
  lstamp(5);
}

// ------------------------------------------------------------------------------------------
// grid (heads, k), 256 threads; sequence length R = 64 (the CR-MSA grid is always 8x8 regions).
__global__ void __launch_bounds__(256) landmark_attn_kernel(const float* __restrict__ lqkv,
                                                            __half* __restrict__ lo, int D,
                                                            int heads, float scale) {
  constexpr int R = 64, CH = 32;
  __shared__ float qs[R][CH + 1];
  __shared__ float ks[R][CH + 1];
  __shared__ float ps[R][R + 1];
  const int h = blockIdx.x, n = blockIdx.y, dh = D / heads;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const size_t ld = 3 * (size_t)D;
  const float* base = lqkv + (size_t)n * R * ld + h * dh;

  float s[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
  for (int c0 = 0; c0 < dh; c0 += CH) {
    __syncthreads();
    for (int i = tid; i < R * CH; i += 256) {
      int r = i / CH, c = i - r * CH;
      qs[r][c] = __ldg(base + (size_t)r * ld + c0 + c);
      ks[r][c] = __ldg(base + (size_t)r * ld + D + c0 + c);
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < CH; ++c) {
      float qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { qv[i] = qs[ty * 4 + i][c]; kv[i] = ks[tx * 4 + i][c]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ps[ty * 4 + i][tx * 4 + j] = s[i][j] * scale;
  __syncthreads();
  {  // row softmax: warp w owns rows 8w..8w+7
    int warp = tid >> 5, lane = tid & 31;
    for (int r = warp * 8; r < warp * 8 + 8; ++r) {
      float a = ps[r][lane], b = ps[r][lane + 32];
      float mx = warp_max(fmaxf(a, b));
      a = __expf(a - mx);
      b = __expf(b - mx);
      float inv = 1.f / warp_sum(a + b);
      ps[r][lane] = a * inv;
      ps[r][lane + 32] = b * inv;
    }
  }
  // O[:, c0:c0+32] = P @ V[:, c0:c0+32]; V chunk staged in qs
  for (int c0 = 0; c0 < dh; c0 += CH) {
    __syncthreads();
    for (int i = tid; i < R * CH; i += 256) {
      int r = i / CH, c = i - r * CH;
      qs[r][c] = __ldg(base + (size_t)r * ld + 2 * D + c0 + c);
    }
    __syncthreads();
    float o[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = 0.f;
#pragma unroll 8
    for (int j = 0; j < R; ++j) {
      float v0 = qs[j][tx * 2], v1 = qs[j][tx * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float p = ps[ty * 4 + i][j];
        o[i][0] = fmaf(p, v0, o[i][0]);
        o[i][1] = fmaf(p, v1, o[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<uint32_t*>(lo + ((size_t)n * R + ty * 4 + i) * D + h * dh + c0 + tx * 2) =
          pack_h2(o[i][0], o[i][1]);
  }
}

// ------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) crmsa_dispatch_kernel(
    const float* __restrict__ x1, const float* __restrict__ x0, const float* __restrict__ logits,
    const float2* __restrict__ rstat, const float* __restrict__ lm, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ out, Grid grid, int k) {
  constexpr int D = 128 * V;
  pdl_launch_dependents();
  pdl_wait();
  int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (tok >= grid.L) return;
  int slot = grid.token_to_slot(tok);
  int rho = slot / grid.P;
  // the long-latency row loads go first; the (dependent, L2-resident) weight chain overlaps them
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i)
    v[i] = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)tok * D) + lane + 32 * i);
  float4 u0[V];
  if (x0) {
#pragma unroll
    for (int i = 0; i < V; ++i)
      u0[i] = __ldg(reinterpret_cast<const float4*>(x0 + (size_t)tok * D) + lane + 32 * i);
  }
  // dispatch weight of landmark n for this token: softmax over the k logits x min-max over the
  // region.  Lane n (< k) owns landmark n; the weight is broadcast with a shuffle when used.
  float lg = lane < k ? __ldg(logits + (size_t)slot * k + lane) : -INFINITY;
  float2 mm = lane < k ? __ldg(rstat + (size_t)rho * k + lane) : make_float2(0.f, 1.f);
  float mx = warp_max(lg);
  float ex = lane < k ? __expf(lg - mx) : 0.f;
  float inv = 1.f / warp_sum(ex);
  float my_w = lane < k ? ex * inv * ((lg - mm.x) / (mm.y - mm.x + 1e-8f)) : 0.f;
  if (x0) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x += u0[i].x; v[i].y += u0[i].y; v[i].z += u0[i].z; v[i].w += u0[i].w;
    }
  }
  for (int n = 0; n < k; ++n) {
    const float4* lrow = reinterpret_cast<const float4*>(lm + ((size_t)n * grid.R + rho) * D);
    float w = __shfl_sync(0xffffffffu, my_w, n);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 u = __ldg(lrow + lane + 32 * i);
      v[i].x = fmaf(w, u.x, v[i].x); v[i].y = fmaf(w, u.y, v[i].y);
      v[i].z = fmaf(w, u.z, v[i].z); v[i].w = fmaf(w, u.w, v[i].w);
    }
  }
  float* orow = out + (size_t)tok * D;
  if (!gamma) {
#pragma unroll
    for (int i = 0; i < V; ++i) reinterpret_cast<float4*>(orow)[lane + 32 * i] = v[i];
    return;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  float rstd = rsqrtf(warp_sum(q) * (1.f / D) + kLnEps);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
    o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
    o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
    o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
    reinterpret_cast<float4*>(orow)[lane + 32 * i] = o;
  }
}

#define RRT_DISPATCH_V(D, ...)                           \
  switch ((D) / 128) {                                   \
    case 1: { constexpr int V = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int V = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int V = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int V = 4; __VA_ARGS__; break; } \
    case 6: { constexpr int V = 6; __VA_ARGS__; break; } \
    case 8: { constexpr int V = 8; __VA_ARGS__; break; } \
    default: return cudaErrorInvalidValue;               \
  }
}  // namespace

cudaError_t launch_crmsa_stats_logits(const float* x1, const float* gamma, const float* beta,
                                      const float* phi, float2* stats, float* logits,
                                      const Grid& grid, int D, int k, cudaStream_t stream) {
  if (D % 128 || k > 32) return cudaErrorInvalidValue;
  int blocks = (grid.Np + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;  // grid-stride over rows: phi is staged once per block
  size_t smem = phi ? (size_t)D * k * sizeof(float) : 0;
  RRT_DISPATCH_V(D, {
    static DeviceOnce configured;  // per instantiation: the attribute call is slow (~25 us)
    if (configured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(crmsa_stats_logits_kernel<V>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
    }
    crmsa_stats_logits_kernel<V><<<blocks, 256, smem, stream>>>(x1, gamma, beta, phi, stats, logits, grid, k);
  });
  return cudaGetLastError();
}

cudaError_t launch_crmsa_mlp_logits(const float* hidden, const float* w2, float* logits, int Np,
                                    int Dh, int k, cudaStream_t stream) {
  crmsa_mlp_logits_kernel<<<(Np + 7) / 8, 256, 0, stream>>>(hidden, w2, logits, Np, Dh, k);
  return cudaGetLastError();
}

cudaError_t launch_crmsa_combine(const float* x1, const float* gamma, const float* beta,
                                 const float2* stats, const float* logits, __half* landmarks,
                                 float2* rstat, const Grid& grid, int D, int k,
                                 cudaStream_t stream) {
  if (D % 128 || k < 1 || k > 16) return cudaErrorInvalidValue;
  size_t smem = ((((size_t)grid.P * k + 3) & ~(size_t)3) + (size_t)8 * k * 128) * sizeof(float);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  dim3 g(D / 128, grid.R);
#define RRT_COMBINE(KM)                                                                          \
  {                                                                                              \
    static DeviceOnce configured;                                                              \
    if (configured.needed()) {                                                                           \
      cudaError_t e = cudaFuncSetAttribute(crmsa_combine_kernel<KM>,                             \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return e;                                                            \
    }                                                                                            \
    crmsa_combine_kernel<KM><<<g, 256, smem, stream>>>(x1, gamma, beta, stats, logits, landmarks, \
                                                       rstat, grid, D, k);                       \
  }
  if (k <= 4) RRT_COMBINE(4) else if (k <= 8) RRT_COMBINE(8) else RRT_COMBINE(16)
#undef RRT_COMBINE
  return cudaGetLastError();
}

static size_t landmarks_smem_bytes(const Grid& grid, int D, int k) {
  const int V = D / 128, RG = 16 / (V > 0 ? V : 1);
  size_t fl = 2 * (size_t)D + (size_t)k * D + (((size_t)grid.P * k + 3) & ~(size_t)3) +
              2 * (size_t)((grid.P + 1) & ~1) + ((grid.P + 3) & ~3) + 32 + (size_t)RG * k * D;
  return fl * sizeof(float);
}

bool crmsa_landmarks_supported(const Grid& grid, int D, int k) {
  const int V = D / 128;
  if (D % 128 || k < 1 || k > RRT_MAX_K_DEV) return false;
  if (V != 1 && V != 2 && V != 4 && V != 8) return false;
  return landmarks_smem_bytes(grid, D, k) <= 227 * 1024;
}

cudaError_t launch_crmsa_landmarks(const float* x1, const float* gamma, const float* beta,
                                   const float* phi, float* logits, __half* landmarks,
                                   float2* rstat, const Grid& grid, int D, int k,
                                   cudaStream_t stream) {
  if (!crmsa_landmarks_supported(grid, D, k)) return cudaErrorInvalidValue;
  {
    static long long* last = nullptr;
    if (last != g_attn_trace) {  // debug hook: mirror the host pointer into the device symbol
      cudaMemcpyToSymbolAsync(g_attn_trace_dev, &g_attn_trace, sizeof(g_attn_trace), 0,
                              cudaMemcpyHostToDevice, stream);
      last = g_attn_trace;
    }
  }
  const int V = D / 128;
  size_t smem = landmarks_smem_bytes(grid, D, k);
#define RRT_LM(VV, KM)                                                                           \
  {                                                                                              \
    static DeviceOnce configured; /* per instantiation: the attribute call is slow (~25 us) */ \
    if (configured.needed()) {                                                                           \
      cudaError_t e = cudaFuncSetAttribute(crmsa_landmarks_kernel<VV, KM>,                       \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return e;                                                            \
    }                                                                                            \
    crmsa_landmarks_kernel<VV, KM><<<grid.R, 512, smem, stream>>>(x1, gamma, beta, phi, logits,  \
                                                                  landmarks, rstat, grid, k);    \
  }
#define RRT_LM_K(VV) \
  { if (k <= 4) RRT_LM(VV, 4) else if (k <= 8) RRT_LM(VV, 8) else RRT_LM(VV, 16) }
  switch (V) {
    case 1: RRT_LM_K(1) break;
    case 2: RRT_LM_K(2) break;
    case 4: RRT_LM_K(4) break;
    default: RRT_LM_K(8) break;
  }
#undef RRT_LM_K
#undef RRT_LM
  return cudaGetLastError();
}

// The split front end (default): row statistics / logits on all SMs, then the folded combine.
cudaError_t launch_crmsa_front_split(const float* x1, const float* gamma, const float* beta,
                                     const float* phi, float2* stats, float* logits,
                                     __half* landmarks, float2* rstat, const Grid& grid, int D, int k,
                                     cudaStream_t stream, const float* rs_part, int rs_parts) {
  if (D % 128 || k < 1 || k > RRT_MAX_K_DEV) return cudaErrorInvalidValue;
  const int V = D / 128;
  if (V != 1 && V != 2 && V != 4 && V != 8) return cudaErrorNotSupported;
  const int KM = k <= 4 ? 4 : (k <= 8 ? 8 : 16);
  if (rs_part && (k > 4 || !phi || rs_parts < 1)) return cudaErrorInvalidValue;
  if (!rs_part) {
    size_t smem = ((size_t)KM * D + 2 * KM) * sizeof(float);
    int blocks = (grid.Np + 15) / 16;
    if (blocks > 148 * 4) blocks = 148 * 4;
#define RRT_RS(VV, KK)                                                                            \
  {                                                                                                \
    static DeviceOnce configured;                                                                \
    if (configured.needed()) {                                                                             \
      cudaError_t e = cudaFuncSetAttribute(crmsa_rowstats_kernel<VV, KK>,                          \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
      if (e != cudaSuccess) return e;                                                              \
    }                                                                                              \
    prefer_max_shared(crmsa_rowstats_kernel<VV, KK>);                                                      \
    cudaError_t le = launch_chain_kernel(crmsa_rowstats_kernel<VV, KK>, dim3(blocks), dim3(256), smem, stream, x1, \
                                         gamma, beta, phi, stats, logits, grid, k);                 \
    if (le != cudaSuccess) return le;                                                              \
  }
#define RRT_RS_K(VV) { if (KM == 4) RRT_RS(VV, 4) else if (KM == 8) RRT_RS(VV, 8) else RRT_RS(VV, 16) }
    switch (V) {
      case 1: RRT_RS_K(1) break;
      case 2: RRT_RS_K(2) break;
      case 4: RRT_RS_K(4) break;
      default: RRT_RS_K(8) break;
    }
#undef RRT_RS_K
#undef RRT_RS
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  {
    size_t smem = ((size_t)grid.P * KM + ((grid.P + 3) & ~3) + (size_t)8 * KM * 128 + 2 * KM + 2 * (size_t)grid.P) *
                  sizeof(float);
    if (smem > 227 * 1024) return cudaErrorNotSupported;
    dim3 g(D / 128, grid.R);
#define RRT_C2(KK, FP)                                                                             \
  {                                                                                                \
    static DeviceOnce configured;                                                                \
    if (configured.needed()) {                                                                             \
      cudaError_t e = cudaFuncSetAttribute(crmsa_combine2_kernel<KK, FP>,                          \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return e;                                                              \
    }                                                                                              \
    prefer_max_shared(crmsa_combine2_kernel<KK, FP>);                                              \
    cudaError_t le = launch_chain_kernel(crmsa_combine2_kernel<KK, FP>, g, dim3(256), smem, stream, x1, gamma, beta, \
                                         (const float2*)stats, (const float*)logits, landmarks, rstat, grid, D, k, \
                                         rs_part, rs_parts, phi, stats, logits);                   \
    if (le != cudaSuccess) return le;                                                              \
  }
    if (rs_part) RRT_C2(4, true)
    else if (KM == 4) RRT_C2(4, false) else if (KM == 8) RRT_C2(8, false) else RRT_C2(16, false)
#undef RRT_C2
  }
  return cudaGetLastError();
}

cudaError_t launch_landmark_attention(const float* lqkv, __half* lo, int k, int R, int D, int heads,
                                      cudaStream_t stream) {
  if (R != 64 || heads <= 0 || D % heads || (D / heads) % 32) return cudaErrorInvalidValue;
  float scale = 1.f / sqrtf((float)(D / heads));
  landmark_attn_kernel<<<dim3(heads, k), 256, 0, stream>>>(lqkv, lo, D, heads, scale);
  return cudaGetLastError();
}

cudaError_t launch_crmsa_dispatch(const float* x1, const float* x0, const float* logits,
                                  const float2* rstat, const float* lm, const float* gamma,
                                  const float* beta, float* out, const Grid& grid, int D, int k,
                                  cudaStream_t stream) {
  if (D % 128 || k < 1 || k > RRT_MAX_K_DEV) return cudaErrorInvalidValue;
  if (grid.L == 0) return cudaSuccess;
  int blocks = (grid.L + 7) / 8;
  RRT_DISPATCH_V(D, prefer_max_shared(crmsa_dispatch_kernel<V>);
                 return launch_chain_kernel(crmsa_dispatch_kernel<V>, dim3(blocks), dim3(256), 0, stream, x1, x0,
                                            logits, rstat, lm, gamma, beta, out, grid, k));
  return cudaGetLastError();
}

}  // namespace rrt
