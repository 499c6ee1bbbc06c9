// Backward of the CR-MSA block (+ the encoder's final LayerNorm), fp32 streaming kernels around the
// landmark-MHA backward (which reuses the GEMM and attention-backward kernels on k*64 rows).
// Autograd of modules/rmsa.py:290-337 in the rank-k form the forward kernels use:
//   l[p,n]  = z[p,:] . phi[:,n]                z = LN_cr(x1) (0 on pad slots)
//   cw      = softmax_p(l)      L[n,:]  = sum_p cw[p,n] z[p,:]          (combine)
//   dp      = softmax_n(l)      mm[p,n] = (l - lo_n) / (hi_n - lo_n + 1e-8)
//   y[p,:]  = sum_n dp*mm * L'[n,:]            L' = MHA(L)               (dispatch)
//   out     = LN_f(x1 + y (+ x0))
// Two token-parallel passes with the per-region reductions between them:
//   crmsa_dispatch_bwd_kernel  dh = LN_f backward, dw[p,n] = dh . L'_n, dL'[n,:] += w dh,
//                              d lo_n / d hi_n partial sums, final-norm gamma/beta gradients
//   (landmark MHA backward:    dL' -> dL)
//   crmsa_combine_bwd_kernel   dl (three paths), dz, dphi, LN_cr backward -> dx1 = dh + ...
// torch.min/max(dim) send their gradient to ONE arg-min/max element; here every element equal to the
// extremum receives it.  Ties only occur among zero pad slots, whose logit gradient is discarded.
#include "backward.cuh"
#include "kernels.cuh"

namespace rrt {
namespace {

template <int V>
__device__ __forceinline__ void load_row(float4 (&v)[V], const float* row, int lane) {
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(row) + lane + 32 * i);
}
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
}
__device__ __forceinline__ void axpy4(float4& y, float a, float4 x) {
  y.x = fmaf(a, x.x, y.x); y.y = fmaf(a, x.y, y.y); y.z = fmaf(a, x.z, y.z); y.w = fmaf(a, x.w, y.w);
}

// CTA-wide sum of per-lane column partials part[V] (columns 4*(lane+32i)..+3) over the 8 warps, then
// one atomicAdd per column: dst[c * stride] += sum.  `red` is [8][D] shared floats.
template <int V>
__device__ __forceinline__ void cta_colsum_atomic(const float4 (&part)[V], float* red, float* dst,
                                                  int stride, int warp, int lane) {
  constexpr int D = 128 * V;
#pragma unroll
  for (int i = 0; i < V; ++i)
    *reinterpret_cast<float4*>(red + warp * D + 4 * (lane + 32 * i)) = part[i];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * D + c];
    if (s != 0.f) atomicAdd(dst + (size_t)c * stride, s);
  }
  __syncthreads();
}

// grid (R, chunks), 256 threads; smem: Ls[KMAX][D] | red[8][D]
template <int V, int KMAX>
__global__ void __launch_bounds__(256) crmsa_dispatch_bwd_kernel(
    const float* __restrict__ x1, const float* __restrict__ x0, const float* __restrict__ logits,
    const float2* __restrict__ rstat, const float* __restrict__ lmp, const float* __restrict__ gamma_f,
    const float* __restrict__ dout, float* __restrict__ dh, float* __restrict__ dw,
    float* __restrict__ dLp, float2* __restrict__ rgrad, float* __restrict__ dgamma_f,
    float* __restrict__ dbeta_f, Grid grid, int k, int tpc) {
  constexpr int D = 128 * V;
  extern __shared__ __align__(16) float smem[];
  float* Ls = smem;
  float* red = Ls + KMAX * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rho = blockIdx.x, P = grid.P;
  for (int i = tid; i < k * (D / 4); i += 256) {
    int n = i / (D / 4), c4 = i - n * (D / 4);
    reinterpret_cast<float4*>(Ls + n * D)[c4] =
        __ldg(reinterpret_cast<const float4*>(lmp + ((size_t)n * grid.R + rho) * D) + c4);
  }
  __syncthreads();

  float4 acc[KMAX][V], dg[V], db[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < KMAX; ++n) acc[n][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float my_dlo = 0.f, my_dhi = 0.f;
  const float2 mm = lane < k ? __ldg(rstat + (size_t)rho * k + lane) : make_float2(0.f, 1.f);
  const float rng = mm.y - mm.x + 1e-8f;
  int p_end = (blockIdx.y + 1) * tpc;
  if (p_end > P) p_end = P;
  for (int p = blockIdx.y * tpc + warp; p < p_end; p += 8) {
    const int slot = rho * P + p;
    const int tok = grid.slot_to_token(slot);
    if (tok >= grid.L) continue;
    float4 v[V], d[V];
    load_row<V>(v, x1 + (size_t)tok * D, lane);
    load_row<V>(d, dout + (size_t)tok * D, lane);
    if (x0) {
      float4 u[V];
      load_row<V>(u, x0 + (size_t)tok * D, lane);
#pragma unroll
      for (int i = 0; i < V; ++i) { v[i].x += u[i].x; v[i].y += u[i].y; v[i].z += u[i].z; v[i].w += u[i].w; }
    }
    const float lg = lane < k ? __ldg(logits + (size_t)slot * k + lane) : -INFINITY;
    const float mx = warp_max(lg);
    const float ex = lane < k ? __expf(lg - mx) : 0.f;
    const float dpn = ex / warp_sum(ex);
    const float mmn = lane < k ? (lg - mm.x) / rng : 0.f;
    const float my_w = dpn * mmn;
#pragma unroll
    for (int n = 0; n < KMAX; ++n) {
      if (n < k) {
        const float w = __shfl_sync(0xffffffffu, my_w, n);
#pragma unroll
        for (int i = 0; i < V; ++i)
          axpy4(v[i], w, *reinterpret_cast<const float4*>(Ls + n * D + 4 * (lane + 32 * i)));
      }
    }
    if (gamma_f) {  // backward of out = LN_f(h), v = h
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mean = warp_sum(s) * (1.f / D);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += dot4(v[i], v[i]);
      }
      const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + kLnEps);
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
        dg[i].x = fmaf(d[i].x, v[i].x, dg[i].x); dg[i].y = fmaf(d[i].y, v[i].y, dg[i].y);
        dg[i].z = fmaf(d[i].z, v[i].z, dg[i].z); dg[i].w = fmaf(d[i].w, v[i].w, dg[i].w);
        db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma_f) + lane + 32 * i);
        d[i].x *= gm.x; d[i].y *= gm.y; d[i].z *= gm.z; d[i].w *= gm.w;
        m1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
        m2 += dot4(d[i], v[i]);
      }
      m1 = warp_sum(m1) * (1.f / D);
      m2 = warp_sum(m2) * (1.f / D);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        d[i].x = rstd * (d[i].x - m1 - v[i].x * m2); d[i].y = rstd * (d[i].y - m1 - v[i].y * m2);
        d[i].z = rstd * (d[i].z - m1 - v[i].z * m2); d[i].w = rstd * (d[i].w - m1 - v[i].w * m2);
      }
    }
    // d = dh
#pragma unroll
    for (int i = 0; i < V; ++i) reinterpret_cast<float4*>(dh + (size_t)tok * D)[lane + 32 * i] = d[i];
    float my_dw = 0.f;
#pragma unroll
    for (int n = 0; n < KMAX; ++n) {
      if (n < k) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i)
          s += dot4(d[i], *reinterpret_cast<const float4*>(Ls + n * D + 4 * (lane + 32 * i)));
        s = warp_sum(s);
        if (lane == n) my_dw = s;
        const float w = __shfl_sync(0xffffffffu, my_w, n);
#pragma unroll
        for (int i = 0; i < V; ++i) axpy4(acc[n][i], w, d[i]);
      }
    }
    if (lane < k) {
      dw[(size_t)slot * k + lane] = my_dw;
      const float dmm = my_dw * dpn;
      my_dlo += dmm * (mmn - 1.f) / rng;
      my_dhi -= dmm * mmn / rng;
    }
  }
  if (lane < k) {
    if (my_dlo != 0.f) atomicAdd(&rgrad[(size_t)rho * k + lane].x, my_dlo);
    if (my_dhi != 0.f) atomicAdd(&rgrad[(size_t)rho * k + lane].y, my_dhi);
  }
#pragma unroll
  for (int n = 0; n < KMAX; ++n)
    if (n < k) cta_colsum_atomic<V>(acc[n], red, dLp + ((size_t)n * grid.R + rho) * D, 1, warp, lane);
  if (gamma_f) {
    cta_colsum_atomic<V>(dg, red, dgamma_f, 1, warp, lane);
    cta_colsum_atomic<V>(db, red, dbeta_f, 1, warp, lane);
  }
}

// grid (R, chunks), 256 threads; smem: dLs[KMAX][D] | Ph[KMAX][D] | red[8][D] | rs[4][KMAX]
// MODE 0: logits = z phi (linear landmark directions).
// crmsa_mlp (logits = W2 tanh(W1 z), modules/rmsa.py:248-252,305) takes two passes around the MLP backward:
// MODE 1 only emits the logit gradient dl[slot, n] (-> dlogits), MODE 2 replaces the dl . phi term of dz by the
// MLP's input gradient (dzx16: fp16 rows in slot order, scaled by amax_x) and finishes the LayerNorm backward.
template <int V, int KMAX, int MODE>
__global__ void __launch_bounds__(256) crmsa_combine_bwd_kernel(
    const float* __restrict__ x1, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ phi, const float* __restrict__ logits, const float2* __restrict__ rstat,
    const __half* __restrict__ lm16, const __half* __restrict__ dlm16,
    const uint32_t* __restrict__ amax_l, const float* __restrict__ dw, const float2* __restrict__ rgrad,
    const float* __restrict__ dh, float dh_weight, float* __restrict__ dx1, float* __restrict__ dphi,
    float* __restrict__ dgamma, float* __restrict__ dbeta, uint32_t* __restrict__ amax_out, Grid grid,
    int k, int tpc, float* __restrict__ dlogits, const __half* __restrict__ dzx16,
    const uint32_t* __restrict__ amax_x) {
  constexpr int D = 128 * V;
  extern __shared__ __align__(16) float smem[];
  float* dLs = smem;              // unscaled dL rows of this region
  float* Ph = dLs + KMAX * D;     // phi^T
  float* red = Ph + KMAX * D;     // [8][D]
  float* rs = red + 8 * D;        // [0]: softmax_p max, [1]: 1/sum, [2]: c_n = dL_n . L_n
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rho = blockIdx.x, P = grid.P;
  const float inv_s = amax_l ? grad_inv_scale(__ldg(amax_l)) : 1.f;
  for (int i = tid; i < k * D; i += 256) {
    int n = i / D, c = i - n * D;
    dLs[i] = __half2float(dlm16[((size_t)n * grid.R + rho) * D + c]) * inv_s;
    if (MODE == 0) Ph[i] = __ldg(phi + (size_t)c * k + n);
  }
  __syncthreads();
  for (int n = warp; n < k; n += 8) {
    float mx = -INFINITY;
    for (int p = lane; p < P; p += 32) mx = fmaxf(mx, __ldg(logits + ((size_t)rho * P + p) * k + n));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) sum += __expf(__ldg(logits + ((size_t)rho * P + p) * k + n) - mx);
    sum = warp_sum(sum);
    float c = 0.f;
    for (int j = lane; j < D; j += 32)
      c = fmaf(dLs[n * D + j], __half2float(lm16[((size_t)n * grid.R + rho) * D + j]), c);
    c = warp_sum(c);
    if (lane == 0) { rs[n] = mx; rs[KMAX + n] = 1.f / sum; rs[2 * KMAX + n] = c; }
  }
  __syncthreads();

  float4 aphi[KMAX][V], dg[V], db[V], gm[V], bt[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    bt[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
#pragma unroll
    for (int n = 0; n < KMAX; ++n) aphi[n][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float2 mm = lane < k ? __ldg(rstat + (size_t)rho * k + lane) : make_float2(0.f, 1.f);
  const float2 rg = lane < k ? __ldg(rgrad + (size_t)rho * k + lane) : make_float2(0.f, 0.f);
  const float rng = mm.y - mm.x + 1e-8f;
  const float pmx = lane < k ? rs[lane] : 0.f, pinv = lane < k ? rs[KMAX + lane] : 0.f;
  const float cn = lane < k ? rs[2 * KMAX + lane] : 0.f;
  float amax = 0.f;
  int p_end = (blockIdx.y + 1) * tpc;
  if (p_end > P) p_end = P;
  for (int p = blockIdx.y * tpc + warp; p < p_end; p += 8) {
    const int slot = rho * P + p;
    const int tok = grid.slot_to_token(slot);
    if (tok >= grid.L) continue;
    float4 zh[V], z[V];
    load_row<V>(zh, x1 + (size_t)tok * D, lane);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (zh[i].x + zh[i].y) + (zh[i].z + zh[i].w);
    const float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      zh[i].x -= mean; zh[i].y -= mean; zh[i].z -= mean; zh[i].w -= mean;
      q += dot4(zh[i], zh[i]);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + kLnEps);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      zh[i].x *= rstd; zh[i].y *= rstd; zh[i].z *= rstd; zh[i].w *= rstd;
      z[i].x = fmaf(zh[i].x, gm[i].x, bt[i].x); z[i].y = fmaf(zh[i].y, gm[i].y, bt[i].y);
      z[i].z = fmaf(zh[i].z, gm[i].z, bt[i].z); z[i].w = fmaf(zh[i].w, gm[i].w, bt[i].w);
    }
    const float lg = lane < k ? __ldg(logits + (size_t)slot * k + lane) : -INFINITY;
    const float cw = lane < k ? __expf(lg - pmx) * pinv : 0.f;
    const float mxk = warp_max(lg);
    const float ex = lane < k ? __expf(lg - mxk) : 0.f;
    const float dpn = ex / warp_sum(ex);
    const float mmn = lane < k ? (lg - mm.x) / rng : 0.f;
    const float dwn = lane < k ? __ldg(dw + (size_t)slot * k + lane) : 0.f;
    const float ddp = dwn * mmn, dmm = dwn * dpn;
    const float sdp = warp_sum(dpn * ddp);
    float dcw = 0.f;
#pragma unroll
    for (int n = 0; n < KMAX; ++n) {
      if (n < k) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i)
          t += dot4(z[i], *reinterpret_cast<const float4*>(dLs + n * D + 4 * (lane + 32 * i)));
        t = warp_sum(t);
        if (lane == n) dcw = t;
      }
    }
    float dl = 0.f;
    if (lane < k) {
      dl = cw * (dcw - cn) + dpn * (ddp - sdp) + dmm / rng;
      if (lg == mm.x) dl += rg.x;
      if (lg == mm.y) dl += rg.y;
    }
    if (MODE == 1) {  // the MLP backward runs between the passes: only the logit gradient leaves this one
      if (lane < k) dlogits[(size_t)slot * k + lane] = dl;
      continue;
    }
    float4 dz[V];
#pragma unroll
    for (int i = 0; i < V; ++i) dz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 2) {  // d z through the MLP (dgrad GEMM of phi.0), unscaled
      const float inv_x = grad_inv_scale(__ldg(amax_x));
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 e = unpack_h4(__ldg(reinterpret_cast<const uint2*>(dzx16 + (size_t)slot * D) + lane + 32 * i));
        dz[i] = make_float4(e.x * inv_x, e.y * inv_x, e.z * inv_x, e.w * inv_x);
      }
    }
#pragma unroll
    for (int n = 0; n < KMAX; ++n) {
      if (n < k) {
        const float cwn = __shfl_sync(0xffffffffu, cw, n), dln = __shfl_sync(0xffffffffu, dl, n);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          axpy4(dz[i], cwn, *reinterpret_cast<const float4*>(dLs + n * D + 4 * (lane + 32 * i)));
          if (MODE == 0) {
            axpy4(dz[i], dln, *reinterpret_cast<const float4*>(Ph + n * D + 4 * (lane + 32 * i)));
            axpy4(aphi[n][i], dln, z[i]);
          }
        }
      }
    }
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      dg[i].x = fmaf(dz[i].x, zh[i].x, dg[i].x); dg[i].y = fmaf(dz[i].y, zh[i].y, dg[i].y);
      dg[i].z = fmaf(dz[i].z, zh[i].z, dg[i].z); dg[i].w = fmaf(dz[i].w, zh[i].w, dg[i].w);
      db[i].x += dz[i].x; db[i].y += dz[i].y; db[i].z += dz[i].z; db[i].w += dz[i].w;
      dz[i].x *= gm[i].x; dz[i].y *= gm[i].y; dz[i].z *= gm[i].z; dz[i].w *= gm[i].w;
      m1 += (dz[i].x + dz[i].y) + (dz[i].z + dz[i].w);
      m2 += dot4(dz[i], zh[i]);
    }
    m1 = warp_sum(m1) * (1.f / D);
    m2 = warp_sum(m2) * (1.f / D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(dh + (size_t)tok * D) + lane + 32 * i);
      float4 o;
      o.x = fmaf(dh_weight, r.x, rstd * (dz[i].x - m1 - zh[i].x * m2));
      o.y = fmaf(dh_weight, r.y, rstd * (dz[i].y - m1 - zh[i].y * m2));
      o.z = fmaf(dh_weight, r.z, rstd * (dz[i].z - m1 - zh[i].z * m2));
      o.w = fmaf(dh_weight, r.w, rstd * (dz[i].w - m1 - zh[i].w * m2));
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
      reinterpret_cast<float4*>(dx1 + (size_t)tok * D)[lane + 32 * i] = o;
    }
  }
  if (MODE == 1) return;
  if (amax_out) {
    amax = warp_max(amax);
    if (lane == 0 && amax > 0.f) atomic_amax(amax_out, amax);
  }
  if (MODE == 0) {
#pragma unroll
    for (int n = 0; n < KMAX; ++n)
      if (n < k) cta_colsum_atomic<V>(aphi[n], red, dphi + n, k, warp, lane);
  }
  cta_colsum_atomic<V>(dg, red, dgamma, 1, warp, lane);
  cta_colsum_atomic<V>(db, red, dbeta, 1, warp, lane);
}

// Backward of logits = W2 tanh(pre) for one slot per warp: dpre[slot, j] = (sum_n dl[slot, n] W2[n, j]) (1 - h^2),
// dW2[n, j] += dl[slot, n] h[slot, j].  h = tanh(pre) is the forward's hidden activation (fp32).  H4 = D / 4.
template <int KMAX>
__global__ void __launch_bounds__(256) crmsa_mlp_hidden_bwd_kernel(const float* __restrict__ dlogits,
                                                                   const float* __restrict__ hidden,
                                                                   const float* __restrict__ w2,
                                                                   float* __restrict__ dpre, float* __restrict__ dw2,
                                                                   int rows, int H4, int k) {
  extern __shared__ __align__(16) float smem[];
  float* W = smem;             // [k][H4]
  float* acc = W + KMAX * H4;  // [k][H4] CTA partial of dW2
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < k * H4; i += 256) { W[i] = __ldg(w2 + i); acc[i] = 0.f; }
  __syncthreads();
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float dl = lane < k ? __ldg(dlogits + (size_t)r * k + lane) : 0.f;
    if (__all_sync(0xffffffffu, dl == 0.f)) {   // pad slots and tokens whose logits carry no gradient
      for (int j = lane; j < H4; j += 32) dpre[(size_t)r * H4 + j] = 0.f;
      continue;
    }
    for (int j = lane; j < H4; j += 32) {
      const float h = __ldg(hidden + (size_t)r * H4 + j);
      float g = 0.f;
#pragma unroll
      for (int n = 0; n < KMAX; ++n)
        if (n < k) {
          const float d = __shfl_sync(0xffffffffu, dl, n);
          g = fmaf(d, W[n * H4 + j], g);
          atomicAdd(&acc[n * H4 + j], d * h);
        }
      dpre[(size_t)r * H4 + j] = g * (1.f - h * h);
    }
  }
  __syncthreads();
  for (int i = tid; i < k * H4; i += 256)
    if (acc[i] != 0.f) atomicAdd(dw2 + i, acc[i]);
}

// Backward of the landmark MHA core for head dims the tensor-core attention backward does not cover
// (crmsa_heads = 1: head_dim = D, the reference's BRCA-R50 / LUAD-PLIP recipes): fp32, 64 landmarks per sequence.
//   P = softmax(scale Q K^T),  dV = P^T dO,  dP = dO V^T,  dS = P (dP - rowsum(dP P)),  dQ = scale dS K,  dK = scale dS^T Q
// grid (heads, k), 256 threads = a 16 x 16 grid of 4 x 4 register tiles; head_dim walked in chunks of 32.
// lqkv: fp32 [k*64, 3D] (the forward's tape); dO16 / dqkv16: fp16 rows in the scaled gradient domain (the kernel
// is linear in dO, so the scale passes through).
__global__ void __launch_bounds__(256) landmark_attn_bwd_kernel(const float* __restrict__ lqkv,
                                                                const __half* __restrict__ dO16,
                                                                __half* __restrict__ dqkv16, int D, int heads,
                                                                float scale) {
  constexpr int R = 64, CH = 32;
  extern __shared__ __align__(16) float lab_smem[];   // 49 KB: dynamic (above the 48 KB static limit)
  float (*a_s)[CH + 1] = reinterpret_cast<float (*)[CH + 1]>(lab_smem);
  float (*b_s)[CH + 1] = reinterpret_cast<float (*)[CH + 1]>(lab_smem + R * (CH + 1));
  float (*ps)[R + 1] = reinterpret_cast<float (*)[R + 1]>(lab_smem + 2 * R * (CH + 1));               // P
  float (*ds)[R + 1] = reinterpret_cast<float (*)[R + 1]>(lab_smem + 2 * R * (CH + 1) + R * (R + 1));  // dP, then dS
  const int h = blockIdx.x, n = blockIdx.y, dh = D / heads;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const size_t ld = 3 * (size_t)D;
  const float* base = lqkv + (size_t)n * R * ld + h * dh;
  const __half* dob = dO16 + (size_t)n * R * D + h * dh;
  __half* dqb = dqkv16 + (size_t)n * R * ld + h * dh;

  float s[4][4], g[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = g[i][j] = 0.f;
  // S = Q K^T and dP = dO V^T, accumulated over the head dim
  for (int c0 = 0; c0 < dh; c0 += CH) {
    __syncthreads();
    for (int i = tid; i < R * CH; i += 256) {
      int r = i / CH, c = i - r * CH;
      a_s[r][c] = __ldg(base + (size_t)r * ld + c0 + c);
      b_s[r][c] = __ldg(base + (size_t)r * ld + D + c0 + c);
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < CH; ++c) {
      float qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { qv[i] = a_s[ty * 4 + i][c]; kv[i] = b_s[tx * 4 + i][c]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qv[i], kv[j], s[i][j]);
    }
    __syncthreads();
    for (int i = tid; i < R * CH; i += 256) {
      int r = i / CH, c = i - r * CH;
      a_s[r][c] = __half2float(dob[(size_t)r * D + c0 + c]);
      b_s[r][c] = __ldg(base + (size_t)r * ld + 2 * D + c0 + c);
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < CH; ++c) {
      float ov[4], vv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { ov[i] = a_s[ty * 4 + i][c]; vv[i] = b_s[tx * 4 + i][c]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) g[i][j] = fmaf(ov[i], vv[j], g[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ps[ty * 4 + i][tx * 4 + j] = s[i][j] * scale;
      ds[ty * 4 + i][tx * 4 + j] = g[i][j];
    }
  __syncthreads();
  {  // rows: P = softmax(S), dS = P (dP - sum_j dP_j P_j); warp w owns rows 8w..8w+7
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp * 8; r < warp * 8 + 8; ++r) {
      float a = ps[r][lane], b = ps[r][lane + 32];
      const float mx = warp_max(fmaxf(a, b));
      a = __expf(a - mx);
      b = __expf(b - mx);
      const float inv = 1.f / warp_sum(a + b);
      a *= inv;
      b *= inv;
      const float da = ds[r][lane], db = ds[r][lane + 32];
      const float dot = warp_sum(da * a + db * b);
      ps[r][lane] = a;
      ps[r][lane + 32] = b;
      ds[r][lane] = a * (da - dot);
      ds[r][lane + 32] = b * (db - dot);
    }
  }
  // per head-dim chunk: dQ = scale dS K, dK = scale dS^T Q, dV = P^T dO   (thread: 4 rows x 2 columns)
  for (int c0 = 0; c0 < dh; c0 += CH) {
    float dq[4][2], dk[4][2], dv[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) dq[i][0] = dq[i][1] = dk[i][0] = dk[i][1] = dv[i][0] = dv[i][1] = 0.f;
    __syncthreads();
    for (int i = tid; i < R * CH; i += 256) {   // a_s = K chunk, b_s = Q chunk
      int r = i / CH, c = i - r * CH;
      a_s[r][c] = __ldg(base + (size_t)r * ld + D + c0 + c);
      b_s[r][c] = __ldg(base + (size_t)r * ld + c0 + c);
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < R; ++j) {
      const float k0 = a_s[j][tx * 2], k1 = a_s[j][tx * 2 + 1], q0 = b_s[j][tx * 2], q1 = b_s[j][tx * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dsr = ds[ty * 4 + i][j];   // dS[row][j]
        const float dst = ds[j][ty * 4 + i];   // dS[j][row]  (dS^T)
        dq[i][0] = fmaf(dsr, k0, dq[i][0]); dq[i][1] = fmaf(dsr, k1, dq[i][1]);
        dk[i][0] = fmaf(dst, q0, dk[i][0]); dk[i][1] = fmaf(dst, q1, dk[i][1]);
      }
    }
    __syncthreads();
    for (int i = tid; i < R * CH; i += 256) {   // a_s = dO chunk
      int r = i / CH, c = i - r * CH;
      a_s[r][c] = __half2float(dob[(size_t)r * D + c0 + c]);
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < R; ++j) {
      const float o0 = a_s[j][tx * 2], o1 = a_s[j][tx * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pt = ps[j][ty * 4 + i];    // P[j][row]  (P^T)
        dv[i][0] = fmaf(pt, o0, dv[i][0]); dv[i][1] = fmaf(pt, o1, dv[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half* row = dqb + (size_t)(ty * 4 + i) * ld + c0 + tx * 2;
      *reinterpret_cast<uint32_t*>(row) = pack_h2(dq[i][0] * scale, dq[i][1] * scale);
      *reinterpret_cast<uint32_t*>(row + D) = pack_h2(dk[i][0] * scale, dk[i][1] * scale);
      *reinterpret_cast<uint32_t*>(row + 2 * D) = pack_h2(dv[i][0], dv[i][1]);
    }
  }
}

int token_chunks(const Grid& g) {
  int chunks = (2 * 148 + g.R - 1) / g.R;
  if (chunks > (g.P + 7) / 8) chunks = (g.P + 7) / 8;
  return chunks < 1 ? 1 : chunks;
}

#define RRT_CRB_DISPATCH(D, k, ...)                                                  \
  switch ((D) / 128) {                                                               \
    case 1: { constexpr int V = 1; if ((k) <= 4) { constexpr int KM = 4; __VA_ARGS__; } else { constexpr int KM = 8; __VA_ARGS__; } break; } \
    case 2: { constexpr int V = 2; if ((k) <= 4) { constexpr int KM = 4; __VA_ARGS__; } else { constexpr int KM = 8; __VA_ARGS__; } break; } \
    case 4: { constexpr int V = 4; if ((k) <= 4) { constexpr int KM = 4; __VA_ARGS__; } else { constexpr int KM = 8; __VA_ARGS__; } break; } \
    case 8: { constexpr int V = 8; if ((k) <= 4) { constexpr int KM = 4; __VA_ARGS__; } else { constexpr int KM = 8; __VA_ARGS__; } break; } \
    default: return cudaErrorInvalidValue;                                           \
  }

template <typename K>
cudaError_t set_smem(K kern, size_t bytes) {
  return bytes > 48 * 1024
             ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)
             : cudaSuccess;
}

}  // namespace

bool crmsa_backward_supported(int D, int k) {
  return (D == 128 || D == 256 || D == 512 || D == 1024) && k >= 1 && k <= 8;
}

cudaError_t launch_crmsa_dispatch_bwd(const float* x1, const float* x0, const float* logits,
                                      const float2* rstat, const float* lmp, const float* gamma_f,
                                      const float* dout, float* dh, float* dw, float* dLp,
                                      float2* rgrad, float* dgamma_f, float* dbeta_f, const Grid& grid,
                                      int D, int k, cudaStream_t stream) {
  if (!crmsa_backward_supported(D, k)) return cudaErrorInvalidValue;
  const int chunks = token_chunks(grid), tpc = (grid.P + chunks - 1) / chunks;
  dim3 gr(grid.R, chunks);
  RRT_CRB_DISPATCH(D, k, {
    size_t smem = (size_t)(KM + 8) * D * sizeof(float);
    auto kern = crmsa_dispatch_bwd_kernel<V, KM>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<gr, 256, smem, stream>>>(x1, x0, logits, rstat, lmp, gamma_f, dout, dh, dw, dLp, rgrad,
                                    dgamma_f, dbeta_f, grid, k, tpc);
  });
  return cudaGetLastError();
}

cudaError_t launch_crmsa_combine_bwd(const float* x1, const float* gamma, const float* beta,
                                     const float* phi, const float* logits, const float2* rstat,
                                     const __half* lm16, const __half* dlm16, const uint32_t* amax_l,
                                     const float* dw, const float2* rgrad, const float* dh,
                                     float dh_weight, float* dx1, float* dphi, float* dgamma,
                                     float* dbeta, uint32_t* amax_out, const Grid& grid, int D, int k,
                                     cudaStream_t stream, int mode, float* dlogits, const __half* dzx16,
                                     const uint32_t* amax_x) {
  if (!crmsa_backward_supported(D, k) || mode < 0 || mode > 2) return cudaErrorInvalidValue;
  if ((mode == 0 && !phi) || (mode == 1 && !dlogits) || (mode == 2 && (!dzx16 || !amax_x))) return cudaErrorInvalidValue;
  const int chunks = token_chunks(grid), tpc = (grid.P + chunks - 1) / chunks;
  dim3 gr(grid.R, chunks);
#define RRT_CRB_LAUNCH(MODE_)                                                                                   \
  {                                                                                                             \
    auto kern = crmsa_combine_bwd_kernel<V, KM, MODE_>;                                                         \
    cudaError_t e = set_smem(kern, smem);                                                                       \
    if (e != cudaSuccess) return e;                                                                             \
    kern<<<gr, 256, smem, stream>>>(x1, gamma, beta, phi, logits, rstat, lm16, dlm16, amax_l, dw, rgrad, dh,    \
                                    dh_weight, dx1, dphi, dgamma, dbeta, amax_out, grid, k, tpc, dlogits,       \
                                    dzx16, amax_x);                                                             \
  }
  RRT_CRB_DISPATCH(D, k, {
    size_t smem = ((size_t)(2 * KM + 8) * D + 4 * KM) * sizeof(float);
    if (mode == 0) RRT_CRB_LAUNCH(0) else if (mode == 1) RRT_CRB_LAUNCH(1) else RRT_CRB_LAUNCH(2)
  });
#undef RRT_CRB_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_landmark_attention_bwd(const float* lqkv, const __half* dO16, __half* dqkv16, int k, int R, int D,
                                          int heads, cudaStream_t stream) {
  if (R != 64 || heads <= 0 || D % heads || (D / heads) % 32) return cudaErrorInvalidValue;
  const float scale = 1.f / sqrtf((float)(D / heads));
  const size_t smem = (size_t)(2 * 64 * 33 + 2 * 64 * 65) * sizeof(float);
  cudaError_t e = set_smem(landmark_attn_bwd_kernel, smem);
  if (e != cudaSuccess) return e;
  landmark_attn_bwd_kernel<<<dim3(heads, k), 256, smem, stream>>>(lqkv, dO16, dqkv16, D, heads, scale);
  return cudaGetLastError();
}

cudaError_t launch_crmsa_mlp_hidden_bwd(const float* dlogits, const float* hidden, const float* w2, float* dpre,
                                        float* dw2, int rows, int H4, int k, cudaStream_t stream) {
  if (k < 1 || k > 8 || H4 < 1 || rows < 1) return cudaErrorInvalidValue;
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (k <= 4) {
    size_t smem = (size_t)2 * 4 * H4 * sizeof(float);
    crmsa_mlp_hidden_bwd_kernel<4><<<blocks, 256, smem, stream>>>(dlogits, hidden, w2, dpre, dw2, rows, H4, k);
  } else {
    size_t smem = (size_t)2 * 8 * H4 * sizeof(float);
    crmsa_mlp_hidden_bwd_kernel<8><<<blocks, 256, smem, stream>>>(dlogits, hidden, w2, dpre, dw2, rows, H4, k);
  }
  return cudaGetLastError();
}

}  // namespace rrt
