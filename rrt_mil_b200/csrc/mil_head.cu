// the two layers the reference's RRTMIL wraps around the encoder
// (modules/rrt.py:204-246): patch_to_emb = Linear(input_dim,512)+act in front, DAttention pooling
// (modules/datten.py:5-38,85-101) + predictor Linear behind.  The two Linear layers run on the
// tcgen05 GEMM (gemm_tcgen05.cu); this file holds the streaming pieces of the pooling head:
//   scores   s_i = hidden_i . w2 (+ b2)                                  warp per token
//   partial  per block of 64 tokens: m_b = max s, z_b = sum exp(s - m_b), v_b = sum exp(s - m_b) h_i
//   final    M = max m_b, Z = sum z_b exp(m_b - M), pooled = sum v_b exp(m_b - M) / Z,
//            logits = pooled . Wp^T + bp
//   weights  a_i = exp(s_i - M) / Z   (only when the caller asks for the attention map)
#include "kernels.cuh"

namespace rrt {
namespace {

// hidden: [L, ld]; gated (AttentionGated, modules/datten.py:66-70): columns [hid, 2 hid) hold the sigmoid
// branch and multiply the first hid columns before the dot product with w2 (= attention_c.weight)
__global__ void __launch_bounds__(256) pool_scores_kernel(const float* __restrict__ hidden,
                                                          const float* __restrict__ w2,
                                                          const float* __restrict__ b2,
                                                          float* __restrict__ scores, int L, int hid, int ld,
                                                          int gated) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= L) return;
  float d = 0.f;
  for (int c = lane; c < hid; c += 32) {
    float v = __ldg(hidden + (size_t)row * ld + c);
    if (gated) v *= __ldg(hidden + (size_t)row * ld + hid + c);
    d = fmaf(v, __ldg(w2 + c), d);
  }
  d = warp_sum(d);
  if (lane == 0) scores[row] = d + (b2 ? __ldg(b2) : 0.f);
}

// training forward through nn.GELU: the GEMM stored the PRE-activation; keep it (the backward needs it, the
// output alone does not determine gelu') and activate.  buf: [L, ld], columns [0, n) are touched; pre: [L, n]
__global__ void __launch_bounds__(256) gelu_keep_pre_kernel(float* __restrict__ buf, float* __restrict__ pre,
                                                            size_t rows, int n4, int ld4) {
  const size_t total = rows * (size_t)n4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / n4;
    const int c = (int)(i - r * n4);
    float4 z = reinterpret_cast<float4*>(buf)[r * ld4 + c];
    if (pre != buf) reinterpret_cast<float4*>(pre)[i] = z;
    reinterpret_cast<float4*>(buf)[r * ld4 + c] = make_float4(gelu_fwd(z.x), gelu_fwd(z.y), gelu_fwd(z.z), gelu_fwd(z.w));
  }
}

constexpr int kPoolRows = 64;

// part[b] = (m_b, z_b, pad, pad, v_b[D])
__global__ void __launch_bounds__(256) pool_partial_kernel(const float* __restrict__ h,
                                                           const float* __restrict__ scores,
                                                           float* __restrict__ part, int L, int D) {
  __shared__ float sw[kPoolRows];
  __shared__ float red[8];
  const int r0 = blockIdx.x * kPoolRows, n = min(kPoolRows, L - r0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float s = tid < n ? __ldg(scores + r0 + tid) : -INFINITY;
  float m = warp_max(s);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  float e = tid < n ? __expf(s - m) : 0.f;
  if (tid < kPoolRows) sw[tid] = e;
  float z = warp_sum(e);
  __syncthreads();
  if (lane == 0) red[warp] = z;
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * (4 + D);
  if (tid == 0) {
    float zs = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) zs += red[w];
    out[0] = m;
    out[1] = zs;
  }
  // weighted column sums: thread = (float4 column, row phase); 8 independent row loads in flight per thread.
  // (one thread per column walking the 64 rows one after the other took 19 us at L = 9000: 1 TB/s)
  __shared__ float4 colred[256];
  const int nq = D / 4;                               // float4 columns
  const int phases = nq >= 256 ? 1 : 256 / nq;        // row phases sharing a column (nq = 128 -> 2)
  for (int q0 = 0; q0 < nq; q0 += 256 / phases) {
    const int q = q0 + tid % (256 / phases), ph = tid / (256 / phases);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) {
      int i = ph;
      for (; i + 7 * phases < n; i += 8 * phases) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = __ldg(reinterpret_cast<const float4*>(h + (size_t)(r0 + i + u * phases) * D) + q);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float wgt = sw[i + u * phases];
          acc.x = fmaf(wgt, v[u].x, acc.x); acc.y = fmaf(wgt, v[u].y, acc.y);
          acc.z = fmaf(wgt, v[u].z, acc.z); acc.w = fmaf(wgt, v[u].w, acc.w);
        }
      }
      for (; i < n; i += phases) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(h + (size_t)(r0 + i) * D) + q);
        const float wgt = sw[i];
        acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
        acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
      }
    }
    if (phases > 1) {
      colred[tid] = acc;
      __syncthreads();
      if (ph == 0 && q < nq) {
        for (int o = 1; o < phases; ++o) {
          const float4 t = colred[tid + o * (256 / phases)];
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
      }
      __syncthreads();
    }
    if (ph == 0 && q < nq) *reinterpret_cast<float4*>(out + 4 + 4 * q) = acc;
  }
}

// one block; mz[0] = M, mz[1] = Z
__global__ void __launch_bounds__(512) pool_final_kernel(const float* __restrict__ part, int nblocks,
                                                         int D, const float* __restrict__ pred_w,
                                                         const float* __restrict__ pred_b,
                                                         int n_classes, float* __restrict__ pooled,
                                                         float* __restrict__ logits,
                                                         float* __restrict__ mz) {
  extern __shared__ float sm[];  // [nblocks] scale factors | [D] pooled
  float* scale = sm;
  float* pv = sm + nblocks;
  __shared__ float red[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY;
  for (int b = tid; b < nblocks; b += blockDim.x) m = fmaxf(m, part[(size_t)b * (4 + D)]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float z = 0.f;
  for (int b = tid; b < nblocks; b += blockDim.x) {
    float sc = __expf(part[(size_t)b * (4 + D)] - m);
    scale[b] = sc;
    z = fmaf(part[(size_t)b * (4 + D) + 1], sc, z);
  }
  z = warp_sum(z);
  if (lane == 0) red[warp] = z;
  __syncthreads();
  z = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) z += red[w];
  const float inv = 1.f / z;
  __syncthreads();   // scale[] complete
  // pooled[c] = sum_b part[b][c] * scale[b] / Z: thread = (float4 column, block phase), 4 loads in flight
  // (one thread per column over all blocks: 33 us at L = 9000, a single CTA waiting on 141 dependent loads)
  {
    const int nq = D / 4, nthreads = blockDim.x;
    const int lanes_q = nq < nthreads ? nq : nthreads, phases = nthreads / lanes_q;
    __shared__ float4 fred[512];
    for (int q0 = 0; q0 < nq; q0 += lanes_q) {
      const int q = q0 + tid % lanes_q, ph = tid / lanes_q;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < nq && ph < phases) {
        int b = ph;
        for (; b + 3 * phases < nblocks; b += 4 * phases) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v[u] = *reinterpret_cast<const float4*>(part + (size_t)(b + u * phases) * (4 + D) + 4 + 4 * q);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float sc = scale[b + u * phases];
            acc.x = fmaf(v[u].x, sc, acc.x); acc.y = fmaf(v[u].y, sc, acc.y);
            acc.z = fmaf(v[u].z, sc, acc.z); acc.w = fmaf(v[u].w, sc, acc.w);
          }
        }
        for (; b < nblocks; b += phases) {
          const float4 v = *reinterpret_cast<const float4*>(part + (size_t)b * (4 + D) + 4 + 4 * q);
          const float sc = scale[b];
          acc.x = fmaf(v.x, sc, acc.x); acc.y = fmaf(v.y, sc, acc.y);
          acc.z = fmaf(v.z, sc, acc.z); acc.w = fmaf(v.w, sc, acc.w);
        }
      }
      fred[tid] = acc;
      __syncthreads();
      if (ph == 0 && q < nq) {
        for (int o = 1; o < phases; ++o) {
          const float4 t = fred[tid + o * lanes_q];
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        pv[4 * q] = acc.x; pv[4 * q + 1] = acc.y; pv[4 * q + 2] = acc.z; pv[4 * q + 3] = acc.w;   // pv: 4-byte aligned only
        reinterpret_cast<float4*>(pooled)[q] = acc;
      }
      __syncthreads();
    }
  }
  if (tid == 0) { mz[0] = m; mz[1] = z; }
  __syncthreads();
  if (pred_w)
    for (int j = warp; j < n_classes; j += (blockDim.x >> 5)) {
      float d = 0.f;
      for (int c = lane; c < D; c += 32) d = fmaf(pv[c], __ldg(pred_w + (size_t)j * D + c), d);
      d = warp_sum(d);
      if (lane == 0) logits[j] = d + (pred_b ? __ldg(pred_b + j) : 0.f);
    }
}

__global__ void pool_weights_kernel(const float* __restrict__ scores, const float* __restrict__ mz,
                                    float* __restrict__ attn, int L, int raw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  float s = __ldg(scores + i);
  attn[i] = raw ? s : __expf(s - mz[0]) / mz[1];
}
}  // namespace

// hid = row length of the hidden buffer (2 x the attention width for the gated head)
size_t attn_pool_scratch_floats(int L, int D, int hid) {
  size_t nblocks = (L + kPoolRows - 1) / kPoolRows;
  return (size_t)L * hid + (size_t)L + nblocks * (4 + D) + 4;
}

cudaError_t launch_gelu_keep_pre(float* buf, float* pre, size_t rows, int n, int ld, cudaStream_t stream) {
  if (n % 4 || ld % 4 || !pre) return cudaErrorInvalidValue;
  if (rows == 0) return cudaSuccess;
  size_t items = rows * (size_t)(n / 4);
  int blocks = (int)((items + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  gelu_keep_pre_kernel<<<blocks, 256, 0, stream>>>(buf, pre, rows, n / 4, ld / 4);
  return cudaGetLastError();
}

// hidden: [L, ld] fp32 (already activated; ld = hid, or 2 hid when gated); h: [L, D] fp32.  scratch:
// attn_pool_scratch_floats minus the hidden buffer, laid out as scores[L] | part[nblocks][4+D] | mz[4]
cudaError_t launch_attn_pool(const float* h, const float* hidden, const float* w2, const float* b2,
                             const float* pred_w, const float* pred_b, int n_classes, float* scratch,
                             float* pooled, float* logits, float* attn, int attn_raw, int L, int D,
                             int hid, bool gated, cudaStream_t stream) {
  if (L < 1 || D % 4 || n_classes < 0) return cudaErrorInvalidValue;
  const int nblocks = (L + kPoolRows - 1) / kPoolRows;
  float* scores = scratch;
  float* part = scores + (((size_t)L + 3) & ~(size_t)3);
  float* mz = part + (size_t)nblocks * (4 + D);
  pool_scores_kernel<<<(L + 7) / 8, 256, 0, stream>>>(hidden, w2, b2, scores, L, hid, gated ? 2 * hid : hid,
                                                      gated ? 1 : 0);
  pool_partial_kernel<<<nblocks, 256, 0, stream>>>(h, scores, part, L, D);
  size_t smem = ((size_t)nblocks + D) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    static DeviceOnce configured;
    if (configured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(pool_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           200 * 1024);
      if (e != cudaSuccess) return e;
    }
  }
  pool_final_kernel<<<1, 512, smem, stream>>>(part, nblocks, D, pred_w, pred_b, n_classes, pooled, logits, mz);
  if (attn) pool_weights_kernel<<<(L + 255) / 256, 256, 0, stream>>>(scores, mz, attn, L, attn_raw);
  return cudaGetLastError();
}

}  // namespace rrt

// ================================================================================================
// Backward of the pooling head (the RRTMIL train step).  Autograd of
// modules/datten.py:28-38 + modules/rrt.py:241:
//   logits = pooled Wp^T + bp,  pooled = sum_l a_l h_l,  a = softmax_L(s),  s_l = hid_l . w2 + b2,
//   hid = act(h W1^T + b1)
//   dpooled = Wp^T dlogits;  dWp = dlogits (x) pooled;  dbp = dlogits
//   da_l = a_l (h_l . dpooled - pooled . dpooled);  dh_l = a_l dpooled  (+ the path through the score MLP)
//   dw2 = sum_l da_l hid_l;  db2 = sum_l da_l;  dhid_l = da_l w2 * act'(hid_l)
// and dW1 / db1 / the MLP part of dh come from the shared linear-layer backward (tcgen05 GEMMs).
namespace rrt {
namespace {

// one block: dpooled[D] | cdot (= pooled . dpooled) into `out`; dpred_w, dpred_b
__global__ void __launch_bounds__(512) pool_bwd_head_kernel(const float* __restrict__ dlogits,
                                                            const float* __restrict__ pred_w,
                                                            const float* __restrict__ pooled,
                                                            int n_classes, int D, float* __restrict__ out,
                                                            float* __restrict__ dpred_w,
                                                            float* __restrict__ dpred_b) {
  __shared__ float red[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float dot = 0.f;
  for (int c = tid; c < D; c += blockDim.x) {
    float acc = 0.f;
    const float pc = __ldg(pooled + c);
    for (int j = 0; j < n_classes; ++j) {
      const float g = __ldg(dlogits + j);
      acc = fmaf(g, __ldg(pred_w + (size_t)j * D + c), acc);
      dpred_w[(size_t)j * D + c] = g * pc;
    }
    out[c] = acc;
    dot = fmaf(acc, pc, dot);
  }
  if (dpred_b)
    for (int j = tid; j < n_classes; j += blockDim.x) dpred_b[j] = __ldg(dlogits + j);
  dot = warp_sum(dot);
  if (lane == 0) red[warp] = dot;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    out[D] = s;
  }
}

// warp per token (grid-stride).  V = D / 128; hid <= 128 * HV.  act: kActRelu | kActTanh | kActGelu | kActNone.
// hidden holds the values the score layer consumed (after the activation and, when drop.on(), after the MLP's
// nn.Dropout, whose mask is regenerated here: index = l * ld + column, stream RRT_DROP_STREAM_POOL).
// pre: [L, hid] pre-activations, GELU only.  GATED (hid == 128, ld == 256): columns 128.. are the sigmoid branch:
//   s = sum_c a_c b_c w2_c  ->  da_c = ds w2_c b_c,  db_c = ds w2_c a_c,  dz_b = db m_b sig(1 - sig)
template <int V, int HV, bool GATED>
__global__ void __launch_bounds__(256) pool_bwd_rows_kernel(
    const float* __restrict__ h, const float* __restrict__ hidden, const float* __restrict__ scores,
    const float* __restrict__ mz, const float* __restrict__ dp_cdot, const float* __restrict__ w2, int act,
    const float* __restrict__ pre, Dropout drop, float* __restrict__ dh, float* __restrict__ dhid,
    float* __restrict__ dw2, float* __restrict__ db2, uint32_t* __restrict__ amax, int L, int hid, int ld) {
  constexpr int D = 128 * V;
  __shared__ float s_dw2[128 * HV];
  __shared__ float s_db2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wpb = blockDim.x >> 5;
  for (int i = tid; i < hid; i += blockDim.x) s_dw2[i] = 0.f;
  if (tid == 0) s_db2 = 0.f;
  __syncthreads();
  float4 dp[V];
#pragma unroll
  for (int i = 0; i < V; ++i) dp[i] = __ldg(reinterpret_cast<const float4*>(dp_cdot) + lane + 32 * i);
  const float cdot = __ldg(dp_cdot + D), M = __ldg(mz), invZ = 1.f / __ldg(mz + 1);
  float w2r[HV * 4], acc_dw2[HV * 4];
#pragma unroll
  for (int j = 0; j < HV * 4; ++j) {
    const int c = 4 * lane + 128 * (j / 4) + (j & 3);
    w2r[j] = c < hid ? __ldg(w2 + c) : 0.f;
    acc_dw2[j] = 0.f;
  }
  // derivative of (dropout o act) wrt the pre-activation, from the stored value hv, its mask factor m
  // (0 or 1/(1-p); 1 without dropout) and, for GELU, the pre-activation z
  auto dact = [&](float hv, float m, float z) -> float {
    if (act == kActRelu) return hv != 0.f ? m : 0.f;
    if (act == kActTanh) { const float t = m != 0.f ? hv / m : 0.f; return m * (1.f - t * t); }
    if (act == kActGelu) return m * gelu_grad(z);
    return m;
  };
  const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
  float acc_db2 = 0.f, amx = 0.f;
  for (int l = blockIdx.x * wpb + warp; l < L; l += gridDim.x * wpb) {
    const float4* hrow = reinterpret_cast<const float4*>(h + (size_t)l * D);
    float4 hv[V];
    float ds = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      hv[i] = __ldg(hrow + lane + 32 * i);
      ds += hv[i].x * dp[i].x + hv[i].y * dp[i].y + hv[i].z * dp[i].z + hv[i].w * dp[i].w;
    }
    ds = warp_sum(ds);
    const float a = __expf(__ldg(scores + l) - M) * invZ;
    const float da = a * (ds - cdot);
    float4* drow = reinterpret_cast<float4*>(dh + (size_t)l * D);
#pragma unroll
    for (int i = 0; i < V; ++i)
      drow[lane + 32 * i] = make_float4(a * dp[i].x, a * dp[i].y, a * dp[i].z, a * dp[i].w);
    acc_db2 += da;   // identical on every lane
    if (GATED) {
      const int c = 4 * lane;
      const size_t ia = (size_t)l * ld + c, ib = ia + hid;
      const float4 ha4 = __ldg(reinterpret_cast<const float4*>(hidden + ia));
      const float4 hb4 = __ldg(reinterpret_cast<const float4*>(hidden + ib));
      const float4 ma4 = drop.on() ? dropout_scale4(drop, ia) : one4;
      const float4 mb4 = drop.on() ? dropout_scale4(drop, ib) : one4;
      const float4 pz4 = act == kActGelu ? __ldg(reinterpret_cast<const float4*>(pre + (size_t)l * hid + c)) : one4;
      const float ha[4] = {ha4.x, ha4.y, ha4.z, ha4.w}, hb[4] = {hb4.x, hb4.y, hb4.z, hb4.w};
      const float ma[4] = {ma4.x, ma4.y, ma4.z, ma4.w}, mb[4] = {mb4.x, mb4.y, mb4.z, mb4.w};
      const float pz[4] = {pz4.x, pz4.y, pz4.z, pz4.w};
      float oa[4], ob[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc_dw2[e] = fmaf(da, ha[e] * hb[e], acc_dw2[e]);
        const float dg = da * w2r[e];
        oa[e] = dg * hb[e] * dact(ha[e], ma[e], pz[e]);
        const float sg = mb[e] != 0.f ? hb[e] / mb[e] : 0.f;
        ob[e] = dg * ha[e] * mb[e] * sg * (1.f - sg);
        amx = fmaxf(amx, fmaxf(fabsf(oa[e]), fabsf(ob[e])));
      }
      *reinterpret_cast<float4*>(dhid + ia) = make_float4(oa[0], oa[1], oa[2], oa[3]);
      *reinterpret_cast<float4*>(dhid + ib) = make_float4(ob[0], ob[1], ob[2], ob[3]);
    } else {
#pragma unroll
      for (int q = 0; q < HV; ++q) {
        const int c = 4 * lane + 128 * q;
        if (c < hid) {
          const size_t idx = (size_t)l * ld + c;
          const float4 hd = __ldg(reinterpret_cast<const float4*>(hidden + idx));
          const float4 m4 = drop.on() ? dropout_scale4(drop, idx) : one4;
          const float4 pz4 = act == kActGelu ? __ldg(reinterpret_cast<const float4*>(pre + (size_t)l * hid + c)) : one4;
          const float hv4[4] = {hd.x, hd.y, hd.z, hd.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
          const float pz[4] = {pz4.x, pz4.y, pz4.z, pz4.w};
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc_dw2[4 * q + e] = fmaf(da, hv4[e], acc_dw2[4 * q + e]);
            o[e] = da * w2r[4 * q + e] * dact(hv4[e], mm[e], pz[e]);
            amx = fmaxf(amx, fabsf(o[e]));
          }
          *reinterpret_cast<float4*>(dhid + idx) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < HV * 4; ++j) {
    const int c = 4 * lane + 128 * (j / 4) + (j & 3);
    if (c < hid && acc_dw2[j] != 0.f) atomicAdd(&s_dw2[c], acc_dw2[j]);
  }
  if (lane == 0 && acc_db2 != 0.f) atomicAdd(&s_db2, acc_db2);
  amx = warp_max(amx);
  if (lane == 0 && amx > 0.f) atomicMax(amax, __float_as_uint(amx));
  __syncthreads();
  for (int i = tid; i < hid; i += blockDim.x)
    if (s_dw2[i] != 0.f) atomicAdd(dw2 + i, s_dw2[i]);
  if (tid == 0 && db2 && s_db2 != 0.f) atomicAdd(db2, s_db2);
}

// dh[i] += inv_scale(amax) * dz16[i]
__global__ void __launch_bounds__(256) add_scaled_f16_kernel(float* __restrict__ dh, const __half* __restrict__ dz,
                                                             size_t n4, const uint32_t* __restrict__ amax) {
  const uint32_t ab = __ldg(amax);
  uint32_t eb = (ab & 0x7fffffffu) >> 23;
  eb = (eb == 0u || eb >= 255u) ? 135u : (eb < 20u ? 20u : (eb > 240u ? 240u : eb));
  const float inv = __uint_as_float((eb - 8u) << 23);   // == grad_inv_scale (backward.cuh)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dh)[i];
    const float4 b = unpack_h4(__ldg(reinterpret_cast<const uint2*>(dz) + i));
    a.x = fmaf(inv, b.x, a.x); a.y = fmaf(inv, b.y, a.y); a.z = fmaf(inv, b.z, a.z); a.w = fmaf(inv, b.w, a.w);
    reinterpret_cast<float4*>(dh)[i] = a;
  }
}
}  // namespace

cudaError_t launch_pool_bwd_head(const float* dlogits, const float* pred_w, const float* pooled, int n_classes,
                                 int D, float* dp_cdot, float* dpred_w, float* dpred_b, cudaStream_t stream) {
  pool_bwd_head_kernel<<<1, 512, 0, stream>>>(dlogits, pred_w, pooled, n_classes, D, dp_cdot, dpred_w, dpred_b);
  return cudaGetLastError();
}

cudaError_t launch_pool_bwd_rows(const float* h, const float* hidden, const float* scores, const float* mz,
                                 const float* dp_cdot, const float* w2, int act, const float* pre,
                                 const Dropout& drop, bool gated, float* dh, float* dhid, float* dw2, float* db2,
                                 uint32_t* amax, int L, int D, int hid, cudaStream_t stream) {
  if (D % 128 || D > 1024 || hid % 4 || hid > 256) return cudaErrorInvalidValue;
  if (gated && hid != 128) return cudaErrorInvalidValue;
  if (act == kActGelu && !pre) return cudaErrorInvalidValue;
  const int ld = gated ? 2 * hid : hid;
  int blocks = (L + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
#define RRT_PB_ARGS h, hidden, scores, mz, dp_cdot, w2, act, pre, drop, dh, dhid, dw2, db2, amax, L, hid, ld
#define RRT_PB(VV)                                                                        \
  {                                                                                       \
    if (gated) pool_bwd_rows_kernel<VV, 1, true><<<blocks, 256, 0, stream>>>(RRT_PB_ARGS); \
    else if (hid <= 128) pool_bwd_rows_kernel<VV, 1, false><<<blocks, 256, 0, stream>>>(RRT_PB_ARGS); \
    else pool_bwd_rows_kernel<VV, 2, false><<<blocks, 256, 0, stream>>>(RRT_PB_ARGS);     \
  }
  switch (D / 128) {
    case 1: RRT_PB(1) break;
    case 2: RRT_PB(2) break;
    case 4: RRT_PB(4) break;
    case 8: RRT_PB(8) break;
    default: return cudaErrorInvalidValue;
  }
#undef RRT_PB
#undef RRT_PB_ARGS
  return cudaGetLastError();
}

cudaError_t launch_add_scaled_f16(float* dh, const __half* dz, size_t n, const uint32_t* amax,
                                  cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  add_scaled_f16_kernel<<<blocks, 256, 0, stream>>>(dh, dz, n / 4, amax);
  return cudaGetLastError();
}

}  // namespace rrt
