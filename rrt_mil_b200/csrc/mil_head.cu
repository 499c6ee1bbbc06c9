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

__global__ void __launch_bounds__(256) pool_scores_kernel(const float* __restrict__ hidden,
                                                          const float* __restrict__ w2,
                                                          const float* __restrict__ b2,
                                                          float* __restrict__ scores, int L, int hid) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= L) return;
  float d = 0.f;
  for (int c = lane; c < hid; c += 32) d = fmaf(__ldg(hidden + (size_t)row * hid + c), __ldg(w2 + c), d);
  d = warp_sum(d);
  if (lane == 0) scores[row] = d + (b2 ? __ldg(b2) : 0.f);
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
  for (int c = tid * 4; c < D; c += 256 * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n; ++i) {
      float4 v = __ldg(reinterpret_cast<const float4*>(h + (size_t)(r0 + i) * D + c));
      float wgt = sw[i];
      acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
      acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + 4 + c) = acc;
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
  for (int c = tid; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc = fmaf(part[(size_t)b * (4 + D) + 4 + c], scale[b], acc);
    acc *= inv;
    pv[c] = acc;
    pooled[c] = acc;
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

size_t attn_pool_scratch_floats(int L, int D, int hid) {
  size_t nblocks = (L + kPoolRows - 1) / kPoolRows;
  return (size_t)L * hid + (size_t)L + nblocks * (4 + D) + 4;
}

// hidden: [L, hid] fp32 (already activated); h: [L, D] fp32.  scratch: attn_pool_scratch_floats minus
// the hidden buffer, laid out as scores[L] | part[nblocks][4+D] | mz[4]
cudaError_t launch_attn_pool(const float* h, const float* hidden, const float* w2, const float* b2,
                             const float* pred_w, const float* pred_b, int n_classes, float* scratch,
                             float* pooled, float* logits, float* attn, int attn_raw, int L, int D,
                             int hid, cudaStream_t stream) {
  if (L < 1 || D % 4 || n_classes < 0) return cudaErrorInvalidValue;
  const int nblocks = (L + kPoolRows - 1) / kPoolRows;
  float* scores = scratch;
  float* part = scores + (((size_t)L + 3) & ~(size_t)3);
  float* mz = part + (size_t)nblocks * (4 + D);
  pool_scores_kernel<<<(L + 7) / 8, 256, 0, stream>>>(hidden, w2, b2, scores, L, hid);
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
