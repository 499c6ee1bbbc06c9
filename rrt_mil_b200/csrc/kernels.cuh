// Host-side launchers of the RRTEncoder kernels.  Every launcher only enqueues on `stream`.
#pragma once
#include "common.cuh"

namespace rrt {

// ---- ln_partition.cu ------------------------------------------------------------------
// z[slot,:] = LayerNorm(x[token(slot),:]) for real tokens, 0 for pad slots; rows in region-major
// slot order (modules/rrt.py:123 + modules/rmsa.py:199-215 in one pass).
cudaError_t launch_ln_partition(const float* x, const float* gamma, const float* beta, float* z,
                                const Grid& grid, int D, bool round_tf32, cudaStream_t stream);
// out = LayerNorm(x) row-wise, token order.
cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float* out,
                             int L, int D, cudaStream_t stream);

// ---- gemm_mma.cu ----------------------------------------------------------------------
enum GemmEpilogueMode {
  kEpiStore = 0,        // c[m,n] = acc + bias[n]
  kEpiTanh = 1,         // c[m,n] = tanh(acc + bias[n])
  kEpiResidualUnpart = 2  // row m is a region slot: out[token(m),n] = resid[token(m),n] + acc + bias[n]
};
struct GemmEpilogue {
  int mode = kEpiStore;
  const float* bias = nullptr;   // [N] or null
  const float* resid = nullptr;  // [L, N] (mode 2)
  Grid grid{};                   // (mode 2)
};
// c[M,N] (ld = N) = a[M,K] @ w[N,K]^T (+epilogue).  K % 32 == 0.
cudaError_t launch_gemm_mma(const float* a, const float* w, float* c, int M, int N, int K,
                            const GemmEpilogue& epi, cudaStream_t stream);

// ---- gemm_tcgen05.cu ------------------------------------------------------------------
// Same contract as launch_gemm_mma (modes kEpiStore / kEpiResidualUnpart) on tcgen05 + TMA + TMEM.
// a and w must hold tf32-representable values (launch_round_tf32 / producing kernels round).
bool gemm_tcgen05_supported(int M, int N, int K);
cudaError_t launch_gemm_tcgen05(const float* a, const float* w, float* c, int M, int N, int K,
                                const GemmEpilogue& epi, cudaStream_t stream);
// dst[i] = round-to-nearest tf32 of src[i] (kept in fp32 containers); n % 4 == 0
cudaError_t launch_round_tf32(const float* src, float* dst, size_t n, cudaStream_t stream);

// ---- rmsa_attn.cu ---------------------------------------------------------------------
// Per (region, head): O = softmax(Q' K^T) V with Q' = scale * (Q + dwconv1d_P(Q; taps_h))
// (modules/rmsa.py:100-122 with the EPEG conv moved onto Q).
// qkv: [Np, 3D] slot order, row layout (3, heads, d).  o: [Np, D] slot order, (heads, d).
// taps: [heads, epeg_k] or null.
// round_out: store O rounded to tf32 (it is the A operand of the tcgen05 projection GEMM).
cudaError_t launch_rmsa_attention(const float* qkv, const float* taps, float* o, const Grid& grid,
                                  int D, int heads, int epeg_k, bool round_out,
                                  cudaStream_t stream);

// ---- rmsa_attn_f16.cu -----------------------------------------------------------------
// Same contract for regions of <= 256 tokens: whole region resident in smem as fp16 (10-bit mantissa,
// as tf32), ldmatrix + mma.sync m16n8k16, one warp per 16 query rows.
bool rmsa_attention_f16_supported(const Grid& grid, int D, int heads);
cudaError_t launch_rmsa_attention_f16(const float* qkv, const float* taps, float* o,
                                      const Grid& grid, int D, int heads, int epeg_k,
                                      bool round_out, cudaStream_t stream);

// ---- crmsa.cu -------------------------------------------------------------------------
// Per padded slot of the CR-MSA grid: LayerNorm statistics of x1 (mean, rstd; rstd = 0 marks a pad
// slot) and, when phi != null, logits[slot, n] = LN(x1)[slot,:] . phi[:, n].
cudaError_t launch_crmsa_stats_logits(const float* x1, const float* gamma, const float* beta,
                                      const float* phi, float2* stats, float* logits,
                                      const Grid& grid, int D, int k, cudaStream_t stream);
// logits[slot, n] = hidden[slot, :] . w2[n, :]   (crmsa_mlp second layer, no bias)
cudaError_t launch_crmsa_mlp_logits(const float* hidden, const float* w2, float* logits, int Np,
                                    int Dh, int k, cudaStream_t stream);
// Per region: softmax over P / min / max of the logits, landmarks[n, rho, :] = sum_p cw[n,p] z2[p,:].
// rstat[rho, n] = (min, max).
cudaError_t launch_crmsa_combine(const float* x1, const float* gamma, const float* beta,
                                 const float2* stats, const float* logits, float* landmarks,
                                 float2* rstat, const Grid& grid, int D, int k, bool round_out,
                                 cudaStream_t stream);
// MHA core over the landmarks: batch = k, sequence = R (64), heads, head_dim = D/heads, plain
// softmax(q k^T * scale) v.  lqkv: [k*R, 3D] rows (n, rho); lo: [k*R, D].
cudaError_t launch_landmark_attention(const float* lqkv, float* lo, int k, int R, int D, int heads,
                                      bool round_out, cudaStream_t stream);
// out[t,:] = LN_final( x1[t,:] + sum_n w[t,n] * lm[n, rho(t), :] (+ x0[t,:]) )   (LN optional)
cudaError_t launch_crmsa_dispatch(const float* x1, const float* x0, const float* logits,
                                  const float2* rstat, const float* lm, const float* gamma,
                                  const float* beta, float* out, const Grid& grid, int D, int k,
                                  cudaStream_t stream);
// out = LN(x1 (+ x0)) -- the tail when cr_msa is off.
cudaError_t launch_add_layernorm(const float* x1, const float* x0, const float* gamma,
                                 const float* beta, float* out, int L, int D, cudaStream_t stream);

}  // namespace rrt
