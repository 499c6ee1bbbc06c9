// Host-side launchers of the RRTEncoder kernels.  Every launcher only enqueues on `stream`.
//
// Data flow of one bag (fp32 = residual stream / statistics, f16 = internal activations, which
// carry the same 10 mantissa bits as tf32):
//   x(fp32) -ln_partition-> z(f16) -gemm-> qkv(f16) -attention-> o(f16) -gemm+residual-> x1(fp32)
//   x1 -stats/logits-> -combine-> landmarks(f16) -gemm-> lqkv(fp32) -attn-> lo(f16) -gemm-> lm'(fp32)
//   x1, lm' -dispatch + final LayerNorm-> out(fp32)
#pragma once
#include "common.cuh"

namespace rrt {

// ---- ln_partition.cu ------------------------------------------------------------------
// z[slot,:] = LayerNorm(x[token(slot),:]) for real tokens, 0 for pad slots; rows in region-major
// slot order (modules/rrt.py:123 + modules/rmsa.py:199-215 in one pass).
cudaError_t launch_ln_partition(const float* x, const float* gamma, const float* beta, __half* z,
                                const Grid& grid, int D, cudaStream_t stream);
// out = LayerNorm(x) row-wise, token order.
cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float* out,
                             int L, int D, cudaStream_t stream);
// out = LN(x1 (+ x0)) -- the tail when cr_msa is off.
cudaError_t launch_add_layernorm(const float* x1, const float* x0, const float* gamma,
                                 const float* beta, float* out, int L, int D, cudaStream_t stream);

// ---- GEMM epilogues -------------------------------------------------------------------
enum GemmEpilogueMode {
  kEpiStore = 0,          // c[m,n] = acc + bias[n]
  kEpiTanh = 1,           // c[m,n] = tanh(acc + bias[n])
  kEpiResidualUnpart = 2, // row m is a region slot: out[token(m),n] = resid[token(m),n] + acc + bias[n]
  // 3 is internal (store + run-time activation)
  kEpiResidualUnpartDrop = 4, // training: ... + dropout(acc + bias[n]) (GemmEpilogue::drop)
  kEpiAtomicAdd = 5           // split-K: c[m,n] += acc (fp32 red.global.add; c zeroed by the caller, no bias)
};
enum GemmActivation { kActNone = 0, kActRelu = 1, kActGelu = 2, kActTanh = 3, kActSigmoid = 4 };
struct GemmEpilogue {
  int mode = kEpiStore;
  int act = kActNone;            // applied to acc + bias (tcgen05 kernel, store mode)
  int act2 = kActNone;           // columns >= act_split (a multiple of 32; 0 = no split) use act2 instead:
  int act_split = 0;             //   the gated pooling head runs [W_a; W_b] as one GEMM (act | sigmoid)
  const float* bias = nullptr;   // [N] or null
  const float* resid = nullptr;  // [L, N] (mode 2)
  Grid grid{};                   // (mode 2)
  Dropout drop{};                // (mode 4) mask index = token * N + n
  // residual modes, optional side output for the CR-MSA block that consumes the rows this GEMM writes:
  // rs_part [L][N/128][8] receives, per token and 128-column part, sum x, sum x^2 and sum_c x_c gamma_c phi[c,n]
  // (n < rs_k <= 4); crmsa finishes LayerNorm statistics and logits from them instead of re-reading the rows.
  // Only when gemm_tcgen05_rowstat_parts(M, N) != 0.
  float* rs_part = nullptr;
  const float* rs_gamma = nullptr;  // [N]  cr_msa.norm.weight
  const float* rs_phi = nullptr;    // [N, rs_k]
  int rs_k = 0;
  // kEpiAtomicAdd, optional: amax word of a scaled fp16 gradient operand (backward.cuh); partial sums are multiplied
  // by its (power-of-two) inverse scale before they are added
  const uint32_t* unscale_amax = nullptr;
};
int gemm_tcgen05_rowstat_parts(int M, int N);

// ---- gemm_tcgen05.cu ------------------------------------------------------------------
// c[M,N] = a[M,K] @ w[N,K]^T (+epilogue) on tcgen05 + TMA + TMEM; fp16 operands, fp32 accumulate.
// c is fp16 (out_f16, kEpiStore only) or fp32.  K % 64 == 0, N % 4 == 0, 16-byte aligned pointers.
bool gemm_tcgen05_supported(int M, int N, int K);
cudaError_t launch_gemm_tcgen05(const __half* a, const __half* w, void* c, bool out_f16, int M,
                                int N, int K, const GemmEpilogue& epi, cudaStream_t stream);
// weight gradient: dw[C_out, C_in] (fp32, zeroed) += dy[rows, C_out]^T @ act[rows, C_in] (fp16, row-major)
// input gradient: d_in[rows, C_in] (fp16) = dy[rows, C_out] (fp16) @ w16[C_out, C_in] (fp16 row-major: MN-major B)
cudaError_t launch_gemm_tcgen05_dgrad(const __half* dy, const __half* w16, __half* d_in, int rows, int C_out,
                                      int C_in, cudaStream_t stream);
// unscale_amax != null: dy is a scaled gradient (backward.cuh); dw receives the unscaled sums
cudaError_t launch_gemm_tcgen05_wgrad(const __half* dy, const __half* act, float* dw, int rows, int C_out,
                                      int C_in, cudaStream_t stream, const uint32_t* unscale_amax = nullptr);
void set_gemm_cluster_mode(int mode);
// at most n SMs for the bag-sized GEMMs launched by THIS host thread from now on (0 = all)
void set_attn_sm_cap(int n);  // persistent-grid cap of the tcgen05 attention kernel (0 = none), per host thread
void set_gemm_sm_cap(int n);  // debug/tuning: 22 = 2x2 clusters, 21 = 2x1, 11 = none
extern long long* g_gemm_trace;  // debug: device buffer [8 CTAs][16] of clock64 stamps, or null
// dst[i] = fp16(src[i]) round-to-nearest, saturating; n % 4 == 0
cudaError_t launch_convert_f16(const float* src, __half* dst, size_t n, cudaStream_t stream);

// dst[i] = fp32(src[i]) for fp16 (bf16 = false) or bf16 rows; n % 4 == 0, src 8-byte / dst 16-byte aligned
cudaError_t launch_widen_f32(const void* src, bool bf16, float* dst, size_t n, cudaStream_t stream);

// ---- gemm_mma.cu ----------------------------------------------------------------------
// fp32-in / fp32-out c = a @ w^T (+epilogue) on mma.sync tf32: the general-purpose linear of the
// C ABI (rrt_linear_forward) for callers whose operands are not fp16.  K % 32 == 0.
cudaError_t launch_gemm_mma(const float* a, const float* w, float* c, int M, int N, int K,
                            const GemmEpilogue& epi, cudaStream_t stream);

// ---- rmsa_attn_f16.cu / rmsa_attn.cu ---------------------------------------------------
// Per (region, head): O = softmax(Q' K^T) V with Q' = scale * (Q + dwconv1d_P(Q; taps_h))
// (modules/rmsa.py:100-122 with the EPEG conv moved onto Q).
// qkv: [Np, 3D] f16, slot order, row layout (3, heads, d).  o: [Np, D] f16, (heads, d).
// taps: [heads, epeg_k] fp32 or null.
// Region-resident kernel (P <= 256) and the flash-style fallback for larger regions.
extern long long* g_attn_trace;  // debug: device buffer [8 CTAs][8] of clock64 stamps, or null
bool rmsa_attention_f16_supported(const Grid& grid, int D, int heads);
cudaError_t launch_rmsa_attention_f16(const __half* qkv, const float* taps, __half* o,
                                      const Grid& grid, int D, int heads, int epeg_k,
                                      cudaStream_t stream);
cudaError_t launch_rmsa_attention(const __half* qkv, const float* taps, __half* o, const Grid& grid,
                                  int D, int heads, int epeg_k, cudaStream_t stream);
// ---- rmsa_attn_tc05.cu: same contract, S and O products on tcgen05 (head_dim 64, P <= 256) ----
extern int g_attn_tc05;  // 1 (default): auto (regions > 128 tokens), 2: wherever supported, 0: never
bool rmsa_attention_tc05_supported(const Grid& grid, int D, int heads, int epeg_k /* 0 = no EPEG */);
cudaError_t launch_rmsa_attention_tc05(const __half* qkv, const float* taps, __half* o,
                                       const Grid& grid, int D, int heads, int epeg_k,
                                       cudaStream_t stream);

// ---- crmsa.cu -------------------------------------------------------------------------
// Per padded slot of the CR-MSA grid: LayerNorm statistics of x1 (mean, rstd; rstd = 0 marks a pad
// slot) and, when phi != null, logits[slot, n] = LN(x1)[slot,:] . phi[:, n].
cudaError_t launch_crmsa_stats_logits(const float* x1, const float* gamma, const float* beta,
                                      const float* phi, float2* stats, float* logits,
                                      const Grid& grid, int D, int k, cudaStream_t stream);
// logits[slot, n] = hidden[slot, :] . w2[n, :]   (crmsa_mlp second layer, no bias)
cudaError_t launch_crmsa_mlp_logits(const float* hidden, const float* w2, float* logits, int Np,
                                    int Dh, int k, cudaStream_t stream);
// Per region: softmax over P / min / max of the logits, landmarks[n, rho, :] = sum_p cw[n,p] z2[p,:]
// (f16: the A operand of the landmark QKV GEMM).  rstat[rho, n] = (min, max).
cudaError_t launch_crmsa_combine(const float* x1, const float* gamma, const float* beta,
                                 const float2* stats, const float* logits, __half* landmarks,
                                 float2* rstat, const Grid& grid, int D, int k,
                                 cudaStream_t stream);
// Fused front end (stats + logits + combine) with one CTA per region; cudaErrorInvalidValue when
// the region does not fit shared memory (callers then use the split kernels above).
// phi == null: `logits` already holds the crmsa_mlp logits and is only read.
bool crmsa_landmarks_supported(const Grid& grid, int D, int k);
cudaError_t launch_crmsa_landmarks(const float* x1, const float* gamma, const float* beta,
                                   const float* phi, float* logits, __half* landmarks,
                                   float2* rstat, const Grid& grid, int D, int k,
                                   cudaStream_t stream);
// Split front end (default): all-SM row statistics / logits kernel + folded combine kernel.
// phi == null: `logits` already holds the crmsa_mlp logits (only the statistics are computed).
cudaError_t launch_crmsa_front_split(const float* x1, const float* gamma, const float* beta,
                                     const float* phi, float2* stats, float* logits,
                                     __half* landmarks, float2* rstat, const Grid& grid, int D, int k,
                                     cudaStream_t stream, const float* rs_part = nullptr, int rs_parts = 0);
// epeg_variants.cu: EPEG ablations (modules/rmsa.py:72-87,104-129), inference only
cudaError_t launch_epeg_value_pe(const __half* qkv, const float* w, const float* bias, __half* pe, const Grid& grid,
                                 int D, int heads, int k, int kw, cudaStream_t stream);
cudaError_t launch_epeg_value_add(__half* dst, int ld, int col0, const __half* pe, int rows, int D,
                                  cudaStream_t stream);
bool rmsa_attention_epeg2d_supported(const Grid& grid, int D, int heads, int k);
cudaError_t launch_rmsa_attention_epeg2d(const __half* qkv, const float* taps, __half* o, const Grid& grid, int D,
                                         int heads, int k, cudaStream_t stream);
// landmark_chain.cu: the whole landmark MHA (QKV projection, attention, output projection) as ONE cluster
// kernel, CTA = head; head_dim 64, heads <= 8.  lm / lo f16 [k*64, D], lout fp32 [k*64, D]; lqkv (nullable):
// f16 [k*64, 3D] copy of the projected q|k|v rows for the training tape.
bool landmark_chain_supported(int k, int D, int heads);
cudaError_t launch_landmark_chain(const __half* lm, const __half* wq, const __half* wp, const float* qkv_b,
                                  const float* proj_b, __half* lqkv, __half* lo, float* lout, int k, int D,
                                  int heads, cudaStream_t stream);
// MHA core over the landmarks: batch = k, sequence = R (64), heads, head_dim = D/heads, plain
// softmax(q k^T * scale) v, fp32 math.  lqkv: [k*R, 3D] fp32 rows (n, rho); lo: [k*R, D] f16.
cudaError_t launch_landmark_attention(const float* lqkv, __half* lo, int k, int R, int D, int heads,
                                      cudaStream_t stream);
// out[t,:] = LN_final( x1[t,:] + sum_n w[t,n] * lm[n, rho(t), :] (+ x0[t,:]) )   (LN optional)
cudaError_t launch_crmsa_dispatch(const float* x1, const float* x0, const float* logits,
                                  const float2* rstat, const float* lm, const float* gamma,
                                  const float* beta, float* out, const Grid& grid, int D, int k,
                                  cudaStream_t stream);

// ---- peg.cu: PEG / PPEG depthwise-conv positional encodings ------------------
// out[L,D] = x + sum_j conv_kj(grid(x)) (+ biases); w/b: the 1 (PEG) or 3 (PPEG: k, 5, 3) reference
// Conv2d weights [D,1,k,k] ([D,1,k,1] with conv_1d) and biases (nullable).  scratch: peg_scratch_floats.
size_t peg_scratch_floats(int D, int peg_k, bool ppeg, bool conv_1d);
cudaError_t launch_peg(const float* x, float* out, int L, int D, int peg_k, bool ppeg, bool conv_1d,
                       const float* const* w, const float* const* b, float* scratch, cudaStream_t stream);
// backward of launch_peg: dx (+ dres), dw[j] / db[j] in the reference's parameter layout; weff = the forward's
// folded kernel (its scratch), scratch = another peg_scratch_floats floats
cudaError_t launch_peg_backward(const float* x, const float* dy, const float* dres, float* dx, int L, int D,
                                int peg_k, bool ppeg, bool conv_1d, const float* weff, float* scratch,
                                float* const* dw, float* const* db, cudaStream_t stream);

// ---- optim.cu: multi-tensor Adam / AdamW step, torch.optim semantics ----------
cudaError_t launch_adam(float* const* p, const float* const* g, float* const* m, float* const* v,
                        const long long* n, int count, float lr, float beta1, float beta2, float eps,
                        float wd, bool decoupled, long long step, float grad_scale, int* launches,
                        cudaStream_t stream, const float* bc_dev = nullptr);

// ---- mil_head.cu: DAttention pooling + predictor behind the encoder ---------
size_t attn_pool_scratch_floats(int L, int D, int hid);
cudaError_t launch_attn_pool(const float* h, const float* hidden, const float* w2, const float* b2,
                             const float* pred_w, const float* pred_b, int n_classes, float* scratch,
                             float* pooled, float* logits, float* attn, int attn_raw, int L, int D,
                             int hid, bool gated, cudaStream_t stream);
// buf[r, 0:n] (row pitch ld) holds pre-activations: copy them to pre [rows, n] and apply nn.GELU in place
cudaError_t launch_gelu_keep_pre(float* buf, float* pre, size_t rows, int n, int ld, cudaStream_t stream);

// backward of the pooling head (mil_head.cu): dp_cdot = dpooled[D] | pooled . dpooled
cudaError_t launch_pool_bwd_head(const float* dlogits, const float* pred_w, const float* pooled, int n_classes,
                                 int D, float* dp_cdot, float* dpred_w, float* dpred_b, cudaStream_t stream);
cudaError_t launch_pool_bwd_rows(const float* h, const float* hidden, const float* scores, const float* mz,
                                 const float* dp_cdot, const float* w2, int act, const float* pre,
                                 const Dropout& drop, bool gated, float* dh, float* dhid, float* dw2, float* db2,
                                 uint32_t* amax, int L, int D, int hid, cudaStream_t stream);
cudaError_t launch_add_scaled_f16(float* dh, const __half* dz, size_t n, const uint32_t* amax,
                                  cudaStream_t stream);

// ---- backward pass (backward.cu, rmsa_attn_bwd.cu, crmsa_bwd.cu; conventions in backward.cuh) ----
// amax words: IEEE bits of max|g| of a stage's fp32 input gradient (zero-initialised, atomicMax);
// consumers derive the power-of-two fp16 scale S from them.
cudaError_t launch_amax(const float* x, size_t n, uint32_t* amax, cudaStream_t stream);
// g: fp32 [grid.L, C] token order -> rows: fp16 [M, C] slot order, * S(amax), pad slots 0 (grid.H == 0:
// identity, M == grid.L); rowsT: fp16 [C, M64] transpose; colsum[c] += sum of the unscaled column
// drop.on(): g is the gradient wrt the OUTPUT of a dropout; it is multiplied by the forward's mask first
// mask_src (nullable, [grid.L, C]): g is multiplied by mask_scale where mask_src != 0 and by 0 elsewhere (the
// backward of ReLU followed by dropout, read off the layer's OUTPUT: out != 0 <=> pre-activation > 0 and kept)
cudaError_t launch_grad_partition(const float* g, const Grid& grid, int M, int C, const uint32_t* amax,
                                  __half* rows, __half* rowsT, float* colsum, cudaStream_t stream,
                                  const Dropout& drop = Dropout{}, const float* mask_src = nullptr,
                                  float mask_scale = 1.f, int mask_mode = 0);
// out = a (+ b), fp32, n % 4 == 0
cudaError_t launch_add2(float* out, const float* a, const float* b, size_t n, cudaStream_t stream);
// x[i] *= mask(i) / (1-p) in place, x fp32 [rows, C] (the landmark projection output in training mode)
cudaError_t launch_dropout_inplace(float* x, size_t n, const Dropout& drop, cudaStream_t stream);
// mask(i)/(1-p) of n elements as fp32 (parity tests: the oracle consumes the same mask)
cudaError_t launch_dropout_mask(float* out, size_t n, const Dropout& drop, cudaStream_t stream);
// in: fp16 [M, C] -> outT fp16 [C, M64] (columns >= M zero); colsum[c] += column sum * 1/S(amax)
cudaError_t launch_transpose_f16(const __half* in, int M, int C, __half* outT, float* colsum,
                                 const uint32_t* amax, cudaStream_t stream);
// LayerNorm backward of y = LN(x (+ x_add)); dz fp16 slot-order scaled (dz_f16) or fp32 token order
cudaError_t launch_ln_backward(const float* x, const float* x_add, const float* gamma, const void* dz,
                               bool dz_f16, const uint32_t* amax_in, const float* dres,
                               const float* dres2, float* dx, float* dgamma, float* dbeta,
                               uint32_t* amax_out, const Grid& grid, int D, cudaStream_t stream);
cudaError_t launch_wt_convert(const float* w, __half* wT, int N, int K, cudaStream_t stream);
// attention core backward: qkv/o as written by the forward, dO fp16 [R*P, D] (scaled) ->
// dqkv fp16 [R*P, 3D] (scaled), dtaps[heads, epeg_k] += unscaled tap gradients (taps may be null)
bool rmsa_attention_bwd_supported(int P, int D, int heads, int epeg_k);
cudaError_t launch_rmsa_attention_bwd(const __half* qkv, const __half* o, const __half* dO,
                                      const float* taps, __half* dqkv, float* dtaps,
                                      const uint32_t* amax, int R, int P, int D, int heads, int epeg_k,
                                      cudaStream_t stream);
bool crmsa_backward_supported(int D, int k);
cudaError_t launch_crmsa_dispatch_bwd(const float* x1, const float* x0, const float* logits,
                                      const float2* rstat, const float* lmp, const float* gamma_f,
                                      const float* dout, float* dh, float* dw, float* dLp,
                                      float2* rgrad, float* dgamma_f, float* dbeta_f, const Grid& grid,
                                      int D, int k, cudaStream_t stream);
cudaError_t launch_crmsa_combine_bwd(const float* x1, const float* gamma, const float* beta,
                                     const float* phi, const float* logits, const float2* rstat,
                                     const __half* lm16, const __half* dlm16, const uint32_t* amax_l,
                                     const float* dw, const float2* rgrad, const float* dh,
                                     float dh_weight, float* dx1, float* dphi, float* dgamma,
                                     float* dbeta, uint32_t* amax_out, const Grid& grid, int D, int k,
                                     cudaStream_t stream, int mode = 0, float* dlogits = nullptr,
                                     const __half* dzx16 = nullptr, const uint32_t* amax_x = nullptr);
// backward of launch_landmark_attention (fp32 tape, any head_dim % 32 == 0): dqkv16 [k*64, 3D] fp16 in the scaled
// gradient domain of dO16
cudaError_t launch_landmark_attention_bwd(const float* lqkv, const __half* dO16, __half* dqkv16, int k, int R, int D,
                                          int heads, cudaStream_t stream);
// crmsa_mlp: backward of logits = W2 tanh(pre) (dpre fp32 [rows, H4], dW2 [k, H4] accumulated); mode 1 / 2 of
// launch_crmsa_combine_bwd run before / after it (csrc/crmsa_bwd.cu)
cudaError_t launch_crmsa_mlp_hidden_bwd(const float* dlogits, const float* hidden, const float* w2, float* dpre,
                                        float* dw2, int rows, int H4, int k, cudaStream_t stream);

}  // namespace rrt
