/*
 * rrt_b200.h -- C ABI of the B200-native RRTEncoder hot path (librrt_b200.so).
 *
 * The reference (DearCaat/RRT-MIL) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it offers is the nn.Module call  RRTEncoder(...).forward(x:[1,N,D]) -> [1,N,D]
 * (/root/reference/modules/rrt.py:133-202).  This header is what a binding for that call
 * binds instead of the ATen operator sequence underneath it: plain pointers and sizes, no
 * torch types.  rrt_mil_b200/cabi.py is the ctypes binding; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every pointer named *_dev / x / out / weights is a DEVICE pointer (fp32, row-major,
 *     channel fastest) unless the name says host; `stream` is a cudaStream_t passed as void*;
 *   - calls only enqueue work on `stream`; they never synchronise the device (the *_host entry
 *     point synchronises `stream` once, at the end, because its result lives in host memory);
 *   - the caller owns every buffer, including the workspace (size from rrt_workspace_bytes);
 *   - return value 0 = ok, negative = error (RRT_E_*), text from rrt_last_error();
 *   - there is no CPU fallback: without a CUDA device every compute entry returns RRT_E_CUDA.
 *
 * Threading / devices
 *   - every entry point may be called from several host threads at once, on different streams and on different
 *     devices of one process (cudaSetDevice first: a call runs on the device that is current in the calling
 *     thread, and `stream` and all pointers must belong to it).  The library keeps no state that is shared between
 *     calls except caches keyed by (thread, device): the internal lane streams / events of
 *     rrt_encoder_forward_batch and the encoded TMA descriptors are per host thread and per device;
 *   - the rrt_debug_* knobs and the stage timer are process-wide and meant for single-threaded tools;
 *   - rrt_last_error() is per host thread.
 */
#ifndef RRT_B200_H_
#define RRT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RRT_API
#else
#define RRT_API __attribute__((visibility("default")))
#endif

#define RRT_ABI_VERSION 8
#define RRT_MAX_RMSA_LAYERS 8 /* n_layers-1 R-MSA TransLayers (modules/rrt.py:143) */
#define RRT_MAX_CRMSA_K 16    /* crmsa_k landmarks per region                         */
#define RRT_MAX_EPEG_K 63     /* odd EPEG kernel length                               */
#define RRT_MAX_LANES 8       /* concurrent bags inside rrt_encoder_forward_batch     */

enum {
  RRT_OK = 0,
  RRT_E_INVALID = -1,     /* bad argument / unsupported configuration */
  RRT_E_WORKSPACE = -2,   /* workspace too small                      */
  RRT_E_CUDA = -3         /* CUDA runtime error (text in rrt_last_error) */
};

/* math_mode: what the tensor-core contractions compute in.  I/O is always fp32. */
enum {
  RRT_MATH_F16 = 0 /* fp16 operands (10 mantissa bits, like tf32), fp32 accumulate / softmax /
                      LayerNorm / residual stream: meets the 1e-3 rel "fp32" bar of the north star */
};

/* Constructor options of RRTEncoder that reach the hot path
 * (modules/rrt.py:134; same names, same defaults on the Python side). */
typedef struct rrt_config {
  int32_t dim;              /* mlp_dim; multiple of 128, <= 1024                       */
  int32_t n_rmsa_layers;    /* n_layers - 1                                            */
  int32_t n_heads;          /* heads of the R-MSA layers                               */
  int32_t region_num;       /* regions per side of the R-MSA grid                      */
  int32_t region_size;      /* >0: fixed region side instead (modules/rmsa.py:177-182) */
  int32_t min_region_num;   /* "give up region attention" escape, modules/rmsa.py:193  */
  double min_region_ratio;  /*   "                                                     */
  int32_t epeg;             /* EPEG conv on the logit map on/off                       */
  int32_t epeg_k;           /* odd kernel length                                       */
  int32_t qkv_bias;
  int32_t cr_msa;           /* CR-MSA layer on/off                                     */
  int32_t crmsa_k;
  int32_t crmsa_heads;
  int32_t crmsa_mlp;        /* phi = Linear-tanh-Linear instead of a [D,k] matrix      */
  int32_t all_shortcut;
  int32_t math_mode;        /* RRT_MATH_*                                              */
  /* ablation positional encoding (modules/rrt.py:150-160, modules/emb_position.py:24-82) */
  int32_t pos;              /* RRT_POS_NONE / RRT_POS_PEG / RRT_POS_PPEG                */
  int32_t pos_pos;          /* -1: before the first layer; 0: before R-MSA layer 1      */
  int32_t peg_k;            /* odd kernel size of the first conv                        */
  int32_t peg_1d;           /* (k,1) column kernels instead of k x k                    */
  /* ablation FFN of every TransLayer (modules/rrt.py:25-41,106,128-129): x += fc2(act(fc1(norm2(x)))) */
  int32_t ffn;              /* on/off                                                   */
  int32_t ffn_act;          /* RRT_ACT_GELU (ffn_act='gelu', default) or RRT_ACT_RELU    */
  int32_t ffn_hidden;       /* int(dim * mlp_ratio); multiple of 64                     */
  /* EPEG ablation variants (modules/rmsa.py:72-87,104-129); inference only */
  int32_t epeg_type;        /* RRT_EPEG_ATTN (default) / RRT_EPEG_VALUE_BF / RRT_EPEG_VALUE_AF */
  int32_t epeg_2d;          /* k x k kernel instead of (k, 1)                           */
} rrt_config;
enum { RRT_EPEG_ATTN = 0, RRT_EPEG_VALUE_BF = 1, RRT_EPEG_VALUE_AF = 2 };
enum { RRT_POS_NONE = 0, RRT_POS_PEG = 1, RRT_POS_PPEG = 2 };

/* One InnerAttention (modules/rmsa.py:56-89).  qkv_b may be NULL (qkv_bias=False);
 * pe_w is the depthwise EPEG kernel (= pe.weight, contiguous) or NULL: [heads, epeg_k] for the default
 * epeg_type 'attn' ([heads, k, k] with epeg_2d), [D, epeg_k] resp. [D, k, k] for the value variants.
 * pe.bias is not needed for 'attn' (constant along the softmax axis); the value variants
 * read it from pe_b ([D] or NULL). */
typedef struct rrt_attn_weights {
  const float* qkv_w;  /* [3D, D] */
  const float* qkv_b;  /* [3D] or NULL */
  const float* proj_w; /* [D, D] */
  const float* proj_b; /* [D] */
  const float* pe_w;   /* [heads, epeg_k] or NULL */
  /* Optional fp16 shadows of qkv_w / proj_w (rrt_convert_f16), the form the tcgen05 GEMMs
   * consume.  NULL: the library converts into its workspace on every call (correct, a few us
   * slower).  The Python binding keeps shadows and refreshes them when a parameter changes. */
  const void* qkv_w_f16;  /* [3D, D] fp16 */
  const void* proj_w_f16; /* [D, D] fp16 */
  const float* pe_b;      /* [D] or NULL: pe.bias of the value variants of EPEG */
} rrt_attn_weights;

/* One Mlp + its pre-norm (TransLayer.norm2 / TransLayer.mlp, modules/rrt.py:47,106); all NULL when ffn = 0. */
typedef struct rrt_ffn_weights {
  const float* norm_w;  /* norm2.weight [D] */
  const float* norm_b;
  const float* fc1_w;   /* mlp.fc1.weight [H, D] */
  const float* fc1_b;   /* [H] */
  const float* fc2_w;   /* mlp.fc2.weight [D, H] */
  const float* fc2_b;   /* [D] */
  const void* fc1_w_f16; /* optional fp16 shadows */
  const void* fc2_w_f16;
} rrt_ffn_weights;

/* state_dict of one RRTEncoder, by reference name. */
typedef struct rrt_weights {
  const float* norm_w; /* norm.weight [D] (final LayerNorm) */
  const float* norm_b;
  const float* layer_norm_w[RRT_MAX_RMSA_LAYERS]; /* layers.i.norm.weight */
  const float* layer_norm_b[RRT_MAX_RMSA_LAYERS];
  rrt_attn_weights layer_attn[RRT_MAX_RMSA_LAYERS]; /* layers.i.attn.attn.* */
  const float* cr_norm_w; /* cr_msa.norm.* */
  const float* cr_norm_b;
  const float* cr_phi;    /* cr_msa.attn.phi [D, k]           (crmsa_mlp = 0) */
  const float* cr_phi_w1; /* cr_msa.attn.phi.0.weight [D/4,D] (crmsa_mlp = 1) */
  const float* cr_phi_w2; /* cr_msa.attn.phi.2.weight [k,D/4] (crmsa_mlp = 1) */
  const void* cr_phi_w1_f16; /* optional fp16 shadow of cr_phi_w1 */
  rrt_attn_weights cr_attn; /* cr_msa.attn.attn.* (pe_w NULL) */
  /* pos_embedding.proj / proj1 / proj2 .weight [D,1,k,k] ([D,1,k,1] with peg_1d) and .bias [D] or NULL;
   * PEG uses index 0 only, PPEG all three (kernel sizes peg_k, 5, 3) */
  const float* pos_w[3];
  const float* pos_b[3];
  rrt_ffn_weights layer_ffn[RRT_MAX_RMSA_LAYERS]; /* layers.i.norm2.* / layers.i.mlp.* */
  rrt_ffn_weights cr_ffn;                          /* cr_msa.norm2.* / cr_msa.mlp.*     */
} rrt_weights;

/* ---- housekeeping ------------------------------------------------------------------- */
RRT_API int rrt_abi_version(void);
RRT_API const char* rrt_last_error(void);

/* Padded square grid of a bag of L tokens: H = side, rs = region side
 * (modules/rmsa.py:175-198).  Pure host arithmetic. */
RRT_API int rrt_grid_geometry(int64_t L, int32_t region_num, int32_t region_size,
                              int32_t min_region_num, double min_region_ratio, int32_t* H,
                              int32_t* rs);

/* Bytes of device workspace rrt_encoder_forward needs for a bag of L tokens. */
RRT_API int rrt_workspace_bytes(const rrt_config* cfg, int64_t L, size_t* bytes);

/* ---- the hot path ------------------------------------------------------------------- */
/* RRTEncoder.forward for one bag, eval mode: x [L, D] -> out [L, D]
 * (modules/rrt.py:165-202).  x and out may not alias. */
RRT_API int rrt_encoder_forward(const rrt_config* cfg, const rrt_weights* w, const float* x,
                                float* out, int64_t L, void* workspace, size_t workspace_bytes,
                                void* stream);

/* n_bags independent bags (one bag per forward, results exactly as n_bags calls of
 * rrt_encoder_forward).  xs / outs / Ls are HOST arrays of device pointers / lengths.
 * Bags are independent, so when the workspace holds room for several bags
 * (workspace_bytes >= lanes * rrt_workspace_bytes(longest bag), lanes <= RRT_MAX_LANES) the library
 * runs `lanes` bags concurrently on internal streams that fork from and join back into `stream`
 * (event fork/join: legal inside CUDA-graph capture); the small latency-bound kernels of one bag
 * then fill the SMs the other bags leave idle.  With room for one bag the bags run back to back. */
RRT_API int rrt_encoder_forward_batch(const rrt_config* cfg, const rrt_weights* w,
                                      const float* const* xs, float* const* outs,
                                      const int64_t* Ls, int32_t n_bags, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* Same call with HOST bags (pinned or pageable): copies x_host to x_dev, runs the forward,
 * copies the result back to out_host, all on `stream`, then synchronises `stream`.
 * x_dev / out_dev are caller-owned device staging buffers of L*D floats each. */
RRT_API int rrt_encoder_forward_host(const rrt_config* cfg, const rrt_weights* w,
                                     const float* x_host, float* out_host, float* x_dev,
                                     float* out_dev, int64_t L, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* ---- the blocks of the path, exported so each can be checked against the oracle -------- */
/* x1 = x + RegionAttntion(LayerNorm(x))   (modules/rrt.py:123-125, modules/rmsa.py:204-230) */
RRT_API int rrt_rmsa_block_forward(const rrt_config* cfg, const float* norm_w,
                                   const float* norm_b, const rrt_attn_weights* attn,
                                   const float* x, float* x1, int64_t L, void* workspace,
                                   size_t workspace_bytes, void* stream);

/* y = x1 + CrossRegionAttntion(LayerNorm(x1)) (+ x0 if cfg->all_shortcut and x0 != NULL);
 * out = LayerNorm(y; final_norm) if final_norm_w != NULL else y
 * (modules/rrt.py:190-195, modules/rmsa.py:290-337). */
RRT_API int rrt_crmsa_block_forward(const rrt_config* cfg, const rrt_weights* w,
                                    const float* x1, const float* x0, float* out, int64_t L,
                                    int32_t apply_final_norm, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* Debug: when non-NULL, the first 8 CTAs of every tcgen05 GEMM launch write clock64 stamps of their
 * pipeline phases into device_buffer[launch % 8][8][16] (int64).  NULL switches the trace off. */
RRT_API int rrt_debug_set_gemm_trace(void* device_buffer);

/* Debug: same for the region-resident attention kernel: device_buffer[8][8] (int64). */
RRT_API int rrt_debug_set_attn_trace(void* device_buffer);

/* Debug / tuning: R-MSA attention core.  1 (default) = auto: the tcgen05 / TMEM kernel (head_dim 64) for regions
 * of more than 128 tokens, the mma.sync kernel otherwise; 2 = the tcgen05 kernel wherever it is supported;
 * 0 = the mma.sync kernel only.  Results agree within the parity tolerance. */
RRT_API int rrt_debug_set_attention_kernel(int32_t mode);

/* Debug / tuning: kernel variant of the bag-sized tcgen05 GEMMs.  11 = single-CTA 128x256 tiles
 * (default, fastest at these sizes); 2 = CTA pairs (cta_group::2, M=256 tiles); 21 / 22 = single-CTA
 * tiles with 2x1 / 2x2 cluster TMA multicast; 128 / 256 = tile width of the GEMMs that have at most one
 * 256-column tile per CTA (proj; default 256).  Results do not depend on it. */
RRT_API int rrt_debug_set_gemm_cluster(int32_t mode);

/* Debug / measurement only: bit i set = the kernels of stage i (rrt_stage_name order) are NOT launched.
 * Results are then wrong by construction; tools/marginal_cost.py uses it to time what each stage costs
 * while several bags are in flight (where per-kernel intervals overlap and cannot be summed). */
RRT_API int rrt_debug_skip_stages(uint32_t mask);

/* dst[i] = fp16(src[i]), round to nearest, saturating at +-65504.  dst is n fp16 values; n % 4 == 0. */
RRT_API int rrt_convert_f16(const float* src, void* dst, int64_t n, void* stream);

/* dst[i] = fp32(src[i]) for n fp16 (src_is_bf16 = 0) or bf16 values, n % 4 == 0.  The encoder's I/O contract is
 * fp32; a host running under fp16 / bf16 autocast (the reference's --amp, main.py:101-102,439) hands it half rows,
 * which are widened here (exact) before the path runs. */
RRT_API int rrt_widen_f32(const void* src, int32_t src_is_bf16, float* dst, int64_t n, void* stream);

/* The tcgen05/TMA/TMEM GEMM every linear layer of the path runs on, exposed for parity tests:
 * c[M,N] (fp32) = a[M,K] (fp16) @ w[N,K]^T (fp16) + bias[N] (fp32, may be NULL).
 * 16-byte aligned pointers; K % 64 == 0, N % 4 == 0. */
RRT_API int rrt_linear_f16_forward(const void* a_f16, const void* w_f16, const float* bias, float* c,
                                   int64_t M, int32_t N, int32_t K, void* stream);

/* c[M,N] = a[M,K] @ w[N,K]^T + bias[N]   (nn.Linear; bias may be NULL).  K % 32 == 0. */
RRT_API int rrt_linear_forward(const float* a, const float* w, const float* bias, float* c,
                               int64_t M, int32_t N, int32_t K, void* stream);

/* ---- SURVEY.md 8(f) "next" rows f1 / f2: the layers RRTMIL wraps around the encoder ------------ */
/* activation codes of the two entries below */
enum { RRT_ACT_NONE = 0, RRT_ACT_RELU = 1, RRT_ACT_GELU = 2, RRT_ACT_TANH = 3 };
/* flag OR-ed into `act` of rrt_attn_pool_forward / _backward: the gated head (AttentionGated,
 * modules/datten.py:40-83).  w1 is then [2*hid, dim] = [attention_a.0.weight; attention_b.0.weight], b1
 * [2*hid] likewise, w2 = attention_c.weight [hid]:  A = (act(h Wa^T + ba) * sigmoid(h Wb^T + bb)) w2 + b2.
 * Workspace sizes are asked for with 2*hid in the `hid` argument. */
#define RRT_ACT_GATED 0x100

/* Device workspace (bytes) of rrt_patch_embed_forward / rrt_attn_pool_forward (hid = rows of w1). */
RRT_API int rrt_mil_head_workspace_bytes(int64_t L, int32_t in_dim, int32_t dim, int32_t hid,
                                         size_t* bytes);

/* patch_to_emb: out[L, out_dim] = act(x[L, in_dim] @ w[out_dim, in_dim]^T + b)   (modules/rrt.py:208-217,
 * 228).  w_f16 = optional fp16 shadow of w (rrt_convert_f16) or NULL.  in_dim % 64 == 0.
 * drop_p / seed: RRTMIL.dp (nn.Dropout(0.25), modules/rrt.py:215,229) in training mode, counter-based mask
 * stream RRT_DROP_STREAM_PATCH over [L, out_dim]; 0 = eval.  After the call the workspace starts with the
 * fp16 copy of x, which rrt_patch_embed_backward reads as its tape.
 * pre (nullable, [L, out_dim] fp32): with act = RRT_ACT_GELU it receives the pre-activations x w^T + b, which
 * the backward needs (gelu' is not a function of the output); ignored for the other activations. */
#define RRT_DROP_STREAM_PATCH 65
RRT_API int rrt_patch_embed_forward(const float* x, int64_t L, int32_t in_dim, int32_t out_dim,
                                    const float* w, const float* b, const void* w_f16, int32_t act,
                                    float* out, void* workspace, size_t workspace_bytes, float drop_p,
                                    uint64_t seed, float* pre, void* stream);
/* Backward of patch_to_emb (+ dp): dout [L, out_dim] = gradient wrt the forward's `out`; dw [out_dim, in_dim],
 * db [out_dim] (nullable) are overwritten.  act = RRT_ACT_RELU (mask read off `out`), RRT_ACT_GELU (needs
 * `pre` of the forward; dropout mask regenerated) or RRT_ACT_NONE (dropout mask regenerated from drop_p /
 * seed).  tape = the forward's workspace, untouched since.  No gradient wrt x (the bag's features are data).
 * workspace >= 512 + L * out_dim * 2 bytes (256-aligned). */
RRT_API int rrt_patch_embed_backward(const float* dout, const float* out, const float* pre, int64_t L,
                                     int32_t in_dim, int32_t out_dim, int32_t act, float drop_p, uint64_t seed,
                                     const void* tape, size_t tape_bytes, float* dw, float* db,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* DAttention pooling + predictor (modules/datten.py:5-38,85-101, modules/rrt.py:221-241):
 *   A = act(h @ w1^T + b1) @ w2^T + b2  [L];  a = softmax_L(A);  pooled[dim] = a @ h;
 *   logits[n_classes] = pooled @ pred_w^T + pred_b   (pred_w may be NULL: pooling only).
 * attn (optional, [L]) receives a, or the raw scores A when attn_raw != 0 (the reference's no_norm).
 * b1 / b2 / pred_b / w1_f16 may be NULL.  act may carry RRT_ACT_GATED (above).
 * drop_p / seed: the nn.Dropout(0.25) inside the score MLP (da_dropout=True, modules/datten.py:20-21,58-60)
 * in training mode: counter-based mask, stream RRT_DROP_STREAM_POOL over the hidden buffer [L, rows of w1]
 * (gated: independent masks for the two branches, as in the reference); 0 = eval.
 * pre (nullable, [L, hid]): with act = RRT_ACT_GELU it receives the pre-activations of the act branch. */
#define RRT_DROP_STREAM_POOL 66
RRT_API int rrt_attn_pool_forward(const float* h, int64_t L, int32_t dim, int32_t hid,
                                  const float* w1, const float* b1, const void* w1_f16, int32_t act,
                                  const float* w2, const float* b2, const float* pred_w,
                                  const float* pred_b, int32_t n_classes, float* pooled,
                                  float* logits, float* attn, int32_t attn_raw, float drop_p, uint64_t seed,
                                  float* pre, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training: forward with a tape + backward (autograd of modules/rrt.py:165-202) ---------------
 * Dropout is not applied (drop_out = 0 or eval mode); the Python module raises for active dropout.
 * rrt_encoder_forward_train computes exactly what rrt_encoder_forward computes and keeps, in the
 * caller-owned `tape` (rrt_train_tape_bytes), what the backward pass re-reads: per R-MSA layer the
 * LayerNorm output, q/k/v, the attention output (fp16) and the layer output (fp32); for CR-MSA the
 * logits, min/max, landmarks and the landmark MHA's q/k/v, o and output; and, when the caller passes
 * no fp16 weight shadows, the fp16 copies of the GEMM weights the forward made (the backward's
 * input-gradient GEMMs read them as they are).  The attention probabilities are recomputed.  rrt_encoder_backward takes d(loss)/d(out) and writes d(loss)/dx and
 * every parameter gradient.  Gradient buffers follow rrt_weights (same shapes, fp32) and MUST be
 * zero-initialised by the caller (bias / LayerNorm / tap / phi gradients are accumulated with
 * atomics; weight gradients are overwritten).  Tensor-core operands of the backward GEMMs are fp16
 * with automatic per-stage power-of-two scaling (csrc/backward.cuh), accumulators fp32.
 * crmsa_mlp (logits = phi.2 tanh(phi.0 z)): gradients of both weights; needs dim 512 or 1024.
 * PEG / PPEG (pos != none): gradients of the conv kernels and biases; PPEG needs bags of >= 37 tokens.
 * Not covered (RRT_E_INVALID): crmsa_k > 8, R-MSA head_dim other than 32 / 64, CR-MSA head_dim 128, regions > 256
 * tokens, FFN, the EPEG ablation variants. */
typedef struct rrt_attn_grads {
  float* qkv_w;  /* [3D, D] */
  float* qkv_b;  /* [3D] or NULL */
  float* proj_w; /* [D, D] */
  float* proj_b; /* [D] */
  float* pe_w;   /* [heads, epeg_k] or NULL (pe.bias has an exactly zero gradient) */
} rrt_attn_grads;

typedef struct rrt_grads {
  float* norm_w;
  float* norm_b;
  float* layer_norm_w[RRT_MAX_RMSA_LAYERS];
  float* layer_norm_b[RRT_MAX_RMSA_LAYERS];
  rrt_attn_grads layer_attn[RRT_MAX_RMSA_LAYERS];
  float* cr_norm_w;
  float* cr_norm_b;
  float* cr_phi; /* [D, k] */
  rrt_attn_grads cr_attn;
  float* cr_phi_w1; /* crmsa_mlp: phi.0.weight [D/4, D] */
  float* cr_phi_w2; /* crmsa_mlp: phi.2.weight [k, D/4] */
  float* pos_w[3];  /* PEG: proj.weight; PPEG: proj / proj1 / proj2 .weight ([D,1,k,k] or [D,1,k,1]); overwritten */
  float* pos_b[3];  /* the matching biases [D] or NULL */
} rrt_grads;

RRT_API int rrt_train_tape_bytes(const rrt_config* cfg, int64_t L, size_t* bytes);
/* drop_p / seed: training-mode proj_drop of every InnerAttention (modules/rmsa.py:70,132; the
 * reference's drop_out, default 0.1; 0 = eval-mode arithmetic).  The mask is counter-based (splitmix64 of
 * seed, mask stream and element index; csrc/common.cuh) and is regenerated by rrt_encoder_backward, which
 * must be given the same drop_p and seed.  Mask streams: R-MSA layer i -> i (mask over [L, D], token
 * order), landmark MHA of CR-MSA -> RRT_DROP_STREAM_CRMSA (mask over [k*64, D]).  It is the reference's
 * dropout in distribution, not torch's Philox stream; rrt_dropout_mask exposes it for parity tests. */
#define RRT_DROP_STREAM_CRMSA 64
/* branch_scale (HOST array of n_rmsa_layers + 1 floats: the R-MSA layers, then CR-MSA; NULL = all ones):
 * stochastic depth of this step (drop_path, modules/rrt.py:102,125: x + drop_path(attn(norm(x))) with a batch
 * of one bag): 0 = the block's branch is dropped (x passes through), otherwise the factor 1 / keep_prob on the
 * branch.  The caller draws the Bernoulli variables and passes the same array to rrt_encoder_backward. */
RRT_API int rrt_encoder_forward_train(const rrt_config* cfg, const rrt_weights* w, const float* x,
                                      float* out, int64_t L, void* tape, size_t tape_bytes,
                                      float drop_p, uint64_t seed, const float* branch_scale, void* stream);
/* out[i] = keep(i) / (1 - drop_p) for i < n (n % 4 == 0): the factor the forward multiplies element i by. */
RRT_API int rrt_dropout_mask(float* out, int64_t n, float drop_p, uint64_t seed, uint32_t mask_stream,
                             void* stream);
/* RRT_OK when rrt_encoder_backward covers this configuration AND this bag length, else RRT_E_INVALID with the
 * reason in rrt_last_error().  The limits depend on the bag, not only on the configuration: the R-MSA backward
 * keeps a region resident (regions of at most 256 tokens: N <= 16384 at region_num = 8), R-MSA head_dim must be
 * 32 or 64, CR-MSA head_dim anything but 128 (crmsa_heads = 1 runs an fp32 landmark backward), crmsa_k <= 8,
 * crmsa_mlp needs dim 512 / 1024, no PEG / PPEG / FFN.
 * Callers check it BEFORE the taped forward, so that a training loop fails at the first call and not inside
 * loss.backward(). */
RRT_API int rrt_backward_supported(const rrt_config* cfg, int64_t L);
RRT_API int rrt_backward_workspace_bytes(const rrt_config* cfg, int64_t L, size_t* bytes);
/* x: the forward input; dout: d(loss)/d(out) [L, D]; dx: d(loss)/dx [L, D] (may not alias dout). */
RRT_API int rrt_encoder_backward(const rrt_config* cfg, const rrt_weights* w, const float* x,
                                 const float* dout, int64_t L, const void* tape, size_t tape_bytes,
                                 const rrt_grads* grads, float* dx, void* workspace,
                                 size_t workspace_bytes, float drop_p, uint64_t seed, const float* branch_scale,
                                 void* stream);

/* Building blocks of the backward pass, exposed for the parity tests.
 * rrt_attention_backward: qkv [R*P, 3D] fp16 and o [R*P, D] fp16 as the forward wrote them,
 * d_o [R*P, D] fp16 -> d_qkv [R*P, 3D] fp16, d_taps [heads, epeg_k] fp32 (+=; taps/d_taps NULL: no EPEG).
 * rrt_layernorm_backward: y = LayerNorm(x): dy [L, D] fp32 -> dx, dgamma (+=), dbeta (+=). */
/* The weight-gradient GEMM, exposed for parity tests: dw[c_out, c_in] = dy[rows, c_out]^T @ act[rows, c_in]
 * (fp16 row-major inputs read MN-major by tcgen05 -- no transposed copies -- fp32 split-K accumulation). */
RRT_API int rrt_linear_wgrad_f16(const void* dy_f16, const void* act_f16, float* dw, int64_t rows,
                                 int32_t c_out, int32_t c_in, void* stream);
RRT_API int rrt_attention_backward(const void* qkv, const void* o, const void* d_o, const float* taps,
                                   void* d_qkv, float* d_taps, int32_t R, int32_t P, int32_t dim,
                                   int32_t heads, int32_t epeg_k, void* stream);
RRT_API int rrt_layernorm_backward(const float* x, const float* gamma, const float* dy, float* dx,
                                   float* dgamma, float* dbeta, int64_t L, int32_t dim, void* stream);

/* Backward of rrt_attn_pool_forward (autograd of modules/datten.py:28-38 + the predictor): dlogits
 * [n_classes] -> dh [L, dim] (overwritten) and the gradients of w1 [hid, dim], b1 [hid] (nullable), w2 [hid],
 * b2 [1] (nullable), pred_w [n_classes, dim], pred_b [n_classes] (nullable), all overwritten.  pooled = the
 * forward's output; tape = the forward's workspace, untouched since.  act = relu | tanh | gelu (needs `pre`) |
 * none, optionally | RRT_ACT_GATED (hid = 128; w1 / dw1 [256, dim], b1 / db1 [256]).  drop_p / seed as in the
 * forward.  Workspace: rrt_mil_head_backward_workspace_bytes with hid = rows of w1. */
RRT_API int rrt_mil_head_backward_workspace_bytes(int64_t L, int32_t dim, int32_t hid, size_t* bytes);
RRT_API int rrt_attn_pool_backward(const float* h, int64_t L, int32_t dim, int32_t hid, const float* w1,
                                   int32_t act, const float* w2, const float* pred_w, int32_t n_classes,
                                   const float* pooled, const float* dlogits, float drop_p, uint64_t seed,
                                   const float* pre, const void* tape, size_t tape_bytes, float* dh, float* dw1,
                                   float* db1, float* dw2, float* db2, float* dpred_w, float* dpred_b,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- optimizer step of the training harness (main.py:224-233: torch.optim.Adam, lr 2e-4, wd 1e-5) ----
 * One launch updates every tensor of the list with torch.optim.Adam (decoupled = 0: L2 weight decay
 * added to the gradient) or AdamW (decoupled = 1) semantics, step = 1, 2, ...:
 *   g = grad * grad_scale;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
 * All pointers are device fp32 arrays of n elements; the struct array itself is host memory. */
/* Per-step state in DEVICE memory, for training steps captured in a CUDA graph (values passed by value would be
 * frozen into the graph):  struct { uint64_t seed; float bc1; float bc2_rsqrt; }  (16 bytes, 16-byte aligned).
 * While set (process-wide -- torch runs the backward on its autograd thread; NULL switches it off), every dropout mask of the training entry points adds
 * `seed` to the seed argument of the call when the mask is evaluated, and rrt_adam_step takes its bias
 * corrections bc1 = 1 - beta1^t, bc2_rsqrt = 1 / sqrt(1 - beta2^t) from the buffer instead of from `step`.
 * The caller updates the buffer (one 16-byte copy on the stream) before every replay. */
RRT_API int rrt_set_step_state(const void* device_state);

typedef struct rrt_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
} rrt_adam_tensor;
RRT_API int rrt_adam_step(const rrt_adam_tensor* tensors, int32_t n_tensors, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int32_t decoupled, int64_t step,
                          float grad_scale, void* stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------- */
/* Kernel launches issued by this library in this process so far. */
RRT_API int64_t rrt_launch_count(void);

/* Stage timing: when enabled, every kernel of rrt_encoder_forward (and of the block entry points)
 * is bracketed by CUDA events on the launching stream.  rrt_stage_timing_read sums the finished
 * intervals per stage since the last enable; the caller synchronises the stream first.  Stages
 * are numbered 0 .. rrt_stage_count()-1 in pipeline order; rrt_stage_name gives the kernel role. */
RRT_API int rrt_stage_timing_enable(int32_t on);
RRT_API int32_t rrt_stage_count(void);
RRT_API const char* rrt_stage_name(int32_t stage);
RRT_API int rrt_stage_timing_read(int32_t stage, double* total_ms, int64_t* launches);

/* PEG / PPEG alone (parity tests): out[L,D] = pos_embedding(x[L,D]); ppeg != 0: three convs (k, 5, 3).
 * w / b: arrays of 3 device pointers as in rrt_weights.pos_w / pos_b.  x != out. */
RRT_API int rrt_peg_forward(const float* x, float* out, int64_t L, int32_t dim, int32_t peg_k,
                            int32_t ppeg, int32_t peg_1d, const float* const* w, const float* const* b,
                            void* stream);

/* out[L,D] = LayerNorm(x[L,D]) (eps 1e-5). */
RRT_API int rrt_layernorm_forward(const float* x, const float* gamma, const float* beta,
                                  float* out, int64_t L, int32_t D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RRT_B200_H_ */
