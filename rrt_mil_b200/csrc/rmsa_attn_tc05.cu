// R-MSA attention core on the 5th-gen tensor cores (head_dim 64, regions of up to 256 tokens):
// one CTA per (region, head), two CTAs per SM when the accumulators fit 256 TMEM columns.
//
//   load     K, V tiles [P16 x 64] f16 by TMA (SWIZZLE_128B); Q rows with the EPEG halo by cp.async
//   EPEG     Q' = scale*log2e * (Q + dwconv1d_P(Q; taps_h)) as a banded-Toeplitz product on
//            mma.sync (as in rmsa_attn_f16.cu); the result is written to shared memory directly in
//            the UMMA K-major 128B-swizzled layout
//   S        tcgen05.mma.kind::f16  M=128, N=P16, K=64:  S[128 x P16] -> TMEM   (A = Q', B = K)
//   softmax  one thread per query row (TMEM lane): two passes of tcgen05.ld over the row, exp2,
//            P (f16) written to shared memory in the K-major swizzled layout, row sum kept in a register
//   O        tcgen05.mma  M=128, N=64, K=P16:  O[128 x 64] -> TMEM   (A = P, B = V consumed MN-major)
//   out      tcgen05.ld O, * 1/rowsum, f16 row -> global
// Regions with more than 128 tokens take a second M=128 block (rows 128..P-1) that reuses the S
// columns, the P buffer and the O columns once the first block has drained them.
// (modules/rmsa.py:103-122; SURVEY.md 0.2-1 for the EPEG-on-Q identity.)
#include "kernels.cuh"
#include "sm100.cuh"

namespace rrt {
namespace {
using namespace sm100;

constexpr int HD = 64;
constexpr int LDH = HD + 8;  // halo'd raw Q rows (cp.async, ldmatrix-friendly skew)

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
// 32 TMEM lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// instruction descriptor: f16 operands, fp32 accumulate, A K-major, B K-major (b_mn = 0) or MN-major
__device__ __forceinline__ uint32_t idesc_f16(int M, int N, int b_mn) {
  return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major operand, 128-byte swizzle: rows of 64 contiguous MN elements (128 B) per K index, 8 K
// indices per 1024-byte atom; SBO = distance between 8-row K groups, LBO = distance between
// 64-element MN chunks (unused for N = 64)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tstamp(long long* tr, int slot) {
  if (tr && blockIdx.x < 8 && blockIdx.y == 0 && threadIdx.x == 0) tr[blockIdx.x * 8 + slot] = clock64();
}
// finer stamps of CTA 0: thread 0 (block-0 softmax warp) -> row 8, thread 128 (block-1 warp) -> row 9
__device__ __forceinline__ void fstamp(long long* tr, int slot) {
  if (tr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x == 0 || threadIdx.x == 128))
    tr[(8 + (threadIdx.x >> 7)) * 8 + slot] = clock64();
}

struct AttnTcParams {
  long long* trace;
  const __half* qkv;
  const float* taps;
  __half* o;
  Grid grid;
  int D, epeg_k, q_rows, P16, tmem_cols, o_col;
  float qscale;
};

__global__ void __launch_bounds__(512) rmsa_attn_tc05_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                             AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int P = p.grid.P, P16 = p.P16;
  const int nkc = (P16 + 63) / 64;                    // 64-key chunks of the P operand
  // layout (1024-byte aligned pieces): Q' | K | V | P (the raw halo'd Q aliases the head of P)
  uint8_t* sQp = smem;                                 // [P16][128 B] swizzled
  uint8_t* sK = sQp + (size_t)P16 * 128;
  uint8_t* sV = sK + (size_t)P16 * 128;
  uint8_t* sP = sV + (size_t)P16 * 128;                // [nkc][128][128 B] swizzled
  const size_t p_bytes = (size_t)nkc * 128 * 128;
  const size_t qraw_bytes = (size_t)p.q_rows * LDH * 2;
  __half* Qs = reinterpret_cast<__half*>(sP);          // raw Q, dead before P is written
  uint8_t* tail = sP + (p_bytes > qraw_bytes ? p_bytes : ((qraw_bytes + 1023) & ~(size_t)1023));
  float* Ts = reinterpret_cast<float*>(tail);          // [epeg_k] (<= 63)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 256);  // [0] TMA, [1] S0, [2] O0, [3] S1, [4] O1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int rho = blockIdx.x, h = blockIdx.y;
  const int pad = p.taps ? p.epeg_k / 2 : 0;
  const size_t ld = 3 * (size_t)p.D;
  const __half* base = p.qkv + (size_t)rho * P * ld + h * HD;

  if (tid == 0) {
    prefetch_tensormap(&tmQKV);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  tstamp(p.trace, 0);

  // ---- loads: K and V tiles by TMA, halo'd raw Q rows by cp.async -----------------------------
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 2u * (uint32_t)P16 * 128u);
    tma_load_2d(sK, &tmQKV, &bars[0], p.D + h * HD, rho * P);
    tma_load_2d(sV, &tmQKV, &bars[0], 2 * p.D + h * HD, rho * P);
  }
  for (int i = tid; i < p.q_rows * (HD / 8); i += blockDim.x) {
    int r = i / (HD / 8), c = (i - r * (HD / 8)) * 8;
    int pp = r - pad;
    bool ok = pp >= 0 && pp < P;
    cp_async16(Qs + (size_t)r * LDH + c, base + (size_t)(ok ? pp : 0) * ld + c, ok);
  }
  cp_async_commit();
  if (p.taps)
    for (int i = tid; i < p.epeg_k; i += blockDim.x) Ts[i] = __ldg(p.taps + h * p.epeg_k + i);
  cp_async_wait<0>();
  __syncthreads();
  tstamp(p.trace, 1);

  // ---- EPEG (Toeplitz on mma.sync): warp w -> rows 16w..16w+15 of Q', written swizzled ----------
  if (16 * warp < P16) {
    const int i0 = 16 * warp;
    float qacc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) qacc[i][e] = 0.f;
    const int nkcq = p.taps ? (16 + p.epeg_k - 1 + 15) / 16 : 1;
    for (int kc = 0; kc < nkcq; ++kc) {
      uint32_t ca[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ro = g + (e & 1) * 8;
        const int co = 16 * kc + 2 * t + (e >> 1) * 8;
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int d = co + u - ro;
          float x = (p.taps && d >= 0 && d < p.epeg_k) ? Ts[d] : 0.f;
          v[u] = x + (d == pad ? 1.f : 0.f);
        }
        ca[e] = pack_h2(v[0], v[1]);
      }
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {
        uint32_t b[4];
        ldsm_x4_trans(b, Qs + (size_t)(i0 + 16 * kc + (lane & 7) + ((lane >> 3) & 1) * 8) * LDH +
                             np * 16 + (lane >> 4) * 8);
        mma_f16_16x8x16(qacc[2 * np], ca, b[0], b[1]);
        mma_f16_16x8x16(qacc[2 * np + 1], ca, b[2], b[3]);
      }
    }
    // accumulator fragment (rows g, g+8; columns 8*nt + 2t, +1) -> K-major swizzled rows of 128 B
#pragma unroll
    for (int nt = 0; nt < HD / 8; ++nt) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int row = i0 + g + hh * 8;
        uint8_t* dst = sQp + (size_t)row * 128 + ((nt ^ (row & 7)) * 16) + t * 4;
        *reinterpret_cast<uint32_t*>(dst) =
            pack_h2(qacc[nt][hh * 2] * p.qscale, qacc[nt][hh * 2 + 1] * p.qscale);
      }
    }
  }
  fence_proxy_async_smem();  // Q' (generic-proxy writes) must be visible to the tensor core
  __syncthreads();           // also: every warp is done with the raw Q that aliases the P buffer

  const int nblk = P > 128 ? 2 : 1;
  const uint32_t idesc_s = idesc_f16(128, P16, 0);
  const uint32_t idesc_o = idesc_f16(128, HD, 1);
  const int ksteps = P16 / 16;
  const uint32_t tS = tmem_base, tO = tmem_base + (uint32_t)p.o_col;

  auto issue_s = [&](int blk, uint64_t* bar) {  // S = Q'[blk] K^T
    const uint64_t ad = umma_desc_k_sw128(smem_u32(sQp + (size_t)blk * 128 * 128));
    const uint64_t bd = umma_desc_k_sw128(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) umma_f16(tS, ad + 2 * k, bd + 2 * k, idesc_s, k != 0);
    umma_commit(bar);
  };
  auto issue_o = [&](uint64_t* bar) {  // O = P V
    for (int s = 0; s < ksteps; ++s) {
      const uint64_t ad = umma_desc_k_sw128(smem_u32(sP + (size_t)(s >> 2) * 16384)) + 2 * (s & 3);
      const uint64_t bd = umma_desc_mn_sw128(smem_u32(sV + (size_t)s * 2048));
      umma_f16(tO, ad, bd, idesc_o, s != 0);
    }
    umma_commit(bar);
  };
  // one thread per query row of block `blk`; returns the row's sum of exponentials
  auto softmax_row = [&](int blk, uint64_t* bar_s, uint64_t* bar_p_free) -> float {
    const int quad = warp & 3;
    const uint32_t trow = tS + ((uint32_t)(quad * 32) << 16);
    fstamp(p.trace, blk == 0 ? 0 : 0);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    fstamp(p.trace, 1);
    float mx = -INFINITY;
    for (int c0 = 0; c0 < P16; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x16(trow + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j < P) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
    fstamp(p.trace, 2);
    if (bar_p_free) mbar_wait(bar_p_free, 0);  // the previous block's P has been consumed
    fstamp(p.trace, 3);
    const int rloc = quad * 32 + lane;         // row within the block
    float sum = 0.f;
    for (int c0 = 0; c0 < P16; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x16(trow + c0, r);
      tmem_ld_wait();
      float e[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        e[j] = (c0 + j < P) ? fast_exp2(__uint_as_float(r[j]) - mx) : 0.f;
        sum += e[j];
      }
      uint8_t* rowp = sP + (size_t)(c0 >> 6) * 16384 + (size_t)rloc * 128;
      const int jj = (c0 & 63) >> 3;  // first of the two 16-byte chunks (8 keys each)
      *reinterpret_cast<uint4*>(rowp + (((jj) ^ (rloc & 7)) * 16)) =
          make_uint4(pack_h2(e[0], e[1]), pack_h2(e[2], e[3]), pack_h2(e[4], e[5]), pack_h2(e[6], e[7]));
      *reinterpret_cast<uint4*>(rowp + (((jj + 1) ^ (rloc & 7)) * 16)) =
          make_uint4(pack_h2(e[8], e[9]), pack_h2(e[10], e[11]), pack_h2(e[12], e[13]), pack_h2(e[14], e[15]));
    }
    fstamp(p.trace, 4);
    tc_fence_before();
    fence_proxy_async_smem();
    return sum;
  };
  auto store_o = [&](int blk, uint64_t* bar_o, float sum) {
    const int quad = warp & 3;
    const uint32_t trow = tO + ((uint32_t)(quad * 32) << 16);
    fstamp(p.trace, 5);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    fstamp(p.trace, 6);
    const int q = blk * 128 + quad * 32 + lane;
    const float inv = 1.f / sum;
    __half* orow = p.o + ((size_t)rho * P + q) * p.D + h * HD;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x16(trow + c0, r);
      tmem_ld_wait();
      if (q < P) {
        uint4 a = make_uint4(pack_h2(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv),
                             pack_h2(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv),
                             pack_h2(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv),
                             pack_h2(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv));
        uint4 b = make_uint4(pack_h2(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv),
                             pack_h2(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv),
                             pack_h2(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv),
                             pack_h2(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv));
        *reinterpret_cast<uint4*>(orow + c0) = a;
        *reinterpret_cast<uint4*>(orow + c0 + 8) = b;
      }
    }
    tc_fence_before();
  };

  tstamp(p.trace, 2);
  // ---- block 0: S0 -> softmax0 ------------------------------------------------------------------
  mbar_wait(&bars[0], 0);  // K and V have landed (every thread observes the TMA barrier)
  tstamp(p.trace, 3);
  if (tid == 0) {
    tc_fence_after();
    issue_s(0, &bars[1]);
  }
  float sum0 = 0.f, sum1 = 0.f;
  if (warp < 4) sum0 = softmax_row(0, &bars[1], nullptr);
  tstamp(p.trace, 4);
  __syncthreads();
  // ---- O0 (and S1 for the second block) ----------------------------------------------------------
  if (tid == 0) {
    tc_fence_after();
    issue_o(&bars[2]);
    if (nblk > 1) issue_s(1, &bars[3]);
  }
  if (warp < 4) store_o(0, &bars[2], sum0);
  else if (nblk > 1 && warp < 8) sum1 = softmax_row(1, &bars[3], &bars[2]);
  tstamp(p.trace, 5);
  __syncthreads();
  tstamp(p.trace, 6);
  if (nblk > 1) {
    if (tid == 0) {
      tc_fence_after();
      issue_o(&bars[4]);
    }
    if (warp >= 4 && warp < 8) store_o(1, &bars[4], sum1);
  }
  tc_fence_before();
  __syncthreads();
  tstamp(p.trace, 7);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn2() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}
}  // namespace

int g_attn_tc05 = 0;  // rrt_debug_set_attention_kernel: 1 = tcgen05 core, 0 = mma.sync core

bool rmsa_attention_tc05_supported(const Grid& grid, int D, int heads) {
  return heads > 0 && D / heads == HD && D % heads == 0 && grid.P >= 16 && grid.P <= 256 &&
         heads <= 65535;
}

cudaError_t launch_rmsa_attention_tc05(const __half* qkv, const float* taps, __half* o,
                                       const Grid& grid, int D, int heads, int epeg_k,
                                       cudaStream_t stream) {
  if (!rmsa_attention_tc05_supported(grid, D, heads)) return cudaErrorInvalidValue;
  EncodeTiledFn fn = encode_fn2();
  if (!fn) return cudaErrorUnknown;
  AttnTcParams p;
  p.trace = g_attn_trace;
  p.qkv = qkv; p.taps = taps; p.o = o; p.grid = grid; p.D = D; p.epeg_k = epeg_k;
  p.P16 = (grid.P + 15) / 16 * 16;
  const int W = p.P16 / 16;
  const int pad = taps ? epeg_k / 2 : 0;
  const int nkcq = taps ? (16 + epeg_k - 1 + 15) / 16 : 1;
  p.q_rows = 16 * (W - 1) + 16 * nkcq;
  if (p.q_rows < 16 * W + 2 * pad) p.q_rows = 16 * W + 2 * pad;
  p.o_col = p.P16 <= 192 ? 192 : 256;
  p.tmem_cols = p.P16 <= 192 ? 256 : 512;
  p.qscale = 1.4426950408889634f / sqrtf((float)HD);
  // K / V tiles: [P16 rows x 64 halves] boxes of the [Np, 3D] f16 qkv tensor
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)(3 * D), (cuuint64_t)grid.Np};
  cuuint64_t strides[1] = {(cuuint64_t)(3 * D) * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)HD, (cuuint32_t)p.P16};
  cuuint32_t estr[2] = {1, 1};
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(qkv), dims, strides, box, estr,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorUnknown;
  const int nkc = (p.P16 + 63) / 64;
  size_t p_bytes = (size_t)nkc * 16384, qraw = ((size_t)p.q_rows * LDH * 2 + 1023) & ~(size_t)1023;
  size_t smem = 1024 + 3 * (size_t)p.P16 * 128 + (p_bytes > qraw ? p_bytes : qraw) + 512;
  // the second M-block reads 128 rows of Q' starting at row 128: keep that window inside the buffer
  if (grid.P > 128 && 3 * (size_t)p.P16 * 128 < 256 * 128) return cudaErrorInvalidValue;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rmsa_attn_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(rmsa_attn_tc05_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
  }
  int warps = W < 8 ? 8 : W;  // softmax needs warps 0..7 (TMEM lane quadrants of both blocks)
  dim3 g(grid.R, heads);
  rmsa_attn_tc05_kernel<<<g, 32 * warps, smem, stream>>>(tm, p);
  return cudaGetLastError();
}

}  // namespace rrt
