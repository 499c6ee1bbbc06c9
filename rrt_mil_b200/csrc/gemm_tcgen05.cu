// Linear layers of the path (QKV projection, output projection, landmark projections, crmsa_mlp
// phi.0) on the 5th-gen tensor cores:
//   C[M,N] = A[M,K] @ W[N,K]^T (+bias, +epilogue),  fp16 operands, fp32 accumulate in TMEM.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warps 0, 3  TMA producers  : 128x64 (A) resp. BNx64 (W) fp16 tiles, SWIZZLE_128B, smem ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma.kind::f16 (M128 N=BN K16),
//                                 tcgen05.commit releases smem stages / publishes accumulators
//   warp 2      TMEM allocator : two 128xBN fp32 accumulators (double buffered, so the epilogue of
//                                 tile i overlaps the main loop of tile i+1)
//   warps 4..11 epilogue       : two warps per TMEM lane quadrant, software-pipelined tcgen05.ld
//                                 32x32b -> registers (accumulator released as soon as it is in
//                                 registers) -> per-warp smem transpose -> coalesced row segments to
//                                 global (+bias, tanh, residual scatter), fp16 or fp32 output
// fp16 carries the same 10 mantissa bits as tf32 (parity bar: 1e-3 rel) at twice the tensor rate
// and half the operand traffic; with fp32 (tf32) tiles this kernel was bound by L2->SMEM operand
// bytes (profiles/r01_ncu_full_v2_tf32_summary.csv).
#include "kernels.cuh"
#include "sm100.cuh"

namespace rrt {
namespace {
using namespace sm100;

constexpr int BM = 128, BK = 64;  // BK fp16 = 128 bytes = one swizzle atom
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_LD = 32;                              // floats per scratch row; 16-B slots XOR-swizzled
constexpr int kEpiWarps = 8;                            // two per TMEM lane quadrant
constexpr int EPI_BYTES = kEpiWarps * 32 * EPI_LD * 4;  // one 32x32 transpose scratch per warp
constexpr int BAR_BYTES = 256;
constexpr int NTHREADS = 32 * (4 + kEpiWarps);

// Tile configuration: 128 x BN output tile, STAGES-deep operand ring, two BN-column accumulators.
//   BN = 256: bag-sized GEMMs (M ~ 10^4): fewest operand bytes per MAC
//   BN =  64: landmark GEMMs (M = k*64 rows): 4x more CTAs, 4x shorter MMA chain per tile
template <int BN_>
struct TileCfg {
  static constexpr int STAGES = BN_ == 256 ? 4 : 8;
  static constexpr int B_BYTES = BN_ * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "dynamic shared memory budget exceeded");
  static constexpr int TMEM_COLS = 2 * BN_ < 32 ? 32 : 2 * BN_;
};

struct Tc05Params {
  int M, N, K;
  const float* bias;
  void* C;  // OutT[M or L, N]
  const float* resid;
  Grid grid;
  long long* trace;  // debug: per-CTA clock64 stamps (tools/gemm_trace.py), null in production
};

// trace slots: 0 start, 1 setup done, 2 first TMA issued, 3 last TMA issued, 4 first operands landed,
// 5 last MMA committed, 6 first accumulator ready (epilogue), 7 epilogue of first tile done,
// 8 epilogue of last tile done, 9 kernel end
__device__ __forceinline__ void stamp(const Tc05Params& p, int slot) {
  if (p.trace && blockIdx.x < 8) p.trace[blockIdx.x * 16 + slot] = clock64();
}

__device__ __forceinline__ void store_out4(float* base, size_t off, float4 v) {
  *reinterpret_cast<float4*>(base + off) = v;
}
__device__ __forceinline__ void store_out4(__half* base, size_t off, float4 v) {
  *reinterpret_cast<uint2*>(base + off) = pack_h4(v);
}

template <int MODE, int BN, typename OutT>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB, Tc05Params p) {
  using Cfg = TileCfg<BN>;
  constexpr int STAGES = Cfg::STAGES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int TMEM_COLS = Cfg::TMEM_COLS;
  constexpr int NCHUNK = BN / 32;                 // 32-column chunks of the accumulator
  constexpr int NC = NCHUNK / 2;                  // chunks per epilogue warp (two warps per quadrant)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  float* sEpi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(p, 0);

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmA);
  if (warp == 3 && lane == 0) prefetch_tensormap(&tmB);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);   // the A producer and the B producer each arrive once (+ their bytes)
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(p, 1);

  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = ((p.M + BM - 1) / BM) * tiles_n;
  const int KB = p.K / BK;

  if (warp == 0 || warp == 3) {
    if (lane == 0) {  // ===== TMA producers: warp 0 streams A tiles, warp 3 streams W tiles =====
      const bool is_a = warp == 0;
      int s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          if (is_a) {
            mbar_arrive_expect_tx(&full[s], A_BYTES);
            tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], kb * BK, m0);
            if (tile == (int)blockIdx.x && kb == 0) stamp(p, 2);
          } else {
            mbar_arrive_expect_tx(&full[s], B_BYTES);
            tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], kb * BK, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
      if (is_a) stamp(p, 3);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc(kFmtF16, BM, BN);
      int s = 0, ph = 0, acc = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (tile == (int)blockIdx.x && kb == 0) stamp(p, 4);
          const uint64_t ad = umma_desc_k_sw128(smem_u32(sA + s * A_BYTES));
          const uint64_t bd = umma_desc_k_sw128(smem_u32(sB + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // 16 fp16 = 32 bytes per MMA along K: +2 in 16-B units
            umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);  // smem stage reusable once these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
      stamp(p, 5);
    }
    __syncwarp();
  } else if (warp >= 4) {  // ===== epilogue: 8 warps, two per TMEM lane quadrant =====
    const int ew = warp - 4;
    const int quad = ew & 3;          // TMEM lanes [32*quad, 32*quad+32) are readable by this warp
    const int c_begin = (ew >> 2) * NC;  // first accumulator chunk of this warp
    float* scratch = sEpi + ew * 32 * EPI_LD;
    OutT* const out = reinterpret_cast<OutT*>(p.C);
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    int acc = 0, aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      // output row of each of the 8 row groups this lane stores (mode 2: region slot -> token)
      long long orow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int gr = m0 + quad * 32 + 4 * i + sub_r;
        long long o = -1;
        if (gr < p.M) {
          if (MODE == kEpiResidualUnpart) {
            int tok = p.grid.slot_to_token(gr);
            if (tok < p.grid.L) o = tok;
          } else {
            o = gr;
          }
        }
        orow[i] = o;
      }
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      if (tile == (int)blockIdx.x && threadIdx.x == 128) stamp(p, 6);
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c_begin * 32;
      uint32_t r[2][32];
      tmem_ld_32x32(t_addr, r[0]);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        tmem_ld_wait();  // chunk j is in registers
        if (j + 1 < NC) {
          tmem_ld_32x32(t_addr + (j + 1) * 32, r[(j + 1) & 1]);  // overlaps the stores of chunk j
        } else {
          // the whole accumulator slice of this warp has left TMEM: release it to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        const int gc = n0 + (c_begin + j) * 32 + sub_c;
        const bool col_ok = gc < p.N;
        float4 v[8];
        if (MODE == kEpiResidualUnpart) {
          // residual rows: all eight loads in flight while the chunk is transposed through smem
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (orow[i] >= 0 && col_ok)
              v[i] = __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)orow[i] * p.N + gc));
          }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(p.bias + gc));
        const uint32_t* rr = r[j & 1];
#pragma unroll
        for (int q = 0; q < 8; ++q)  // row = lane; 16-byte slot q lands at slot q ^ (row & 7)
          *reinterpret_cast<float4*>(scratch + lane * EPI_LD + 4 * (q ^ (lane & 7))) =
              make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]),
                          __uint_as_float(rr[4 * q + 2]), __uint_as_float(rr[4 * q + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + sub_r;
          float4 a = *reinterpret_cast<const float4*>(scratch + rl * EPI_LD +
                                                      4 * ((lane & 7) ^ (rl & 7)));
          if (MODE == kEpiResidualUnpart) {
            v[i].x += a.x + bv.x; v[i].y += a.y + bv.y; v[i].z += a.z + bv.z; v[i].w += a.w + bv.w;
          } else {
            v[i] = make_float4(a.x + bv.x, a.y + bv.y, a.z + bv.z, a.w + bv.w);
            if (MODE == kEpiTanh)
              v[i] = make_float4(tanhf(v[i].x), tanhf(v[i].y), tanhf(v[i].z), tanhf(v[i].w));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (orow[i] >= 0 && col_ok) store_out4(out, (size_t)orow[i] * p.N + gc, v[i]);
        __syncwarp();
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
      if (threadIdx.x == 128) stamp(p, tile == (int)blockIdx.x ? 7 : 8);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(p, 9);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void convert_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride)
    reinterpret_cast<uint2*>(dst)[i] = pack_h4(__ldg(reinterpret_cast<const float4*>(src) + i));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major fp16 [rows, K] -> tiles of box_rows x 64 halves (128 bytes), 128-byte swizzle
bool make_map(CUtensorMap* m, const __half* base, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
}  // namespace

cudaError_t launch_convert_f16(const float* src, __half* dst, size_t n, cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  convert_f16_kernel<<<blocks, 256, 0, stream>>>(src, dst, n4);
  return cudaGetLastError();
}

long long* g_gemm_trace = nullptr;  // debug hook (rrt_debug_set_gemm_trace)

bool gemm_tcgen05_supported(int M, int N, int K) {
  return M >= 1 && N >= 4 && (N % 4) == 0 && K >= BK && (K % BK) == 0;
}

namespace {
template <int MODE, int BN, typename OutT>
cudaError_t launch_cfg(const __half* a, const __half* w, const Tc05Params& p, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  CUtensorMap tmA, tmB;
  if (!make_map(&tmA, a, p.M, p.K, BM) || !make_map(&tmB, w, p.N, p.K, BN)) return cudaErrorUnknown;
  static bool configured = false;  // per instantiation; one process drives one GPU
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_f16_tcgen05_kernel<MODE, BN, OutT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_f16_tcgen05_kernel<MODE, BN, OutT><<<grid, NTHREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}

template <int MODE, typename OutT>
cudaError_t launch_mode(const __half* a, const __half* w, const Tc05Params& p, cudaStream_t stream) {
  // small problems: narrower tiles so that more SMs share the (latency-bound) work
  const bool narrow = ((p.M + BM - 1) / BM) * ((p.N + 255) / 256) < sm_count() / 2;
  return narrow ? launch_cfg<MODE, 64, OutT>(a, w, p, stream) : launch_cfg<MODE, 256, OutT>(a, w, p, stream);
}
}  // namespace

cudaError_t launch_gemm_tcgen05(const __half* a, const __half* w, void* c, bool out_f16, int M, int N,
                                int K, const GemmEpilogue& epi, cudaStream_t stream) {
  if (!gemm_tcgen05_supported(M, N, K)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
       reinterpret_cast<uintptr_t>(c)) & 15)
    return cudaErrorInvalidValue;
  Tc05Params p;
  p.M = M; p.N = N; p.K = K;
  p.bias = epi.bias; p.C = c; p.resid = epi.resid; p.grid = epi.grid;
  p.trace = g_gemm_trace;
  if (epi.mode == kEpiResidualUnpart)
    return out_f16 ? cudaErrorInvalidValue : launch_mode<kEpiResidualUnpart, float>(a, w, p, stream);
  if (epi.mode == kEpiTanh)
    return out_f16 ? cudaErrorInvalidValue : launch_mode<kEpiTanh, float>(a, w, p, stream);
  return out_f16 ? launch_mode<kEpiStore, __half>(a, w, p, stream)
                 : launch_mode<kEpiStore, float>(a, w, p, stream);
}

}  // namespace rrt
