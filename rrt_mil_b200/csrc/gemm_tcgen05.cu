// Bag-sized linear layers (QKV projection, output projection) on the 5th-gen tensor cores:
//   C[M,N] = A[M,K] @ W[N,K]^T (+bias, +epilogue),  tf32 operands, fp32 accumulate in TMEM.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   : 128x32 (A) and 256x32 (W) fp32 tiles, SWIZZLE_128B, 4-stage ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma.kind::tf32 (M128 N256 K8),
//                                 tcgen05.commit releases smem stages / publishes accumulators
//   warp 2      TMEM allocator : 512 columns = two 128x256 fp32 accumulators (double buffered, so
//                                 the epilogue of tile i overlaps the main loop of tile i+1)
//   warps 4..7  epilogue       : tcgen05.ld 32x32b -> registers -> per-warp smem transpose ->
//                                 coalesced 128-B row segments to global (+bias, +residual scatter)
// Operands must already be tf32-representable (cvt.rna done by the producing kernel /
// round_tf32_kernel); the tensor core ignores the 13 low mantissa bits.
#include "kernels.cuh"
#include "sm100.cuh"

namespace rrt {
namespace {
using namespace sm100;

constexpr int BM = 128, BK = 32;  // BK fp32 = 128 bytes = one swizzle atom
constexpr int A_BYTES = BM * BK * 4;
constexpr int EPI_LD = 36;                              // floats per scratch row (32 + 4 pad)
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;          // 4 epilogue warps
constexpr int BAR_BYTES = 256;
constexpr int NTHREADS = 256;

// Tile configuration: 128 x BN output tile, STAGES-deep operand ring, two BN-column accumulators.
//   BN = 256: bag-sized GEMMs (M ~ 10^4): fewest operand bytes per MAC
//   BN =  64: landmark GEMMs (M = k*64 rows): 4x more CTAs, 4x shorter MMA chain per tile
template <int BN_>
struct TileCfg {
  static constexpr int STAGES = BN_ == 256 ? 4 : 8;
  static constexpr int B_BYTES = BN_ * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 2 * BN_ < 32 ? 32 : 2 * BN_;
};

struct Tc05Params {
  int M, N, K;
  const float* bias;
  float* C;
  const float* resid;
  Grid grid;
};

template <int MODE, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tf32_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, Tc05Params p) {
  using Cfg = TileCfg<BN>;
  constexpr int STAGES = Cfg::STAGES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int TMEM_COLS = Cfg::TMEM_COLS;
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

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = ((p.M + BM - 1) / BM) * tiles_n;
  const int KB = p.K / BK;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], kb * BK, m0);
          tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], kb * BK, n0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc(kFmtTF32, BM, BN);
      int s = 0, ph = 0, acc = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t ad = umma_desc_k_sw128(smem_u32(sA + s * A_BYTES));
          const uint64_t bd = umma_desc_k_sw128(smem_u32(sB + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)  // 8 tf32 = 32 bytes per MMA along K: +2 in 16-B units
            umma_tf32(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);  // smem stage reusable once these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {  // ===== epilogue =====
    const int ew = warp - 4;  // TMEM lane quadrant this warp may read
    float* scratch = sEpi + ew * 32 * EPI_LD;
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    int acc = 0, aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      // output row of each of the 8 row groups this lane stores (mode 1: region slot -> token)
      long long orow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int gr = m0 + ew * 32 + 4 * i + sub_r;
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
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (n0 + c * 32 >= p.N) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(scratch + lane * EPI_LD + 4 * q) =
              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                          __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        __syncwarp();
        const int gc = n0 + c * 32 + sub_c;
        const bool col_ok = gc < p.N;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(p.bias + gc));
        float4 v[8];
        if (MODE == kEpiResidualUnpart) {
          // all eight residual loads in flight before any dependent add / store
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (orow[i] >= 0 && col_ok)
              v[i] = __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)orow[i] * p.N + gc));
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 a = *reinterpret_cast<const float4*>(scratch + (4 * i + sub_r) * EPI_LD + sub_c);
          if (MODE == kEpiResidualUnpart) {
            v[i].x += a.x + bv.x; v[i].y += a.y + bv.y; v[i].z += a.z + bv.z; v[i].w += a.w + bv.w;
          } else {
            v[i] = make_float4(a.x + bv.x, a.y + bv.y, a.z + bv.z, a.w + bv.w);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (orow[i] >= 0 && col_ok)
            *reinterpret_cast<float4*>(p.C + (size_t)orow[i] * p.N + gc) = v[i];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
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

// row-major fp32 [rows, K] -> tiles of box_rows x 32 floats, 128-byte swizzle
bool make_map(CUtensorMap* m, const float* base, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
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

cudaError_t launch_round_tf32(const float* src, float* dst, size_t n, cudaStream_t stream) {
  if (n % 4) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  round_tf32_kernel<<<blocks, 256, 0, stream>>>(src, dst, n4);
  return cudaGetLastError();
}

bool gemm_tcgen05_supported(int M, int N, int K) {
  return M >= 1 && N >= 4 && (N % 4) == 0 && K >= BK && (K % BK) == 0;
}

namespace {
template <int MODE, int BN>
cudaError_t launch_cfg(const float* a, const float* w, const Tc05Params& p, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  CUtensorMap tmA, tmB;
  if (!make_map(&tmA, a, p.M, p.K, BM) || !make_map(&tmB, w, p.N, p.K, BN)) return cudaErrorUnknown;
  static bool configured = false;  // per instantiation; one process drives one GPU
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_tcgen05_kernel<MODE, BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tf32_tcgen05_kernel<MODE, BN><<<grid, NTHREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_gemm_tcgen05(const float* a, const float* w, float* c, int M, int N, int K,
                                const GemmEpilogue& epi, cudaStream_t stream) {
  if (!gemm_tcgen05_supported(M, N, K)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
       reinterpret_cast<uintptr_t>(c)) & 15)
    return cudaErrorInvalidValue;
  if (epi.mode == kEpiTanh) return cudaErrorInvalidValue;
  Tc05Params p;
  p.M = M; p.N = N; p.K = K;
  p.bias = epi.bias; p.C = c; p.resid = epi.resid; p.grid = epi.grid;
  // small problems: narrower tiles so that more SMs share the (latency-bound) work
  const bool narrow = ((M + BM - 1) / BM) * ((N + 255) / 256) < sm_count() / 2;
  if (epi.mode == kEpiResidualUnpart)
    return narrow ? launch_cfg<kEpiResidualUnpart, 64>(a, w, p, stream)
                  : launch_cfg<kEpiResidualUnpart, 256>(a, w, p, stream);
  return narrow ? launch_cfg<kEpiStore, 64>(a, w, p, stream)
                : launch_cfg<kEpiStore, 256>(a, w, p, stream);
}

}  // namespace rrt
