// Small/odd-shape linear layer: c = a @ w^T (+bias, +epilogue) on the legacy tensor path
// (mma.sync m16n8k8 tf32, fp32 accumulate, cp.async 3-stage pipeline).  Used for the
// latency-bound landmark GEMMs of CR-MSA (M = k*64 rows) and as the bring-up GEMM; the
// bag-sized GEMMs (QKV, proj) run on tcgen05 (gemm_tcgen05.cu).
#include "kernels.cuh"

namespace rrt {

namespace {
constexpr int BN = 128, BK = 32, LDS = BK + 4, STAGES = 3, NTHREADS = 256;

template <int BM>
__global__ void __launch_bounds__(NTHREADS) gemm_tf32_mma_kernel(const float* __restrict__ A,
                                                                 const float* __restrict__ W,
                                                                 float* __restrict__ C, int M,
                                                                 int N, int K, GemmEpilogue epi) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                      // [STAGES][BM][LDS]
  float* Bs = smem + STAGES * BM * LDS;  // [STAGES][BN][LDS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  constexpr int WM = BM / 2, MT = WM / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[MT][4][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  auto load_stage = [&](int stage, int k0) {
    for (int i = tid; i < BM * 8; i += NTHREADS) {
      int r = i >> 3, c = (i & 7) * 4;
      bool ok = (m0 + r) < M;
      cp_async16(&As[(stage * BM + r) * LDS + c], A + (size_t)(ok ? m0 + r : 0) * K + k0 + c, ok);
    }
    for (int i = tid; i < BN * 8; i += NTHREADS) {
      int r = i >> 3, c = (i & 7) * 4;
      bool ok = (n0 + r) < N;
      cp_async16(&Bs[(stage * BN + r) * LDS + c], W + (size_t)(ok ? n0 + r : 0) * K + k0 + c, ok);
    }
  };

  const int KT = K / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s * BK);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    int nk = kt + STAGES - 1;
    if (nk < KT) load_stage(nk % STAGES, nk * BK);
    cp_async_commit();
    const float* as = As + (kt % STAGES) * BM * LDS;
    const float* bs = Bs + (kt % STAGES) * BN * LDS;
#pragma unroll
    for (int kk = 0; kk < BK / 8; ++kk) {
      uint32_t af[MT][4], bf[4][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float* p = as + (wm * WM + mt * 16 + g) * LDS + kk * 8 + t;
        af[mt][0] = tf32_bits(p[0]);
        af[mt][1] = tf32_bits(p[8 * LDS]);
        af[mt][2] = tf32_bits(p[4]);
        af[mt][3] = tf32_bits(p[8 * LDS + 4]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* p = bs + (wn * 32 + nt * 8 + g) * LDS + kk * 8 + t;
        bf[nt][0] = tf32_bits(p[0]);
        bf[nt][1] = tf32_bits(p[4]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32_16x8x8(acc[mt][nt], af[mt], bf[nt]);
    }
  }
  cp_async_wait<0>();

  // epilogue
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      int r = m0 + wm * WM + mt * 16 + g + half * 8;
      if (r >= M) continue;
      size_t orow = (size_t)r;
      if (epi.mode == kEpiResidualUnpart) {
        int tok = epi.grid.slot_to_token(r);
        if (tok >= epi.grid.L) continue;
        orow = (size_t)tok;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        int c = n0 + wn * 32 + nt * 8 + 2 * t;
        if (c >= N) continue;
        float v0 = acc[mt][nt][half * 2 + 0], v1 = acc[mt][nt][half * 2 + 1];
        if (epi.bias) { v0 += __ldg(epi.bias + c); v1 += __ldg(epi.bias + c + 1); }
        if (epi.mode == kEpiTanh) { v0 = tanhf(v0); v1 = tanhf(v1); }
        if (epi.mode == kEpiResidualUnpart) {
          float2 rr = __ldg(reinterpret_cast<const float2*>(epi.resid + orow * N + c));
          v0 += rr.x; v1 += rr.y;
        }
        *reinterpret_cast<float2*>(C + orow * N + c) = make_float2(v0, v1);
      }
    }
  }
}

template <int BM>
cudaError_t launch(const float* a, const float* w, float* c, int M, int N, int K,
                   const GemmEpilogue& epi, cudaStream_t stream) {
  size_t smem = (size_t)STAGES * (BM + BN) * LDS * sizeof(float);
  static DeviceOnce configured;  // per template instance
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_mma_kernel<BM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_tf32_mma_kernel<BM><<<grid, NTHREADS, smem, stream>>>(a, w, c, M, N, K, epi);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_gemm_mma(const float* a, const float* w, float* c, int M, int N, int K,
                            const GemmEpilogue& epi, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (K <= 0 || K % BK || (N & 1)) return cudaErrorInvalidValue;
  if (M <= 1024) return launch<64>(a, w, c, M, N, K, epi, stream);
  return launch<128>(a, w, c, M, N, K, epi, stream);
}

}  // namespace rrt
