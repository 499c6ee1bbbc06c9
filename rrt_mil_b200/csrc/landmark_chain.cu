// CR-MSA landmark MHA as ONE kernel (modules/rmsa.py:318-321: the nn-style MHA over the k landmark batches of
// 64 region landmarks each):   lout = proj( softmax(q k^T / sqrt(hd)) v ),  [q|k|v] = lm Wqkv^T + b
//
// Round 1 ran it as three launches (narrow tcgen05 GEMM -> mma.sync attention -> narrow tcgen05 GEMM): 0.43 GFLOP
// on grids of 48 / 24 / 16 CTAs, ~30 us of launch + pipeline ramp for ~3 us of work.  Here: one launch, one
// thread-block CLUSTER per 128-row tile (= two landmark batches), CTA = head (head_dim 64, heads <= 8):
//   phase 1  lm[tile rows x D] . Wqkv[head rows]^T on tcgen05 (M128 N192 K16; TMA ring of [128 x 64] lm boxes and
//            3 x [64 x 64] weight boxes, SWIZZLE_128B, 4 stages), accumulator in TMEM -> +bias -> f16 q|k|v tile
//            in shared memory
//   phase 2  the tile's 2 batches x 4 row blocks = 8 units, one per warp: S = q k^T (mma.sync m16n8k16), softmax in
//            registers, O = P v  ->  lo[T, D] (f16, global)
//   cluster barrier (release / acquire + proxy fences): every head's columns of the tile's lo rows are visible to
//            the TMA of every CTA of the cluster
//   phase 3  CTA j computes output columns [64 j, 64 j + 64) of the tile's rows: lo . Wp[rows 64 j..]^T on tcgen05
//            (M128 N64; the ring re-cut into 6 stages of 24 KB), +bias -> lout fp32
// The kernel is latency-bound (8 k-blocks of 40 KB, then 8 of 24 KB per CTA): the ring depth, not the tensor
// core, sets its time.
// The training forward uses it too: it then also stores the f16 q|k|v rows the backward re-reads (lqkv).
#include <cuda.h>
#include "kernels.cuh"
#include "mma_f16.cuh"
#include "sm100.cuh"

namespace rrt {
// gemm_tcgen05.cu: cached CUtensorMap of a row-major fp16 [rows, K] tensor, boxes of box_rows x 64, 128-B swizzle
bool tc05_make_kmajor_map(CUtensorMap* m, const __half* base, int rows, int K, int box_rows);

namespace {
using namespace sm100;

constexpr int HD = 64;                       // head_dim
constexpr int STAGES = 4;                    // phase 1: [A | B1] stages of 40 KB
constexpr int STAGES3 = 6;                   // phase 3: [A | B3] stages of 24 KB in the same 160 KB
constexpr int A_BYTES = 128 * 64 * 2;        // [128 rows x 64 k] f16
constexpr int B1_BYTES = 192 * 64 * 2;       // q, k, v weight rows of one head
constexpr int B3_BYTES = 64 * 64 * 2;        // proj weight rows of one column slice
constexpr int STAGE_BYTES = A_BYTES + B1_BYTES;
constexpr int STAGE3_BYTES = A_BYTES + B3_BYTES;
static_assert(STAGES3 * STAGE3_BYTES <= STAGES * STAGE_BYTES, "phase-3 ring must fit in the phase-1 ring");
constexpr int LDQ = 3 * HD + 8;              // halves per row of the q|k|v tile (pitch 400 B: ldmatrix conflict-free)
constexpr int TILE_BYTES = 128 * LDQ * 2;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + TILE_BYTES + (192 + 64) * 4 + 256;
constexpr int TMEM_COLS = 256;               // [0, 192) phase-1 accumulator, [192, 256) phase-3 accumulator
constexpr int NTHREADS = 256;

struct LmParams {
  int T, D, heads;
  const float* qkv_b;   // [3 D] or null
  const float* proj_b;  // [D] or null
  __half* lqkv;         // [T, 3 D] or null: copy of q|k|v for the training tape
  __half* lo;           // [T, D]
  float* lout;          // [T, D]
  float scale_log2;     // head_dim^-0.5 * log2(e)
};

__global__ void __launch_bounds__(NTHREADS, 1)
landmark_chain_kernel(const __grid_constant__ CUtensorMap tmLm, const __grid_constant__ CUtensorMap tmWq,
                      const __grid_constant__ CUtensorMap tmLo, const __grid_constant__ CUtensorMap tmWp,
                      LmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* ring = smem;
  __half* tile = reinterpret_cast<__half*>(smem + STAGES * STAGE_BYTES);
  float* sbias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + TILE_BYTES);   // [192] qkv | [64] proj
  uint64_t* full = reinterpret_cast<uint64_t*>(sbias + 256);
  uint64_t* empty = full + STAGES;
  uint64_t* full3 = empty + STAGES;
  uint64_t* empty3 = full3 + STAGES3;
  uint64_t* acc_full = empty3 + STAGES3;  // [2]: phase 1, phase 3
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = (int)cluster_ctarank();   // head of phases 1-2, output column slice of phase 3
  const int mt = blockIdx.x / p.heads;    // 128-row tile of this cluster
  const int T = p.T, D = p.D, KB = D / 64;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmLm); prefetch_tensormap(&tmWq); }
  if (warp == 3 && lane == 0) { prefetch_tensormap(&tmLo); prefetch_tensormap(&tmWp); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < STAGES3; ++i) { mbar_init(&full3[i], 1); mbar_init(&empty3[i], 1); }
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  // weights only: biases of this head / this column slice (may run ahead of the predecessor)
  if (tid < 192) sbias[tid] = p.qkv_b ? __ldg(p.qkv_b + (tid >> 6) * D + h * HD + (tid & 63)) : 0.f;
  else sbias[tid] = p.proj_b ? __ldg(p.proj_b + h * HD + (tid - 192)) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ================= phase 1: q|k|v of (tile mt, head h) =============================================
  if (tid == 0) {
    int s = 0, ph = 0;
    // the weight boxes of the first stages do not depend on the predecessor: issue them before the PDL wait
    const int pre = KB < STAGES ? KB : STAGES;
    for (int kb = 0; kb < pre; ++kb) {
      mbar_arrive_expect_tx(&full[kb], STAGE_BYTES);
#pragma unroll
      for (int j = 0; j < 3; ++j)
        tma_load_2d(ring + kb * STAGE_BYTES + A_BYTES + j * (64 * 64 * 2), &tmWq, &full[kb], kb * 64, j * D + h * HD);
    }
    pdl_wait();   // lm is the predecessor's output
    for (int kb = 0; kb < KB; ++kb) {
      uint8_t* st = ring + s * STAGE_BYTES;
      if (kb >= pre) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tma_load_2d(st + A_BYTES + j * (64 * 64 * 2), &tmWq, &full[s], kb * 64, j * D + h * HD);
      }
      tma_load_2d(st, &tmLm, &full[s], kb * 64, mt * 128);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (tid == 32) {
    constexpr uint32_t idesc = umma_idesc(kFmtF16, 128, 192);
    int s = 0, ph = 0;
    for (int kb = 0; kb < KB; ++kb) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint64_t ad = umma_desc_k_sw128(smem_u32(ring + s * STAGE_BYTES));
      const uint64_t bd = umma_desc_k_sw128(smem_u32(ring + s * STAGE_BYTES + A_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
      umma_commit(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    umma_commit(&acc_full[0]);
  }
  if (warp >= 4) {   // accumulator -> (+bias) f16 q|k|v tile; thread = row
    const int quad = warp & 3, row = quad * 32 + lane;
    mbar_wait(&acc_full[0], 0);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 6; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(t_addr + c * 32, r);
      tmem_ld_wait();
      uint4* dst = reinterpret_cast<uint4*>(tile + (size_t)row * LDQ + c * 32);
      // tape copy: chunk c = section c / 2 (q, k, v) of this head, columns (c & 1) * 32 .. + 31
      const int grow = mt * 128 + row;
      uint4* gdst = (p.lqkv && grow < T)
                        ? reinterpret_cast<uint4*>(p.lqkv + (size_t)grow * 3 * D + (c >> 1) * D + h * HD + (c & 1) * 32)
                        : nullptr;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = pack_h2(__uint_as_float(r[8 * q + 2 * e]) + sbias[c * 32 + 8 * q + 2 * e],
                         __uint_as_float(r[8 * q + 2 * e + 1]) + sbias[c * 32 + 8 * q + 2 * e + 1]);
        dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
        if (gdst) gdst[q] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  // the phase-3 weight boxes do not depend on anything this kernel computes: stream them under the attention
  // (every phase-1 MMA has completed -- acc_full -- so the ring is free)
  if (tid == 0) {
    const int pre3 = KB < STAGES3 ? KB : STAGES3;
    for (int kb = 0; kb < pre3; ++kb) {
      mbar_arrive_expect_tx(&full3[kb], STAGE3_BYTES);
      tma_load_2d(ring + kb * STAGE3_BYTES + A_BYTES, &tmWp, &full3[kb], kb * 64, h * HD);
    }
  }
  // ================= phase 2: attention, unit = (batch of the tile, 16-row block) = warp ================
  {
    const int bb = warp >> 2, r0 = bb * 64 + (warp & 3) * 16;   // rows of this warp inside the tile
    if (mt * 128 + bb * 64 < T) {
      const int g = lane >> 2, t = lane & 3;
      const __half* Q = tile;                 // columns [0, 64)
      const __half* K = tile + HD;            // columns [64, 128), rows of batch bb
      const __half* V = tile + 2 * HD;
      uint32_t qa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) load_a_rowmajor<LDQ>(qa[ks], Q, r0, ks, lane);
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          load_b_nk<LDQ>(b, K, bb * 64 + np * 16, ks, lane);
          mma_f16_16x8x16(s[2 * np], qa[ks], b[0], b[1]);
          mma_f16_16x8x16(s[2 * np + 1], qa[ks], b[2], b[3]);
        }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[nt][e] *= p.scale_log2;
          mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
        }
      float l[2] = {0.f, 0.f};
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = fast_exp2(s[nt][e] - mx[e >> 1]);
          l[e >> 1] += pv;
          s[nt][e] = pv;
        }
      float o[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // 16 keys per step
        uint32_t pa[4] = {pack_h2(s[2 * j][0], s[2 * j][1]), pack_h2(s[2 * j][2], s[2 * j][3]),
                          pack_h2(s[2 * j + 1][0], s[2 * j + 1][1]), pack_h2(s[2 * j + 1][2], s[2 * j + 1][3])};
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t b[4];
          load_b_kn<LDQ>(b, V, bb * 64 + j * 16, np * 16, lane);
          mma_f16_16x8x16(o[2 * np], pa, b[0], b[1]);
          mma_f16_16x8x16(o[2 * np + 1], pa, b[2], b[3]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        l[hh] += __shfl_xor_sync(0xffffffffu, l[hh], 1);
        l[hh] += __shfl_xor_sync(0xffffffffu, l[hh], 2);
        const float inv = 1.f / l[hh];
        __half* orow = p.lo + (size_t)(mt * 128 + r0 + g + hh * 8) * D + h * HD + 2 * t;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd)
          *reinterpret_cast<uint32_t*>(orow + nd * 8) = pack_h2(o[nd][hh * 2] * inv, o[nd][hh * 2 + 1] * inv);
      }
    }
  }

  // ================= every head's lo columns of this tile -> visible to the cluster's TMA ============
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  cluster_sync_all();
  asm volatile("fence.proxy.async;" ::: "memory");

  // ================= phase 3: output columns [64 h, 64 h + 64) of the tile's rows ====================
  if (tid == 0) {
    int s = 0, ph = 0;
    const int pre3 = KB < STAGES3 ? KB : STAGES3;
    for (int kb = 0; kb < KB; ++kb) {
      uint8_t* st = ring + s * STAGE3_BYTES;
      if (kb >= pre3) {
        mbar_wait(&empty3[s], ph ^ 1);
        mbar_arrive_expect_tx(&full3[s], STAGE3_BYTES);
        tma_load_2d(st + A_BYTES, &tmWp, &full3[s], kb * 64, h * HD);
      }
      tma_load_2d(st, &tmLo, &full3[s], kb * 64, mt * 128);
      if (++s == STAGES3) { s = 0; ph ^= 1; }
    }
  } else if (tid == 32) {
    constexpr uint32_t idesc = umma_idesc(kFmtF16, 128, 64);
    int s = 0, ph = 0;
    for (int kb = 0; kb < KB; ++kb) {
      mbar_wait(&full3[s], ph);
      tc_fence_after();
      const uint64_t ad = umma_desc_k_sw128(smem_u32(ring + s * STAGE3_BYTES));
      const uint64_t bd = umma_desc_k_sw128(smem_u32(ring + s * STAGE3_BYTES + A_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 192, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
      umma_commit(&empty3[s]);
      if (++s == STAGES3) { s = 0; ph ^= 1; }
    }
    umma_commit(&acc_full[1]);
  }
  if (warp >= 4) {
    const int quad = warp & 3;
    mbar_wait(&acc_full[1], 0);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + 192 + ((uint32_t)(quad * 32) << 16);
    const int row = mt * 128 + quad * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(t_addr + c * 32, r);
      tmem_ld_wait();
      if (row < T) {
        float4* dst = reinterpret_cast<float4*>(p.lout + (size_t)row * D + h * HD + c * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          dst[q] = make_float4(__uint_as_float(r[4 * q]) + sbias[192 + c * 32 + 4 * q],
                               __uint_as_float(r[4 * q + 1]) + sbias[192 + c * 32 + 4 * q + 1],
                               __uint_as_float(r[4 * q + 2]) + sbias[192 + c * 32 + 4 * q + 2],
                               __uint_as_float(r[4 * q + 3]) + sbias[192 + c * 32 + 4 * q + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}
}  // namespace

bool landmark_chain_supported(int k, int D, int heads) {
  return k >= 1 && k <= RRT_MAX_K_DEV && heads >= 1 && heads <= 8 && D == heads * HD;
}

// lm [k*64, D] f16 (landmarks, batch-major), wq [3D, D] / wp [D, D] f16; lo [k*64, D] f16 and lout [k*64, D] fp32 out
cudaError_t launch_landmark_chain(const __half* lm, const __half* wq, const __half* wp, const float* qkv_b,
                                  const float* proj_b, __half* lqkv, __half* lo, float* lout, int k, int D,
                                  int heads, cudaStream_t stream) {
  if (!landmark_chain_supported(k, D, heads)) return cudaErrorInvalidValue;
  const int T = k * 64;
  CUtensorMap tmLm, tmWq, tmLo, tmWp;
  if (!tc05_make_kmajor_map(&tmLm, lm, T, D, 128) || !tc05_make_kmajor_map(&tmWq, wq, 3 * D, D, 64) ||
      !tc05_make_kmajor_map(&tmLo, lo, T, D, 128) || !tc05_make_kmajor_map(&tmWp, wp, D, D, 64))
    return cudaErrorUnknown;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(landmark_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  LmParams p;
  p.T = T; p.D = D; p.heads = heads;
  p.qkv_b = qkv_b; p.proj_b = proj_b; p.lqkv = lqkv; p.lo = lo; p.lout = lout;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(heads * ((T + 127) / 128));   // one cluster of `heads` CTAs per 128-row tile
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = heads; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl || g_pdl_light) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, landmark_chain_kernel, tmLm, tmWq, tmLo, tmWp, p);
}

}  // namespace rrt
