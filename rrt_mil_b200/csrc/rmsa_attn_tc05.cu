// R-MSA attention core on the 5th-gen tensor cores (head_dim 64, regions of up to 256 tokens):
// a PERSISTENT, warp-specialised kernel, one CTA per SM, each CTA walking (region, head) items with TWO items
// in flight (softmax warpgroups 0 and 1 take alternate items, as in FlashAttention-4's ping-pong).
//
//   warp 8       issuer (one thread).  TMA: raw Q rows (with the EPEG halo), K and V head slices of an item
//                straight out of the [R][P][3D] view of the QKV GEMM's output (3-D tensor maps, SWIZZLE_128B;
//                rows outside the region are zero-filled by the TMA unit = the conv's zero padding and the
//                key padding), up to three items ahead.  MMA: S = Q' K^T (tcgen05.mma M=128, N=P16, K=64,
//                operands in smem), O = P V (A = P read from TENSOR MEMORY, B = V consumed MN-major from
//                smem) and the row sums l = P 1 (B = a tile of ones).  TMEM allocator.
//   warps 9..11  helpers (mma.sync): EPEG Q' = scale*log2e * (Q + dwconv1d_P(Q; taps_h)) as a banded-Toeplitz
//                product, written to shared memory in the UMMA K-major swizzled layout; and the region's TAIL
//                rows (query rows >= 128: 16 of them at P = 144), 16 rows per warp step with ldmatrix on the
//                swizzled K / V tiles and an online softmax in registers -- a second 128-row tcgen05 block
//                for 16 rows would cost a whole TMEM slot and a fifth, half-empty softmax warp on one scheduler.
//   warps 0..3   softmax + epilogue of query rows 0..127 of the EVEN items (one thread per row = TMEM lane):
//   warps 4..7   ... of the ODD items.  Row max, exp2; P is written back to TMEM as packed f16 over the dead
//                S columns (tcgen05.st), so it never touches shared memory; O (fp32, TMEM) / l -> f16 ->
//                swizzled staging tile -> TMA store.
// TMEM per warpgroup: [ S: P16 fp32 columns | P over S[0, P16/2) | l (16) | O (64) over the tail of S ].
// All hand-offs are mbarriers (no __syncthreads after the prologue).
// (modules/rmsa.py:103-122; SURVEY.md 0.2-1 for the EPEG-on-Q identity.)
#include "kernels.cuh"
#include "mma_f16.cuh"
#include "sm100.cuh"

namespace rrt {
namespace {
using namespace sm100;

constexpr int HD = 64;
// 16 warps = 512 threads, 128 registers per thread.  The helpers are single warps running long dependent
// chains on the legacy tensor path (ldmatrix -> mma.sync -> shuffle -> exp2 -> mma.sync; measured ~620 cycles per
// 16-row EPEG tile, ~6 k cycles per 16-row tail tile): there are seven of them so that their latency is hidden by
// each other, not by the softmax warps waiting on them.
constexpr int kIssuerWarp = 8;
constexpr int kFirstHelper = 9, kHelpers = 7;
constexpr int kThreads = 32 * 16;
constexpr int kStageBytes = 32 * 128;  // per softmax warp: 32 output rows x 64 f16, swizzled, source of the TMA store

enum Bar {
  kKvFull = 0, kKvEmpty = 3, kQFull = 6, kQEmpty = 8, kQpFull = 10, kQpEmpty = 12,
  kSFull = 14, kPReady = 16, kOFull = 18, kODone = 20, kNumBars = 22
};

// The kernel has four warp roles with disjoint code; all of it has to stay resident in the SM's instruction
// cache (a first version with every wait loop and issue lambda inlined at each call site was 65 KB of SASS and
// was slower for it).  Hence: one call site per helper function.
__device__ __forceinline__ void mbar_wait_c(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned t0 = (unsigned)clock();
  while (!mbar_try_wait(bar, parity)) {
    if ((unsigned)clock() - t0 > 4000000000u) __trap();  // ~2 s of SM clocks: a pipeline bug, not a wait
  }
}

// ---- PTX not in sm100.cuh ------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A is [128 lanes x K] packed f16 (two per 32-bit column), K-major
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float tmem_ld_x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// instruction descriptor: f16 operands, fp32 accumulate, A K-major, B K-major (b_mn = 0) or MN-major (1)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_mn) {
  return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major B operand (one 128-byte swizzled row of 64 elements per K index, 8 K indices per 1024-byte atom):
// SBO = distance between 8-row K groups, LBO = distance to the next 64-element MN chunk.  O and the row sums
// come out of ONE product with N = 80: columns 0..63 = V, columns 64..79 = a chunk of ones (the A operand P is
// read from tensor memory once; a separate product for the sums would read it twice, and that read is the cost)
__device__ __forceinline__ uint64_t umma_desc_v_sw128(uint32_t smem_addr, uint32_t lbo_bytes = 1024) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct AttnTcParams {
  const float* taps;   // [heads, epeg_k] or null
  __half* o;           // [R*P, D] (tail rows are stored directly)
  long long* trace;    // debug: clock64 stamps of CTA 0 (tools/attn_probe.py), or null
  int P, P16, heads, D, epeg_k, pad, nkc, q_rows, ntail, n_items;
  int kv_stages, q_stages;                 // ring depths (3 / 2 when shared memory allows)
  int tail_helpers, epeg_helpers;          // 2 + 5 dedicated warps (<= 2 tail tiles per item), else all 7 do both
  int tail_part;                           // min(ntail, tail_helpers): tail helpers that work on any one item
  int sep_o;                               // 1: O | l in columns of their own (P16 <= 160): S of item n + 2 is issued
                                           // right behind O of item n instead of after its epilogue
  uint32_t slot_stride, o_off;             // TMEM columns of one warpgroup's slot: S, P at 0; O | l at o_off
  uint32_t tmem_cols, q_bytes, kv_bytes, qp_bytes;
  float qscale;
};

__device__ __forceinline__ void stamp(const AttnTcParams& p, int n, int slot) {
  if (p.trace && blockIdx.x == 0 && n < 7) p.trace[n * 8 + slot] = clock64();
}
// rows 8..15: helper 0 (per step), rows 16..23: issuer (per item)
__device__ __forceinline__ void stamp2(const AttnTcParams& p, int base, int n, int slot) {
  if (p.trace && blockIdx.x == 0 && n < 8) p.trace[(base + n) * 8 + slot] = clock64();
}

// byte offset of 16-byte chunk `chunk` of row `row` in a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

// ---- helper-warp work ---------------------------------------------------------------------------------------
// EPEG of the 16 query rows i0..i0+15: Q' = qscale * (Q + conv(Q)) -> Q' buffer (K-major, swizzled)
__device__ __forceinline__ void epeg_tile(const AttnTcParams& p, const uint8_t* q, uint8_t* qp, const float* Tw,
                                          const uint32_t (&ca01)[2][4], int i0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float qacc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) qacc[i][e] = 0.f;
  for (int kc = 0; kc < p.nkc; ++kc) {
    // A fragment of the Toeplitz band: element (row i, halo row r) = taps[r - i] (+1 at r - i = pad); the
    // first two 16-column steps (all of it for epeg_k <= 17) are precomputed per item
    uint32_t ca[4];
    if (kc < 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) ca[e] = ca01[kc][e];
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ro = g + (e & 1) * 8;
        const int co = 16 * kc + 2 * t + (e >> 1) * 8;
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int d = co + u - ro;
          v[u] = ((d >= 0 && d < 64) ? Tw[d] : 0.f) + (d == p.pad ? 1.f : 0.f);
        }
        ca[e] = pack_h2(v[0], v[1]);
      }
    }
    const int row = i0 + 16 * kc + (lane & 7) + ((lane >> 3) & 1) * 8;  // halo row
#pragma unroll
    for (int np = 0; np < HD / 16; ++np) {
      uint32_t b[4];
      ldsm_x4_trans(b, q + swz(row, np * 2 + (lane >> 4)));
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
      *reinterpret_cast<uint32_t*>(qp + swz(row, nt) + t * 4) =
          pack_h2(qacc[nt][hh * 2] * p.qscale, qacc[nt][hh * 2 + 1] * p.qscale);
    }
  }
}

// softmax(Q' K^T) V of the 16 query rows r0..r0+15 on mma.sync, 8 * NT keys per online-softmax step
template <int NT>
__device__ __forceinline__ void tail_tile(const AttnTcParams& p, const uint8_t* qp, const uint8_t* k,
                                          const uint8_t* v, __half* o_item, int r0, int lane, uint64_t* qp_free) {
  const int g = lane >> 2, t = lane & 3, P = p.P;
  uint32_t qa[HD / 16][4];
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) ldsm_x4(qa[ks], qp + swz(r0 + (lane & 15), ks * 2 + (lane >> 4)));
  if (qp_free) {  // Q' is in registers: the buffer may take the EPEG of item n + 2 while this tile runs
    __syncwarp();
    if (lane == 0) mbar_arrive(qp_free);
  }
  float oacc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
#pragma unroll 1
  for (int k0 = 0; k0 < p.P16; k0 += 8 * NT) {
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b[4];
        ldsm_x4(b, k + swz(k0 + np * 16 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)));
        mma_f16_16x8x16(s[2 * np], qa[ks], b[0], b[1]);
        mma_f16_16x8x16(s[2 * np + 1], qa[ks], b[2], b[3]);
      }
    }
    if (k0 + 8 * NT > P) {  // tile padding keys (zero-pad TOKENS are real keys and stay unmasked)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k0 + nt * 8 + 2 * t + (e & 1) >= P) s[nt][e] = -INFINITY;
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
      mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
    }
    const float corr[2] = {fast_exp2(m_run[0] - mx[0]), fast_exp2(m_run[1] - mx[1])};
    m_run[0] = mx[0];
    m_run[1] = mx[1];
    l_run[0] *= corr[0];
    l_run[1] *= corr[1];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      oacc[i][0] *= corr[0]; oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1]; oacc[i][3] *= corr[1];
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = fast_exp2(s[nt][e] - mx[e >> 1]);
        l_run[e >> 1] += pv;
        s[nt][e] = pv;
      }
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {  // 16 keys per step: two S n-tiles form one A fragment
      const uint32_t pa[4] = {pack_h2(s[2 * j][0], s[2 * j][1]), pack_h2(s[2 * j][2], s[2 * j][3]),
                              pack_h2(s[2 * j + 1][0], s[2 * j + 1][1]), pack_h2(s[2 * j + 1][2], s[2 * j + 1][3])};
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {
        uint32_t b[4];
        ldsm_x4_trans(b, v + swz(k0 + j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 2 + (lane >> 4)));
        mma_f16_16x8x16(oacc[2 * np], pa, b[0], b[1]);
        mma_f16_16x8x16(oacc[2 * np + 1], pa, b[2], b[3]);
      }
    }
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int q = r0 + g + hh * 8;
    if (q >= P) continue;
    const float inv = 1.f / l_run[hh];
    __half* orow = o_item + (size_t)q * p.D + 2 * t;
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
      *reinterpret_cast<uint32_t*>(orow + nd * 8) = pack_h2(oacc[nd][hh * 2] * inv, oacc[nd][hh * 2 + 1] * inv);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
rmsa_attn_tc05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const __grid_constant__ CUtensorMap tmO, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQp = smem;                                             // 2 x [max(P16,128)][128 B] swizzled
  uint8_t* sQraw = sQp + 2 * (size_t)p.qp_bytes;                   // q_stages x [q_rows][128 B] swizzled (TMA)
  uint8_t* sK = sQraw + (size_t)p.q_stages * p.q_bytes;            // kv_stages x [P16][128 B]
  uint8_t* sV = sK + (size_t)p.kv_stages * p.kv_bytes;             // kv_stages x [P16][128 B]
  uint8_t* sStage = sV + (size_t)p.kv_stages * p.kv_bytes;         // 8 x [32][128 B] swizzled
  uint8_t* sOnes = sStage + 8 * kStageBytes;                       // [16][128 B] of f16 1.0: B operand of the row sums
  float* sT = reinterpret_cast<float*>(sOnes + 2048);              // [kHelpers][64] taps
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + 8 * 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  if (tid == 0) stamp(p, 6, 6);
  if (tid == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmKV);
    prefetch_tensormap(&tmO);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bars[kKvFull + i], 1);
      mbar_init(&bars[kKvEmpty + i], 1 + p.tail_part);   // O of the item committed + the tail helpers that have a tile of it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[kQFull + i], 1);
      mbar_init(&bars[kQEmpty + i], p.epeg_helpers);
      mbar_init(&bars[kQpFull + i], p.epeg_helpers);
      mbar_init(&bars[kQpEmpty + i], 1 + p.tail_part);   // S of the item committed + those tail helpers have their A fragments
      mbar_init(&bars[kSFull + i], 1);
      mbar_init(&bars[kPReady + i], 4);
      mbar_init(&bars[kOFull + i], 1);
      mbar_init(&bars[kODone + i], 4);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 2048 / 4; i += kThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  if (warp == kIssuerWarp) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything below reads what the QKV GEMM wrote / writes what the proj GEMM reads
  if (tid == 0) stamp(p, 6, 7);

  const int P = p.P, P16 = p.P16;
  const int first = blockIdx.x, step = gridDim.x;
  const int total = first < p.n_items ? (p.n_items - first + step - 1) / step : 0;  // items of this CTA
  const int KS = p.kv_stages, QS = p.q_stages;

  if (warp == kIssuerWarp) {
    // ===== issuer: TMA loads and every tcgen05.mma ============================================================
    if (lane == 0) {
      const uint32_t idesc_s = idesc_f16(128, P16, 0);
      const uint32_t idesc_o = idesc_f16(128, HD + 16, 1);
      const uint32_t ones_a = smem_u32(sOnes);
      const int ksteps = P16 / 16;
      // Start-up: this thread requests item 0 only; the first fill of the other ring slots (items 1 .. stages - 1)
      // is issued by the first tail helper, which has nothing to do until the first S operands exist.  The single
      // issuing thread needs ~800 cycles per TMA request with cold code: eight of them in front of the first S held
      // it 2.6 k cycles behind its operands (tools/attn_probe.py --trace).
      int nlq = total < QS ? total : QS, nlk = total < KS ? total : KS;
      // (region, head) of the next Q / K|V request, advanced by one grid stride per item: no divisions in the loop
      const int d_rho = step / p.heads, d_h = step - d_rho * p.heads;
      int rho_q, h_q, rho_k, h_k;
      {
        const int iq = first + nlq * step, ik = first + nlk * step;
        rho_q = iq / p.heads; h_q = iq - rho_q * p.heads;
        rho_k = ik / p.heads; h_k = ik - rho_k * p.heads;
      }
      // ring stage / phase of the next request (sq, uq | sk, uk), of the item whose O is next (so) and of the item
      // whose S is next (ss, us): counters instead of runtime % and / by the ring depths
      int sq = nlq == QS ? 0 : nlq, uq = nlq == QS ? 1 : 0;
      int sk = nlk == KS ? 0 : nlk, uk = nlk == KS ? 1 : 0;
      int so = 0, ss = 0, us = 0;
      if (total > 0) {
        const int rho = first / p.heads, h = first - rho * p.heads;
        mbar_arrive_expect_tx(&bars[kQFull], p.q_bytes);
        tma_load_3d(sQraw, &tmQ, &bars[kQFull], h * HD, -p.pad, rho);
        mbar_arrive_expect_tx(&bars[kKvFull], 2 * p.kv_bytes);
        tma_load_3d(sK, &tmKV, &bars[kKvFull], p.D + h * HD, 0, rho);
        tma_load_3d(sV, &tmKV, &bars[kKvFull], 2 * p.D + h * HD, 0, rho);
      }
      for (int j = 0; j < total + 2; ++j) {   // step j: O of item j - 2, loads, S of item j
        const int n = j - 2;
        if (n >= 0) {
          stamp2(p, 16, n, 0);
          const int slot = n & 1, u = n >> 1;
          mbar_wait_c(&bars[kPReady + slot], u & 1);
          if (p.sep_o && u > 0) mbar_wait_c(&bars[kODone + slot], (u - 1) & 1);  // epilogue of item n - 2 has read O
          tc_fence_after();
          stamp2(p, 16, n, 1);
          const uint32_t tP = tmem_base + (uint32_t)slot * p.slot_stride;
          const uint32_t tO = tP + p.o_off;
          // O | l = P [V | 1]: the second 64-column chunk of every K step is the same 2 KB tile of ones (LBO is
          // per descriptor), operands advance by constants
          uint32_t va = smem_u32(sV + (size_t)so * p.kv_bytes), ta = tP;
          for (int ks = 0; ks < ksteps; ++ks) {
            umma_f16_ts(tO, ta, umma_desc_v_sw128(va, ones_a - va), idesc_o, ks != 0);
            va += 2048;
            ta += 8;
          }
          umma_commit(&bars[kOFull + slot]);
          umma_commit(&bars[kKvEmpty + so]);  // K and V of this stage are consumed (by the tensor core)
          if (++so == KS) so = 0;
          stamp2(p, 16, n, 2);
        }
        // S of item j goes out right behind O of item j - 2 when its operands were requested in an earlier step
        // (the usual case from j = 2 on); otherwise the loads first.  (One code copy of each: the flag only
        // orders the two sections.)
        // (step 0 keeps the loads first: nothing can be requested yet, but the section's first, instruction-cache-cold
        // pass costs ~2 k cycles, and there it runs while the first EPEG is still in progress)
        const bool s_first = j > 0 && j < total && nlq > j && nlk > j;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          if ((pass == 0) == s_first) {
            if (j < total) {
              // the slot's S / P columns are free: aliased layout -> once the epilogue of item n has read O (which
              // lives over S); separate layout -> already (the tensor pipe runs O of item n, which reads P, before
              // this S: same issuing thread, in order; every softmax thread has read S: p_ready)
              if (n >= 0 && !p.sep_o) mbar_wait_c(&bars[kODone + (n & 1)], (n >> 1) & 1);
              if (n >= 0) stamp2(p, 16, n, 3);
              if (j < 2) stamp2(p, 24, j, 0);
              mbar_wait_c(&bars[kQpFull + (j & 1)], (j >> 1) & 1);
              if (j < 2) stamp2(p, 24, j, 1);
              mbar_wait_c(&bars[kKvFull + ss], us);
              if (j < 2) stamp2(p, 24, j, 2);
              tc_fence_after();
              const uint64_t ad = umma_desc_k_sw128(smem_u32(sQp + (size_t)(j & 1) * p.qp_bytes));
              const uint64_t kd = umma_desc_k_sw128(smem_u32(sK + (size_t)ss * p.kv_bytes));
              if (++ss == KS) { ss = 0; us ^= 1; }
              const uint32_t tS = tmem_base + (uint32_t)(j & 1) * p.slot_stride;
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) umma_f16(tS, ad + 2 * k, kd + 2 * k, idesc_s, k != 0);
              umma_commit(&bars[kSFull + (j & 1)]);
              umma_commit(&bars[kQpEmpty + (j & 1)]);  // the S product has read rows 0..127 of Q'
              if (j < 2) stamp2(p, 24, j, 3);
              if (n >= 0) stamp2(p, 16, n, 4);
            }
          } else {
            // loads: those of item j must go out now (S of item j is next: wait for the ring slot if need be);
            // further ahead only into ring slots that are free already, so that a slow tail never blocks the MMAs.
            // Item by item (Q, then K | V of the same item): at kernel start every CTA requests up to three items
            // at once (21 MB over the chip), and the first S waits for whatever was queued ahead of its K tile
            for (;;) {
              const bool q_ok = nlq < total && nlq <= j + QS, k_ok = nlk < total && nlk <= j + KS;
              const bool q_turn = q_ok && (nlq <= nlk || !k_ok);
              bool did = false;
              if (q_turn) {
                const int st = sq;
                bool go = true;
                if (nlq <= j) mbar_wait_c(&bars[kQEmpty + st], uq ^ 1);
                else go = mbar_test(&bars[kQEmpty + st], uq ^ 1);
                if (go) {
                  mbar_arrive_expect_tx(&bars[kQFull + st], p.q_bytes);
                  tma_load_3d(sQraw + (size_t)st * p.q_bytes, &tmQ, &bars[kQFull + st], h_q * HD, -p.pad, rho_q);
                  ++nlq;
                  if (++sq == QS) { sq = 0; uq ^= 1; }
                  rho_q += d_rho; h_q += d_h;
                  if (h_q >= p.heads) { h_q -= p.heads; ++rho_q; }
                  did = true;
                }
              }
              if (!did && k_ok) {
                const int st = sk;
                bool go = true;
                if (nlk <= j) mbar_wait_c(&bars[kKvEmpty + st], uk ^ 1);
                else go = mbar_test(&bars[kKvEmpty + st], uk ^ 1);
                if (go) {
                  mbar_arrive_expect_tx(&bars[kKvFull + st], 2 * p.kv_bytes);
                  tma_load_3d(sK + (size_t)st * p.kv_bytes, &tmKV, &bars[kKvFull + st], p.D + h_k * HD, 0, rho_k);
                  tma_load_3d(sV + (size_t)st * p.kv_bytes, &tmKV, &bars[kKvFull + st], 2 * p.D + h_k * HD, 0, rho_k);
                  ++nlk;
                  if (++sk == KS) { sk = 0; uk ^= 1; }
                  rho_k += d_rho; h_k += d_h;
                  if (h_k >= p.heads) { h_k -= p.heads; ++rho_k; }
                  did = true;
                }
              }
              if (!did) break;
            }
            if (n >= 0) stamp2(p, 16, n, 5);
          }
        }
      }
    }
  } else if (warp >= kFirstHelper) {
    // ===== helpers (mma.sync): tail rows (>= 128) and EPEG of every item ========================================
    // With at most two tail tiles per item (P <= 160) the roles are DEDICATED: warps 9, 10 take the tails (a tile
    // is ~6 k cycles of dependent mma.sync work), warps 11..15 the EPEG -- with a shared split the helper that
    // held a tail delayed its share of the EPEG of item n + 2, and with it S of item n + 2, by that much.  Larger
    // regions have more tail work than two warps can carry: there all seven helpers do both.
    const int w = warp - kFirstHelper, g = lane >> 2, t = lane & 3;
    const bool shared_roles = p.tail_helpers == kHelpers;
    const bool does_tail = shared_roles || w < p.tail_helpers, does_epeg = shared_roles || w >= p.tail_helpers;
    const int wt = w, we = shared_roles ? w : w - p.tail_helpers;   // index within the role group
    float* Tw = sT + w * 64;
    const int ntiles = P16 / 16;
    int cur_h = -1;
    uint32_t ca01[2][4];
    if (w == 0 && lane == 0) {   // first fill of ring slots 1 .. (see the issuer's start-up)
      for (int i = 1; i < total && (i < QS || i < KS); ++i) {
        const int item = first + i * step, rho = item / p.heads, h = item - rho * p.heads;
        if (i < QS) {
          mbar_arrive_expect_tx(&bars[kQFull + i], p.q_bytes);
          tma_load_3d(sQraw + (size_t)i * p.q_bytes, &tmQ, &bars[kQFull + i], h * HD, -p.pad, rho);
        }
        if (i < KS) {
          mbar_arrive_expect_tx(&bars[kKvFull + i], 2 * p.kv_bytes);
          tma_load_3d(sK + (size_t)i * p.kv_bytes, &tmKV, &bars[kKvFull + i], p.D + h * HD, 0, rho);
          tma_load_3d(sV + (size_t)i * p.kv_bytes, &tmKV, &bars[kKvFull + i], 2 * p.D + h * HD, 0, rho);
        }
      }
    }
    for (int stp = 0; stp < total + 2; ++stp) {   // step: tail rows of item stp - 2, then EPEG of item stp
      const int n = stp - 2;
      // Only the helpers that own a tail tile of item n take part in its hand-offs (min(ntail, tail_helpers) of them:
      // the barrier counts).  When every tail helper had to arrive for every item, the one busy with the 7 k-cycle tail
      // of item n - 1 held the Q' buffer and the K | V stage of item n (and with them EPEG and S of item n + 2).
      if (n >= 0 && does_tail && (wt + n) % p.tail_helpers < p.ntail) {
        // both waits also order this helper's arrivals on the "empty" barriers behind the previous use of the stage
        mbar_wait_c(&bars[kQpFull + (n & 1)], (n >> 1) & 1);
        mbar_wait_c(&bars[kKvFull + n % KS], (n / KS) & 1);
        if (w == 0 && lane == 0) stamp2(p, 8, n, 1);
        bool qp_released = false;  // the Q' buffer is handed back as soon as this warp's last tile has its A fragments
        const int item = first + n * step, rho = item / p.heads, h = item - rho * p.heads;
        for (int tt = (wt + n) % p.tail_helpers; tt < p.ntail; tt += p.tail_helpers) {   // rotates over the tail helpers
          const bool last = tt + p.tail_helpers >= p.ntail;
          tail_tile<2>(p, sQp + (size_t)(n & 1) * p.qp_bytes, sK + (size_t)(n % KS) * p.kv_bytes,
                       sV + (size_t)(n % KS) * p.kv_bytes, p.o + (size_t)rho * P * p.D + h * HD, 128 + 16 * tt,
                       lane, last ? &bars[kQpEmpty + (n & 1)] : nullptr);
          qp_released = qp_released || last;
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[kKvEmpty + n % KS]);
          if (!qp_released) mbar_arrive(&bars[kQpEmpty + (n & 1)]);
        }
        if (w == 0 && lane == 0) stamp2(p, 8, n, 2);
      }
      if (stp < total && does_epeg) {
        const int m = stp;
        const int item = first + m * step, h = item % p.heads;
        if (h != cur_h) {   // (the head is the same for every item of a CTA when the grid is a multiple of heads)
          cur_h = h;
          __syncwarp();
          Tw[lane] = (p.taps && lane < p.epeg_k) ? __ldg(p.taps + h * p.epeg_k + lane) : 0.f;
          Tw[lane + 32] = (p.taps && lane + 32 < p.epeg_k) ? __ldg(p.taps + h * p.epeg_k + lane + 32) : 0.f;
          __syncwarp();
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ro = g + (e & 1) * 8;
              const int co = 16 * kc + 2 * t + (e >> 1) * 8;
              float v[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int d = co + u - ro;
                v[u] = ((d >= 0 && d < 64) ? Tw[d] : 0.f) + (d == p.pad ? 1.f : 0.f);
              }
              ca01[kc][e] = pack_h2(v[0], v[1]);
            }
        }
        const int qs = m % QS;
        mbar_wait_c(&bars[kQFull + qs], (m / QS) & 1);
        mbar_wait_c(&bars[kQpEmpty + (m & 1)], ((m >> 1) & 1) ^ 1);
        if (we == 0 && lane == 0) stamp2(p, 8, m, 3);
        for (int tile = (we + m) % p.epeg_helpers; tile < ntiles; tile += p.epeg_helpers)
          epeg_tile(p, sQraw + (size_t)qs * p.q_bytes, sQp + (size_t)(m & 1) * p.qp_bytes, Tw, ca01, 16 * tile, lane);
        fence_proxy_async_smem();  // Q' (generic-proxy writes) must be visible to the tensor core
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[kQpFull + (m & 1)]);
          mbar_arrive(&bars[kQEmpty + qs]);
        }
        if (we == 0 && lane == 0) stamp2(p, 8, m, 4);
      }
    }
  } else {
    // ===== softmax + epilogue of rows 0..127: warpgroup wg takes the items n = wg, wg + 2, ... ==================
    const int wg = warp >> 2, quad = warp & 3;
    const uint32_t tS = tmem_base + (uint32_t)wg * p.slot_stride + ((uint32_t)(quad * 32) << 16);
    const uint32_t tP = tS, tO = tS + p.o_off, tL = tO + HD;
    const int row0 = quad * 32;
    const bool active = row0 < P;
    uint8_t* stg = sStage + (size_t)warp * kStageBytes;
    for (int n = wg; n < total; n += 2) {
      const int ph = (n >> 1) & 1;
      const int item = first + n * step, rho = item / p.heads, h = item - rho * p.heads;
      if (tid == 0) stamp(p, n >> 1, 0);
      mbar_wait_c(&bars[kSFull + wg], ph);
      tc_fence_after();
      if (tid == 0) stamp(p, n >> 1, 1);
      if (active) {
        // key columns >= P are tile padding (K rows zero-filled by TMA): overwrite their scores with -inf once,
        // so that neither pass needs a mask (exp2(-inf - max) = 0)
        if (P16 != P) {
          for (int c = P; c < P16; ++c) tmem_st_x1(tS + c, 0xff800000u);
          tmem_st_wait();
        }
        uint32_t ra[32], rb[32];
        const int G = P16 >> 5, rem16 = P16 & 16;
        // ---- pass 1: row maximum; the load of the next 32 columns is in flight while this group is folded
        float mx = -INFINITY;
        if (G > 0) tmem_ld_x32(tS, rb);
#pragma unroll 1
        for (int gi = 0; gi < G; ++gi) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) ra[j] = rb[j];
          if (gi + 1 < G) tmem_ld_x32(tS + 32 * (gi + 1), rb);
          else if (rem16) tmem_ld_x16(tS + 32 * G, rb);
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(ra[j]));
        }
        if (rem16) {
          if (G == 0) tmem_ld_x16(tS, rb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(rb[j]));
        }
        if (tid == 0) stamp(p, n >> 1, 2);
        // ---- pass 2: p = exp2(s - max) -> packed f16 over the dead S columns (row sums: the tensor core)
        if (G > 0) tmem_ld_x32(tS, rb);
#pragma unroll 1
        for (int gi = 0; gi < G; ++gi) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) ra[j] = rb[j];
          if (gi + 1 < G) tmem_ld_x32(tS + 32 * (gi + 1), rb);
          else if (rem16) tmem_ld_x16(tS + 32 * G, rb);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pk[j] = pack_h2(fast_exp2(__uint_as_float(ra[2 * j]) - mx), fast_exp2(__uint_as_float(ra[2 * j + 1]) - mx));
          tmem_st_x16(tP + 16 * gi, pk);
        }
        if (rem16) {
          if (G == 0) tmem_ld_x16(tS, rb);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pk[j] = pack_h2(fast_exp2(__uint_as_float(rb[2 * j]) - mx), fast_exp2(__uint_as_float(rb[2 * j + 1]) - mx));
          tmem_st_x8(tP + 16 * G, pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[kPReady + wg]);
      if (tid == 0) stamp(p, n >> 1, 3);
      mbar_wait_c(&bars[kOFull + wg], ph);
      tc_fence_after();
      if (tid == 0) stamp(p, n >> 1, 4);
      if (active) {
        if (lane == 0) tma_store_wait_read<0>();  // the previous store out of this tile has read it
        __syncwarp();
        uint32_t ro[32];
        tmem_ld_x32(tO, ro);
        const float l = tmem_ld_x1(tL);
        tmem_ld_wait();
        const float inv = 1.f / l;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half) {
            tmem_ld_x32(tO + 32, ro);
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v = make_uint4(pack_h2(__uint_as_float(ro[8 * j]) * inv, __uint_as_float(ro[8 * j + 1]) * inv),
                                 pack_h2(__uint_as_float(ro[8 * j + 2]) * inv, __uint_as_float(ro[8 * j + 3]) * inv),
                                 pack_h2(__uint_as_float(ro[8 * j + 4]) * inv, __uint_as_float(ro[8 * j + 5]) * inv),
                                 pack_h2(__uint_as_float(ro[8 * j + 6]) * inv, __uint_as_float(ro[8 * j + 7]) * inv));
            *reinterpret_cast<uint4*>(stg + swz(lane, 4 * half + j)) = v;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {  // rows >= P of the box are outside the [R][P][D] view: clipped by the TMA unit
          tma_store_3d(&tmO, stg, h * HD, row0, rho);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[kODone + wg]);
      if (tid == 0) stamp(p, n >> 1, 5);
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
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

// [R][P][ld] f16 view of a row-major [R*P, ld] tensor, boxes of (64 columns, rows, 1 region), SWIZZLE_128B.
// Loads: rows outside [0, P) read as 0.  Stores: rows outside are clipped.
bool make_map3(CUtensorMap* tm, const __half* base, int ld, int P, int R, int rows) {
  EncodeTiledFn fn = encode_fn2();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)P, (cuuint64_t)R};
  cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(__half), (cuuint64_t)P * ld * sizeof(__half)};
  cuuint32_t box[3] = {(cuuint32_t)HD, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// host-side cache of the encoded maps (cuTensorMapEncodeTiled costs ~1 us each; workspaces are reused)
struct MapKey3 { const void* base; int ld, P, R, rows; };
struct MapSlot3 { MapKey3 k; CUtensorMap m; bool valid; };
bool cached_map3(CUtensorMap* tm, const __half* base, int ld, int P, int R, int rows) {
  static thread_local MapSlot3 slots[96];
  static thread_local int next = 0;
  for (auto& s : slots)
    if (s.valid && s.k.base == base && s.k.ld == ld && s.k.P == P && s.k.R == R && s.k.rows == rows) {
      *tm = s.m;
      return true;
    }
  if (!make_map3(tm, base, ld, P, R, rows)) return false;
  MapSlot3& s = slots[next];
  next = (next + 1) % 96;
  s.k = MapKey3{base, ld, P, R, rows};
  s.m = *tm;
  s.valid = true;
  return true;
}

uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

struct Geometry {
  int P16, pad, nkc, q_rows, ntail, kv_stages, q_stages, sep_o;
  uint32_t o_off, slot_stride, tmem_cols, qp_bytes;
  size_t smem;
};
bool geometry(const Grid& grid, bool epeg, int epeg_k, Geometry* g) {
  const int P = grid.P;
  g->P16 = (P + 15) / 16 * 16;
  const int W = g->P16 / 16;
  g->pad = epeg ? epeg_k / 2 : 0;
  g->nkc = epeg ? (16 + epeg_k - 1 + 15) / 16 : 1;
  g->q_rows = 16 * (W - 1) + 16 * g->nkc;
  if (g->q_rows < 16 * W + 2 * g->pad) g->q_rows = (16 * W + 2 * g->pad + 7) / 8 * 8;
  g->ntail = P > 128 ? (P - 128 + 15) / 16 : 0;
  // TMEM slot of one warpgroup: [ S: P16 | P over S[0, P16/2) | O (64) and l (16) ]; O | l behind S when two such
  // slots fit the 512 columns (P16 <= 160), else over the tail of S
  const uint32_t p16 = (uint32_t)g->P16;
  const uint32_t sep_stride = round_up(round_up(p16, 32) + 80, 32);
  g->sep_o = 2 * sep_stride <= 512 ? 1 : 0;
  if (g->sep_o) {
    g->o_off = round_up(p16, 32);
    g->slot_stride = sep_stride;
  } else {
    const uint32_t o_min = p16 / 2 > (p16 > 80 ? p16 - 80 : 0) ? p16 / 2 : p16 - 80;
    g->o_off = round_up(o_min, 32);
    g->slot_stride = round_up(g->o_off + 80 > p16 ? g->o_off + 80 : p16, 32);
  }
  uint32_t cols = 32;
  while (cols < 2 * g->slot_stride) cols *= 2;
  g->tmem_cols = cols;
  g->qp_bytes = (uint32_t)(g->P16 > 128 ? g->P16 : 128) * 128u;
  if (g->q_rows > 256 || g->P16 > 256 || cols > 512) return false;
  const size_t fixed = 1024 + 2 * (size_t)g->qp_bytes + 8 * kStageBytes + 2048 + 8 * 64 * 4 + kNumBars * 8 + 128;
  const int tries[3][2] = {{3, 2}, {2, 2}, {2, 1}};  // (kv_stages, q_stages), deepest first
  for (auto& tr : tries) {
    g->kv_stages = tr[0];
    g->q_stages = tr[1];
    g->smem = fixed + (size_t)g->q_stages * g->q_rows * 128 + 2 * (size_t)g->kv_stages * g->P16 * 128;
    if (g->smem <= 227 * 1024) return true;
  }
  return false;
}
}  // namespace

// rrt_debug_set_attention_kernel / RRT_ATTN: 1 = auto (default): this kernel for regions of more than 128 tokens
// (N = 9000: 21.0 vs 23.2 us, N = 50000 / region_num 16: 105 vs 156 us), the mma.sync kernel for smaller regions
// (N = 512: 8.6 vs 14.8 us -- per-item pipeline overhead with little tensor work to hide it); 2 = this kernel
// wherever it is supported; 0 = mma.sync only
int g_attn_tc05 = 1;
// Persistent-grid cap of this kernel while several bags are in flight (api.cu sets it per call, per host thread).
// With 8 bags in flight 64 CTAs x 8 items beat 128 x 4: the start-up of a CTA (first loads + first EPEG, ~6 k cycles
// before the first S) is paid once per CTA, and the SMs left over run the other bags' kernels meanwhile.
// Measured, us per bag, 8 lanes: no cap 61.9, 103 SMs 61.6, 86 60.7, 74 60.7, 64 59.9, 52 59.9.
static thread_local int g_attn_sm_cap = 0;
void set_attn_sm_cap(int n) { g_attn_sm_cap = n; }

bool rmsa_attention_tc05_supported(const Grid& grid, int D, int heads, int epeg_k) {
  if (!(heads > 0 && D % heads == 0 && D / heads == HD && grid.P >= 1 && grid.P <= 256 && D % 8 == 0)) return false;
  if (epeg_k > 63) return false;
  Geometry g;
  return geometry(grid, epeg_k > 0, epeg_k, &g);
}

cudaError_t launch_rmsa_attention_tc05(const __half* qkv, const float* taps, __half* o,
                                       const Grid& grid, int D, int heads, int epeg_k,
                                       cudaStream_t stream) {
  const bool epeg = taps != nullptr;
  if (!rmsa_attention_tc05_supported(grid, D, heads, epeg ? epeg_k : 0)) return cudaErrorInvalidValue;
  Geometry g;
  geometry(grid, epeg, epeg_k, &g);
  AttnTcParams p;
  p.taps = taps; p.o = o; p.trace = g_attn_trace;
  p.P = grid.P; p.P16 = g.P16; p.heads = heads; p.D = D;
  p.epeg_k = epeg ? epeg_k : 0; p.pad = g.pad; p.nkc = g.nkc; p.q_rows = g.q_rows; p.ntail = g.ntail;
  p.n_items = grid.R * heads;
  p.kv_stages = g.kv_stages; p.q_stages = g.q_stages;
  p.sep_o = g.sep_o;
  p.tail_helpers = g.ntail <= 2 ? 2 : kHelpers;
  p.epeg_helpers = g.ntail <= 2 ? kHelpers - 2 : kHelpers;
  p.tail_part = g.ntail < p.tail_helpers ? g.ntail : p.tail_helpers;
  p.slot_stride = g.slot_stride; p.o_off = g.o_off; p.tmem_cols = g.tmem_cols;
  p.q_bytes = (uint32_t)g.q_rows * 128u; p.kv_bytes = (uint32_t)g.P16 * 128u; p.qp_bytes = g.qp_bytes;
  p.qscale = 1.4426950408889634f / sqrtf((float)HD);
  CUtensorMap tmQ, tmKV, tmO;
  if (!cached_map3(&tmQ, qkv, 3 * D, grid.P, grid.R, g.q_rows) ||
      !cached_map3(&tmKV, qkv, 3 * D, grid.P, grid.R, g.P16) || !cached_map3(&tmO, o, D, grid.P, grid.R, 32))
    return cudaErrorUnknown;
  static DeviceOnce configured;
  static int sms[32];
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rmsa_attn_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(rmsa_attn_tc05_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms[dev & 31], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
  }
  // persistent grid: every CTA walks the same number of items (+-1); the smallest grid that keeps the
  // number of rounds of a full-chip grid leaves the remaining SMs to whatever else is in flight
  int nsm = sms[dev & 31] > 0 ? sms[dev & 31] : 148;
  // SM cap: RRT_ATTN_SMS, else what the batch entry point set for the number of bags in flight (set_attn_sm_cap)
  static const int env_cap = [] { const char* e = getenv("RRT_ATTN_SMS"); return e ? atoi(e) : -1; }();
  const int cap = env_cap >= 0 ? env_cap : g_attn_sm_cap;
  if (cap > 0 && cap < nsm) nsm = cap;
  const int rounds = (p.n_items + nsm - 1) / nsm;
  const int ctas = (p.n_items + rounds - 1) / rounds;
  return launch_chain_kernel(rmsa_attn_tc05_kernel, dim3(ctas), dim3(kThreads), g.smem, stream, tmQ, tmKV, tmO, p);
}

}  // namespace rrt
