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
#include <cstdlib>
#include "backward.cuh"
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
//   BN = 128: bag-sized GEMMs with a narrow output (proj: N = D): at 256 columns every CTA would own
//             ONE tile and run prologue, main loop and epilogue back to back; two 128-column tiles
//             per CTA let the second main loop hide the first (HBM-bound, residual) epilogue
//   BN =  64: landmark GEMMs (M = k*64 rows): 4x more CTAs, 4x shorter MMA chain per tile
//   RESID (the residual-scatter epilogues): the epilogue, not the main loop, bounds those GEMMs -- 256 KB of fp32
//             residual in / x1 out per 128 x 256 tile against 5.7 k cycles of MMAs -- so part of the operand ring
//             is given to a per-warp ring of kResidRing residual chunks filled by cp.async up to a tile ahead
constexpr int kResidRing = 3;                  // 32 x 32 fp32 chunks in flight per epilogue warp
constexpr int kResidChunkBytes = 32 * 32 * 4;
template <int BN_, bool RESID_ = false>
struct TileCfg {
  static constexpr int STAGES =
      RESID_ ? (BN_ == 256 ? 2 : (BN_ == 128 ? 3 : 4)) : (BN_ == 256 ? 4 : (BN_ == 128 ? 6 : 8));
  static constexpr int B_BYTES = BN_ * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RESID_BYTES = RESID_ ? kResidRing * kEpiWarps * kResidChunkBytes : 0;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + RESID_BYTES + BAR_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "dynamic shared memory budget exceeded");
  static constexpr int TMEM_COLS = 2 * BN_ < 32 ? 32 : 2 * BN_;
};

struct Tc05Params {
  int M, N, K;
  const float* bias;
  void* C;  // OutT[M or L, N]
  const float* resid;
  Grid grid;
  int act;           // GemmActivation, store mode only
  int act2, act_split;  // columns >= act_split (multiple of 32, 0 = none) use act2
  // residual modes, optional: per-row partial LayerNorm / CR-MSA logit sums of the rows this GEMM writes
  // (GemmEpilogue::rs_part); rs_gamma [N], rs_phi [N, rs_k]
  float* rs_part;
  const float* rs_gamma;
  const float* rs_phi;
  int rs_k;
  Dropout drop;      // kEpiResidualUnpartDrop only
  int ksplit;        // >= 1: the K loop of every tile is cut into ksplit work items (kEpiAtomicAdd)
  // kEpiAtomicAdd, optional: amax word of the scaled fp16 gradient operand; every partial sum is multiplied by
  // grad_inv_scale (a power of two: exact) before it is added, so the result needs no separate unscale pass
  const uint32_t* unscale_amax;
  long long* trace;  // debug: per-CTA clock64 stamps (tools/gemm_trace.py), null in production
};

// internal compile-time mode: kEpiStore with a run-time activation.  Kept apart from kEpiStore so that
// the bag-sized QKV / landmark GEMMs do not carry the erff/tanhf code in their epilogue (measured:
// 20.9 -> 31.5 us for the QKV GEMM when the switch sat in the common kernel).
constexpr int kEpiStoreAct = 3;
__host__ __device__ constexpr bool is_store_mode(int mode) { return mode == kEpiStore || mode == kEpiStoreAct; }
__host__ __device__ constexpr bool is_resid_mode(int mode) {
  return mode == kEpiResidualUnpart || mode == kEpiResidualUnpartDrop;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case kActRelu: return fmaxf(v, 0.f);
    case kActGelu: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));  // nn.GELU() (erf form)
    case kActTanh: return tanhf(v);
    case kActSigmoid: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

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
// split-K partial sums: one 16-byte reduction per 4 outputs (no return value: fire and forget)
__device__ __forceinline__ void red_add4(float* base, size_t off, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + off), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add4(__half*, size_t, float4) {}

// ---- residual-scatter epilogue (kEpiResidualUnpart / kEpiResidualUnpartDrop) of the single-CTA kernel ------------
// Residual prefetch cursor of one epilogue warp: walks the (tile, chunk) sequence the warp will consume,
// kResidRing chunks ahead of the consumer, and cp.asyncs each 32 x 32 fp32 residual chunk into the next slot of the
// warp's ring.  Every 16-byte piece is written and later read by the SAME thread (row group i -> row 4i + lane / 8,
// columns 4 * (lane % 8)): no cross-thread synchronisation, only cp.async.wait_group.  One commit group per call
// (empty past the last chunk), so that wait_group<kResidRing - 1> always means "the chunk being consumed has landed".
// Region slot -> token row for the 8 row groups of a lane (slots first, first + 4, ..., first + 28; -1 = pad slot or
// outside the bag).  The three integer divisions of Grid::slot_to_token are paid ONCE (out of line: inlined at every
// call site they were 3.4 k of the kernel's SASS); the other seven rows follow by stepping (region, row, column)
// with carries.  With eight full conversions per tile in the cursor and eight more at the top of the epilogue a tile
// cost ~8 k cycles of divisions on the epilogue warps' critical path (tools/gemm_trace.py under the SM cap).
struct SlotPos { int rr, rc, pr, pc; };
__device__ __noinline__ SlotPos resid_slot_pos(int slot, int P, int g, int rs) {
  SlotPos s;
  const int rho = slot / P, p = slot - rho * P;
  s.rr = rho / g; s.rc = rho - s.rr * g;
  s.pr = p / rs;  s.pc = p - s.pr * rs;
  return s;
}
__device__ __forceinline__ void resid_tokens8(const Grid& g, int first, int M, int (&tok)[8]) {
  SlotPos s = resid_slot_pos(first, g.P, g.g, g.rs);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = (s.rr * g.rs + s.pr) * g.H + s.rc * g.rs + s.pc;
    tok[i] = (first + 4 * i < M && t < g.L) ? t : -1;
    s.pc += 4;
    while (s.pc >= g.rs) { s.pc -= g.rs; ++s.pr; }
    while (s.pr >= g.rs) {
      s.pr -= g.rs;
      if (++s.rc == g.g) { s.rc = 0; ++s.rr; }
    }
  }
}
// single slot (the row-statistics record of a thread's row)
__device__ __forceinline__ int resid_token_of_slot(const Grid& g, int slot, int M) {
  if (slot >= M) return -1;
  const SlotPos s = resid_slot_pos(slot, g.P, g.g, g.rs);
  const int t = (s.rr * g.rs + s.pr) * g.H + s.rc * g.rs + s.pc;
  return t < g.L ? t : -1;
}

template <int BN>
struct ResidCursor {
  static constexpr int NC = (BN / 32) / 2;
  uint8_t* ring;     // this warp's kResidRing chunks
  int wt, step, num_work, tiles_nc, ci, cj, CM, CN;
  int j, slot;       // chunk within the cursor's tile; ring slot of the next chunk
  int tok[8];        // token rows of the cursor's tile for this lane's 8 row groups (-1: pad / outside)
  int n0;
  __device__ __forceinline__ void load_tile(const Tc05Params& p, int quad, int lane) {
    if (wt >= num_work) return;
    const int m0 = ((wt / tiles_nc) * CM + ci) * BM;
    n0 = ((wt % tiles_nc) * CN + cj) * BN;
    resid_tokens8(p.grid, m0 + quad * 32 + (lane >> 3), p.M, tok);
  }
  __device__ __forceinline__ void issue(const Tc05Params& p, int ew, int lane) {
    if (wt < num_work) {
      const int gc = n0 + ((ew >> 2) * NC + j) * 32 + (lane & 7) * 4;
      uint8_t* dst = ring + slot * kResidChunkBytes + lane * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool ok = tok[i] >= 0 && gc < p.N;
        cp_async16(dst + i * 512, ok ? p.resid + (size_t)tok[i] * p.N + gc : p.resid, ok);
      }
      if (++slot == kResidRing) slot = 0;
      if (++j == NC) {
        j = 0;
        wt += step;
        load_tile(p, ew & 3, lane);
      }
    }
    cp_async_commit();
  }
};

// One 128 x BN accumulator for one of the 8 epilogue warps.  The chunk loop is ROLLED and the row-statistics pass
// reads gamma (.) phi from shared memory: the unrolled form of this epilogue was 10 k SASS instructions (157 KB; with
// dropout 188 KB) and ran out of the instruction cache -- ncu's warp-state samples of the 8 epilogue warps were
// dominated by no_inst (profiles/r02f_*): 28 k cycles per tile against 5.7 k for the main loop.
template <int MODE, int BN, typename ReleaseFn>
__device__ __forceinline__ void epilogue_tile_resid(const Tc05Params& p, uint32_t tmem_acc, uint64_t* tfull_bar,
                                                    uint32_t tfull_parity, int m0, int n0, int ew, int lane,
                                                    float* scratch, bool first, ReleaseFn release,
                                                    ResidCursor<BN>& cur, int& cons_slot) {
  constexpr int NC = (BN / 32) / 2;  // chunks per epilogue warp
  const int quad = ew & 3;           // TMEM lanes [32*quad, 32*quad+32) are readable by this warp
  const int c_begin = (ew >> 2) * NC;
  float* const out = reinterpret_cast<float*>(p.C);
  const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
  // output row of each of the 8 row groups this lane stores (region slot -> token; -1: pad row)
  int orow[8];
  resid_tokens8(p.grid, m0 + quad * 32 + sub_r, p.M, orow);
  // row statistics of the rows being written (thread = row `lane` of this warp's 32): sum, sum of squares and up
  // to four dot products with gamma (.) phi[:, n] over this warp's NC * 32 columns
  float rs_sum = 0.f, rs_sq = 0.f, rs_dot[4] = {0.f, 0.f, 0.f, 0.f};
  const bool row_stats = p.rs_part != nullptr;
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
  if (first && threadIdx.x == 128) stamp(p, 6);
  const uint32_t t_addr = tmem_acc + ((uint32_t)(quad * 32) << 16) + c_begin * 32;
  uint32_t r[32];
  tmem_ld_32x32(t_addr, r);
#pragma unroll 1
  for (int j = 0; j < NC; ++j) {
    const int gc = n0 + (c_begin + j) * 32 + sub_c;
    const bool col_ok = gc < p.N;
    // bias of this lane's 4 columns and gamma_c * phi[c, 0..3] of column c = `lane` of this chunk: fetched here, used
    // after the transpose resp. at the end of the chunk
    const float4 bv = (p.bias && col_ok) ? __ldg(reinterpret_cast<const float4*>(p.bias + gc))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    float gm = 0.f;   // (raw loads here, the product where the table is written: no scoreboard wait at the top)
    float4 ph = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_stats) {
      const int gcl = n0 + (c_begin + j) * 32 + lane;
      gm = __ldg(p.rs_gamma + gcl);
      if (p.rs_k == 4) {
        ph = __ldg(reinterpret_cast<const float4*>(p.rs_phi) + gcl);
      } else {
        const float* pp = p.rs_phi + (size_t)gcl * p.rs_k;
        ph.x = __ldg(pp);
        if (p.rs_k > 1) ph.y = __ldg(pp + 1);
        if (p.rs_k > 2) ph.z = __ldg(pp + 2);
      }
    }
    tmem_ld_wait();  // chunk j is in registers
    if (first && threadIdx.x == 128 && j < 2) stamp(p, 10 + 3 * j);
#pragma unroll
    for (int q = 0; q < 8; ++q)  // row = lane; 16-byte slot q lands at slot q ^ (row & 7)
      *reinterpret_cast<float4*>(scratch + lane * EPI_LD + 4 * (q ^ (lane & 7))) =
          make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                      __uint_as_float(r[4 * q + 3]));
    if (j + 1 < NC) {
      tmem_ld_32x32(t_addr + (j + 1) * 32, r);  // overlaps the rest of chunk j
    } else {
      // the whole accumulator slice of this warp has left TMEM: release it to the MMA warp
      tc_fence_before();
      __syncwarp();
      release();
    }
    __syncwarp();
    if (first && threadIdx.x == 128 && j < 2) stamp(p, 11 + 3 * j);
    // this thread's 8 pieces of the residual chunk have landed in the ring
    cp_async_wait<kResidRing - 1>();
    uint8_t* slot = cur.ring + cons_slot * kResidChunkBytes;
    if (++cons_slot == kResidRing) cons_slot = 0;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(slot + i * 512 + lane * 16);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = 4 * i + sub_r;
      const float4 a = *reinterpret_cast<const float4*>(scratch + rl * EPI_LD + 4 * ((lane & 7) ^ (rl & 7)));
      if (MODE == kEpiResidualUnpartDrop) {  // x1 = x + dropout(o Wp^T + b)   (modules/rmsa.py:131-132)
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        if (orow[i] >= 0) m = dropout_scale4(p.drop, (unsigned long long)orow[i] * p.N + gc);
        v[i].x = fmaf(a.x + bv.x, m.x, v[i].x); v[i].y = fmaf(a.y + bv.y, m.y, v[i].y);
        v[i].z = fmaf(a.z + bv.z, m.z, v[i].z); v[i].w = fmaf(a.w + bv.w, m.w, v[i].w);
      } else {
        v[i].x += a.x + bv.x; v[i].y += a.y + bv.y; v[i].z += a.z + bv.z; v[i].w += a.w + bv.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (orow[i] >= 0 && col_ok) store_out4(out, (size_t)orow[i] * p.N + gc, v[i]);
    if (first && threadIdx.x == 128 && j < 2) stamp(p, 12 + 3 * j);
    if (row_stats) {
      // hand the finished values back through the scratch tile (same swizzled slots they were read from); the
      // 32 x 4 table of gamma (.) phi goes into the ring slot just consumed (each lane overwrites one of ITS pieces)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + sub_r;
        *reinterpret_cast<float4*>(scratch + rl * EPI_LD + 4 * ((lane & 7) ^ (rl & 7))) = v[i];
      }
      *reinterpret_cast<float4*>(slot + lane * 16) = make_float4(gm * ph.x, gm * ph.y, gm * ph.z, gm * ph.w);
      __syncwarp();
#pragma unroll 2
      for (int q = 0; q < 8; ++q) {
        const float4 xv = *reinterpret_cast<const float4*>(scratch + lane * EPI_LD + 4 * (q ^ (lane & 7)));
        rs_sum += (xv.x + xv.y) + (xv.z + xv.w);
        rs_sq = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, fmaf(xv.z, xv.z, fmaf(xv.w, xv.w, rs_sq))));
        const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 g = *reinterpret_cast<const float4*>(slot + (4 * q + e) * 16);  // broadcast read
          rs_dot[0] = fmaf(xe[e], g.x, rs_dot[0]);
          rs_dot[1] = fmaf(xe[e], g.y, rs_dot[1]);
          rs_dot[2] = fmaf(xe[e], g.z, rs_dot[2]);
          rs_dot[3] = fmaf(xe[e], g.w, rs_dot[3]);
        }
      }
    }
    __syncwarp();  // every lane is done with the scratch tile and the borrowed ring slot
    cur.issue(p, ew, lane);  // the slot just consumed takes the chunk kResidRing ahead
  }
  if (row_stats) {
    // one 32-byte record per (token, 128-column part): [sum, sum sq, dot_0..3, -, -]
    const int tok = resid_token_of_slot(p.grid, m0 + quad * 32 + lane, p.M);
    if (tok >= 0) {
      constexpr int PW = NC * 32;  // columns per part
      const int parts = p.N / PW, part = (n0 + c_begin * 32) / PW;
      float4* rec = reinterpret_cast<float4*>(p.rs_part + ((size_t)tok * parts + part) * 8);
      rec[0] = make_float4(rs_sum, rs_sq, rs_dot[0], rs_dot[1]);
      rec[1] = make_float4(rs_dot[2], rs_dot[3], 0.f, 0.f);
    }
  }
}

// ---- TMA-store epilogue with a run-time activation (kEpiStoreAct: patch_to_emb, the pooling head's score layers) -----
// Same data path as the kEpiStore branch of epilogue_tile (thread = row, +bias, activation, swizzled staging tile,
// TMA store), but with the chunk loop ROLLED: unrolled, the four activation bodies x 32 elements x NC chunks made this
// kernel 146 KB of SASS, and its epilogue warps starved on instruction fetch like the residual epilogue's did.
template <int BN, typename OutT, typename ReleaseFn>
__device__ __forceinline__ void epilogue_tile_store_act(const Tc05Params& p, const CUtensorMap* tmC, uint32_t tmem_acc,
                                                        uint64_t* tfull_bar, uint32_t tfull_parity, int m0, int n0,
                                                        int ew, int lane, float* scratch, ReleaseFn release) {
  constexpr int NC = (BN / 32) / 2;
  constexpr int ROWB = 32 * (int)sizeof(OutT);      // 64 B (f16) or 128 B (fp32) per row
  constexpr int NCH = ROWB / 16;                     // 16-byte chunks per row
  constexpr int NBUF = (2 * 32 * ROWB <= 32 * EPI_LD * 4) ? 2 : 1;  // staging tiles in this warp's 4 KB scratch
  constexpr int EPC = 16 / (int)sizeof(OutT);       // elements per 16-byte chunk
  const int quad = ew & 3, c_begin = (ew >> 2) * NC;
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
  const uint32_t t_addr = tmem_acc + ((uint32_t)(quad * 32) << 16) + c_begin * 32;
  uint32_t r[32];
  tmem_ld_32x32(t_addr, r);
  int buf = 0;
#pragma unroll 1
  for (int j = 0; j < NC; ++j) {
    const int gcl = n0 + (c_begin + j) * 32 + lane;   // lane l keeps the bias of column l, broadcast by shuffle
    const float bias_l = (p.bias && gcl < p.N) ? __ldg(p.bias + gcl) : 0.f;
    tmem_ld_wait();
    float fv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) fv[i] = __uint_as_float(r[i]) + __shfl_sync(0xffffffffu, bias_l, i);
    if (j + 1 < NC) {
      tmem_ld_32x32(t_addr + (j + 1) * 32, r);  // overlaps the activation and the store of chunk j
    } else {
      tc_fence_before();
      __syncwarp();
      release();
    }
    const int act_j = (p.act_split > 0 && n0 + (c_begin + j) * 32 >= p.act_split) ? p.act2 : p.act;
    if (act_j == kActRelu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) fv[i] = fmaxf(fv[i], 0.f);
    } else if (act_j == kActGelu) {  // (full unrolls: fv stays in registers; only the taken branch is fetched)
#pragma unroll
      for (int i = 0; i < 32; ++i) fv[i] = gelu_fwd(fv[i]);
    } else if (act_j == kActTanh) {
#pragma unroll
      for (int i = 0; i < 32; ++i) fv[i] = tanhf(fv[i]);
    } else if (act_j == kActSigmoid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) fv[i] = 1.f / (1.f + __expf(-fv[i]));
    }
    uint8_t* stage = reinterpret_cast<uint8_t*>(scratch) + buf * (32 * ROWB);
    if (j >= NBUF) {  // the store that last read this buffer must have drained it
      if (lane == 0) tma_store_wait_read<NBUF - 1>();
      __syncwarp();
    }
    uint8_t* rowp = stage + lane * ROWB;
    const int sw = ROWB == 128 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
    for (int q = 0; q < NCH; ++q) {
      const float* f = fv + q * EPC;
      uint4 pk;
      if (sizeof(OutT) == 2) {
        pk = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4 % EPC], f[5 % EPC]),
                        pack_h2(f[6 % EPC], f[7 % EPC]));
      } else {
        pk = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
      }
      *reinterpret_cast<uint4*>(rowp + 16 * (q ^ sw)) = pk;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, stage, n0 + (c_begin + j) * 32, m0 + quad * 32);
      tma_store_commit();
    }
    if (++buf == NBUF) buf = 0;
  }
  if (lane == 0) tma_store_wait_read<0>();  // both staging buffers are free again before the next tile reuses them
  __syncwarp();
}

// Epilogue of one 128 x BN accumulator for one of the 8 epilogue warps (two per TMEM lane quadrant):
// software-pipelined tcgen05.ld, accumulator released (release()) as soon as this warp's slice is in
// registers, then either TMA tile stores (kEpiStore) or transpose + st.global (tanh, split-K reduction).  The
// residual-scatter modes have their own epilogue (epilogue_tile_resid) and the activation-store mode too.
template <int MODE, int BN, typename OutT, typename ReleaseFn>
__device__ __forceinline__ void epilogue_tile(const Tc05Params& p, const CUtensorMap* tmC,
                                              uint32_t tmem_acc, uint64_t* tfull_bar,
                                              uint32_t tfull_parity, int m0, int n0, int ew, int lane,
                                              float* scratch, bool first, ReleaseFn release) {
  static_assert(!is_resid_mode(MODE), "residual modes: epilogue_tile_resid");
  constexpr bool kTmaStore = is_store_mode(MODE);
  constexpr int NC = (BN / 32) / 2;  // chunks per epilogue warp
  const int quad = ew & 3;           // TMEM lanes [32*quad, 32*quad+32) are readable by this warp
  const int c_begin = (ew >> 2) * NC;
  OutT* const out = reinterpret_cast<OutT*>(p.C);
  const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    // output row of each of the 8 row groups this lane stores (-1: outside the matrix)
    int orow[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gr = m0 + quad * 32 + 4 * i + sub_r;
      orow[i] = gr < p.M ? gr : -1;
    }
    // TMA-store mode: thread = row needs the bias of all 32 columns of a chunk; lane l keeps
    // column l of every chunk and the value is broadcast with a shuffle when used
    float bias_l[NC];
    if (kTmaStore) {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int gc = n0 + (c_begin + j) * 32 + lane;
        bias_l[j] = (p.bias && gc < p.N) ? __ldg(p.bias + gc) : 0.f;
      }
    }
    // bias of this warp's chunks: fetched while the MMA warp is still producing the accumulator
    float4 bias_r[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int gc = n0 + (c_begin + j) * 32 + sub_c;
      bias_r[j] = (p.bias && gc < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + gc))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 v[8];
    mbar_wait(tfull_bar, tfull_parity);
    tc_fence_after();
    if (first && threadIdx.x == 128) stamp(p, 6);
    const uint32_t t_addr = tmem_acc + ((uint32_t)(quad * 32) << 16) + c_begin * 32;
    uint32_t r[2][32];
    tmem_ld_32x32(t_addr, r[0]);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      tmem_ld_wait();  // chunk j is in registers
      if (first && threadIdx.x == 128 && j < 2) stamp(p, 10 + 3 * j);
      if (j + 1 < NC) {
        tmem_ld_32x32(t_addr + (j + 1) * 32, r[(j + 1) & 1]);  // overlaps the stores of chunk j
      } else {
        // the whole accumulator slice of this warp has left TMEM: release it to the MMA warp
        tc_fence_before();
        __syncwarp();
        release();
      }
      const uint32_t* rr = r[j & 1];
      if (kTmaStore) {
        // thread = output row; +bias, convert, write the 32-column row segment into this warp's
        // staging buffer in the TMA swizzle of its row pitch, then one lane issues the tile store
        constexpr int ROWB = 32 * (int)sizeof(OutT);      // 64 B (f16) or 128 B (fp32) per row
        constexpr int NCH = ROWB / 16;                     // 16-byte chunks per row
        // the 4 KB scratch of this warp holds two f16 staging tiles or one fp32 tile
        constexpr int NBUF = (2 * 32 * ROWB <= 32 * EPI_LD * 4) ? 2 : 1;
        uint8_t* stage = reinterpret_cast<uint8_t*>(scratch) + (j % NBUF) * (32 * ROWB);
        if (j >= NBUF) {  // the store that last read this buffer must have drained it
          if (lane == 0) tma_store_wait_read<NBUF - 1>();
          __syncwarp();
        }
        uint8_t* rowp = stage + lane * ROWB;
        // swizzle: 16-B chunk index ^= bits of the row (128B pattern: row%8; 64B pattern: (row/2)%4)
        const int sw = ROWB == 128 ? (lane & 7) : ((lane >> 1) & 3);
        float fv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) fv[i] = __uint_as_float(rr[i]) + __shfl_sync(0xffffffffu, bias_l[j], i);
        if (MODE == kEpiStoreAct) {
          // ONE branch per 32-column chunk (a per-element switch kept erff / tanhf on every element's path:
          // 70 us for the 9000 x 1024 x 512 patch_to_emb GEMM with a ReLU)
          const int act_j = (p.act_split > 0 && n0 + (c_begin + j) * 32 >= p.act_split) ? p.act2 : p.act;
          if (act_j == kActRelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) fv[i] = fmaxf(fv[i], 0.f);
          } else if (act_j == kActGelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) fv[i] = gelu_fwd(fv[i]);
          } else if (act_j == kActTanh) {
#pragma unroll
            for (int i = 0; i < 32; ++i) fv[i] = tanhf(fv[i]);
          } else if (act_j == kActSigmoid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) fv[i] = 1.f / (1.f + __expf(-fv[i]));
          }
        }
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
          constexpr int EPC = 16 / (int)sizeof(OutT);      // elements per 16-byte chunk
          const float* f = fv + q * EPC;
          uint4 pk;
          if (sizeof(OutT) == 2) {
            pk = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4 % EPC], f[5 % EPC]),
                            pack_h2(f[6 % EPC], f[7 % EPC]));
          } else {
            pk = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                            __float_as_uint(f[3]));
          }
          *reinterpret_cast<uint4*>(rowp + 16 * (q ^ sw)) = pk;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmC, stage, n0 + (c_begin + j) * 32, m0 + quad * 32);
          tma_store_commit();
        }
        if (first && threadIdx.x == 128 && j < 2) stamp(p, 11 + 3 * j);
      } else {
      const int gc = n0 + (c_begin + j) * 32 + sub_c;
      const bool col_ok = gc < p.N;
      const float4 bv = bias_r[j];
#pragma unroll
      for (int q = 0; q < 8; ++q)  // row = lane; 16-byte slot q lands at slot q ^ (row & 7)
        *reinterpret_cast<float4*>(scratch + lane * EPI_LD + 4 * (q ^ (lane & 7))) =
            make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]),
                        __uint_as_float(rr[4 * q + 2]), __uint_as_float(rr[4 * q + 3]));
      __syncwarp();
      if (first && threadIdx.x == 128 && j < 2) stamp(p, 11 + 3 * j);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + sub_r;
        float4 a = *reinterpret_cast<const float4*>(scratch + rl * EPI_LD +
                                                    4 * ((lane & 7) ^ (rl & 7)));
        v[i] = make_float4(a.x + bv.x, a.y + bv.y, a.z + bv.z, a.w + bv.w);
        if (MODE == kEpiTanh)
          v[i] = make_float4(tanhf(v[i].x), tanhf(v[i].y), tanhf(v[i].z), tanhf(v[i].w));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (orow[i] >= 0 && col_ok) {
          if (MODE == kEpiAtomicAdd) {
            if (p.unscale_amax) {
              const float us = grad_inv_scale(__ldg(p.unscale_amax));
              v[i].x *= us; v[i].y *= us; v[i].z *= us; v[i].w *= us;
            }
            red_add4(out, (size_t)orow[i] * p.N + gc, v[i]);
          }
          else store_out4(out, (size_t)orow[i] * p.N + gc, v[i]);
        }
      if (first && threadIdx.x == 128 && j < 2) stamp(p, 12 + 3 * j);
      __syncwarp();
      }  // !kTmaStore
    }
    if (kTmaStore) {  // both staging buffers are free again before the next tile reuses them
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
    }
}

// CM x CN: thread-block cluster shape.  CTA (ci, cj) of a cluster computes output tile
// (m-block mc*CM+ci, n-tile nc*CN+cj).  The A tile of a cluster row is needed by its CN CTAs and the
// W tile of a cluster column by its CM CTAs: every CTA TMA-loads a 1/CN slice of its A tile and a
// 1/CM slice of its W tile and MULTICASTS them to the CTAs that share them, which divides the L2->SMEM
// operand traffic (the measured bound of the 1x1 kernel) by up to 2 for a 2x2 cluster.
// MNMAJOR: both operands arrive MN-major -- A = a[K, M], W = w[K, N] row-major, i.e. C = a^T @ w with the
// contraction over ROWS (the weight-gradient GEMM: dW[C_out, C_in] = dY[tokens, C_out]^T act[tokens, C_in],
// straight from the row-major activations, no transposed copies).  Tiles are staged as 64 x 64 TMA boxes
// (sm100.cuh::umma_desc_mn_sw128).  MNMAJOR = 2: only W is MN-major -- C = a @ w with w[K, N] row-major (the
// input-gradient GEMM d_in = dY W straight from the forward's fp16 weight, no transposed copy).
template <int MODE, int BN, typename OutT, int CM, int CN, int MNMAJOR = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_f16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, Tc05Params p) {
  using Cfg = TileCfg<BN, is_resid_mode(MODE)>;
  constexpr int CSIZE = CM * CN;
  // plain stores (kEpiStore): the epilogue hands each 32x32 chunk to the TMA engine; the scatter /
  // tanh epilogues keep st.global (rows are permuted resp. the layer is tiny)
  constexpr bool kTmaStore = is_store_mode(MODE);
  constexpr int STAGES = Cfg::STAGES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int TMEM_COLS = Cfg::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  float* sEpi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint8_t* sResid = smem + STAGES * STAGE_BYTES + EPI_BYTES;  // residual modes: kEpiWarps x kResidRing chunks
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES + Cfg::RESID_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) stamp(p, 0);
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmA);
  if (warp == 3 && lane == 0) prefetch_tensormap(&tmB);
  if (kTmaStore && warp == 2 && lane == 0) prefetch_tensormap(&tmC);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);   // the A producer and the B producer each arrive once (+ their bytes)
      mbar_init(&empty[i], CM + CN - 1);  // MMA commits of every CTA that multicasts into this stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  if (CSIZE > 1) cluster_sync_all(); else __syncthreads();  // barriers of every peer are initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(p, 1);
  // PDL (common.cuh): barriers, TMEM and descriptor prefetch above ran while the predecessor was finishing;
  // nothing below may touch its output before it has completed
  pdl_wait();

  // cluster-tile schedule: every CTA of a cluster walks the same list, so that the multicast
  // producers / consumers of the peers stay in lock step stage by stage
  const int crank = CSIZE > 1 ? (int)cluster_ctarank() : 0;
  const int ci = crank % CM, cj = crank / CM;
  const int cluster_id = blockIdx.x / CSIZE, num_clusters = gridDim.x / CSIZE;
  const int tiles_mc = ((p.M + BM - 1) / BM + CM - 1) / CM;
  const int tiles_nc = ((p.N + BN - 1) / BN + CN - 1) / CN;
  const int num_ctiles = tiles_mc * tiles_nc;
  const int KB = (p.K + BK - 1) / BK;  // (K % 64 != 0 only with MNMAJOR: TMA zero-fills the missing rows)
  // split-K (kEpiAtomicAdd): work item wt = (tile wt % num_ctiles, K slice wt / num_ctiles of KS)
  const int KS = MODE == kEpiAtomicAdd ? p.ksplit : 1;
  const int num_work = num_ctiles * KS;
  // peers that share my A tile (same ci) / my W tile (same cj); rank = ci + CM * cj
  uint16_t mask_a = 0, mask_b = 0;
#pragma unroll
  for (int j = 0; j < CN; ++j) mask_a |= (uint16_t)(1u << (ci + CM * j));
#pragma unroll
  for (int i = 0; i < CM; ++i) mask_b |= (uint16_t)(1u << (i + CM * cj));

  if (warp == 0 || warp == 3) {
    if (lane == 0) {  // ===== TMA producers: warp 0 streams A slices, warp 3 streams W slices =====
      const bool is_a = warp == 0;
      int s = 0, ph = 0;
      for (int wt = cluster_id; wt < num_work; wt += num_clusters) {
        const int ct = wt % num_ctiles, ks = wt / num_ctiles;
        const int m0 = ((ct / tiles_nc) * CM + ci) * BM, n0 = ((ct % tiles_nc) * CN + cj) * BN;
        for (int kb = ks * KB / KS, kb1 = (ks + 1) * KB / KS; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);  // every CTA that reads or refills this stage has released it
          if (is_a) {
            mbar_arrive_expect_tx(&full[s], A_BYTES);  // my slice + the peers' slices of MY tile
            uint8_t* dst = sA + s * A_BYTES + cj * (A_BYTES / CN);
            if (MNMAJOR == 1) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)  // 64 (M) x 64 (K rows) boxes, 8 KB each
                tma_load_2d(dst + j * 8192, &tmA, &full[s], m0 + 64 * j, kb * BK);
            } else if (CN > 1) tma_load_2d_mcast(dst, &tmA, &full[s], kb * BK, m0 + cj * (BM / CN), mask_a);
            else tma_load_2d(dst, &tmA, &full[s], kb * BK, m0);
            if (wt == cluster_id && kb == 0) stamp(p, 2);
          } else {
            mbar_arrive_expect_tx(&full[s], B_BYTES);
            uint8_t* dst = sB + s * B_BYTES + ci * (B_BYTES / CM);
            if (MNMAJOR) {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(dst + j * 8192, &tmB, &full[s], n0 + 64 * j, kb * BK);
            } else if (CM > 1) tma_load_2d_mcast(dst, &tmB, &full[s], kb * BK, n0 + ci * (BN / CM), mask_b);
            else tma_load_2d(dst, &tmB, &full[s], kb * BK, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
      if (is_a) stamp(p, 3);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc(kFmtF16, BM, BN) |
                                 (MNMAJOR == 1 ? kIdescMnMajorAB : (MNMAJOR == 2 ? kIdescMnMajorB : 0u));
      int s = 0, ph = 0, acc = 0, aph = 0;
      for (int wt = cluster_id; wt < num_work; wt += num_clusters) {
        const int ks = wt / num_ctiles;
        mbar_wait(&tempty[acc], aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int kb0 = ks * KB / KS;
        for (int kb = kb0, kb1 = (ks + 1) * KB / KS; kb < kb1; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (wt == cluster_id && kb == kb0) stamp(p, 4);
          const uint64_t ad = MNMAJOR == 1 ? umma_desc_mn_sw128(smem_u32(sA + s * A_BYTES))
                                      : umma_desc_k_sw128(smem_u32(sA + s * A_BYTES));
          const uint64_t bd = MNMAJOR ? umma_desc_mn_sw128(smem_u32(sB + s * B_BYTES))
                                      : umma_desc_k_sw128(smem_u32(sB + s * B_BYTES));
          // K-major: 16 fp16 = 32 bytes per MMA along K (+2 in 16-B units); MN-major: 16 rows of 128 B (+128)
          constexpr int kstep_a = MNMAJOR == 1 ? 128 : 2, kstep_b = MNMAJOR ? 128 : 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16(d_tmem, ad + kstep_a * k, bd + kstep_b * k, idesc, ((kb - kb0) | k) != 0);
          // smem stage reusable once these MMAs have read it: tell every CTA that writes into it
          if (CSIZE > 1) umma_commit_mcast(&empty[s], (uint16_t)(mask_a | mask_b));
          else umma_commit(&empty[s]);
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
    float* scratch = sEpi + ew * 32 * EPI_LD;
    int acc = 0, aph = 0;
    ResidCursor<BN> cur;
    int cons_slot = 0;
    if (is_resid_mode(MODE)) {  // the first kResidRing residual chunks are in flight while the first main loop runs
      cur.ring = sResid + (size_t)ew * kResidRing * kResidChunkBytes;
      cur.wt = cluster_id; cur.step = num_clusters; cur.num_work = num_work; cur.tiles_nc = tiles_nc;
      cur.ci = ci; cur.cj = cj; cur.CM = CM; cur.CN = CN;
      cur.j = 0; cur.slot = 0;
      cur.load_tile(p, ew & 3, lane);
#pragma unroll 1
      for (int r = 0; r < kResidRing; ++r) cur.issue(p, ew, lane);
    }
    for (int wt = cluster_id; wt < num_work; wt += num_clusters) {
      const int ct = wt % num_ctiles;
      const int m0 = ((ct / tiles_nc) * CM + ci) * BM, n0 = ((ct % tiles_nc) * CN + cj) * BN;
      uint64_t* te = &tempty[acc];
      if constexpr (is_resid_mode(MODE))
        epilogue_tile_resid<MODE, BN>(p, tmem_base + acc * BN, &tfull[acc], aph, m0, n0, ew, lane, scratch,
                                      wt == cluster_id, [&] { if (lane == 0) mbar_arrive(te); }, cur, cons_slot);
      else if constexpr (MODE == kEpiStoreAct)
        epilogue_tile_store_act<BN, OutT>(p, &tmC, tmem_base + acc * BN, &tfull[acc], aph, m0, n0, ew, lane, scratch,
                                          [&] { if (lane == 0) mbar_arrive(te); });
      else
        epilogue_tile<MODE, BN, OutT>(p, &tmC, tmem_base + acc * BN, &tfull[acc], aph, m0, n0, ew, lane,
                                      scratch, wt == cluster_id, [&] { if (lane == 0) mbar_arrive(te); });
      if (++acc == 2) { acc = 0; aph ^= 1; }
      if (threadIdx.x == 128) stamp(p, wt == cluster_id ? 7 : 8);
    }
    if (is_resid_mode(MODE)) cp_async_wait<0>();
  }

  if (kTmaStore && warp >= 4 && lane == 0) tma_store_wait_all<0>();  // smem must outlive the stores
  tc_fence_before();
  // no CTA may exit while a peer can still multicast into its smem or arrive on its barriers
  if (CSIZE > 1) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) stamp(p, 9);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster compute one 256 x 256 output
// tile with M=256 MMAs issued by the leader CTA.  Each CTA stages its own 128 rows of A and only HALF
// of the W tile (128 of 256 rows); the tensor core reads the other half from the peer's shared
// memory.  Per SM that is 32 KB of operands per k-block instead of 48 KB -- the operand fill rate per
// SM, not L2 bandwidth, was the measured limit of the single-CTA main loop (tools/gemm_trace.py) --
// and the smaller stage buys a 6-deep ring.  Accumulators stay per CTA (rows 0..127 / 128..255 of the
// pair tile), so the epilogue is the same as the single-CTA kernel's.
constexpr int P2_STAGES = 6;
constexpr int P2_BN = 256;
constexpr int P2_BHALF_BYTES = (P2_BN / 2) * BK * 2;
constexpr int P2_STAGE_BYTES = A_BYTES + P2_BHALF_BYTES;
constexpr int P2_SMEM_BYTES = 1024 + P2_STAGES * P2_STAGE_BYTES + EPI_BYTES + BAR_BYTES;
static_assert(P2_SMEM_BYTES <= 227 * 1024, "dynamic shared memory budget exceeded");

template <int MODE, typename OutT>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_f16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA,
                             const __grid_constant__ CUtensorMap tmB,
                             const __grid_constant__ CUtensorMap tmC, Tc05Params p) {
  constexpr int BN = P2_BN, STAGES = P2_STAGES;
  constexpr bool kTmaStore = is_store_mode(MODE);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem;                               // [STAGES][128 x 64] my rows of A
  uint8_t* sB = smem + STAGES * A_BYTES;            // [STAGES][128 x 64] my half of the W tile
  float* sEpi = reinterpret_cast<float*>(smem + STAGES * P2_STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * P2_STAGE_BYTES + EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  if (threadIdx.x == 0) stamp(p, 0);

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmA);
  if (warp == 3 && lane == 0) prefetch_tensormap(&tmB);
  if (kTmaStore && warp == 2 && lane == 0) prefetch_tensormap(&tmC);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);   // leader only: its two producers arm the bytes of BOTH CTAs
      mbar_init(&empty[i], 1);  // both CTAs: multicast commit of the leader's MMA warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);               // both CTAs: multicast commit
      mbar_init(&tempty[i], 2 * kEpiWarps);  // leader only: the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, 2 * BN);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(p, 1);

  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_ptiles = ((p.M + 2 * BM - 1) / (2 * BM)) * tiles_n;
  const int KB = p.K / BK;

  if (warp == 0 || warp == 3) {
    if (lane == 0) {  // ===== TMA producers (both CTAs): warp 0 = my A rows, warp 3 = my W half =====
      const bool is_a = warp == 0;
      int s = 0, ph = 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        const int m0 = (pt / tiles_n) * 2 * BM + (int)crank * BM;
        const int n0 = (pt % tiles_n) * BN + (int)crank * (BN / 2);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          // the leader's barrier collects the bytes of both CTAs; only the leader's producers arm it
          if (crank == 0) mbar_arrive_expect_tx(&full[s], 2 * (is_a ? A_BYTES : P2_BHALF_BYTES));
          if (is_a) {
            tma_load_2d_2cta(sA + s * A_BYTES, &tmA, &full[s], kb * BK, m0);
            if (pt == pair_id && kb == 0) stamp(p, 2);
          } else {
            tma_load_2d_2cta(sB + s * P2_BHALF_BYTES, &tmB, &full[s], kb * BK, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
      if (is_a) stamp(p, 3);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (crank == 0 && lane == 0) {  // ===== MMA issuer: leader CTA only =====
      constexpr uint32_t idesc = umma_idesc(kFmtF16, 2 * BM, BN);
      int s = 0, ph = 0, acc = 0, aph = 0;
      for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
        mbar_wait(&tempty[acc], aph ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (pt == pair_id && kb == 0) stamp(p, 4);
          const uint64_t ad = umma_desc_k_sw128(smem_u32(sA + s * A_BYTES));
          const uint64_t bd = umma_desc_k_sw128(smem_u32(sB + s * P2_BHALF_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_2cta(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2cta(&empty[s], 0b11);  // stage free in both CTAs
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_2cta(&tfull[acc], 0b11);  // accumulator halves complete in both CTAs
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
      stamp(p, 5);
    }
    __syncwarp();
  } else if (warp >= 4) {  // ===== epilogue (both CTAs, own accumulator half) =====
    const int ew = warp - 4;
    float* scratch = sEpi + ew * 32 * EPI_LD;
    int acc = 0, aph = 0;
    for (int pt = pair_id; pt < num_ptiles; pt += num_pairs) {
      const int m0 = (pt / tiles_n) * 2 * BM + (int)crank * BM, n0 = (pt % tiles_n) * BN;
      uint64_t* te = &tempty[acc];
      epilogue_tile<MODE, BN, OutT>(p, &tmC, tmem_base + acc * BN, &tfull[acc], aph, m0, n0, ew, lane,
                                    scratch, pt == pair_id, [&] {
                                      if (lane == 0) {
                                        if (crank == 0) mbar_arrive(te);
                                        else mbar_arrive_remote(te, 0);
                                      }
                                    });
      if (++acc == 2) { acc = 0; aph ^= 1; }
      if (threadIdx.x == 128) stamp(p, pt == pair_id ? 7 : 8);
    }
  }

  if (kTmaStore && warp >= 4 && lane == 0) tma_store_wait_all<0>();
  tc_fence_before();
  cluster_sync_all();  // the peer's smem / barriers / TMEM half stay valid until both are done
  if (threadIdx.x == 0) stamp(p, 9);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 2 * BN);
  }
}

__global__ void convert_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride)
    reinterpret_cast<uint2*>(dst)[i] = pack_h4(__ldg(reinterpret_cast<const float4*>(src) + i));
}

// 16-bit floating point rows (fp16 or bf16: an autocast host feeds the encoder half tensors, main.py:439) -> fp32
template <bool BF16>
__global__ void widen_f32_kernel(const uint2* __restrict__ src, float4* __restrict__ dst, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const uint2 u = __ldg(src + i);
    float4 v;
    if (BF16) {
      v = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                      __uint_as_float(u.y & 0xffff0000u));
    } else {
      v = unpack_h4(u);
    }
    dst[i] = v;
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

// cuTensorMapEncodeTiled costs 1-2 us of host time and a bag needs a dozen descriptors; the same
// buffers (workspace, weight shadows) come back call after call, so encoded maps are kept in a small
// per-thread direct-mapped cache keyed by everything that goes into the descriptor.
struct MapKey {
  const void* base; int a, b, c, kind;
  bool operator==(const MapKey& o) const {
    return base == o.base && a == o.a && b == o.b && c == o.c && kind == o.kind;
  }
};
struct MapSlot { MapKey key; CUtensorMap map; bool valid; };
// 128 sets x 4 ways: a direct-mapped table of 128 slots ping-ponged between colliding keys once 8 lanes x ~10 maps
// were live (two encodes per call on the hot path, and cuTensorMapEncodeTiled inside ncu range replays)
constexpr int kMapSets = 128, kMapWays = 4;
bool cached_map(const MapKey& k, CUtensorMap* out, bool (*encode)(const MapKey&, CUtensorMap*)) {
  static thread_local MapSlot slots[kMapSets * kMapWays] = {};
  static thread_local unsigned char next_way[kMapSets] = {};
  size_t h = (reinterpret_cast<uintptr_t>(k.base) >> 8) * 0x9E3779B97F4A7C15ull;
  h ^= (size_t)k.a * 0x85EBCA6Bull ^ (size_t)k.b * 0xC2B2AE35ull ^ (size_t)k.c * 0x27D4EB2Full ^ (size_t)k.kind;
  const int set = (int)((h >> 20) % kMapSets);
  MapSlot* ways = slots + set * kMapWays;
  for (int w = 0; w < kMapWays; ++w)
    if (ways[w].valid && ways[w].key == k) { *out = ways[w].map; return true; }
  if (!encode(k, out)) return false;
  MapSlot& s = ways[next_way[set]];
  next_way[set] = (unsigned char)((next_way[set] + 1) % kMapWays);
  s.key = k; s.map = *out; s.valid = true;
  return true;
}

// row-major fp16 [rows, K] -> tiles of box_rows x 64 halves (128 bytes), 128-byte swizzle
bool encode_map(const MapKey& k, CUtensorMap* m);
bool make_map(CUtensorMap* m, const __half* base, int rows, int K, int box_rows) {
  return cached_map(MapKey{base, rows, K, box_rows, 0}, m, encode_map);
}
bool encode_map(const MapKey& key, CUtensorMap* m) {
  const __half* base = static_cast<const __half*>(key.base);
  const int rows = key.a, K = key.b, box_rows = key.c;
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

// output [rows, N] of 2- or 4-byte elements -> 32 x 32 element boxes; swizzle = the box row pitch
bool encode_store_map(const MapKey& k, CUtensorMap* m);
bool make_store_map(CUtensorMap* m, void* base, int rows, int N, int elem_bytes) {
  return cached_map(MapKey{base, rows, N, elem_bytes, 1}, m, encode_store_map);
}
bool encode_store_map(const MapKey& key, CUtensorMap* m) {
  void* base = const_cast<void*>(key.base);
  const int rows = key.a, N = key.b, elem_bytes = key.c;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)N * elem_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  elem_bytes == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// row-major fp16 [rows = K, cols = MN] -> 64 x 64 boxes (128-byte row segments), 128-byte swizzle
bool encode_map_mn(const MapKey& k, CUtensorMap* m);
bool make_map_mn(CUtensorMap* m, const __half* base, int k_rows, int mn_cols) {
  return cached_map(MapKey{base, k_rows, mn_cols, 0, 2}, m, encode_map_mn);
}
bool encode_map_mn(const MapKey& key, CUtensorMap* m) {
  const __half* base = static_cast<const __half*>(key.base);
  const int k_rows = key.a, mn_cols = key.b;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)mn_cols, (cuuint64_t)k_rows};
  cuuint64_t strides[1] = {(cuuint64_t)mn_cols * sizeof(__half)};
  cuuint32_t box[2] = {64, 64};
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

cudaError_t launch_widen_f32(const void* src, bool bf16, float* dst, size_t n, cudaStream_t stream) {
  if (n % 4 || ((reinterpret_cast<uintptr_t>(src) & 7) | (reinterpret_cast<uintptr_t>(dst) & 15)))
    return cudaErrorInvalidValue;
  const size_t n4 = n / 4;
  if (n4 == 0) return cudaSuccess;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (bf16) widen_f32_kernel<true><<<blocks, 256, 0, stream>>>(static_cast<const uint2*>(src), reinterpret_cast<float4*>(dst), n4);
  else widen_f32_kernel<false><<<blocks, 256, 0, stream>>>(static_cast<const uint2*>(src), reinterpret_cast<float4*>(dst), n4);
  return cudaGetLastError();
}

// SM cap of the bag-sized GEMMs; the batch entry point sets it while several bags are in flight
static thread_local int g_gemm_sm_cap = 0;
long long* g_gemm_trace = nullptr;  // debug hook (rrt_debug_set_gemm_trace): [8 launches][8 CTAs][16]
static int g_trace_launch = 0;

bool gemm_tcgen05_supported(int M, int N, int K) {
  return M >= 1 && N >= 4 && (N % 4) == 0 && K >= BK && (K % BK) == 0;
}

namespace {
template <int MODE, int BN, typename OutT, int CM, int CN>
cudaError_t launch_cfg(const __half* a, const __half* w, const Tc05Params& p, cudaStream_t stream) {
  using Cfg = TileCfg<BN, is_resid_mode(MODE)>;
  constexpr int CSIZE = CM * CN;
  CUtensorMap tmA, tmB, tmC;
  if (!make_map(&tmA, a, p.M, p.K, BM / CN) || !make_map(&tmB, w, p.N, p.K, BN / CM))
    return cudaErrorUnknown;
  if (is_store_mode(MODE)) {
    if (!make_store_map(&tmC, p.C, p.M, p.N, (int)sizeof(OutT))) return cudaErrorUnknown;
  } else {
    tmC = tmA;  // unused by the st.global epilogues
  }
  auto kern = gemm_f16_tcgen05_kernel<MODE, BN, OutT, CM, CN>;
  static DeviceOnce configured;  // per instantiation; one process drives one GPU
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const int ctiles = (((p.M + BM - 1) / BM + CM - 1) / CM) * (((p.N + BN - 1) / BN + CN - 1) / CN) *
                     (MODE == kEpiAtomicAdd ? p.ksplit : 1);
  int clusters = sm_count() / CSIZE;
  // SM cap (set_gemm_sm_cap / RRT_GEMM_SMS=n, 0 = none): persistent grid of at most n CTAs for the bag-sized
  // GEMMs.  With several bags in flight the GEMMs (tensor / operand-ingest bound, ~7 % of DRAM bandwidth)
  // then share the GPU with the HBM-bound kernels of the other bags instead of taking turns with them, and
  // every CTA pipelines 7 tiles instead of 3 (prologue and last epilogue amortised).  Measured, 16 bags,
  // 4 lanes: 71.2 us/bag on 148 SMs, 70.6 / 69.8 / 69.4 / 68.4 on 132 / 111 / 96 / 74, 67.7 on 64 and 56.
  static const int env_cap = [] { const char* e = getenv("RRT_GEMM_SMS"); return e ? atoi(e) : -1; }();
  // (RRT_GEMM_SMS_RESID: separate tuning knob for the residual-epilogue (proj) GEMM)
  static const int env_cap_resid = [] { const char* e = getenv("RRT_GEMM_SMS_RESID"); return e ? atoi(e) : -1; }();
  const int sm_cap = (is_resid_mode(MODE) && env_cap_resid >= 0 && g_gemm_sm_cap > 0)
                         ? env_cap_resid
                         : (env_cap >= 0 ? env_cap : g_gemm_sm_cap);
  if (sm_cap > 0 && BN == 256 && clusters > sm_cap / CSIZE) clusters = sm_cap / CSIZE > 0 ? sm_cap / CSIZE : 1;
  if (ctiles < clusters) clusters = ctiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CSIZE);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CSIZE > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CSIZE;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  } else if (g_pdl) {  // programmatic dependent launch of the serial chain (common.cuh)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, p);
}

// CM*10 + CN for the bag-sized GEMMs (rrt_debug_set_gemm_cluster).  Measured on B200 (bench_v8):
// 1x1 23.0 us, 2x1 23.4 us, 2x2 38.8 us for the QKV GEMM -- L2 already de-duplicates the operand
// reads of neighbouring CTAs, and the lock-step coupling of a cluster costs more than it saves, so
// clusters stay OFF by default; the code path is kept for tuning.
int g_gemm_cluster = 11;
// 1: bag-sized GEMMs on the cta_group::2 kernel (rrt_debug_set_gemm_cluster(2)).  Measured (bench_v11):
// main loop 5.2k cycles/tile vs 5.7k single-CTA (5.7k IS the cuBLAS-measured tensor rate), but the
// cluster set-up and the longer tail cost more than that buys at 3 tiles per CTA: QKV 23.0 vs 20.9 us.
int g_gemm_pair = 0;
// 1: 128-column tiles for bag-sized GEMMs that would otherwise run one 256-column tile per CTA
// (rrt_debug_set_gemm_cluster(128)).  Measured (s3_marginal): 16 bags, 4 lanes: 74.1 us/bag with 128-column
// proj tiles vs 71.3 with 256 -- two narrow tiles ingest 2 x 256 KB of operands per CTA instead of 384 KB and
// the main loop is operand-ingest bound (~67 B/clk/SM), which costs more than the hidden epilogue saves.
int g_gemm_narrow = 0;

template <int MODE, typename OutT>
cudaError_t launch_pair(const __half* a, const __half* w, const Tc05Params& p, cudaStream_t stream) {
  CUtensorMap tmA, tmB, tmC;
  if (!make_map(&tmA, a, p.M, p.K, BM) || !make_map(&tmB, w, p.N, p.K, P2_BN / 2)) return cudaErrorUnknown;
  if (is_store_mode(MODE)) {
    if (!make_store_map(&tmC, p.C, p.M, p.N, (int)sizeof(OutT))) return cudaErrorUnknown;
  } else {
    tmC = tmA;
  }
  auto kern = gemm_f16_tcgen05_2cta_kernel<MODE, OutT>;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const int ptiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + P2_BN - 1) / P2_BN);
  int pairs = sm_count() / 2;
  static const int env_cap = [] { const char* e = getenv("RRT_GEMM_SMS"); return e ? atoi(e) : -1; }();
  const int sm_cap = env_cap >= 0 ? env_cap : g_gemm_sm_cap;   // see launch_cfg
  if (sm_cap > 0 && pairs > sm_cap / 2) pairs = sm_cap / 2 > 0 ? sm_cap / 2 : 1;
  if (ptiles < pairs) pairs = ptiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pairs * 2);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = P2_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, p);
}

}  // namespace
// landmark_chain.cu builds its operand maps with the same encoder / cache
bool tc05_make_kmajor_map(CUtensorMap* m, const __half* base, int rows, int K, int box_rows) {
  return make_map(m, base, rows, K, box_rows);
}
// Row-statistics side output of the residual epilogue: number of 128-column parts per row, or 0 when the
// launch below would not pick a 256-column tile for this problem (the part width is half a tile).
// Must mirror launch_mode().
int gemm_tcgen05_rowstat_parts(int M, int N) {
  const int tiles_m = (M + BM - 1) / BM, tiles_n256 = (N + 255) / 256;
  if (tiles_m * tiles_n256 < sm_count() / 2) return 0;                               // BN = 64
  if (g_gemm_narrow && tiles_m * tiles_n256 <= sm_count() && N % 128 == 0) return 0;  // BN = 128
  if (N % 128) return 0;
  return N / 128;
}
namespace {
template <int MODE, typename OutT>
cudaError_t launch_mode(const __half* a, const __half* w, const Tc05Params& p, cudaStream_t stream) {
  // small problems: narrower tiles so that more SMs share the (latency-bound) work
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n256 = (p.N + 255) / 256;
  if (tiles_m * tiles_n256 < sm_count() / 2) return launch_cfg<MODE, 64, OutT, 1, 1>(a, w, p, stream);
  // at most one 256-column tile per CTA: halve the tile width so that every CTA pipelines two tiles
  if (g_gemm_narrow && tiles_m * tiles_n256 <= sm_count() && p.N % 128 == 0)
    return launch_cfg<MODE, 128, OutT, 1, 1>(a, w, p, stream);
  if (MODE == kEpiStoreAct) return launch_cfg<MODE, 256, OutT, 1, 1>(a, w, p, stream);  // no tuning variants
  // (the residual epilogues exist for the single-CTA kernel only: their cp.async ring follows its tile schedule)
  // CTA pairs also whenever the SM cap is on (several bags in flight: SM time is what counts there, and a pair moves
  // 512 KB instead of 768 KB per tile through each SM's shared memory): 8 lanes 59.9 -> 59.2 us per bag.  A single bag
  // keeps the single-CTA kernel (20.9 vs 23.0 us: cluster set-up and the longer tail at 3 tiles per CTA).
  if constexpr (!is_resid_mode(MODE) && MODE != kEpiStoreAct) {
    static const bool pair_capped = [] { const char* e = getenv("RRT_GEMM_PAIR_CAPPED"); return !(e && atoi(e) == 0); }();
    if ((g_gemm_pair || (pair_capped && g_gemm_sm_cap > 0)) && tiles_m >= 2) return launch_pair<MODE, OutT>(a, w, p, stream);
  }
  // bag-sized problems: clusters with multicast operand tiles when the tile grid allows it
  if (g_gemm_cluster == 22 && tiles_m >= 2 && tiles_n256 >= 2 && tiles_n256 % 2 == 0)
    return launch_cfg<MODE, 256, OutT, 2, 2>(a, w, p, stream);
  if (g_gemm_cluster >= 21 && tiles_m >= 2) return launch_cfg<MODE, 256, OutT, 2, 1>(a, w, p, stream);
  return launch_cfg<MODE, 256, OutT, 1, 1>(a, w, p, stream);
}
}  // namespace

// dw[C_out, C_in] (fp32, zeroed by the caller) += dy[rows, C_out]^T @ act[rows, C_in]: both operands
// MN-major straight from the row-major activations, split-K over the rows so that every SM has work.
cudaError_t launch_gemm_tcgen05_wgrad(const __half* dy, const __half* act, float* dw, int rows, int C_out,
                                      int C_in, cudaStream_t stream, const uint32_t* unscale_amax) {
  if (rows < 1 || C_out < 8 || C_in < 8 || (C_out % 8) || (C_in % 8)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(dw)) & 15)
    return cudaErrorInvalidValue;
  constexpr int BN = 256;
  using Cfg = TileCfg<BN>;
  Tc05Params p{};
  p.M = C_out; p.N = C_in; p.K = rows;
  p.C = dw;
  p.act = kActNone;
  p.act2 = kActNone;
  p.act_split = 0;
  p.rs_part = nullptr; p.rs_gamma = nullptr; p.rs_phi = nullptr; p.rs_k = 0;
  p.trace = nullptr;
  p.unscale_amax = unscale_amax;
  const int tiles = ((C_out + BM - 1) / BM) * ((C_in + BN - 1) / BN), KB = (rows + BK - 1) / BK;
  int ksplit = sm_count() / tiles;
  if (ksplit > KB) ksplit = KB;
  if (ksplit < 1) ksplit = 1;
  p.ksplit = ksplit;
  CUtensorMap tmA, tmB;
  if (!make_map_mn(&tmA, dy, rows, C_out) || !make_map_mn(&tmB, act, rows, C_in)) return cudaErrorUnknown;
  auto kern = gemm_f16_tcgen05_kernel<kEpiAtomicAdd, BN, float, 1, 1, 1>;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  int ctas = tiles * ksplit;
  if (ctas > sm_count()) ctas = sm_count();
  kern<<<ctas, NTHREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmA, p);
  return cudaGetLastError();
}

// input gradient: d_in[rows, C_in] (fp16) = dy[rows, C_out] (fp16, K-major) @ w16[C_out, C_in] (fp16, row-major =
// MN-major B operand): the forward's fp16 weight as it is, no transposed copy
cudaError_t launch_gemm_tcgen05_dgrad(const __half* dy, const __half* w16, __half* d_in, int rows, int C_out,
                                      int C_in, cudaStream_t stream) {
  constexpr int BN = 256;
  using Cfg = TileCfg<BN>;
  if (rows < 1 || C_out < BK || (C_out % BK) || C_in < 8 || (C_in % 8)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(w16) | reinterpret_cast<uintptr_t>(d_in)) & 15)
    return cudaErrorInvalidValue;
  Tc05Params p{};
  p.M = rows; p.N = C_in; p.K = C_out;
  p.C = d_in;
  p.act = kActNone; p.act2 = kActNone; p.act_split = 0;
  p.ksplit = 1;
  CUtensorMap tmA, tmB, tmC;
  if (!make_map(&tmA, dy, rows, C_out, BM) || !make_map_mn(&tmB, w16, C_out, C_in) ||
      !make_store_map(&tmC, d_in, rows, C_in, 2))
    return cudaErrorUnknown;
  auto kern = gemm_f16_tcgen05_kernel<kEpiStore, BN, __half, 1, 1, 2>;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const int tiles = ((rows + BM - 1) / BM) * ((C_in + BN - 1) / BN);
  int ctas = sm_count();
  if (g_gemm_sm_cap > 0 && ctas > g_gemm_sm_cap) ctas = g_gemm_sm_cap;
  if (tiles < ctas) ctas = tiles;
  kern<<<ctas, NTHREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmC, p);
  return cudaGetLastError();
}

void set_gemm_sm_cap(int n) { g_gemm_sm_cap = n; }

void set_gemm_cluster_mode(int mode) {
  if (mode == 128 || mode == 256) { g_gemm_narrow = mode == 128; return; }
  if (mode == 2) { g_gemm_pair = 1; return; }
  g_gemm_pair = 0;
  g_gemm_cluster = mode;
}

cudaError_t launch_gemm_tcgen05(const __half* a, const __half* w, void* c, bool out_f16, int M, int N,
                                int K, const GemmEpilogue& epi, cudaStream_t stream) {
  if (!gemm_tcgen05_supported(M, N, K)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
       reinterpret_cast<uintptr_t>(c)) & 15)
    return cudaErrorInvalidValue;
  Tc05Params p;
  p.M = M; p.N = N; p.K = K;
  p.bias = epi.bias; p.C = c; p.resid = epi.resid; p.grid = epi.grid;
  p.act = epi.mode == kEpiTanh ? (int)kActTanh : epi.act;
  p.act2 = epi.act2;
  p.act_split = epi.act_split;
  if (p.act_split % 32 || p.act_split < 0) return cudaErrorInvalidValue;
  p.rs_part = nullptr; p.rs_gamma = nullptr; p.rs_phi = nullptr; p.rs_k = 0;
  p.unscale_amax = epi.mode == kEpiAtomicAdd ? epi.unscale_amax : nullptr;
  if (epi.rs_part) {
    if (!is_resid_mode(epi.mode) || !epi.rs_gamma || !epi.rs_phi || epi.rs_k < 1 || epi.rs_k > 4 ||
        gemm_tcgen05_rowstat_parts(M, N) == 0)
      return cudaErrorInvalidValue;
    p.rs_part = epi.rs_part; p.rs_gamma = epi.rs_gamma; p.rs_phi = epi.rs_phi; p.rs_k = epi.rs_k;
  }
  p.trace = g_gemm_trace ? g_gemm_trace + (size_t)(g_trace_launch++ % 8) * 128 : nullptr;
  p.drop = epi.drop;
  p.ksplit = 1;
  if (epi.mode == kEpiAtomicAdd) {
    // c (fp32, zeroed by the caller) += a @ w^T, the K loop of every 128x256 tile cut into as many slices
    // as it takes to give every SM a work item (weight gradients: few output tiles, K = the bag's tokens)
    if (out_f16 || epi.bias) return cudaErrorInvalidValue;
    const int tiles = ((M + BM - 1) / BM) * ((N + 255) / 256), KB = K / BK;
    int ksplit = sm_count() / tiles;
    if (ksplit > KB) ksplit = KB;
    if (ksplit < 1) ksplit = 1;
    p.ksplit = ksplit;
    return launch_cfg<kEpiAtomicAdd, 256, float, 1, 1>(a, w, p, stream);
  }
  if (epi.mode == kEpiResidualUnpartDrop && epi.drop.on())
    return out_f16 ? cudaErrorInvalidValue : launch_mode<kEpiResidualUnpartDrop, float>(a, w, p, stream);
  if (is_resid_mode(epi.mode))
    return out_f16 ? cudaErrorInvalidValue : launch_mode<kEpiResidualUnpart, float>(a, w, p, stream);
  if (p.act != kActNone || (p.act_split > 0 && p.act2 != kActNone))
    return out_f16 ? launch_mode<kEpiStoreAct, __half>(a, w, p, stream)
                   : launch_mode<kEpiStoreAct, float>(a, w, p, stream);
  return out_f16 ? launch_mode<kEpiStore, __half>(a, w, p, stream)
                 : launch_mode<kEpiStore, float>(a, w, p, stream);
}

}  // namespace rrt
