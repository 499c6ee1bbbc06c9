// R-MSA attention core for regions of up to 256 tokens (every configuration the reference ships:
// P = 144 at N = 9000 / region_num 8, P = 196 at N = 50000 / region_num 16), one CTA per
// (region, head), the whole region resident in shared memory (fp16, as written by the QKV GEMM):
//   load   Q rows (with the EPEG halo), K, V of this head: cp.async 16-byte copies, no conversion
//   EPEG   Q' = scale*log2e * (Q + dwconv1d_P(Q; taps_h)) as a banded-Toeplitz product on the
//          tensor path: Q'[16 rows] = C[16 x (16+k-1)] . Q[halo rows], C[i][r] = taps[r-i] (+1 on the
//          diagonal); the fp32 accumulator fragment IS the A fragment of the next product
//   core   one warp per 16 query rows: S = Q' K^T (ldmatrix + mma.sync m16n8k16, fp32 accumulate),
//          online softmax in registers over KV tiles of 48 or 64 keys, O += P V, fp16 O out
// fp16 operands carry the same 10-bit mantissa as tf32; softmax state and accumulators are fp32.
// (modules/rmsa.py:103-122; SURVEY.md 0.2-1 for the EPEG-on-Q identity.)
#include "kernels.cuh"
#include "mma_f16.cuh"

namespace rrt {
long long* g_attn_trace = nullptr;  // debug: clock64 stamps of CTA 0..7 (tools/attn_trace.py)
namespace {
__device__ __forceinline__ void astamp(long long* tr, int slot) {
  if (tr && blockIdx.x + blockIdx.y < 8 && (blockIdx.x == 0 || blockIdx.y == 0) && threadIdx.x == 0)
    tr[(blockIdx.x + blockIdx.y) * 8 + slot] = clock64();
}

// HD: head dim; NT: 8-key n-tiles per KV tile (6 -> 48 keys, 8 -> 64 keys); MAXW: warps per CTA cap
template <int HD, int NT, int MAXW>
__global__ void __launch_bounds__(32 * MAXW) __maxnreg__(MAXW <= 9 ? 96 : 128) rmsa_attn_f16_kernel(const __half* __restrict__ qkv,
                                                                  const float* __restrict__ taps,
                                                                  __half* __restrict__ o, Grid grid,
                                                                  int D, int epeg_k, float qscale,
                                                                  int n_kv_tiles, int q_rows,
                                                                  long long* tr, int heads_fastest) {
  constexpr int LDH = HD + 8;   // halves per smem row: 16-byte row skew keeps ldmatrix conflict-free
  constexpr int KS = HD / 16;   // k16 steps over head_dim
  constexpr int ND = HD / 8;    // 8-wide n-tiles over head_dim
  constexpr int KV = 8 * NT;    // keys per KV tile
  constexpr int C8 = HD / 8;    // 16-byte chunks per row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = grid.P;
  const int pad = taps ? epeg_k / 2 : 0;
  const int pk = n_kv_tiles * KV;  // key rows incl. padding of the last tile
  __half* Qs = reinterpret_cast<__half*>(smem_raw);  // [q_rows][LDH]: row r holds Q[r - pad]
  __half* Ks = Qs + (size_t)q_rows * LDH;             // [pk][LDH]
  __half* Vs = Ks + (size_t)pk * LDH;                 // [pk][LDH]
  float* Ts = reinterpret_cast<float*>(Vs + (size_t)pk * LDH);  // [epeg_k]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  pdl_launch_dependents();
  pdl_wait();
  // heads fastest: the CTAs that run side by side read ADJACENT 128-byte segments of the same qkv rows
  // (one DRAM page / L2 line neighbourhood) instead of the same segment of rows 144 * 3 KB apart
  const int rho = heads_fastest ? blockIdx.y : blockIdx.x, h = heads_fastest ? blockIdx.x : blockIdx.y;
  astamp(tr, 0);
  const size_t ld = 3 * (size_t)D;
  const __half* base = qkv + (size_t)rho * P * ld + h * HD;

  // ---- stage Q (halo), K, V: asynchronous 16-byte copies, zero fill outside the region ----------
  for (int i = tid; i < q_rows * C8; i += blockDim.x) {
    int r = i / C8, c = (i - r * C8) * 8;
    int p = r - pad;
    bool ok = p >= 0 && p < P;
    cp_async16(Qs + (size_t)r * LDH + c, base + (size_t)(ok ? p : 0) * ld + c, ok);
  }
  cp_async_commit();  // group 0: Q (needed first, by the EPEG product)
  for (int i = tid; i < pk * C8; i += blockDim.x) {
    int r = i / C8, c = (i - r * C8) * 8;
    bool ok = r < P;
    cp_async16(Ks + (size_t)r * LDH + c, base + (size_t)(ok ? r : 0) * ld + c + D, ok);
  }
  cp_async_commit();  // group 1: K (first needed by S = Q'K^T)
  for (int i = tid; i < pk * C8; i += blockDim.x) {
    int r = i / C8, c = (i - r * C8) * 8;
    bool ok = r < P;
    cp_async16(Vs + (size_t)r * LDH + c, base + (size_t)(ok ? r : 0) * ld + c + 2 * D, ok);
  }
  cp_async_commit();  // group 2: V (first needed by the first P.V): lands behind the work on Q and K
  if (taps)
    for (int i = tid; i < epeg_k; i += blockDim.x) Ts[i] = __ldg(taps + h * epeg_k + i);
  astamp(tr, 1);
  cp_async_wait<2>();
  __syncthreads();
  astamp(tr, 2);

  // ---- Q' fragments ----------------------------------------------------------------------------
  const int i0 = 16 * warp;  // first query row of this warp (= first halo row of its band)
  uint32_t qa[KS][4];
  if (taps) {
    float qacc[ND][4];
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) qacc[i][e] = 0.f;
    const int nkc = (16 + epeg_k - 1 + 15) / 16;  // k16 steps that cover the band of 16+k-1 rows
    for (int kc = 0; kc < nkc; ++kc) {
      // A fragment of the Toeplitz band: element (row i, halo row r) = taps[r-i] (+1 at r-i = pad)
      uint32_t ca[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ro = g + (e & 1) * 8;                    // row offset within the warp's 16 rows
        const int co = 16 * kc + 2 * t + (e >> 1) * 8;     // halo-row offset of the first column
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int d = co + u - ro;
          float x = (d >= 0 && d < epeg_k) ? Ts[d] : 0.f;
          v[u] = x + (d == pad ? 1.f : 0.f);
        }
        ca[e] = pack_h2(v[0], v[1]);
      }
#pragma unroll
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t b[4];
        ldsm_x4_trans(b, Qs + (size_t)(i0 + 16 * kc + (lane & 7) + ((lane >> 3) & 1) * 8) * LDH +
                             np * 16 + (lane >> 4) * 8);
        mma_f16_16x8x16(qacc[2 * np], ca, b[0], b[1]);
        mma_f16_16x8x16(qacc[2 * np + 1], ca, b[2], b[3]);
      }
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      qa[ks][0] = pack_h2(qacc[2 * ks][0] * qscale, qacc[2 * ks][1] * qscale);
      qa[ks][1] = pack_h2(qacc[2 * ks][2] * qscale, qacc[2 * ks][3] * qscale);
      qa[ks][2] = pack_h2(qacc[2 * ks + 1][0] * qscale, qacc[2 * ks + 1][1] * qscale);
      qa[ks][3] = pack_h2(qacc[2 * ks + 1][2] * qscale, qacc[2 * ks + 1][3] * qscale);
    }
  } else {
    const __half2 sc = __float2half2_rn(qscale);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      ldsm_x4(qa[ks], Qs + (size_t)(i0 + (lane & 15)) * LDH + ks * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __half2 v = __hmul2(*reinterpret_cast<__half2*>(&qa[ks][e]), sc);
        qa[ks][e] = *reinterpret_cast<uint32_t*>(&v);
      }
    }
  }

  astamp(tr, 3);
  cp_async_wait<1>();
  __syncthreads();  // K has landed
  // ---- attention core ---------------------------------------------------------------------------
  float oacc[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[i][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int kt = 0; kt < n_kv_tiles; ++kt) {
    const int kt0 = kt * KV;
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b[4];
        ldsm_x4(b, Ks + (size_t)(kt0 + np * 16 + (lane & 7) + (lane >> 4) * 8) * LDH + ks * 16 +
                       ((lane >> 3) & 1) * 8);
        mma_f16_16x8x16(s[2 * np], qa[ks], b[0], b[1]);
        mma_f16_16x8x16(s[2 * np + 1], qa[ks], b[2], b[3]);
      }
    }
    if (kt0 + KV > P) {  // tile padding keys (zero-pad TOKENS are real keys and stay unmasked)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kt0 + nt * 8 + 2 * t + (e & 1) >= P) s[nt][e] = -INFINITY;
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
    for (int i = 0; i < ND; ++i) {
      oacc[i][0] *= corr[0]; oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1]; oacc[i][3] *= corr[1];
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pv = fast_exp2(s[nt][e] - mx[e >> 1]);
        l_run[e >> 1] += pv;
        s[nt][e] = pv;
      }
    if (kt == 0) {  // V has landed (uniform branch: every warp runs the same KV tiles)
      cp_async_wait<0>();
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {  // 16 keys per step: two S n-tiles form one A fragment
      uint32_t pa[4] = {pack_h2(s[2 * j][0], s[2 * j][1]), pack_h2(s[2 * j][2], s[2 * j][3]),
                        pack_h2(s[2 * j + 1][0], s[2 * j + 1][1]),
                        pack_h2(s[2 * j + 1][2], s[2 * j + 1][3])};
#pragma unroll
      for (int np = 0; np < ND / 2; ++np) {
        uint32_t b[4];
        ldsm_x4_trans(b, Vs + (size_t)(kt0 + j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDH +
                             np * 16 + (lane >> 4) * 8);
        mma_f16_16x8x16(oacc[2 * np], pa, b[0], b[1]);
        mma_f16_16x8x16(oacc[2 * np + 1], pa, b[2], b[3]);
      }
    }
  }

  astamp(tr, 4);
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
  }
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    int q = i0 + g + hh * 8;
    if (q >= P) continue;
    float inv = 1.f / l_run[hh];
    __half* orow = o + ((size_t)rho * P + q) * D + h * HD + 2 * t;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd)
      *reinterpret_cast<uint32_t*>(orow + nd * 8) =
          pack_h2(oacc[nd][hh * 2] * inv, oacc[nd][hh * 2 + 1] * inv);
  }
  astamp(tr, 5);
}

template <int HD, int NT, int MAXW>
cudaError_t launch(const __half* qkv, const float* taps, __half* o, const Grid& grid, int D,
                   int heads, int epeg_k, cudaStream_t stream) {
  const int W = (grid.P + 15) / 16;
  const int KV = 8 * NT;
  const int tiles = (grid.P + KV - 1) / KV;
  const int pad = taps ? epeg_k / 2 : 0;
  // halo'd Q rows: every warp's band [16w, 16w + 16*nkc) must exist (zero filled past the data)
  const int nkc = taps ? (16 + epeg_k - 1 + 15) / 16 : 1;
  int q_rows = 16 * (W - 1) + 16 * nkc;
  if (q_rows < 16 * W + 2 * pad) q_rows = 16 * W + 2 * pad;
  size_t smem = ((size_t)q_rows + 2 * (size_t)tiles * KV) * (HD + 8) * sizeof(__half) +
                (taps ? epeg_k : 0) * sizeof(float) + 16;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(rmsa_attn_f16_kernel<HD, NT, MAXW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(rmsa_attn_f16_kernel<HD, NT, MAXW>,
                               cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
  }
  const float kLog2e = 1.4426950408889634f;
  float qscale = kLog2e / sqrtf((float)HD);
  // tuning knob RRT_ATTN_ORDER=region: regions fastest (the first version's order)
  static const bool region_major = [] { const char* e = getenv("RRT_ATTN_ORDER"); return e && !strcmp(e, "region"); }();
  const int heads_fastest = (!region_major && grid.R <= 65535) ? 1 : 0;
  dim3 g = heads_fastest ? dim3(heads, grid.R) : dim3(grid.R, heads);
  return launch_chain_kernel(rmsa_attn_f16_kernel<HD, NT, MAXW>, g, dim3(32 * W), smem, stream, qkv, taps, o,
                             grid, D, epeg_k, qscale, tiles, q_rows, g_attn_trace, heads_fastest);
}

template <int HD>
cudaError_t launch_hd(const __half* qkv, const float* taps, __half* o, const Grid& grid, int D,
                      int heads, int epeg_k, cudaStream_t stream) {
  // KV tile of 48 or 64 keys, whichever pads the region's keys less; <= 9 warps (P <= 144) gets the
  // tighter register budget so that two CTAs fit the per-scheduler register files
  int pad48 = (grid.P + 47) / 48 * 48, pad64 = (grid.P + 63) / 64 * 64;
  const bool small = grid.P <= 144;
  if (pad48 < pad64)
    return small ? launch<HD, 6, 9>(qkv, taps, o, grid, D, heads, epeg_k, stream)
                 : launch<HD, 6, 16>(qkv, taps, o, grid, D, heads, epeg_k, stream);
  return small ? launch<HD, 8, 9>(qkv, taps, o, grid, D, heads, epeg_k, stream)
               : launch<HD, 8, 16>(qkv, taps, o, grid, D, heads, epeg_k, stream);
}
}  // namespace

bool rmsa_attention_f16_supported(const Grid& grid, int D, int heads) {
  int hd = heads > 0 ? D / heads : 0;
  return grid.P <= 256 && (hd == 32 || hd == 64 || hd == 128) && heads <= 65535;
}

cudaError_t launch_rmsa_attention_f16(const __half* qkv, const float* taps, __half* o,
                                      const Grid& grid, int D, int heads, int epeg_k,
                                      cudaStream_t stream) {
  if (!rmsa_attention_f16_supported(grid, D, heads)) return cudaErrorInvalidValue;
  switch (D / heads) {
    case 32: return launch_hd<32>(qkv, taps, o, grid, D, heads, epeg_k, stream);
    case 64: return launch_hd<64>(qkv, taps, o, grid, D, heads, epeg_k, stream);
    default: return launch_hd<128>(qkv, taps, o, grid, D, heads, epeg_k, stream);
  }
}

}  // namespace rrt
