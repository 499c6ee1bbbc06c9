// Ablation positional encodings of RRTEncoder: PEG / PPEG
// (modules/emb_position.py:24-82).  The bag's tokens are folded row-major into an Hg x Hg grid (the
// tail of the last row is filled by wrapping around to the FIRST tokens, :39,:71; PPEG extends grids
// smaller than 7 x 7 with zero cells, :41-45) and every channel gets a depthwise "same" convolution:
//   PEG :  out = x + conv_k(x)                              PPEG:  out = x + conv_k(x) + conv_5(x) + conv_3(x)
// (k x k kernels, or (k,1) column kernels with peg_1d).  Zero-padded "same" depthwise convolutions of
// different odd sizes add up to ONE K x K convolution (K = the largest), so
//   peg_fold_kernel   folds the 1-3 kernels, the identity tap and the biases into w_eff[K*KW][D], b_eff[D]
//                     (tap-major, channel fastest: the layout the streaming kernel reads coalesced)
//   peg_apply_kernel  one thread per (token, 4 channels): <= K*KW float4 loads of neighbours (L1/L2
//                     resident: every token row is read K*KW times by adjacent threads) and of w_eff.
// Token-major [L, D] in and out, channel fastest: every load and store is a coalesced 16-byte access.
#include <cstdlib>
#include <cstring>
#include "kernels.cuh"

namespace rrt {
namespace {

struct PegGeom {
  int L, D;
  int Hn;  // ceil(sqrt(L)): cells [L, Hn*Hn) wrap to tokens [0, Hn*Hn - L)
  int Hg;  // grid side (7 when PPEG and Hn < 7: cells >= Hn*Hn are zero)
  int K, KW;
};

__global__ void __launch_bounds__(256) peg_fold_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                                       int k0, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, int k1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       int k2, int conv_1d, float* __restrict__ weff,
                                                       float* __restrict__ beff, int D, int K, int KW) {
  // reference weights: [D, 1, k, k] (or [D, 1, k, 1] with conv_1d), one kernel per channel
  const int n = K * KW * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + D; i += gridDim.x * blockDim.x) {
    if (i >= n) {
      const int c = i - n;
      beff[c] = (b0 ? __ldg(b0 + c) : 0.f) + (b1 ? __ldg(b1 + c) : 0.f) + (b2 ? __ldg(b2 + c) : 0.f);
      continue;
    }
    const int c = i % D, tap = i / D;
    const int dy = tap / KW - K / 2, dx = KW == 1 ? 0 : tap % KW - K / 2;
    float v = (dy == 0 && dx == 0) ? 1.f : 0.f;  // the "+ cnn_feat" identity
    const float* ws[3] = {w0, w1, w2};
    const int ks[3] = {k0, k1, k2};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (!ws[j]) continue;
      const int k = ks[j], kw = conv_1d ? 1 : k;
      const int y = dy + k / 2, x = conv_1d ? 0 : dx + k / 2;
      if (y >= 0 && y < k && x >= 0 && x < kw) v += __ldg(ws[j] + ((size_t)c * k + y) * kw + x);
    }
    weff[i] = v;
  }
}

__global__ void __launch_bounds__(256) peg_apply_kernel(const float* __restrict__ x,
                                                        const float* __restrict__ weff,
                                                        const float* __restrict__ beff,
                                                        float* __restrict__ out, PegGeom g) {
  const int q4 = g.D / 4;
  const long long items = (long long)g.L * q4;
  const int wrap_end = g.Hn * g.Hn;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(it / q4), q = (int)(it - (long long)t * q4);
    const int r = t / g.Hg, c = t - r * g.Hg;
    float4 acc = __ldg(reinterpret_cast<const float4*>(beff) + q);
    for (int ky = 0; ky < g.K; ++ky) {
      const int rr = r + ky - g.K / 2;
      if (rr < 0 || rr >= g.Hg) continue;
      for (int kx = 0; kx < g.KW; ++kx) {
        const int cc = g.KW == 1 ? c : c + kx - g.K / 2;
        if (cc < 0 || cc >= g.Hg) continue;
        const int cell = rr * g.Hg + cc;
        const int src = cell < g.L ? cell : (cell < wrap_end ? cell - g.L : -1);
        if (src < 0) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)src * g.D) + q);
        const float4 w = __ldg(reinterpret_cast<const float4*>(weff + (size_t)(ky * g.KW + kx) * g.D) + q);
        acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
        acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(out + (size_t)t * g.D)[q] = acc;
  }
}

// Tiled version: CTA = 8 x 16 output cells x 32 channels.  The (8+K-1) x (16+KW-1) halo of the tile (wrap /
// zero rules applied while filling) and the 32-channel slab of the folded kernel are staged in shared
// memory once (coalesced 128-byte cell rows); a thread owns one channel quad and a horizontal strip of 4
// cells, so every tap is 4 LDS.128 of activations + 1 of weights for 16 FMAs.  Activations are read from
// HBM / L2 (K+7)(KW+15)/128 times instead of K*KW times through L1.
// CK / CKW > 0: compile-time kernel size -- the tap loops unroll and a thread keeps the CKW+3 cells of a halo
// row in registers (sliding window: (CKW+3) + CKW loads per tap row instead of 5 CKW); 0: run-time sizes.
constexpr int kTileH = 8, kTileW = 16, kSlab = 32;
template <int CK, int CKW>
__global__ void __launch_bounds__(256) peg_apply_tiled_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ weff,
                                                              const float* __restrict__ beff,
                                                              float* __restrict__ out, PegGeom g) {
  extern __shared__ __align__(16) float sm[];
  const int K = CK > 0 ? CK : g.K, KW = CK > 0 ? CKW : g.KW;
  const int HH = kTileH + K - 1, HW = kTileW + KW - 1;
  float4* sx = reinterpret_cast<float4*>(sm);                       // [HH][HW][8 quads]
  float4* sw = sx + (size_t)HH * HW * (kSlab / 4);                   // [K*KW][8 quads]
  const int tid = threadIdx.x;
  const int c0 = blockIdx.z * kSlab;                                 // first channel of the slab
  const int r0 = blockIdx.y * kTileH, col0 = blockIdx.x * kTileW;    // first output cell of the tile
  const int py = K / 2, px = KW == 1 ? 0 : K / 2;
  const int wrap_end = g.Hn * g.Hn;
  for (int i = tid; i < HH * HW * (kSlab / 4); i += blockDim.x) {
    const int q = i & 7, cell_i = i >> 3;
    const int hy = cell_i / HW, hx = cell_i - hy * HW;
    const int rr = r0 + hy - py, cc = col0 + hx - px;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr >= 0 && rr < g.Hg && cc >= 0 && cc < g.Hg) {
      const int cell = rr * g.Hg + cc;
      const int src = cell < g.L ? cell : (cell < wrap_end ? cell - g.L : -1);
      if (src >= 0) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)src * g.D + c0) + q);
    }
    sx[i] = v;
  }
  for (int i = tid; i < K * KW * (kSlab / 4); i += blockDim.x)
    sw[i] = __ldg(reinterpret_cast<const float4*>(weff + (size_t)(i >> 3) * g.D + c0) + (i & 7));
  __syncthreads();
  const int q = tid & 7, strip = tid >> 3;          // 32 strips of 4 cells: 8 rows x 4 strips
  const int ty = strip >> 2, tx = (strip & 3) * 4;
  const float4 b = __ldg(reinterpret_cast<const float4*>(beff + c0) + q);
  float4 acc[4] = {b, b, b, b};
  if (CK > 0) {
#pragma unroll
    for (int ky = 0; ky < (CK > 0 ? CK : 1); ++ky) {
      const float4* row = sx + ((size_t)(ty + ky) * HW + tx) * (kSlab / 4) + q;
      float4 xin[(CKW > 0 ? CKW : 1) + 3];
#pragma unroll
      for (int i = 0; i < (CKW > 0 ? CKW : 1) + 3; ++i) xin[i] = row[(size_t)i * (kSlab / 4)];
#pragma unroll
      for (int kx = 0; kx < (CKW > 0 ? CKW : 1); ++kx) {
        const float4 w = sw[(ky * KW + kx) * (kSlab / 4) + q];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = xin[kx + j];
          acc[j].x = fmaf(v.x, w.x, acc[j].x); acc[j].y = fmaf(v.y, w.y, acc[j].y);
          acc[j].z = fmaf(v.z, w.z, acc[j].z); acc[j].w = fmaf(v.w, w.w, acc[j].w);
        }
      }
    }
  } else {
    for (int ky = 0; ky < K; ++ky) {
      const float4* row = sx + ((size_t)(ty + ky) * HW + tx) * (kSlab / 4) + q;
      for (int kx = 0; kx < KW; ++kx) {
        const float4 w = sw[(ky * KW + kx) * (kSlab / 4) + q];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = row[(size_t)(kx + j) * (kSlab / 4)];
          acc[j].x = fmaf(v.x, w.x, acc[j].x); acc[j].y = fmaf(v.y, w.y, acc[j].y);
          acc[j].z = fmaf(v.z, w.z, acc[j].z); acc[j].w = fmaf(v.w, w.w, acc[j].w);
        }
      }
    }
  }
  const int r = r0 + ty;
  if (r < g.Hg) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx + j;
      const long long t = (long long)r * g.Hg + c;
      if (c < g.Hg && t < g.L) reinterpret_cast<float4*>(out + (size_t)t * g.D + c0)[q] = acc[j];
    }
  }
}

int ceil_sqrt_i(long long n) {
  long long r = (long long)floor(sqrt((double)n));
  while (r * r > n) --r;
  while ((r + 1) * (r + 1) <= n) ++r;
  return (int)(r * r == n ? r : r + 1);
}
}  // namespace

size_t peg_scratch_floats(int D, int peg_k, bool ppeg, bool conv_1d) {
  const int K = ppeg && peg_k < 5 ? 5 : peg_k;
  return (size_t)(K * (conv_1d ? 1 : K) + 1) * D;
}

cudaError_t launch_peg(const float* x, float* out, int L, int D, int peg_k, bool ppeg, bool conv_1d,
                       const float* const* w, const float* const* b, float* scratch, cudaStream_t stream) {
  if (D % 4 || L < 1 || peg_k < 1 || peg_k % 2 == 0 || !w[0] || x == out) return cudaErrorInvalidValue;
  if (ppeg && (!w[1] || !w[2])) return cudaErrorInvalidValue;
  PegGeom g;
  g.L = L; g.D = D;
  g.Hn = ceil_sqrt_i(L);
  g.Hg = (ppeg && g.Hn < 7) ? 7 : g.Hn;
  g.K = ppeg && peg_k < 5 ? 5 : peg_k;
  g.KW = conv_1d ? 1 : g.K;
  float* weff = scratch;
  float* beff = scratch + (size_t)g.K * g.KW * D;
  const int n = (g.K * g.KW + 1) * D;
  peg_fold_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w[0], b[0], peg_k, ppeg ? w[1] : nullptr,
                                                       ppeg ? b[1] : nullptr, 5, ppeg ? w[2] : nullptr,
                                                       ppeg ? b[2] : nullptr, 3, conv_1d ? 1 : 0, weff, beff,
                                                       D, g.K, g.KW);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // tiled kernel when the channel count splits into 32-wide slabs and the halo tile fits shared memory;
  // the direct kernel otherwise (RRT_PEG=direct forces it: tuning knob / cross-check)
  const size_t smem = ((size_t)(kTileH + g.K - 1) * (kTileW + g.KW - 1) + (size_t)g.K * g.KW) * kSlab * sizeof(float);
  static const bool direct = [] { const char* e = getenv("RRT_PEG"); return e && !strcmp(e, "direct"); }();
  const int tiles = (g.Hg + kTileH - 1) / kTileH;
  if (!direct && D % kSlab == 0 && smem <= 160 * 1024 && tiles <= 65535 && D / kSlab <= 65535) {
    dim3 grid((g.Hg + kTileW - 1) / kTileW, tiles, D / kSlab);
    auto run = [&](auto kern) -> cudaError_t {
      // once per (kernel instantiation, device): cudaFuncSetAttribute costs ~25 us of host time.
      // MaxShared carve-out: without it the driver keeps a small shared-memory split and ONE 46 KB CTA per SM
      // (measured: 170 us instead of the direct kernel's 95)
      static thread_local const void* seen[16];
      static thread_local int seen_dev[16], n_seen = 0;
      int dev = 0;
      cudaGetDevice(&dev);
      bool done = false;
      for (int i = 0; i < n_seen; ++i) done = done || (seen[i] == (const void*)kern && seen_dev[i] == dev);
      if (!done) {
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (ce == cudaSuccess)
          ce = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared);
        if (ce != cudaSuccess) return ce;
        if (n_seen < 16) { seen[n_seen] = (const void*)kern; seen_dev[n_seen] = dev; ++n_seen; }
      }
      kern<<<grid, 256, smem, stream>>>(x, weff, beff, out, g);
      return cudaGetLastError();
    };
    const int key = g.K * 100 + g.KW;
    switch (key) {
      case 707: return run(peg_apply_tiled_kernel<7, 7>);
      case 505: return run(peg_apply_tiled_kernel<5, 5>);
      case 303: return run(peg_apply_tiled_kernel<3, 3>);
      case 701: return run(peg_apply_tiled_kernel<7, 1>);
      case 501: return run(peg_apply_tiled_kernel<5, 1>);
      default: return run(peg_apply_tiled_kernel<0, 0>);
    }
  }
  long long items = (long long)L * (D / 4);
  long long blocks = (items + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  peg_apply_kernel<<<(int)blocks, 256, 0, stream>>>(x, weff, beff, out, g);
  return cudaGetLastError();
}

// ---- backward (autograd of modules/emb_position.py:36-58,66-82) ------------------------------------------------
// forward:  out[o] = b_eff + sum_tap w_eff[tap] * in(o + off(tap)),  o < L,  in(cell) = x[cell] (cell < L),
//           x[cell - L] (wrap cells L <= cell < Hn^2), 0 otherwise
//   d in(s)      = sum_tap w_eff[tap] * dy[s - off(tap)]            (dy = 0 for cells >= L: the fill is cropped)
//   dx[t]        = d in(t) + d in(L + t) [t < Hn^2 - L]  (+ dres[t])
//   dw_eff[tap]  = sum_{o < L} dy[o] * in(o + off(tap)),   db_eff = sum_o dy[o]
// and every conv of the fold (k; PPEG also 5 and 3) reads ITS taps out of dw_eff; all of them get db_eff.
namespace {
__global__ void __launch_bounds__(256) peg_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ weff,
                                                         const float* __restrict__ dres, float* __restrict__ dx,
                                                         PegGeom g) {
  const int q4 = g.D / 4;
  const long long items = (long long)g.L * q4;
  const int wrap = g.Hn * g.Hn - g.L;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items;
       it += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(it / q4), q = (int)(it - (long long)t * q4);
    float4 acc = dres ? __ldg(reinterpret_cast<const float4*>(dres + (size_t)t * g.D) + q)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int rep = 0; rep < 2; ++rep) {
      if (rep == 1 && t >= wrap) break;
      const int s = rep == 0 ? t : g.L + t;
      const int r = s / g.Hg, c = s - r * g.Hg;
      for (int ky = 0; ky < g.K; ++ky) {
        const int ro = r - (ky - g.K / 2);
        if (ro < 0 || ro >= g.Hg) continue;
        for (int kx = 0; kx < g.KW; ++kx) {
          const int co = g.KW == 1 ? c : c - (kx - g.K / 2);
          if (co < 0 || co >= g.Hg) continue;
          const int o = ro * g.Hg + co;
          if (o >= g.L) continue;
          const float4 v = __ldg(reinterpret_cast<const float4*>(dy + (size_t)o * g.D) + q);
          const float4 w = __ldg(reinterpret_cast<const float4*>(weff + (size_t)(ky * g.KW + kx) * g.D) + q);
          acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
          acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
        }
      }
    }
    reinterpret_cast<float4*>(dx + (size_t)t * g.D)[q] = acc;
  }
}

// grid (K*KW + 1 taps [the last = bias], D / 128 slabs, cell chunks); 256 threads = 8 warps x 32 channel quads
__global__ void __launch_bounds__(256) peg_bwd_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dweff, PegGeom g) {
  __shared__ float4 red[8][32];
  const int tap = blockIdx.x, slab = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = slab * 32 + lane;       // channel quad
  const bool is_bias = tap == g.K * g.KW;
  const int ky = is_bias ? 0 : tap / g.KW, kx = is_bias ? 0 : tap - ky * g.KW;
  const int oy = is_bias ? 0 : ky - g.K / 2, ox = (is_bias || g.KW == 1) ? 0 : kx - g.K / 2;
  const int wrap_end = g.Hn * g.Hn;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q * 4 < g.D) {
    for (int o = blockIdx.z * 8 + warp; o < g.L; o += gridDim.z * 8) {
      const float4 d = __ldg(reinterpret_cast<const float4*>(dy + (size_t)o * g.D) + q);
      float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
      if (!is_bias) {
        const int r = o / g.Hg + oy, c = o % g.Hg + ox;
        if (r < 0 || r >= g.Hg || c < 0 || c >= g.Hg) continue;
        const int cell = r * g.Hg + c;
        const int src = cell < g.L ? cell : (cell < wrap_end ? cell - g.L : -1);
        if (src < 0) continue;
        v = __ldg(reinterpret_cast<const float4*>(x + (size_t)src * g.D) + q);
      }
      acc.x = fmaf(d.x, v.x, acc.x); acc.y = fmaf(d.y, v.y, acc.y);
      acc.z = fmaf(d.z, v.z, acc.z); acc.w = fmaf(d.w, v.w, acc.w);
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && q * 4 < g.D) {
    float4 s = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) { s.x += red[w][lane].x; s.y += red[w][lane].y; s.z += red[w][lane].z; s.w += red[w][lane].w; }
    float* dst = dweff + (size_t)tap * g.D + q * 4;
    atomicAdd(dst, s.x); atomicAdd(dst + 1, s.y); atomicAdd(dst + 2, s.z); atomicAdd(dst + 3, s.w);
  }
}

// dw_j [D, 1, k_j, kw_j] (reference layout) and db_j [D] from dw_eff [K*KW + 1][D]
__global__ void __launch_bounds__(256) peg_bwd_unfold_kernel(const float* __restrict__ dweff, float* __restrict__ dw,
                                                             float* __restrict__ db, int k, int conv_1d, int D, int K,
                                                             int KW) {
  const int kw = conv_1d ? 1 : k, n = D * k * kw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + (db ? D : 0); i += gridDim.x * blockDim.x) {
    if (i >= n) { db[i - n] = dweff[(size_t)K * KW * D + (i - n)]; continue; }
    const int c = i / (k * kw), y = (i / kw) % k, x = i % kw;
    const int ky = y - k / 2 + K / 2, kx = conv_1d ? 0 : x - k / 2 + K / 2;
    dw[i] = dweff[(size_t)(ky * KW + kx) * D + c];
  }
}
}  // namespace

// x: the forward's input [L, D]; dy: gradient wrt the forward's output; dres (nullable): added to dx (the
// all_shortcut gradient); weff: the forward's folded kernel (the scratch of launch_peg, untouched since);
// scratch: peg_scratch_floats floats.  dw[j] / db[j]: gradients in the reference's parameter layout (db nullable).
cudaError_t launch_peg_backward(const float* x, const float* dy, const float* dres, float* dx, int L, int D,
                                int peg_k, bool ppeg, bool conv_1d, const float* weff, float* scratch,
                                float* const* dw, float* const* db, cudaStream_t stream) {
  if (D % 4 || L < 1 || peg_k < 1 || peg_k % 2 == 0 || !dw[0] || !weff || !scratch) return cudaErrorInvalidValue;
  if (ppeg && (!dw[1] || !dw[2])) return cudaErrorInvalidValue;
  PegGeom g;
  g.L = L; g.D = D;
  g.Hn = ceil_sqrt_i(L);
  g.Hg = (ppeg && g.Hn < 7) ? 7 : g.Hn;
  if (g.Hg != g.Hn) return cudaErrorNotSupported;   // PPEG's zero extension of grids below 7 x 7 (L < 37)
  g.K = ppeg && peg_k < 5 ? 5 : peg_k;
  g.KW = conv_1d ? 1 : g.K;
  long long items = (long long)L * (D / 4);
  long long blocks = (items + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  peg_bwd_dx_kernel<<<(int)blocks, 256, 0, stream>>>(dy, weff, dres, dx, g);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int taps = g.K * g.KW + 1;
  e = cudaMemsetAsync(scratch, 0, (size_t)taps * D * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  int zc = (L + 8 * 32 - 1) / (8 * 32);
  if (zc > 64) zc = 64;
  if (zc < 1) zc = 1;
  peg_bwd_dw_kernel<<<dim3(taps, (D + 127) / 128, zc), 256, 0, stream>>>(x, dy, scratch, g);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int ks[3] = {peg_k, 5, 3};
  for (int j = 0; j < (ppeg ? 3 : 1); ++j) {
    const int n = D * ks[j] * (conv_1d ? 1 : ks[j]) + D;
    peg_bwd_unfold_kernel<<<(n + 255) / 256, 256, 0, stream>>>(scratch, dw[j], db ? db[j] : nullptr, ks[j],
                                                               conv_1d ? 1 : 0, D, g.K, g.KW);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace rrt
