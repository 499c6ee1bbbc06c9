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
  long long items = (long long)L * (D / 4);
  long long blocks = (items + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  peg_apply_kernel<<<(int)blocks, 256, 0, stream>>>(x, weff, beff, out, g);
  return cudaGetLastError();
}

}  // namespace rrt
