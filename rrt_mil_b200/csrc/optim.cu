// Optimizer step of the training harness: torch.optim.Adam / AdamW semantics
// (main.py:224-233: Adam, lr 2e-4, weight_decay 1e-5) for a LIST of parameter tensors in ONE launch.
// The tensors' pointers travel in the kernel argument (multi-tensor apply): CTA b owns 4096 consecutive
// elements of the concatenated index space and finds its tensor in the prefix table.  Pure streaming:
// 16 B read of p, g, m, v and 16 B write of p, m, v per 4 elements -> HBM-bound (28 B / element).
#include "kernels.cuh"

namespace rrt {
namespace {
constexpr int kMaxTensors = 48;
constexpr int kChunk = 4096;  // elements per CTA (256 threads x 4 float4)

struct AdamTable {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  long long n[kMaxTensors];
  int chunk_begin[kMaxTensors + 1];  // first CTA of tensor i
  int count;
};

struct AdamHyper {
  float lr, beta1, beta2, eps, wd, grad_scale;
  float bc1, bc2_rsqrt;  // 1 - beta1^t,  1 / sqrt(1 - beta2^t)
  int decoupled;         // AdamW: p *= 1 - lr*wd instead of g += wd*p
  const float* bc_dev;   // optional {bc1, bc2_rsqrt} in device memory (a replayed CUDA graph: the step number changes)
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamHyper& h) {
  g *= h.grad_scale;
  if (h.decoupled) p *= 1.f - h.lr * h.wd;
  else g = fmaf(h.wd, p, g);
  m = fmaf(h.beta1, m, (1.f - h.beta1) * g);
  v = fmaf(h.beta2, v, (1.f - h.beta2) * g * g);
  const float denom = sqrtf(v) * h.bc2_rsqrt + h.eps;
  p -= (h.lr / h.bc1) * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable tb, AdamHyper h) {
  if (h.bc_dev) { h.bc1 = __ldg(h.bc_dev); h.bc2_rsqrt = __ldg(h.bc_dev + 1); }
  int t = 0;
  while (t + 1 < tb.count && (int)blockIdx.x >= tb.chunk_begin[t + 1]) ++t;
  const long long n = tb.n[t];
  const long long base = (long long)(blockIdx.x - tb.chunk_begin[t]) * kChunk;
  float* __restrict__ p = tb.p[t];
  const float* __restrict__ g = tb.g[t];
  float* __restrict__ m = tb.m[t];
  float* __restrict__ v = tb.v[t];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
#pragma unroll
  for (int j = 0; j < kChunk / 1024; ++j) {
    const long long i = base + j * 1024 + threadIdx.x * 4;
    if (i >= n) break;
    if (vec && i + 4 <= n) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
             vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i));
      adam1(pp.x, gg.x, mm.x, vv.x, h); adam1(pp.y, gg.y, mm.y, vv.y, h);
      adam1(pp.z, gg.z, mm.z, vv.z, h); adam1(pp.w, gg.w, mm.w, vv.w, h);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (long long e = i; e < n && e < i + 4; ++e) {
        float pp = p[e], mm = m[e], vv = v[e];
        adam1(pp, g[e], mm, vv, h);
        p[e] = pp; m[e] = mm; v[e] = vv;
      }
    }
  }
}
}  // namespace

cudaError_t launch_adam(float* const* p, const float* const* g, float* const* m, float* const* v,
                        const long long* n, int count, float lr, float beta1, float beta2, float eps,
                        float wd, bool decoupled, long long step, float grad_scale, int* launches,
                        cudaStream_t stream, const float* bc_dev) {
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.wd = wd; h.grad_scale = grad_scale;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_rsqrt = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step)));
  h.decoupled = decoupled ? 1 : 0;
  h.bc_dev = bc_dev;
  *launches = 0;
  for (int first = 0; first < count; first += kMaxTensors) {
    AdamTable tb;
    tb.count = 0;
    int chunks = 0;
    for (int i = first; i < count && tb.count < kMaxTensors; ++i) {
      if (n[i] <= 0) continue;
      const int k = tb.count++;
      tb.p[k] = p[i]; tb.g[k] = g[i]; tb.m[k] = m[i]; tb.v[k] = v[i]; tb.n[k] = n[i];
      tb.chunk_begin[k] = chunks;
      chunks += (int)((n[i] + kChunk - 1) / kChunk);
    }
    tb.chunk_begin[tb.count] = chunks;
    if (chunks == 0) continue;
    adam_kernel<<<chunks, 256, 0, stream>>>(tb, h);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  return cudaSuccess;
}

}  // namespace rrt
