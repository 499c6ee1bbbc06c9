// Shared device/host helpers for the RRTEncoder kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace rrt {

// One-time per-device kernel configuration (cudaFuncSetAttribute is slow, ~25 us, and function
// attributes are per device): `flags` is a static array owned by the call site.
struct DeviceOnce {
  bool done[32] = {};
  bool needed() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 31;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// Tuning knob RRT_CARVEOUT=max: the streaming kernels ask for the same L1 / shared-memory split as the
// tensor-core kernels (maximum shared memory), so that an SM never has to drain to change its carve-out
// between back-to-back kernels.  Measured (s20): 73.1 us/bag with it vs 71.4 without (16 bags, 4 lanes);
// ln_partition 13.5 -> 15.0 us, dispatch 13.3 -> 14.6 us -- the smaller L1 costs the streaming kernels more
// than the reconfiguration saves, so it is OFF by default.
inline void prefer_max_shared_impl(const void* kernel) {
  static const bool on = [] { const char* e = getenv("RRT_CARVEOUT"); return e && !strcmp(e, "max"); }();
  if (!on) return;
  // (kernel, device) pairs already configured; kernels are few, a linear scan is cheaper than a hash
  static thread_local const void* seen[128];
  static thread_local int seen_dev[128];
  static thread_local int n_seen = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < n_seen; ++i)
    if (seen[i] == kernel && seen_dev[i] == dev) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (n_seen < 128) { seen[n_seen] = kernel; seen_dev[n_seen] = dev; ++n_seen; }
}
template <typename K>
inline void prefer_max_shared(K kernel) {
  prefer_max_shared_impl(reinterpret_cast<const void*>(kernel));
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// The kernels of one bag form a strict chain on one stream; between two of them the GPU otherwise idles for
// the grid-launch latency plus the next kernel's prologue (~2-3 us per boundary, ten boundaries per bag).
// Every chain kernel therefore (1) signals `launch_dependents` on entry, so that the NEXT kernel's CTAs may be
// scheduled as soon as all of this kernel's CTAs have started, and (2) executes `griddepcontrol.wait` before
// its first access to global memory that a predecessor may have written or may still read.  `wait` returns
// only when the predecessor grid has COMPLETED and flushed; since the predecessor itself waited for its own
// predecessor, everything earlier in the stream is complete too, so all RAW / WAR / WAW orderings of plain
// stream order are kept.  Both instructions are no-ops for a kernel launched without the attribute.
// The attribute is only set while ONE bag runs at a time (g_pdl): with several bags in flight a pre-launched
// CTA would sit on an SM (the GEMMs take a whole one) that another bag's kernels could be using.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
extern thread_local bool g_pdl;       // api.cu: every chain kernel
extern thread_local bool g_pdl_light;  // api.cu: the streaming / attention kernels only (not the whole-SM GEMMs)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl || g_pdl_light) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default (modules/rrt.py:47,139)
#define RRT_MAX_K_DEV 16  // == RRT_MAX_CRMSA_K in include/rrt_b200.h

// Padded square grid of one bag (modules/rmsa.py:175-198).  Slot s = rho*P + p is the
// region-major position of grid cell (row, col); token t = row*H + col; t >= L is padding.
struct Grid {
  int L;   // real tokens
  int H;   // grid side
  int rs;  // region side
  int g;   // regions per side = H / rs
  int P;   // tokens per region = rs*rs
  int R;   // regions = g*g
  int Np;  // padded tokens = H*H

  __host__ __device__ __forceinline__ int slot_to_token(int s) const {
    int rho = s / P, p = s - rho * P;
    int rr = rho / g, rc = rho - rr * g;
    int pr = p / rs, pc = p - pr * rs;
    return (rr * rs + pr) * H + rc * rs + pc;
  }
  __host__ __device__ __forceinline__ int token_to_slot(int t) const {
    int row = t / H, col = t - row * H;
    int rr = row / rs, pr = row - rr * rs;
    int rc = col / rs, pc = col - rc * rs;
    return (rr * g + rc) * P + pr * rs + pc;
  }
};

// proj_drop of InnerAttention in training mode (modules/rmsa.py:70,132): counter-based, so the
// backward pass regenerates the mask of the forward instead of storing it.  One splitmix64 hash per
// group of 4 consecutive elements (flat index of the [rows, D] tensor the dropout acts on, token
// order), 16 bits per element: an element is dropped when its 16-bit field is < thresh, survivors
// are scaled by 1/(1-p).  oracle/rrt_oracle.py::dropout_mask restates the same rule in numpy.
struct Dropout {
  unsigned long long key = 0;  // seed + stream * golden ratio (host side: dropout_make)
  uint32_t thresh = 0;         // round(p * 65536); 0 = off
  float scale = 1.f;           // 1 / (1 - p)
  // optional: a per-step seed that lives in DEVICE memory and is ADDED to `key` when the mask is evaluated, so
  // that a captured CUDA graph of a training step draws a new mask on every replay (rrt_set_step_state)
  const unsigned long long* seed_dev = nullptr;
  // thresh == 0 with scale != 1: nothing is dropped, everything is scaled (stochastic depth's 1/keep on a branch)
  __host__ __device__ bool on() const { return thresh != 0 || scale != 1.f; }
};
// api.cu (rrt_set_step_state); null = off.  PROCESS-wide, not per thread: torch runs the backward of a step on its
// autograd thread, and forward and backward must evaluate the same masks
extern const unsigned long long* volatile g_step_seed_dev;
inline Dropout dropout_make(float p, unsigned long long seed, unsigned stream_id) {
  Dropout d;
  if (p > 0.f) {
    double t = (double)p * 65536.0 + 0.5;
    d.thresh = t >= 65535.0 ? 65535u : (uint32_t)t;
    d.scale = 1.f / (1.f - p);
    d.key = seed + (unsigned long long)(stream_id + 1) * 0x9E3779B97F4A7C15ull;
    d.seed_dev = g_step_seed_dev;
  }
  return d;
}
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// keep-and-scale factors of the 4 elements at flat indices idx .. idx+3 (idx % 4 == 0)
__device__ __forceinline__ float4 dropout_scale4(const Dropout& d, unsigned long long idx) {
  const unsigned long long key = d.seed_dev ? d.key + __ldg(d.seed_dev) : d.key;
  const unsigned long long h = splitmix64((idx >> 2) + key);
  float4 m;
  m.x = (uint32_t)(h & 0xffffu) >= d.thresh ? d.scale : 0.f;
  m.y = (uint32_t)((h >> 16) & 0xffffu) >= d.thresh ? d.scale : 0.f;
  m.z = (uint32_t)((h >> 32) & 0xffffu) >= d.thresh ? d.scale : 0.f;
  m.w = (uint32_t)(h >> 48) >= d.thresh ? d.scale : 0.f;
  return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// round-to-nearest fp32 -> tf32 (kept in an fp32 container)
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t tf32_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// two fp32 -> packed fp16x2 (lo in bits 0..15), round to nearest, saturating to +-65504 instead of
// overflowing to inf.  fp16 keeps 10 mantissa bits, the same as tf32: it is the storage format of
// every INTERNAL activation (LayerNorm output, q/k/v, attention output, landmarks) and of the
// weight shadows the tensor cores read; the residual stream, statistics and accumulators are fp32.
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint2 pack_h4(float4 v) {
  return make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ float4 unpack_h4(uint2 u) {
  float2 a = unpack_h2(u.x), b = unpack_h2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// 2^x on the SFU in one instruction (exp2f() wraps MUFU.EX2 in three range fix-up instructions);
// softmax arguments are <= 0 and results below 2^-126 may flush to zero
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// nn.GELU() (erf form) and its derivative  Phi(z) + z phi(z)
__device__ __forceinline__ float gelu_fwd(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float z) {
  return 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
}

// D(16x8,f32) += A(16x8,tf32,row) * B(8x8,tf32,col)
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4],
                                                const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

}  // namespace rrt
