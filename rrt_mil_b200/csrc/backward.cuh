// Backward-pass conventions shared by the gradient kernels.
//
// Gradient tensors that feed tensor cores are fp16 (10 mantissa bits, as the forward operands), so
// they are kept in a SCALED domain: every stage records amax = max|g| of its fp32 input gradient
// (atomicMax on the IEEE bits of |g|) and the consumers derive S = 2^e with amax*S in [2^8, 2^9)
// from that word -- a power of two, so scaling is exact and the backward stays linear in the
// upstream gradient.  fp32 outputs (residual-stream gradients, parameter gradients) are unscaled
// with 1/S where they are produced.  This is per-stage automatic loss scaling; nothing to tune.
#pragma once
#include "common.cuh"

namespace rrt {

__host__ __device__ __forceinline__ uint32_t scale_exp_from_amax(uint32_t bits) {
  uint32_t eb = (bits & 0x7fffffffu) >> 23;
  if (eb == 0u || eb >= 255u) return 135u;  // zero / denormal / inf / nan amax: S = 1
  return eb < 20u ? 20u : (eb > 240u ? 240u : eb);
}
// S: amax * S in [256, 512)
__device__ __forceinline__ float grad_scale(uint32_t amax_bits) {
  return __uint_as_float((262u - scale_exp_from_amax(amax_bits)) << 23);
}
__device__ __forceinline__ float grad_inv_scale(uint32_t amax_bits) {
  return __uint_as_float((scale_exp_from_amax(amax_bits) - 8u) << 23);
}
__device__ __forceinline__ void atomic_amax(uint32_t* slot, float v) {
  atomicMax(slot, __float_as_uint(fabsf(v)));
}

}  // namespace rrt
