// Shared device/host helpers for libmop_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mop_b200.h"

namespace mop {

// ---- error plumbing (thread local, no exceptions across the ABI) -----------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MOP_CHECK_CUDA(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::mop::cuda_fail(_e, #expr); \
  } while (0)

#define MOP_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      ::mop::set_error(__VA_ARGS__); \
      return (code);                 \
    }                                \
  } while (0)

int sm_count();

// ---- dtype helpers -----------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

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
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// tanh-approximated GELU exactly as torch: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
__device__ __forceinline__ float gelu_tanh_(float x) {
  const float c = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanhf(c * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float dgelu_tanh_(float x) {
  const float c = 0.7978845608028654f;
  float t = tanhf(c * (x + 0.044715f * x * x * x));
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * c * (1.0f + 3.0f * 0.044715f * x * x);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- attention dropout: counter-based keep mask ----------------------------------------------------------------------------
// keep(b, h, i, j) is a pure function of (seed, offset, b*H + h, i, j): every kernel (forward, dQ, dK/dV, tensor-core or
// fp32-math) regenerates the same mask from the two 64-bit numbers the caller passes, nothing is stored.  A 32-bit avalanche
// hash (xorshift-multiply, "lowbias32") is applied to a per-row key and the column; P(keep) = 1 - p to 2^-32.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
struct Dropout {
  uint32_t thresh;   // keep iff hash >= thresh
  uint32_t key;
  float inv_keep;    // 1 / (1 - p); 0 for p >= 1
  int on;
};
__host__ __device__ inline Dropout make_dropout(float p, uint64_t seed, uint64_t offset) {
  Dropout d;
  d.on = p > 0.f ? 1 : 0;
  const double t = (double)p * 4294967296.0;
  d.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  d.inv_keep = p < 1.f ? 1.f / (1.f - p) : 0.f;
  d.key = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) ^ mix32((uint32_t)offset + 0x9E3779B9u * (uint32_t)(offset >> 32) + 0x7F4A7C15u)));
  return d;
}
__host__ __device__ __forceinline__ uint32_t dropout_row_key(const Dropout& d, uint32_t bh, uint32_t i) {
  return mix32(mix32(d.key ^ (bh * 0x85EBCA6Bu)) + i * 0xC2B2AE35u);
}
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t row_key, uint32_t j, uint32_t thresh) {
  return mix32(row_key ^ (j * 0x9E3779B9u)) >= thresh;
}
// the factor an attention probability is multiplied by: 0 or 1 / (1 - p)
__host__ __device__ __forceinline__ float dropout_factor(const Dropout& d, uint32_t row_key, uint32_t j) {
  return dropout_keep(row_key, j, d.thresh) ? d.inv_keep : 0.f;
}

}  // namespace mop
