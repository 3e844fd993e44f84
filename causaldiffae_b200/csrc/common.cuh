// Shared helpers for libcdae (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cdae.h"

namespace cdae {

void set_error(const char* fmt, ...);

#define CDAE_CHECK_ARG(cond, ...)                                   \
  do { if (!(cond)) { cdae::set_error(__VA_ARGS__); return CDAE_ERR_ARG; } } while (0)
#define CDAE_CHECK_SHAPE(cond, ...)                                 \
  do { if (!(cond)) { cdae::set_error(__VA_ARGS__); return CDAE_ERR_SHAPE; } } while (0)
#define CDAE_CHECK_LAUNCH(name)                                     \
  do { cudaError_t e_ = cudaGetLastError();                         \
       if (e_ != cudaSuccess) { cdae::set_error("%s: %s", name, cudaGetErrorString(e_)); return CDAE_ERR_CUDA; } } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;

// ---- device helpers
// MUFU.EX2 + MUFU.RCP (2 ulp): the results are rounded to bf16 right after, full-precision division buys nothing
__device__ __forceinline__ float sigmoid_f(float u) { return __fdividef(1.0f, 1.0f + __expf(-u)); }
__device__ __forceinline__ float silu_f(float u) { return u * sigmoid_f(u); }

struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(p.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
__device__ __forceinline__ bf16x8 ld8(const void* p) { return *reinterpret_cast<const bf16x8*>(p); }
__device__ __forceinline__ void st8(void* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

// SiLU through one MUFU: with h = u/2, silu(u) = h + h*tanh(h) and silu'(u) = (1 + t + h*(1 - t^2)) / 2, t = tanh(h)
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 tanh2(float2 h) { return make_float2(tanh_fast(h.x), tanh_fast(h.y)); }
__device__ __forceinline__ float2 dsilu_half(float2 h) {
  const float2 kM1 = make_float2(-1.f, -1.f), kHalf = make_float2(0.5f, 0.5f);
  const float2 t = tanh2(h);
  const float2 q = __ffma2_rn(t, t, kM1);                        // t^2 - 1
  const float2 w = __ffma2_rn(__fmul2_rn(h, kM1), q, t);         // t + h*(1 - t^2)
  return __ffma2_rn(w, kHalf, kHalf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cdae
