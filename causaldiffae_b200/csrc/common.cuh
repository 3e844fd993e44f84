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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cdae
