// Fused HBM-bound elementwise / reduction kernels of the diffusion process and the optimizer.
// 128-bit vectorised, grid sized in multiples of the SM count, fp32 arithmetic.
#include "common.cuh"

namespace cdae {

constexpr int kEwThreads = 256;
static inline int ew_grid(int64_t nvec) {
  int64_t blocks = ceil_div(nvec, kEwThreads);
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// ---------------------------------------------------------------- q_sample (ref gaussian_diffusion.py:201-222)
__global__ void __launch_bounds__(kEwThreads) q_sample_kernel(const float4* __restrict__ x0, const float4* __restrict__ nz,
                                                              const int64_t* __restrict__ t, const float* __restrict__ ta,
                                                              const float* __restrict__ tb, float4* __restrict__ out,
                                                              int64_t nvec, int64_t vec_per_sample) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / vec_per_sample;
    const int64_t tt = t[b];
    const float a = __ldg(ta + tt), c = __ldg(tb + tt);
    const float4 x = x0[i], n = nz[i];
    // a*x + c*n with separately rounded products: same association as the reference expression
    out[i] = make_float4(__fadd_rn(__fmul_rn(a, x.x), __fmul_rn(c, n.x)), __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(c, n.y)),
                         __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(c, n.z)), __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(c, n.w)));
  }
}

// ---------------------------------------------------------------- per-sample MSE + gradient (ref gaussian_diffusion.py:847)
__global__ void __launch_bounds__(kEwThreads) mse_kernel(const float4* __restrict__ pred, const float4* __restrict__ tgt,
                                                         float* __restrict__ mse, const float* __restrict__ gscale,
                                                         float4* __restrict__ dpred, int64_t vec_per_sample,
                                                         float inv_per_sample, float gmul) {
  const int64_t b = blockIdx.x;
  const float4* p = pred + b * vec_per_sample;
  const float4* q = tgt + b * vec_per_sample;
  const float gs = (dpred && gscale) ? 2.0f * (gscale[b] * gmul) * inv_per_sample : 0.f;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < vec_per_sample; i += blockDim.x) {
    const float4 a = p[i], c = q[i];
    const float dx = a.x - c.x, dy = a.y - c.y, dz = a.z - c.z, dw = a.w - c.w;
    acc += dx * dx + dy * dy + dz * dz + dw * dw;
    if (dpred) dpred[b * vec_per_sample + i] = make_float4(gs * dx, gs * dy, gs * dz, gs * dw);
  }
  __shared__ float red[kEwThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kEwThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) mse[b] = v * inv_per_sample;
  }
}

// ---------------------------------------------------------------- DDIM step (ref gaussian_diffusion.py:277-285,320-341,506-558)
// coef row (8 floats): {sqrt_recip_ac, sqrt_recipm1_ac, sqrt(ac_prev), sqrt(1-ac_prev-sigma^2), sigma*[t!=0], clip, xstart_mode, 0}
__global__ void __launch_bounds__(kEwThreads) ddim_kernel(const float4* __restrict__ x, const float4* __restrict__ ec,
                                                          const float4* __restrict__ eu, float w, int use_w,
                                                          const float* __restrict__ coef, const int32_t* __restrict__ tidx,
                                                          int tidx_stride, const float4* __restrict__ noise,
                                                          float4* __restrict__ xprev, float4* __restrict__ x0out,
                                                          int64_t nvec, int64_t vec_per_sample) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / vec_per_sample;
    const float* cr = coef + 8 * (int64_t)tidx[b * tidx_stride];
    const float r = cr[0], s = cr[1], sap = cr[2], dir = cr[3], sg = cr[4];
    const bool clip = cr[5] != 0.f, xs_mode = cr[6] != 0.f;
    const float4 xv = x[i];
    float4 e = ec[i];
    if (use_w) {
      const float4 u = eu[i];
      const float w1 = 1.f - w;
      e = make_float4(__fadd_rn(__fmul_rn(w, e.x), __fmul_rn(w1, u.x)), __fadd_rn(__fmul_rn(w, e.y), __fmul_rn(w1, u.y)),
                      __fadd_rn(__fmul_rn(w, e.z), __fmul_rn(w1, u.z)), __fadd_rn(__fmul_rn(w, e.w), __fmul_rn(w1, u.w)));
    }
    float xin[4] = {xv.x, xv.y, xv.z, xv.w}, ein[4] = {e.x, e.y, e.z, e.w}, o[4], x0v[4];
    float nzv[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise && sg != 0.f) { const float4 n = noise[i]; nzv[0] = n.x; nzv[1] = n.y; nzv[2] = n.z; nzv[3] = n.w; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float rx = __fmul_rn(r, xin[k]);
      float x0 = xs_mode ? ein[k] : __fsub_rn(rx, __fmul_rn(s, ein[k]));
      if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
      const float e2 = __fdiv_rn(__fsub_rn(rx, x0), s);
      float m = __fadd_rn(__fmul_rn(x0, sap), __fmul_rn(dir, e2));
      m = __fadd_rn(m, __fmul_rn(sg, nzv[k]));
      o[k] = m; x0v[k] = x0;
    }
    xprev[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (x0out) x0out[i] = make_float4(x0v[0], x0v[1], x0v[2], x0v[3]);
  }
}

// ---------------------------------------------------------------- AdamW + EMA + sum(g^2)  (ref train_util.py:276-303, nn.py:503-513)
// hyper (device): {lr, beta1, beta2, eps, weight_decay, ema_rate, grad_scale}; *step = optimizer steps taken so far (the
// bias corrections are those of step + 1, evaluated in double like torch.optim.AdamW does on the host); guard (optional):
// a device scalar that must be finite for the step to happen (sum of squared gradients, cdae_sumsq) - otherwise the whole
// launch is a no-op, which is the "skip the step" of the reference's optimize_fp16.  G = float (fp32 gradient arena) or
// __nv_bfloat16 (the all-reduced bf16 copy of it).
__device__ __forceinline__ float4 ldg4(const float4* g, int64_t i) { return g[i]; }
__device__ __forceinline__ float4 ldg4(const uint2* g, int64_t i) {
  const uint2 u = g[i];
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}

template <typename GV>
__global__ void __launch_bounds__(kEwThreads) adam_ema_kernel(float4* __restrict__ p, const GV* __restrict__ g,
                                                              float4* __restrict__ m, float4* __restrict__ v,
                                                              float4* __restrict__ ema, const float* __restrict__ hyper,
                                                              const int64_t* __restrict__ step,
                                                              const float* __restrict__ guard,
                                                              float* __restrict__ gsq_out, int64_t nvec) {
  if (guard && !isfinite(*guard)) return;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], er = hyper[5], gscale = hyper[6];
  const double t = (double)(*step + 1);
  const float step_size = (float)((double)lr / (1.0 - pow((double)b1, t)));
  const float bc2s = (float)sqrt(1.0 - pow((double)b2, t));
  float gsq = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = p[i], gv = ldg4(g, i), mv = m[i], vv = v[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w},
          v2[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k] * gscale;
      gsq += gk * gk;
      pp[k] *= 1.f - lr * wd;
      mm[k] = mm[k] + (gk - mm[k]) * (1.f - b1);
      v2[k] = v2[k] * b2 + (1.f - b2) * gk * gk;
      const float denom = sqrtf(v2[k]) / bc2s + eps;
      pp[k] -= step_size * (mm[k] / denom);
    }
    p[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    m[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    v[i] = make_float4(v2[0], v2[1], v2[2], v2[3]);
    if (ema) {
      float4 ev = ema[i];
      ev.x = ev.x * er + pp[0] * (1.f - er); ev.y = ev.y * er + pp[1] * (1.f - er);
      ev.z = ev.z * er + pp[2] * (1.f - er); ev.w = ev.w * er + pp[3] * (1.f - er);
      ema[i] = ev;
    }
  }
  if (gsq_out) {
    __shared__ float red[kEwThreads / 32];
    gsq = warp_sum(gsq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gsq;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t2 = threadIdx.x < kEwThreads / 32 ? red[threadIdx.x] : 0.f;
      t2 = warp_sum(t2);
      if (threadIdx.x == 0) atomicAdd(gsq_out, t2);
    }
  }
}

// runs after the update: the step counter moves on unless the guard vetoed the step
__global__ void adam_tick_kernel(int64_t* step, const float* guard, const float* gsq, float* lognorm) {
  if (!(guard && !isfinite(*guard))) *step += 1;
  if (lognorm && gsq) { lognorm[0] += sqrtf(*gsq); lognorm[1] += 1.f; }     // the logger's running mean of the gradient norm
}

// sum of squares of a flat fp32 buffer accumulated into *out (the non-finite-gradient guard; also a grad-norm probe)
template <typename GV>
__global__ void __launch_bounds__(kEwThreads) sumsq_kernel(const GV* __restrict__ g, float* __restrict__ out, int64_t nvec) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = ldg4(g, i);
    acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  __shared__ float red[kEwThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < kEwThreads / 32 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// fp32 -> bf16 (round to nearest even) copy of a flat buffer: the gradient arena as it goes on the wire
__global__ void __launch_bounds__(kEwThreads) cast_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst,
                                                               int64_t nvec) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = src[i];
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
    dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  }
}

__global__ void __launch_bounds__(kEwThreads) ema_kernel(float4* __restrict__ ema, const float4* __restrict__ p, float er,
                                                         int64_t nvec) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 e = ema[i]; const float4 q = p[i];
    e.x = e.x * er + q.x * (1.f - er); e.y = e.y * er + q.y * (1.f - er);
    e.z = e.z * er + q.z * (1.f - er); e.w = e.w * er + q.w * (1.f - er);
    ema[i] = e;
  }
}

}  // namespace cdae
using namespace cdae;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int cdae_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                             const float* sqrt_1mac, float* x_t, int64_t B, int64_t per_sample, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(x0 && noise && t && sqrt_ac && sqrt_1mac && x_t, "q_sample: null pointer");
  CDAE_CHECK_SHAPE(per_sample % 4 == 0 && aligned16(x0) && aligned16(noise) && aligned16(x_t),
                   "q_sample: per_sample %% 4 and 16-byte alignment required");
  const int64_t nvec = B * per_sample / 4;
  q_sample_kernel<<<ew_grid(nvec), kEwThreads, 0, (cudaStream_t)s>>>((const float4*)x0, (const float4*)noise, t, sqrt_ac,
                                                                    sqrt_1mac, (float4*)x_t, nvec, per_sample / 4);
  CDAE_CHECK_LAUNCH("q_sample_kernel");
  return CDAE_OK;
}

extern "C" int cdae_mse_loss(const float* pred, const float* target, float* mse, const float* gscale, float gmul,
                             float* dpred, int64_t B, int64_t per_sample, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(pred && target && mse, "mse_loss: null pointer");
  CDAE_CHECK_SHAPE(per_sample % 4 == 0 && aligned16(pred) && aligned16(target) && (!dpred || aligned16(dpred)),
                   "mse_loss: per_sample %% 4 and 16-byte alignment required");
  mse_kernel<<<(unsigned)B, kEwThreads, 0, (cudaStream_t)s>>>((const float4*)pred, (const float4*)target, mse, gscale,
                                                               (float4*)dpred, per_sample / 4, 1.0f / (float)per_sample, gmul);
  CDAE_CHECK_LAUNCH("mse_kernel");
  return CDAE_OK;
}

extern "C" int cdae_ddim_step(const float* x, const float* eps_c, const float* eps_u, float w, int use_w,
                              const float* coef_table, const int32_t* t_idx, int t_idx_stride, const float* noise,
                              float* x_prev, float* pred_xstart, int64_t B, int64_t per_sample, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(x && eps_c && coef_table && t_idx && x_prev && (!use_w || eps_u), "ddim_step: null pointer");
  CDAE_CHECK_SHAPE(per_sample % 4 == 0 && aligned16(x) && aligned16(eps_c) && aligned16(x_prev),
                   "ddim_step: per_sample %% 4 and 16-byte alignment required");
  const int64_t nvec = B * per_sample / 4;
  ddim_kernel<<<ew_grid(nvec), kEwThreads, 0, (cudaStream_t)s>>>(
      (const float4*)x, (const float4*)eps_c, (const float4*)eps_u, w, use_w, coef_table, t_idx, t_idx_stride,
      (const float4*)noise, (float4*)x_prev, (float4*)pred_xstart, nvec, per_sample / 4);
  CDAE_CHECK_LAUNCH("ddim_kernel");
  return CDAE_OK;
}

extern "C" int cdae_adam_ema(float* p, const void* g, int g_is_bf16, float* m, float* v, float* ema, const float* hyper,
                             int64_t* step, const float* guard, float* gsq_out, float* lognorm, int64_t n, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(p && g && m && v && hyper && step, "adam_ema: null pointer");
  CDAE_CHECK_SHAPE(n % 4 == 0 && aligned16(p) && aligned16(m) && aligned16(v) && (!ema || aligned16(ema)) &&
                       (reinterpret_cast<uintptr_t>(g) & (g_is_bf16 ? 7 : 15)) == 0,
                   "adam_ema: n %% 4 and 16-byte alignment required (pad the arena)");
  if (g_is_bf16)
    adam_ema_kernel<uint2><<<ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)s>>>((float4*)p, (const uint2*)g, (float4*)m,
                                                                              (float4*)v, (float4*)ema, hyper, step, guard,
                                                                              gsq_out, n / 4);
  else
    adam_ema_kernel<float4><<<ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)s>>>((float4*)p, (const float4*)g, (float4*)m,
                                                                               (float4*)v, (float4*)ema, hyper, step, guard,
                                                                               gsq_out, n / 4);
  CDAE_CHECK_LAUNCH("adam_ema_kernel");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)s>>>(step, guard, gsq_out, lognorm);
  CDAE_CHECK_LAUNCH("adam_tick_kernel");
  return CDAE_OK;
}

extern "C" int cdae_sumsq(const void* g, int g_is_bf16, float* out, int64_t n, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(g && out, "sumsq: null pointer");
  CDAE_CHECK_SHAPE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & (g_is_bf16 ? 7 : 15)) == 0,
                   "sumsq: n %% 4 and 16-byte alignment required");
  if (g_is_bf16) sumsq_kernel<uint2><<<ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)s>>>((const uint2*)g, out, n / 4);
  else sumsq_kernel<float4><<<ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)s>>>((const float4*)g, out, n / 4);
  CDAE_CHECK_LAUNCH("sumsq_kernel");
  return CDAE_OK;
}

extern "C" int cdae_cast_bf16(const float* src, void* dst_bf16, int64_t n, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(src && dst_bf16, "cast_bf16: null pointer");
  CDAE_CHECK_SHAPE(n % 4 == 0 && aligned16(src) && (reinterpret_cast<uintptr_t>(dst_bf16) & 7) == 0,
                   "cast_bf16: n %% 4 and alignment required");
  cast_bf16_kernel<<<ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)s>>>((const float4*)src, (uint2*)dst_bf16, n / 4);
  CDAE_CHECK_LAUNCH("cast_bf16_kernel");
  return CDAE_OK;
}

extern "C" int cdae_ema_update(float* ema, const float* p, float rate, int64_t n, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(ema && p, "ema_update: null pointer");
  CDAE_CHECK_SHAPE(n % 4 == 0 && aligned16(ema) && aligned16(p), "ema_update: n %% 4 and alignment");
  ema_kernel<<<ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)s>>>((float4*)ema, (const float4*)p, rate, n / 4);
  CDAE_CHECK_LAUNCH("ema_kernel");
  return CDAE_OK;
}

extern "C" int cdae_zero(void* p, int64_t bytes, cdae_stream s) {
  CDAE_CHECK_ARG(p || bytes == 0, "zero: null pointer");
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)s);
  if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return CDAE_ERR_CUDA; }
  return CDAE_OK;
}
