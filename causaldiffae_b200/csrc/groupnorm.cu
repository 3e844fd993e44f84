// GroupNorm32 (+ FiLM scale/shift) (+ SiLU), forward and backward, NHWC bf16 activations, fp32 statistics.
// Replaces native_group_norm + casts + sigmoid/mul + FiLM mul/add (ref nn.py:430-437, unet.py:185-198, 223-231).
//
// One thread-block CLUSTER per sample: each CTA owns a slab of pixels, accumulates per-channel partial sums with
// 128-bit loads, the cluster combines them through distributed shared memory, then every CTA normalises its slab
// (the second read of the slab hits L2: the slab was just streamed by the same CTA).  Inputs may be a channel
// concatenation of two tensors (UNet skip connections, unet.py:629) - groups may straddle the boundary.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cdae {

constexpr int kGroups = 32;

struct GnParams {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  int C0, C1, C, HW, S;          // S = CTAs per sample (cluster size)
  int nvec, R;                   // 8-channel vectors per pixel, pixel rows per pass
  const float* gamma; const float* beta;
  const float* film; int film_ld, film_off;
  int silu;
  float* mean; float* rstd;
  // forward
  __nv_bfloat16* y;
  // backward
  const __nv_bfloat16* dy;
  const __nv_bfloat16* dadd;
  __nv_bfloat16* dx0; __nv_bfloat16* dx1; int accumulate_dx;
  float* dgamma; float* dbeta; float* dfilm;
};

__device__ __forceinline__ const __nv_bfloat16* src_ptr(const GnParams& p, int b, int pix, int c) {
  return c < p.C0 ? p.x0 + ((size_t)b * p.HW + pix) * p.C0 + c : p.x1 + ((size_t)b * p.HW + pix) * p.C1 + (c - p.C0);
}

// smem layout: float chan[2][C] | float gpart[2][32] | float gstat[2][32]
__global__ void gn_fwd_kernel(const GnParams p) {
  extern __shared__ float sm[];
  float* chan = sm;
  float* gpart = sm + 2 * p.C;
  float* gstat = gpart + 2 * kGroups;
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const int c = vec * 8;
  const int per = (p.HW + p.S - 1) / p.S;
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;

  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) chan[i] = 0.f;
  __syncthreads();

  float s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; ss[k] = 0.f; }
  for (int pix = p0 + row; pix < p1; pix += p.R) {
    float f[8];
    unpack8(ld8(src_ptr(p, b, pix, c)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] += f[k]; ss[k] += f[k] * f[k]; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { atomicAdd(&chan[c + k], s[k]); atomicAdd(&chan[p.C + c + k], ss[k]); }
  __syncthreads();
  if (threadIdx.x < kGroups) {
    float a = 0.f, q = 0.f;
    for (int k = 0; k < cpg; ++k) { a += chan[threadIdx.x * cpg + k]; q += chan[p.C + threadIdx.x * cpg + k]; }
    gpart[threadIdx.x] = a; gpart[kGroups + threadIdx.x] = q;
  }
  cluster.sync();
  if (threadIdx.x < kGroups) {
    float a = 0.f, q = 0.f;
    for (int r = 0; r < p.S; ++r) {
      const float* rp = cluster.map_shared_rank(gpart, r);
      a += rp[threadIdx.x]; q += rp[kGroups + threadIdx.x];
    }
    const float n = (float)cpg * (float)p.HW;
    const float m = a / n;
    const float var = fmaxf(q / n - m * m, 0.f);
    const float rs = rsqrtf(var + 1e-5f);
    gstat[threadIdx.x] = m; gstat[kGroups + threadIdx.x] = rs;
    if (rank == 0) { p.mean[b * kGroups + threadIdx.x] = m; p.rstd[b * kGroups + threadIdx.x] = rs; }
  }
  cluster.sync();   // remote reads of gpart complete before any CTA may exit; also publishes gstat block-wide

  float A[8], Bc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = c + k, g = ch / cpg;
    const float m = gstat[g], rs = gstat[kGroups + g];
    const float ga = p.gamma[ch], be = p.beta[ch];
    float sc1 = 1.f, sh = 0.f;
    if (p.film) { const float* fr = p.film + (size_t)b * p.film_ld + p.film_off; sc1 = 1.f + fr[ch]; sh = fr[p.C + ch]; }
    A[k] = rs * ga * sc1;
    Bc[k] = (be - m * rs * ga) * sc1 + sh;
  }
  for (int pix = p0 + row; pix < p1; pix += p.R) {
    float f[8];
    unpack8(ld8(src_ptr(p, b, pix, c)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float u = f[k] * A[k] + Bc[k];
      f[k] = p.silu ? silu_f(u) : u;
    }
    st8(p.y + ((size_t)b * p.HW + pix) * p.C + c, pack8(f));
  }
}

// smem: float chan[2][C] (P = sum du, Q = sum du*xhat) | float tot[2][C] | float gs[2][32]
__global__ void gn_bwd_kernel(const GnParams p) {
  extern __shared__ float sm[];
  float* chan = sm;
  float* tot = sm + 2 * p.C;
  float* gs = tot + 2 * p.C;
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const int c = vec * 8;
  const int per = (p.HW + p.S - 1) / p.S;
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;

  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) chan[i] = 0.f;
  __syncthreads();

  float mean[8], rs[8], ga[8], be[8], sc1[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = c + k, g = ch / cpg;
    mean[k] = p.mean[b * kGroups + g]; rs[k] = p.rstd[b * kGroups + g];
    ga[k] = p.gamma[ch]; be[k] = p.beta[ch];
    sc1[k] = 1.f; sh[k] = 0.f;
    if (p.film) { const float* fr = p.film + (size_t)b * p.film_ld + p.film_off; sc1[k] = 1.f + fr[ch]; sh[k] = fr[p.C + ch]; }
  }
  auto du_of = [&](float dyv, float xhat, int k) {
    if (!p.silu) return dyv;
    const float u = (xhat * ga[k] + be[k]) * sc1[k] + sh[k];
    const float sg = sigmoid_f(u);
    return dyv * sg * (1.f + u * (1.f - sg));
  };

  float P[8], Q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { P[k] = 0.f; Q[k] = 0.f; }
  for (int pix = p0 + row; pix < p1; pix += p.R) {
    float f[8], d[8];
    unpack8(ld8(src_ptr(p, b, pix, c)), f);
    unpack8(ld8(p.dy + ((size_t)b * p.HW + pix) * p.C + c), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xhat = (f[k] - mean[k]) * rs[k];
      const float du = du_of(d[k], xhat, k);
      P[k] += du; Q[k] += du * xhat;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { atomicAdd(&chan[c + k], P[k]); atomicAdd(&chan[p.C + c + k], Q[k]); }
  cluster.sync();
  // per-sample channel totals (every CTA computes them; cheap) and group sums
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < p.S; ++r) a += cluster.map_shared_rank(chan, r)[i];
    tot[i] = a;
  }
  cluster.sync();
  if (threadIdx.x < kGroups) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      const int ch = threadIdx.x * cpg + k;
      float kc = p.gamma[ch];
      if (p.film) kc *= 1.f + p.film[(size_t)b * p.film_ld + p.film_off + ch];
      s1 += kc * tot[ch]; s2 += kc * tot[p.C + ch];
    }
    gs[threadIdx.x] = s1; gs[kGroups + threadIdx.x] = s2;
  }
  if (rank == 0) {
    for (int ch = threadIdx.x; ch < p.C; ch += blockDim.x) {
      const float Pc = tot[ch], Qc = tot[p.C + ch];
      float s1c = 1.f;
      if (p.film) {
        const size_t fo = (size_t)b * p.film_ld + p.film_off;
        s1c = 1.f + p.film[fo + ch];
        if (p.dfilm) {
          p.dfilm[fo + ch] += p.gamma[ch] * Qc + p.beta[ch] * Pc;   // d scale  (columns owned by this layer & sample)
          p.dfilm[fo + p.C + ch] += Pc;                              // d shift
        }
      }
      if (p.dgamma) atomicAdd(p.dgamma + ch, s1c * Qc);
      if (p.dbeta) atomicAdd(p.dbeta + ch, s1c * Pc);
    }
  }
  __syncthreads();

  const float inv_n = 1.f / ((float)cpg * (float)p.HW);
  float s1[8], s2[8], kc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (c + k) / cpg;
    s1[k] = gs[g] * inv_n; s2[k] = gs[kGroups + g] * inv_n; kc[k] = ga[k] * sc1[k];
  }
  for (int pix = p0 + row; pix < p1; pix += p.R) {
    float f[8], d[8];
    unpack8(ld8(src_ptr(p, b, pix, c)), f);
    unpack8(ld8(p.dy + ((size_t)b * p.HW + pix) * p.C + c), d);
    __nv_bfloat16* dst = c < p.C0 ? p.dx0 + ((size_t)b * p.HW + pix) * p.C0 + c
                                  : p.dx1 + ((size_t)b * p.HW + pix) * p.C1 + (c - p.C0);
    float o[8];
    const bool acc = (p.accumulate_dx >> (c < p.C0 ? 0 : 1)) & 1;
    if (acc) unpack8(ld8(dst), o);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    }
    if (p.dadd) {
      float a[8];
      unpack8(ld8(p.dadd + ((size_t)b * p.HW + pix) * p.C + c), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += a[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xhat = (f[k] - mean[k]) * rs[k];
      const float du = du_of(d[k], xhat, k);
      o[k] += rs[k] * (kc[k] * du - s1[k] - xhat * s2[k]);
    }
    st8(dst, pack8(o));
  }
}

static int gn_config(GnParams& p, int B, int* threads, int* S_out) {
  p.C = p.C0 + p.C1;
  CDAE_CHECK_SHAPE(p.C % kGroups == 0, "groupnorm: C=%d not a multiple of 32", p.C);
  CDAE_CHECK_SHAPE(p.C0 % 8 == 0 && p.C1 % 8 == 0, "groupnorm: source channel counts must be multiples of 8");
  CDAE_CHECK_SHAPE(p.C <= 2048, "groupnorm: C=%d too large", p.C);
  p.nvec = p.C / 8;
  int R = 512 / p.nvec;
  while (R > 1 && (p.nvec * R) % 32 != 0) --R;
  if (R < 1) R = 1;
  CDAE_CHECK_SHAPE(p.nvec * R <= 1024 && p.nvec * R >= 1, "groupnorm: unsupported channel count %d", p.C);
  p.R = R;
  *threads = p.nvec * R;
  int S = 8;
  while (S > 1 && ((int64_t)p.HW * p.C / S < 16384 || p.HW / S < R)) S >>= 1;
  p.S = S;
  *S_out = S;
  (void)B;
  return CDAE_OK;
}

template <typename K>
static int gn_launch(K kernel, const GnParams& p, int B, int threads, int S, size_t smem, cudaStream_t st, const char* name) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(B * S));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
  if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return CDAE_ERR_CUDA; }
  return CDAE_OK;
}

}  // namespace cdae
using namespace cdae;

extern "C" int cdae_gn_fwd(const void* x0, int C0, const void* x1, int C1, int B, int HW, const float* gamma,
                           const float* beta, const float* film, int film_ld, int film_off, int silu, void* y,
                           float* mean, float* rstd, cdae_stream s) {
  CDAE_CHECK_ARG(x0 && gamma && beta && y && mean && rstd && (C1 == 0 || x1), "gn_fwd: null pointer");
  if (B == 0) return CDAE_OK;
  GnParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1; p.C0 = C0; p.C1 = C1; p.HW = HW;
  p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld; p.film_off = film_off; p.silu = silu;
  p.y = (__nv_bfloat16*)y; p.mean = mean; p.rstd = rstd;
  int threads, S;
  int rc = gn_config(p, B, &threads, &S);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (2 * p.C + 4 * kGroups);
  return gn_launch(gn_fwd_kernel, p, B, threads, S, smem, (cudaStream_t)s, "gn_fwd_kernel");
}

extern "C" int cdae_gn_bwd(const void* dy, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                           const float* gamma, const float* beta, const float* film, int film_ld, int film_off, int silu,
                           const float* mean, const float* rstd, const void* dadd, void* dx0, void* dx1, int accumulate_dx,
                           float* dgamma,
                           float* dbeta, float* dfilm, cdae_stream s) {
  CDAE_CHECK_ARG(dy && x0 && gamma && beta && mean && rstd && dx0 && (C1 == 0 || (x1 && dx1)), "gn_bwd: null pointer");
  if (B == 0) return CDAE_OK;
  GnParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1; p.C0 = C0; p.C1 = C1; p.HW = HW;
  p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld; p.film_off = film_off; p.silu = silu;
  p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd);
  p.dy = (const __nv_bfloat16*)dy; p.dadd = (const __nv_bfloat16*)dadd; p.dx0 = (__nv_bfloat16*)dx0; p.dx1 = (__nv_bfloat16*)dx1; p.accumulate_dx = accumulate_dx;
  p.dgamma = dgamma; p.dbeta = dbeta; p.dfilm = dfilm;
  int threads, S;
  int rc = gn_config(p, B, &threads, &S);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (4 * p.C + 2 * kGroups);
  return gn_launch(gn_bwd_kernel, p, B, threads, S, smem, (cudaStream_t)s, "gn_bwd_kernel");
}
