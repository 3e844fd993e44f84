// GroupNorm32 (+ FiLM scale/shift) (+ SiLU), forward and backward, NHWC bf16 activations, fp32 statistics.
// Replaces native_group_norm + casts + sigmoid/mul + FiLM mul/add (ref nn.py:430-437, unet.py:185-198, 223-231).
//
// One thread-block CLUSTER per sample: each CTA owns a slab of pixels, accumulates per-channel partial sums with
// vector loads (several independent requests in flight per thread), the cluster combines them through distributed
// shared memory, then every CTA normalises its slab (the second read of the slab hits L2: it was just streamed by the
// same CTA).  Inputs may be a channel concatenation of two tensors (UNet skip connections, unet.py:629) - groups may
// straddle the boundary.  CTAs are 256 threads with bounded registers so that 3-4 CTAs share an SM and the reduction /
// cluster-barrier phases of one CTA overlap the streaming phases of the others.
#include <cooperative_groups.h>

#include <mutex>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cdae {

constexpr int kGroups = 32;
constexpr int kGnThreads = 256;

struct GnParams {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  int C0, C1, C, HW, S;          // S = CTAs per unit (cluster size)
  int nvec, R;                   // channel vectors per pixel, pixel rows per pass
  int CC, nchunk;                // resident kernels: channels per unit (whole groups, multiple of 8), units per sample
  const float* gamma; const float* beta;
  const float* film; int film_ld, film_off;
  int silu;
  float* mean; float* rstd;
  __nv_bfloat16* y;              // forward
  const __nv_bfloat16* dy;       // backward
  const __nv_bfloat16* dadd;
  __nv_bfloat16* dx0; __nv_bfloat16* dx1; int accumulate_dx;
  float* dgamma; float* dbeta; float* dfilm;
};

// V bf16 channels per thread: 8 -> 128-bit, 4 -> 64-bit accesses
template <int V> struct VecT;
template <> struct VecT<8> { using type = uint4; };
template <> struct VecT<4> { using type = uint2; };

template <int V>
__device__ __forceinline__ typename VecT<V>::type vraw(const __nv_bfloat16* p) {
  return *reinterpret_cast<const typename VecT<V>::type*>(p);
}
template <int V>
__device__ __forceinline__ void vunpack(const typename VecT<V>::type& raw, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < V / 2; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
template <int V>
__device__ __forceinline__ void vstore(__nv_bfloat16* p, const float* f) {
  typename VecT<V>::type raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < V / 2; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<typename VecT<V>::type*>(p) = raw;
}

// ------------------------------------------------------------------------------------------------ forward
// smem: float chan[2][C] | float gpart[2][32] | float gstat[2][32]
template <int V, int U>
__global__ void __launch_bounds__(kGnThreads, 3) gn_fwd_kernel(const GnParams p) {
  extern __shared__ float sm[];
  float* chan = sm;
  float* gpart = sm + 2 * p.C;
  float* gstat = gpart + 2 * kGroups;
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const bool active = row < p.R;
  const int c = vec * V;
  const int per = (p.HW + p.S - 1) / p.S;
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;

  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) chan[i] = 0.f;
  __syncthreads();

  const bool in0 = c < p.C0;
  const __nv_bfloat16* xbase = in0 ? p.x0 + (size_t)b * p.HW * p.C0 + c : p.x1 + (size_t)b * p.HW * p.C1 + (c - p.C0);
  const int xpitch = in0 ? p.C0 : p.C1;
  if (active) {
    float s[V], ss[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { s[k] = 0.f; ss[k] = 0.f; }
    for (int pix = p0 + row; pix < p1; pix += U * p.R) {
      typename VecT<V>::type v[U];
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (pix + j * p.R < p1) v[j] = vraw<V>(xbase + (size_t)(pix + j * p.R) * xpitch);
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (pix + j * p.R < p1) {
          float f[V];
          vunpack<V>(v[j], f);
#pragma unroll
          for (int k = 0; k < V; ++k) { s[k] += f[k]; ss[k] += f[k] * f[k]; }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) { atomicAdd(&chan[c + k], s[k]); atomicAdd(&chan[p.C + c + k], ss[k]); }
  }
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
  if (!active) return;

  float A[V], Bc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int ch = c + k, g = ch / cpg;
    const float m = gstat[g], rs = gstat[kGroups + g];
    const float ga = p.gamma[ch], be = p.beta[ch];
    float sc1 = 1.f, sh = 0.f;
    if (p.film) { const float* fr = p.film + (size_t)b * p.film_ld + p.film_off; sc1 = 1.f + fr[ch]; sh = fr[p.C + ch]; }
    A[k] = rs * ga * sc1;
    Bc[k] = (be - m * rs * ga) * sc1 + sh;
  }
  __nv_bfloat16* ybase = p.y + (size_t)b * p.HW * p.C + c;
  for (int pix = p0 + row; pix < p1; pix += U * p.R) {
    typename VecT<V>::type v[U];
#pragma unroll
    for (int j = 0; j < U; ++j)
      if (pix + j * p.R < p1) v[j] = vraw<V>(xbase + (size_t)(pix + j * p.R) * xpitch);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (pix + j * p.R < p1) {
        float f[V];
        vunpack<V>(v[j], f);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float u = f[k] * A[k] + Bc[k];
          f[k] = p.silu ? silu_f(u) : u;
        }
        vstore<V>(ybase + (size_t)(pix + j * p.R) * p.C, f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// With xhat = x*a1 + b1, u = xhat*G + Hh (G = gamma*(1+scale), Hh = beta*(1+scale)+shift), du = dy*silu'(u):
//   P = sum du, Q = sum du*xhat per (sample, channel);  s1_g = sum_c G_c P_c, s2_g = sum_c G_c Q_c per group;
//   dx = rstd*(G*du - s1/n - xhat*s2/n);  d shift = P, d scale = gamma*Q + beta*P, d beta += (1+scale)*P, d gamma += (1+scale)*Q.
// smem: float chan[2][C] | float tot[2][C] | float gs[2][32]
template <int V, int U>
__global__ void __launch_bounds__(kGnThreads, 3) gn_bwd_kernel(const GnParams p) {
  extern __shared__ float sm[];
  float* chan = sm;
  float* tot = sm + 2 * p.C;
  float* gs = tot + 2 * p.C;
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const bool active = row < p.R;
  const int c = vec * V;
  const int per = (p.HW + p.S - 1) / p.S;
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;

  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) chan[i] = 0.f;
  __syncthreads();

  float a1[V], b1[V], G[V], Hh[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int ch = c + k, g = ch / cpg;
    const float m = p.mean[b * kGroups + g], rs = p.rstd[b * kGroups + g];
    float sc1 = 1.f, sh = 0.f;
    if (p.film) { const float* fr = p.film + (size_t)b * p.film_ld + p.film_off; sc1 = 1.f + fr[ch]; sh = fr[p.C + ch]; }
    a1[k] = rs; b1[k] = -m * rs;
    G[k] = p.gamma[ch] * sc1; Hh[k] = p.beta[ch] * sc1 + sh;
  }
  auto du_of = [&](float dyv, float xhat, int k) {
    if (!p.silu) return dyv;
    const float u = xhat * G[k] + Hh[k];
    const float sg = sigmoid_f(u);
    return dyv * sg * (1.f + u * (1.f - sg));
  };
  const bool in0 = c < p.C0;
  const __nv_bfloat16* xbase = in0 ? p.x0 + (size_t)b * p.HW * p.C0 + c : p.x1 + (size_t)b * p.HW * p.C1 + (c - p.C0);
  const int xpitch = in0 ? p.C0 : p.C1;
  const __nv_bfloat16* dybase = p.dy + (size_t)b * p.HW * p.C + c;

  if (active) {
    float P[V], Q[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { P[k] = 0.f; Q[k] = 0.f; }
    for (int pix = p0 + row; pix < p1; pix += U * p.R) {
      typename VecT<V>::type vx[U], vd[U];
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (pix + j * p.R < p1) {
          vx[j] = vraw<V>(xbase + (size_t)(pix + j * p.R) * xpitch);
          vd[j] = vraw<V>(dybase + (size_t)(pix + j * p.R) * p.C);
        }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (pix + j * p.R < p1) {
          float f[V], d[V];
          vunpack<V>(vx[j], f); vunpack<V>(vd[j], d);
#pragma unroll
          for (int k = 0; k < V; ++k) {
            const float xhat = f[k] * a1[k] + b1[k];
            const float du = du_of(d[k], xhat, k);
            P[k] += du; Q[k] += du * xhat;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) { atomicAdd(&chan[c + k], P[k]); atomicAdd(&chan[p.C + c + k], Q[k]); }
  }
  cluster.sync();
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
  if (!active) return;

  // K2 = rstd*s1/n, K3 = rstd*s2/n ; dx = rstd*G*du - K2 - xhat*K3
  const float inv_n = 1.f / ((float)cpg * (float)p.HW);
  float K1[V], K2[V], K3[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int g = (c + k) / cpg;
    K1[k] = a1[k] * G[k];
    K2[k] = a1[k] * gs[g] * inv_n; K3[k] = a1[k] * gs[kGroups + g] * inv_n;
  }
  constexpr int U2 = U > 2 ? 2 : U;   // four streams (x, dy, dx, dadd) are live here: halve the batching to stay in registers
  __nv_bfloat16* dxbase = in0 ? p.dx0 + (size_t)b * p.HW * p.C0 + c : p.dx1 + (size_t)b * p.HW * p.C1 + (c - p.C0);
  const __nv_bfloat16* daddbase = p.dadd ? p.dadd + (size_t)b * p.HW * p.C + c : nullptr;
  const bool acc = (p.accumulate_dx >> (in0 ? 0 : 1)) & 1;
  for (int pix = p0 + row; pix < p1; pix += U2 * p.R) {
    typename VecT<V>::type vx[U2], vd[U2], vo[U2], va[U2];
#pragma unroll
    for (int j = 0; j < U2; ++j)
      if (pix + j * p.R < p1) {
        const size_t px = (size_t)(pix + j * p.R);
        vx[j] = vraw<V>(xbase + px * xpitch);
        vd[j] = vraw<V>(dybase + px * p.C);
        if (acc) vo[j] = vraw<V>(dxbase + px * xpitch);
        if (daddbase) va[j] = vraw<V>(daddbase + px * p.C);
      }
#pragma unroll
    for (int j = 0; j < U2; ++j) {
      if (pix + j * p.R < p1) {
        float f[V], d[V], o[V];
        vunpack<V>(vx[j], f); vunpack<V>(vd[j], d);
        if (acc) vunpack<V>(vo[j], o);
        else {
#pragma unroll
          for (int k = 0; k < V; ++k) o[k] = 0.f;
        }
        if (daddbase) {
          float a[V];
          vunpack<V>(va[j], a);
#pragma unroll
          for (int k = 0; k < V; ++k) o[k] += a[k];
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float xhat = f[k] * a1[k] + b1[k];
          const float du = du_of(d[k], xhat, k);
          o[k] += K1[k] * du - K2[k] - xhat * K3[k];
        }
        vstore<V>(dxbase + (size_t)(pix + j * p.R) * xpitch, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ resident kernels
// Work unit = (sample, chunk of CC channels holding whole groups); a cluster of S CTAs shares the unit, each CTA owning
// HW/S pixels.  The CTA's slab is copied global -> SHARED MEMORY with cp.async (the whole slab is in flight at once:
// memory-level parallelism does not cost registers) and stays there between the statistics pass and the apply pass,
// so every tensor crosses HBM (and the L2 -> SM fabric, which on B200 is barely faster than HBM) exactly once:
// forward 2 B read + 2 B written per element, backward 4 B read + 2 B written.  128-bit accesses, 8 channels per
// thread; partial sums go warp shuffle -> shared atomics -> DSMEM across the cluster.  The backward overwrites the dy
// slab with du = dy*silu'(u) (bf16): the sigmoid is evaluated once per element.
constexpr int kResElemsFwd = 32768;     // 64 KB slab (x)
constexpr int kResElemsBwd = 32768;     // 64 KB slab (du)
constexpr int kResThreads = 256;

__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// sigmoid(u) = 0.5*tanh(0.5u) + 0.5 : one MUFU op (the results are rounded to bf16 right after)
__device__ __forceinline__ float sigmoid_t(float u) { return fmaf(0.5f, tanh_fast(0.5f * u), 0.5f); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// sum v over the lanes that share (lane % 8) - valid when nvec == 8
__device__ __forceinline__ float vec8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// bf16 pair -> two fp32 (one shift, one mask)
__device__ __forceinline__ void unpack2(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16); hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ void unpack_u4(const uint4& v, float* f) {
  unpack2(v.x, f[0], f[1]); unpack2(v.y, f[2], f[3]); unpack2(v.z, f[4], f[5]); unpack2(v.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack_u4(const float* f) {
  uint4 r;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(f[0], f[1]); r.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[2], f[3]); r.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[4], f[5]); r.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[6], f[7]); r.w = *reinterpret_cast<uint32_t*>(&h);
  return r;
}

// The inner loops are written for instruction count (the first version of these kernels was ISSUE bound: 21 / 42
// thread instructions per element forward / backward, ncu r1): thread `tid` owns slab entries tid, tid + RN, ... (RN =
// active threads), so every stream is a pointer that advances by a loop-invariant stride; SiLU is evaluated through
// h = u/2: silu(u) = h + h*tanh(h) and silu'(u) = (1 + t + h*(1 - t^2))/2 with t = tanh(h) (one MUFU per element).
// smem: bf16 slab[P][CC] | float chan[2][CC] | float gpart[2][32] | float gstat[2][32]
template <bool SILU>
__global__ void __launch_bounds__(kResThreads, 3) gn_fwd_res_kernel(const GnParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int per = (p.HW + p.S - 1) / p.S;
  uint4* slab = reinterpret_cast<uint4*>(smraw);
  float* chan = reinterpret_cast<float*>(smraw + (size_t)per * p.CC * 2);
  float* gpart = chan + 2 * p.CC;
  float* gstat = gpart + 2 * kGroups;
  cg::cluster_group cluster = cg::this_cluster();
  const int unit = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int b = unit / p.nchunk, chunk = unit % p.nchunk;
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const bool active = row < p.R;
  const int cl = vec * 8, c = chunk * p.CC + cl;              // channel inside the chunk / absolute
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;
  const int ng = p.CC / cpg;                                   // groups in this chunk
  const int RN = p.R * p.nvec;
  const int nit = active ? max(0, (p1 - p0 - row + p.R - 1) / p.R) : 0;   // pixels owned by this thread

  const bool in0 = c < p.C0;
  const __nv_bfloat16* xbase = in0 ? p.x0 + (size_t)b * p.HW * p.C0 + c : p.x1 + (size_t)b * p.HW * p.C1 + (c - p.C0);
  const int xpitch = in0 ? p.C0 : p.C1;
  uint4* const sl = slab + threadIdx.x;
  {
    const __nv_bfloat16* g = xbase + (size_t)(p0 + row) * xpitch;
    const size_t gstep = (size_t)p.R * xpitch;
    uint4* s_ = sl;
    for (int i = 0; i < nit; ++i) { cp_async16(s_, g); s_ += RN; g += gstep; }
  }
  // per-channel affine constants do not depend on the statistics: fetch them while the slab is in flight
  float Gk[8], Hk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = active ? c + k : 0;
    const float ga = p.gamma[ch], be = p.beta[ch];
    float sc1 = 1.f, sh = 0.f;
    if (p.film) { const float* fr = p.film + (size_t)b * p.film_ld + p.film_off; sc1 = 1.f + fr[ch]; sh = fr[p.C + ch]; }
    Gk[k] = ga * sc1; Hk[k] = be * sc1 + sh;
  }
  for (int i = threadIdx.x; i < 2 * p.CC; i += blockDim.x) chan[i] = 0.f;
  cp_async_wait_all();
  __syncthreads();

  float s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; ss[k] = 0.f; }
  {
    const uint4* s_ = sl;
#pragma unroll 4
    for (int i = 0; i < nit; ++i) {
      float f[8];
      unpack_u4(*s_, f);
      s_ += RN;
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += f[k]; ss[k] = fmaf(f[k], f[k], ss[k]); }
    }
  }
  if (p.nvec == 8) {            // warp = 4 pixel rows x 8 vectors: fold the rows before touching shared memory
#pragma unroll
    for (int k = 0; k < 8; ++k) { s[k] = vec8_sum(s[k]); ss[k] = vec8_sum(ss[k]); }
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { atomicAdd(&chan[cl + k], s[k]); atomicAdd(&chan[p.CC + cl + k], ss[k]); }
    }
  } else if (active) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&chan[cl + k], s[k]); atomicAdd(&chan[p.CC + cl + k], ss[k]); }
  }
  __syncthreads();
  if (threadIdx.x < ng) {
    float a = 0.f, q = 0.f;
    for (int k = 0; k < cpg; ++k) { a += chan[threadIdx.x * cpg + k]; q += chan[p.CC + threadIdx.x * cpg + k]; }
    gpart[threadIdx.x] = a; gpart[kGroups + threadIdx.x] = q;
  }
  cluster.sync();
  if (threadIdx.x < ng) {
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
    if (rank == 0) {
      const int g = chunk * ng + threadIdx.x;
      p.mean[b * kGroups + g] = m; p.rstd[b * kGroups + g] = rs;
    }
  }
  cluster.sync();   // remote reads of gpart complete before any CTA may exit; also publishes gstat block-wide
  if (!active) return;

  float A[8], Bc[8];
  const float half = SILU ? 0.5f : 1.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (cl + k) / cpg;
    const float m = gstat[g], rs = gstat[kGroups + g];
    A[k] = half * rs * Gk[k];
    Bc[k] = half * (Hk[k] - m * rs * Gk[k]);
  }
  {
    __nv_bfloat16* y = p.y + (size_t)b * p.HW * p.C + c + (size_t)(p0 + row) * p.C;
    const size_t ystep = (size_t)p.R * p.C;
    const uint4* s_ = sl;
#pragma unroll 2
    for (int i = 0; i < nit; ++i) {
      float f[8];
      unpack_u4(*s_, f);
      s_ += RN;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float h = fmaf(f[k], A[k], Bc[k]);
        f[k] = SILU ? fmaf(h, tanh_fast(h), h) : h;
      }
      *reinterpret_cast<uint4*>(y) = pack_u4(f);
      y += ystep;
    }
  }
}

// smem: bf16 du slab[P][CC] | float chan[2][CC] | float tot[2][CC] | float gs[2][32]
// With u = x*aG + bH (aG = rstd*G, bH = -mean*rstd*G + Hh), du = dy*silu'(u), P = sum du, Qx = sum du*x:
//   Q = sum du*xhat = rstd*Qx - mean*rstd*P ;  dx = K1*du - K2' - x*K3'  with K1 = rstd*G, K3' = rstd^2*s2/n,
//   K2' = rstd*s1/n - mean*rstd^2*s2/n.
// x and dy are streamed through registers (two 128-bit loads per pixel per thread, two pixels in flight); du is kept in
// shared memory, x is read again in the apply pass (an L2 hit: the same CTA streamed it microseconds earlier and the
// in-flight footprint of all CTAs is a few tens of MB).
template <bool SILU>
__global__ void __launch_bounds__(kResThreads, 3) gn_bwd_res_kernel(const GnParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int per = (p.HW + p.S - 1) / p.S;
  uint4* slab = reinterpret_cast<uint4*>(smraw);
  float* chan = reinterpret_cast<float*>(smraw + (size_t)per * p.CC * 2);
  float* tot = chan + 2 * p.CC;
  float* gs = tot + 2 * p.CC;
  cg::cluster_group cluster = cg::this_cluster();
  const int unit = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int b = unit / p.nchunk, chunk = unit % p.nchunk;
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const bool active = row < p.R;
  const int cl = vec * 8, c = chunk * p.CC + cl;
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;
  const int ng = p.CC / cpg;
  const int c0 = chunk * p.CC;
  const int RN = p.R * p.nvec;
  const int nit = active ? max(0, (p1 - p0 - row + p.R - 1) / p.R) : 0;

  for (int i = threadIdx.x; i < 2 * p.CC; i += blockDim.x) chan[i] = 0.f;
  __syncthreads();

  float aGh[8], bHh[8];      // halved affine: h = u/2 = x*aGh + bHh
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = active ? c + k : 0, g = ch / cpg;
    const float m = p.mean[b * kGroups + g], rs = p.rstd[b * kGroups + g];
    float sc1 = 1.f, sh = 0.f;
    if (p.film) { const float* fr = p.film + (size_t)b * p.film_ld + p.film_off; sc1 = 1.f + fr[ch]; sh = fr[p.C + ch]; }
    const float G = p.gamma[ch] * sc1, Hh = p.beta[ch] * sc1 + sh;
    aGh[k] = 0.5f * rs * G; bHh[k] = 0.5f * (Hh - m * rs * G);
  }
  const bool in0 = c < p.C0;
  const __nv_bfloat16* xbase = in0 ? p.x0 + (size_t)b * p.HW * p.C0 + c : p.x1 + (size_t)b * p.HW * p.C1 + (c - p.C0);
  const int xpitch = in0 ? p.C0 : p.C1;
  const size_t xstep = (size_t)p.R * xpitch, dstep = (size_t)p.R * p.C;
  const __nv_bfloat16* const xg0 = xbase + (size_t)(p0 + row) * xpitch;
  uint4* const sl = slab + threadIdx.x;

  float P[8], Qx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { P[k] = 0.f; Qx[k] = 0.f; }
  {
    const __nv_bfloat16* xg = xg0;
    const __nv_bfloat16* dg = p.dy + (size_t)b * p.HW * p.C + c + (size_t)(p0 + row) * p.C;
    uint4* s_ = sl;
    auto one = [&](const uint4& vx, const uint4& vd, uint4* dst) {
      float f[8], d[8];
      unpack_u4(vx, f); unpack_u4(vd, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float du = d[k];
        if (SILU) {
          const float h = fmaf(f[k], aGh[k], bHh[k]);
          const float t = tanh_fast(h);
          const float w = fmaf(h, fmaf(-t, t, 1.f), t);        // t + h*(1 - t^2)
          du = d[k] * fmaf(0.5f, w, 0.5f);
        }
        P[k] += du; Qx[k] = fmaf(du, f[k], Qx[k]);
        d[k] = du;
      }
      *dst = pack_u4(d);
    };
    int i = 0;
    for (; i + 2 <= nit; i += 2) {
      const uint4 vx0 = __ldg(reinterpret_cast<const uint4*>(xg)), vd0 = __ldg(reinterpret_cast<const uint4*>(dg));
      const uint4 vx1 = __ldg(reinterpret_cast<const uint4*>(xg + xstep)), vd1 = __ldg(reinterpret_cast<const uint4*>(dg + dstep));
      one(vx0, vd0, s_); one(vx1, vd1, s_ + RN);
      xg += 2 * xstep; dg += 2 * dstep; s_ += 2 * RN;
    }
    if (i < nit) {
      const uint4 vx0 = __ldg(reinterpret_cast<const uint4*>(xg)), vd0 = __ldg(reinterpret_cast<const uint4*>(dg));
      one(vx0, vd0, s_);
    }
  }
  if (p.nvec == 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { P[k] = vec8_sum(P[k]); Qx[k] = vec8_sum(Qx[k]); }
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { atomicAdd(&chan[cl + k], P[k]); atomicAdd(&chan[p.CC + cl + k], Qx[k]); }
    }
  } else if (active) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&chan[cl + k], P[k]); atomicAdd(&chan[p.CC + cl + k], Qx[k]); }
  }
  cluster.sync();
  for (int i = threadIdx.x; i < 2 * p.CC; i += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < p.S; ++r) a += cluster.map_shared_rank(chan, r)[i];
    tot[i] = a;
  }
  cluster.sync();
  // tot[0][c] = P_c, tot[1][c] = sum du*x ; Q_c = rstd*(Qx_c - mean*P_c)
  if (threadIdx.x < ng) {
    const int g = chunk * ng + threadIdx.x;
    const float m = p.mean[b * kGroups + g], rs = p.rstd[b * kGroups + g];
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      const int lc = threadIdx.x * cpg + k, ch = c0 + lc;
      float kc = p.gamma[ch];
      if (p.film) kc *= 1.f + p.film[(size_t)b * p.film_ld + p.film_off + ch];
      const float Pc = tot[lc], Qc = rs * (tot[p.CC + lc] - m * Pc);
      s1 += kc * Pc; s2 += kc * Qc;
    }
    gs[threadIdx.x] = s1; gs[kGroups + threadIdx.x] = s2;
  }
  if (rank == 0) {
    for (int lc = threadIdx.x; lc < p.CC; lc += blockDim.x) {
      const int ch = c0 + lc, g = ch / cpg;
      const float m = p.mean[b * kGroups + g], rs = p.rstd[b * kGroups + g];
      const float Pc = tot[lc], Qc = rs * (tot[p.CC + lc] - m * Pc);
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
  if (!active) return;

  const float inv_n = 1.f / ((float)cpg * (float)p.HW);
  float K1[8], K2[8], K3[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = c + k, g = ch / cpg, lg = (cl + k) / cpg;
    const float m = p.mean[b * kGroups + g], rs = p.rstd[b * kGroups + g];
    K1[k] = 2.f * aGh[k];
    K3[k] = rs * rs * gs[kGroups + lg] * inv_n;
    K2[k] = rs * gs[lg] * inv_n - m * K3[k];
  }
  const bool acc = (p.accumulate_dx >> (in0 ? 0 : 1)) & 1;
  const bool hasadd = p.dadd != nullptr;
  {
    const __nv_bfloat16* xg = xg0;
    __nv_bfloat16* og = (in0 ? p.dx0 + (size_t)b * p.HW * p.C0 + c : p.dx1 + (size_t)b * p.HW * p.C1 + (c - p.C0)) +
                        (size_t)(p0 + row) * xpitch;
    const __nv_bfloat16* ag = hasadd ? p.dadd + (size_t)b * p.HW * p.C + c + (size_t)(p0 + row) * p.C : nullptr;
    const uint4* s_ = sl;
    auto one = [&](const uint4& vx, const uint4& vdu, const uint4& vo, const uint4& va, __nv_bfloat16* dst) {
      float f[8], d[8], o[8];
      unpack_u4(vx, f); unpack_u4(vdu, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(K1[k], d[k], -fmaf(f[k], K3[k], K2[k]));
      if (acc) {
        float t[8];
        unpack_u4(vo, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += t[k];
      }
      if (hasadd) {
        float t[8];
        unpack_u4(va, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += t[k];
      }
      *reinterpret_cast<uint4*>(dst) = pack_u4(o);
    };
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    int i = 0;
    for (; i + 2 <= nit; i += 2) {
      const uint4 vx0 = __ldg(reinterpret_cast<const uint4*>(xg)), vx1 = __ldg(reinterpret_cast<const uint4*>(xg + xstep));
      uint4 vo0 = z, vo1 = z, va0 = z, va1 = z;
      if (acc) { vo0 = *reinterpret_cast<const uint4*>(og); vo1 = *reinterpret_cast<const uint4*>(og + xstep); }
      if (hasadd) { va0 = __ldg(reinterpret_cast<const uint4*>(ag)); va1 = __ldg(reinterpret_cast<const uint4*>(ag + dstep)); }
      one(vx0, s_[0], vo0, va0, og); one(vx1, s_[RN], vo1, va1, og + xstep);
      xg += 2 * xstep; og += 2 * xstep; s_ += 2 * RN;
      if (hasadd) ag += 2 * dstep;
    }
    if (i < nit) {
      const uint4 vx0 = __ldg(reinterpret_cast<const uint4*>(xg));
      uint4 vo0 = z, va0 = z;
      if (acc) vo0 = *reinterpret_cast<const uint4*>(og);
      if (hasadd) va0 = __ldg(reinterpret_cast<const uint4*>(ag));
      one(vx0, s_[0], vo0, va0, og);
    }
  }
}

// pick the unit decomposition of the resident kernels; returns false when the sample does not fit (huge images)
static bool gn_res_config(GnParams& p, int max_elems, int max_cluster) {
  p.C = p.C0 + p.C1;
  if (p.C % kGroups || p.C0 % 8 || p.C1 % 8) return false;
  const int cpg = p.C / kGroups;
  int base = cpg;                                   // lcm(8, cpg)
  while (base % 8) base += cpg;
  // candidates: multiples of lcm(8, cpg) that divide C.  Prefer the smallest one with >= 128 B rows whose unit fits a
  // cluster; otherwise the largest narrower one that fits.
  auto fits = [&](int m, int* S_out) {
    if (m / 8 > kResThreads || m / cpg > kGroups) return false;
    int S = 1;
    while (S < max_cluster && (int64_t)((p.HW + S - 1) / S) * m > max_elems) S <<= 1;
    *S_out = S;
    return (int64_t)((p.HW + S - 1) / S) * m <= max_elems;
  };
  int CC = 0, S = 1;
  for (int m = base; m <= p.C && !CC; m += base)
    if (p.C % m == 0 && m >= 64 && fits(m, &S)) CC = m;
  if (!CC)
    for (int m = base; m < 64 && m <= p.C; m += base)
      if (p.C % m == 0 && fits(m, &S)) CC = m;       // keeps the largest
  if (!CC) return false;
  fits(CC, &S);
  p.CC = CC; p.nchunk = p.C / CC; p.S = S;
  p.nvec = CC / 8; p.R = kResThreads / p.nvec;
  return true;
}

template <int TAG, typename K>
static int gn_res_launch(K kernel, const GnParams& p, int B, size_t smem, cudaStream_t st, const char* name) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  });
  if (attr_err != cudaSuccess) { set_error("%s attributes: %s", name, cudaGetErrorString(attr_err)); return CDAE_ERR_CUDA; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(B * p.nchunk * p.S));
  cfg.blockDim = dim3((unsigned)((p.nvec * p.R + 31) / 32 * 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p.S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
  if (e != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e)); return CDAE_ERR_CUDA; }
  return CDAE_OK;
}

template <int V>
static int gn_config(GnParams& p, int* S_out) {
  p.C = p.C0 + p.C1;
  CDAE_CHECK_SHAPE(p.C % kGroups == 0, "groupnorm: C=%d not a multiple of 32", p.C);
  CDAE_CHECK_SHAPE(p.C0 % 8 == 0 && p.C1 % 8 == 0, "groupnorm: source channel counts must be multiples of 8");
  p.nvec = p.C / V;
  CDAE_CHECK_SHAPE(p.nvec <= kGnThreads, "groupnorm: C=%d too large for %d-channel vectors", p.C, V);
  p.R = kGnThreads / p.nvec;           // rows of pixels per pass; threads with row >= R idle (C not a power of two)
  int S = 8;
  while (S > 1 && ((int64_t)p.HW * p.C / S < 8192 || p.HW / S < p.R)) S >>= 1;
  p.S = S;
  *S_out = S;
  return CDAE_OK;
}

template <typename K>
static int gn_launch(K kernel, const GnParams& p, int B, int S, size_t smem, cudaStream_t st, const char* name) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(B * S));
  cfg.blockDim = dim3((unsigned)((p.nvec * p.R + 31) / 32 * 32));   // no idle warps when C is not a power of two
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
  if (gn_res_config(p, kResElemsFwd, 8)) {
    const size_t per = (size_t)((HW + p.S - 1) / p.S);
    const size_t smem_res = per * p.CC * 2 + sizeof(float) * (2 * p.CC + 4 * kGroups);
    return silu ? gn_res_launch<0>(gn_fwd_res_kernel<true>, p, B, smem_res, (cudaStream_t)s, "gn_fwd_res_kernel")
                : gn_res_launch<1>(gn_fwd_res_kernel<false>, p, B, smem_res, (cudaStream_t)s, "gn_fwd_res_kernel");
  }
  int S;
  const bool wide = (C0 + C1) > 4 * kGnThreads;     // > 1024 channels: 8-channel vectors keep nvec <= 256
  int rc = wide ? gn_config<8>(p, &S) : gn_config<4>(p, &S);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (2 * p.C + 4 * kGroups);
  if (wide) return gn_launch(gn_fwd_kernel<8, 4>, p, B, S, smem, (cudaStream_t)s, "gn_fwd_kernel");
  return gn_launch(gn_fwd_kernel<4, 4>, p, B, S, smem, (cudaStream_t)s, "gn_fwd_kernel");
}

extern "C" int cdae_gn_bwd(const void* dy, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                           const float* gamma, const float* beta, const float* film, int film_ld, int film_off, int silu,
                           const float* mean, const float* rstd, const void* dadd, void* dx0, void* dx1, int accumulate_dx,
                           float* dgamma, float* dbeta, float* dfilm, cdae_stream s) {
  CDAE_CHECK_ARG(dy && x0 && gamma && beta && mean && rstd && dx0 && (C1 == 0 || (x1 && dx1)), "gn_bwd: null pointer");
  if (B == 0) return CDAE_OK;
  GnParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1; p.C0 = C0; p.C1 = C1; p.HW = HW;
  p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld; p.film_off = film_off; p.silu = silu;
  p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd);
  p.dy = (const __nv_bfloat16*)dy; p.dadd = (const __nv_bfloat16*)dadd; p.dx0 = (__nv_bfloat16*)dx0; p.dx1 = (__nv_bfloat16*)dx1;
  p.accumulate_dx = accumulate_dx;
  p.dgamma = dgamma; p.dbeta = dbeta; p.dfilm = dfilm;
  if (gn_res_config(p, kResElemsBwd, 8)) {
    const size_t per = (size_t)((HW + p.S - 1) / p.S);
    const size_t smem_res = per * p.CC * 2 + sizeof(float) * (4 * p.CC + 2 * kGroups);
    return silu ? gn_res_launch<2>(gn_bwd_res_kernel<true>, p, B, smem_res, (cudaStream_t)s, "gn_bwd_res_kernel")
                : gn_res_launch<3>(gn_bwd_res_kernel<false>, p, B, smem_res, (cudaStream_t)s, "gn_bwd_res_kernel");
  }
  int S;
  const bool wide = (C0 + C1) > 4 * kGnThreads;
  int rc = wide ? gn_config<8>(p, &S) : gn_config<4>(p, &S);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (4 * p.C + 2 * kGroups);
  if (wide) return gn_launch(gn_bwd_kernel<8, 2>, p, B, S, smem, (cudaStream_t)s, "gn_bwd_kernel");
  return gn_launch(gn_bwd_kernel<4, 4>, p, B, S, smem, (cudaStream_t)s, "gn_bwd_kernel");
}
