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

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cdae {

constexpr int kGroups = 32;
constexpr int kGnThreads = 256;

struct GnParams {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  int C0, C1, C, HW, S;          // S = CTAs per unit (cluster size)
  int nvec, R;                   // channel vectors per pixel, pixel rows per pass
  int CC, nchunk;                // pipelined kernels: channels per unit (whole groups, multiple of 8), units per sample
  int per, nunits, ncl, v4, nvp;      // pixels per CTA slab, units, clusters in the persistent grid, 128-bit constant loads ok
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

// ------------------------------------------------------------------------------------------------ pipelined kernels
// Work unit = (sample, chunk of CC channels holding whole groups); a cluster of S CTAs shares the unit, each CTA owning
// HW/S pixels (its "slab").  The kernels are PERSISTENT: a cluster walks units cid, cid + ncl, ... and every CTA keeps a
// two-slab ring in shared memory - while it reduces / normalises the slab of unit i, the slab of unit i+1 is already
// streaming in through cp.async (16 B per request, no registers held), and the stores of unit i-1 drain behind it.
// ncu on the first (one unit per CTA, 64 KB slab) version showed why that matters: load, compute and store phases ran
// back to back in every CTA of a wave, so HBM idled two thirds of the time (45 % of peak) although each tensor crossed
// it only once.  Three such CTAs share an SM (64 KB + statistics each), so cluster barriers of one overlap the math
// of the others.  Every tensor crosses HBM once: forward 2 B read + 2 B written per element, backward 4 B + 2 B.
// Partial sums: registers -> warp shuffles -> per-warp rows in shared memory (plain stores, no atomics when the
// vector count divides 32) -> DSMEM across the cluster (published in a parity-double-buffered block: ONE cluster
// barrier per unit).  SiLU goes through h = u/2: silu(u) = h + h*tanh(h), silu'(u) = (1 + t + h*(1 - t^2))/2 with
// t = tanh(h): one MUFU per element.
constexpr int kPipeElems = 16384;       // 32 KB slab; two of them + statistics = 3 CTAs per SM
constexpr int kResThreads = 256;
constexpr int kResWarps = kResThreads / 32;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

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
// packed fp32 pairs (FADD2 / FFMA2, sm_100): the inner loops are issue bound, one instruction per TWO elements helps
__device__ __forceinline__ void unpack_u4_2(const uint4& v, float2* f) {
  unpack2(v.x, f[0].x, f[0].y); unpack2(v.y, f[1].x, f[1].y); unpack2(v.z, f[2].x, f[2].y); unpack2(v.w, f[3].x, f[3].y);
}
__device__ __forceinline__ uint4 pack_u4_2(const float2* f) {
  uint4 r;
  __nv_bfloat162 h;
  h = __float22bfloat162_rn(f[0]); r.x = *reinterpret_cast<uint32_t*>(&h);
  h = __float22bfloat162_rn(f[1]); r.y = *reinterpret_cast<uint32_t*>(&h);
  h = __float22bfloat162_rn(f[2]); r.z = *reinterpret_cast<uint32_t*>(&h);
  h = __float22bfloat162_rn(f[3]); r.w = *reinterpret_cast<uint32_t*>(&h);
  return r;
}
// eight consecutive floats (two 128-bit loads when the address allows)
__device__ __forceinline__ void load8f(const float* q, bool v4, float* o) {
  if (v4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(q)), b = __ldg(reinterpret_cast<const float4*>(q + 4));
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = __ldg(q + k);
  }
}

// NOTE (r2 experiment, rejected): replacing cluster.sync() below by  __syncthreads(); barrier.cluster.arrive.relaxed;
// barrier.cluster.wait.acquire  removes the MEMBAR.ALL.GPU that the .release arrive compiles to (it drains the cp.async prefetch
// of the next unit: -3 % on this kernel) - but it is WRONG on B200: peers' ld.shared::cluster then occasionally read stale
// partial sums (the 500-step loss-parity run diverged; tools/r2_optgraph_debug.py reproduces it in 4 of 5 runs).  The release
// at cluster scope is required for DSMEM visibility even when every STS of the CTA precedes a __syncthreads().

// Per-thread partial sums a[8], q[8] (channels cl..cl+7 of the chunk) -> wsum[warp][0][CC], wsum[warp][1][CC].
// Lanes with equal (lane % nvp) own the same channels (nvp is a power of two): xor-shuffle them together and let the
// first min(nvp, 32) lanes store - plain stores, no atomics.  When nvp > 32 a warp owns a fixed subset of the channels;
// the rest of its row was zeroed once at kernel start.
__device__ __forceinline__ void warp_partials(float* wsum, int CC, int nvp, bool active, int cl, float* a, float* q) {
  float* wrow = wsum + (threadIdx.x >> 5) * 2 * CC;
  for (int o = nvp; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] += __shfl_xor_sync(0xffffffffu, a[k], o); q[k] += __shfl_xor_sync(0xffffffffu, q[k], o); }
  }
  if (active && (int)(threadIdx.x & 31) < nvp) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { wrow[cl + k] = a[k]; wrow[CC + cl + k] = q[k]; }
  }
}

// Per-channel constants are computed ONCE per unit by the first CC threads and broadcast through shared memory (the
// first pipelined version recomputed them in every thread: ~380 of ~1300 instructions per thread and unit, ncu r1).
// smem: bf16 slab[2][per][CC] | float wsum[8][2][CC] | float gpart[2][2][32] | float ab[2][CC]
template <bool SILU>
__global__ void __launch_bounds__(kResThreads, 3) gn_fwd_pipe_kernel(const GnParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int per = p.per;
  const size_t slab_bytes = (size_t)per * p.CC * 2;
  float* wsum = reinterpret_cast<float*>(smraw + 2 * slab_bytes);
  float* gpart = wsum + kResWarps * 2 * p.CC;
  float* ab = gpart + 4 * kGroups;
  cg::cluster_group cluster = cg::this_cluster();
  const int cid = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int vec = threadIdx.x % p.nvp, row = threadIdx.x / p.nvp;     // nvp = nvec rounded up to a power of two
  const bool active = vec < p.nvec;
  const int cl = active ? vec * 8 : 0;
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;
  const int ng = p.CC / cpg;                                   // groups in a chunk
  const int RN = p.R * p.nvec;
  const int nit = active ? max(0, (p1 - p0 - row + p.R - 1) / p.R) : 0;   // pixels owned by this thread
  const int sidx = row * p.nvec + vec;                                 // this thread's first slab entry (stride RN)
  const float inv_n = 1.f / ((float)cpg * (float)p.HW);
  const bool chan_thread = (int)threadIdx.x < p.CC;                    // owns channel threadIdx.x of the chunk
  const int myg = chan_thread ? (int)threadIdx.x / cpg : 0;
  const float half = SILU ? 0.5f : 1.f;

  auto issue = [&](int u, int stage) {
    const int b = u / p.nchunk, c = (u % p.nchunk) * p.CC + cl;
    const bool in0 = c < p.C0;
    const int xpitch = in0 ? p.C0 : p.C1;
    const __nv_bfloat16* g = (in0 ? p.x0 + (size_t)b * p.HW * p.C0 + c : p.x1 + (size_t)b * p.HW * p.C1 + (c - p.C0)) +
                             (size_t)(p0 + row) * xpitch;
    const size_t gstep = (size_t)p.R * xpitch;
    uint4* s_ = reinterpret_cast<uint4*>(smraw + stage * slab_bytes) + sidx;
    for (int i = 0; i < nit; ++i) { cp_async16(s_, g); s_ += RN; g += gstep; }
  };

  for (int i = threadIdx.x; i < kResWarps * 2 * p.CC; i += blockDim.x) wsum[i] = 0.f;   // entries a warp never owns stay 0
  __syncthreads();
  if (cid < p.nunits) issue(cid, 0);
  cp_async_commit();
  int it = 0;
  for (int u = cid; u < p.nunits; u += p.ncl, ++it) {
    const int st = it & 1;
    if (u + p.ncl < p.nunits) issue(u + p.ncl, st ^ 1);
    cp_async_commit();
    const int b = u / p.nchunk, chunk = u - b * p.nchunk;
    // this thread's channel: affine constants that do not depend on the statistics (fetched while the slab is in flight)
    float gk = 0.f, hk = 0.f;
    if (chan_thread) {
      const int ch = chunk * p.CC + threadIdx.x;
      gk = __ldg(p.gamma + ch); hk = __ldg(p.beta + ch);
      if (p.film) {
        const float* fr = p.film + (size_t)b * p.film_ld + p.film_off;
        const float sc = __ldg(fr + ch), sh = __ldg(fr + p.C + ch);
        gk = fmaf(gk, sc, gk); hk = fmaf(hk, sc, hk) + sh;
      }
    }
    cp_async_wait1();
    __syncthreads();

    const uint4* sl = reinterpret_cast<const uint4*>(smraw + st * slab_bytes) + sidx;
    float s[8], ss[8];
    {
      float2 s2[4], q2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { s2[k] = make_float2(0.f, 0.f); q2[k] = make_float2(0.f, 0.f); }
      const uint4* s_ = sl;
#pragma unroll 4
      for (int i = 0; i < nit; ++i) {
        float2 f[4];
        unpack_u4_2(*s_, f);
        s_ += RN;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s2[k] = __fadd2_rn(s2[k], f[k]); q2[k] = __ffma2_rn(f[k], f[k], q2[k]); }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[2 * k] = s2[k].x; s[2 * k + 1] = s2[k].y; ss[2 * k] = q2[k].x; ss[2 * k + 1] = q2[k].y; }
    }
    warp_partials(wsum, p.CC, p.nvp, active, cl, s, ss);
    __syncthreads();
    float* gp = gpart + st * 2 * kGroups;
    if ((int)threadIdx.x < ng) {
      float a = 0.f, q = 0.f;
      for (int w = 0; w < kResWarps; ++w) {
        const float* wr = wsum + w * 2 * p.CC + threadIdx.x * cpg;
        for (int k = 0; k < cpg; ++k) { a += wr[k]; q += wr[p.CC + k]; }
      }
      gp[threadIdx.x] = a; gp[kGroups + threadIdx.x] = q;
    }
    if (p.S > 1) cluster.sync(); else __syncthreads();
    if (chan_thread) {
      float a = 0.f, q = 0.f;
      if (p.S > 1) {
        for (int r = 0; r < p.S; ++r) {
          const float* rp = cluster.map_shared_rank(gp, r);
          a += rp[myg]; q += rp[kGroups + myg];
        }
      } else { a = gp[myg]; q = gp[kGroups + myg]; }
      const float m = a * inv_n;
      const float var = fmaxf(q * inv_n - m * m, 0.f);
      const float rs = rsqrtf(var + 1e-5f);
      ab[threadIdx.x] = half * rs * gk;
      ab[p.CC + threadIdx.x] = half * (hk - m * rs * gk);
      if (rank == 0 && (int)threadIdx.x == myg * cpg) {
        const int g = chunk * ng + myg;
        p.mean[b * kGroups + g] = m; p.rstd[b * kGroups + g] = rs;
      }
    }
    __syncthreads();

    if (active) {
      float2 A[4], Bc[4];
      *reinterpret_cast<float4*>(A) = *reinterpret_cast<const float4*>(ab + cl);
      *reinterpret_cast<float4*>(A + 2) = *reinterpret_cast<const float4*>(ab + cl + 4);
      *reinterpret_cast<float4*>(Bc) = *reinterpret_cast<const float4*>(ab + p.CC + cl);
      *reinterpret_cast<float4*>(Bc + 2) = *reinterpret_cast<const float4*>(ab + p.CC + cl + 4);
      __nv_bfloat16* y = p.y + (size_t)b * p.HW * p.C + chunk * p.CC + cl + (size_t)(p0 + row) * p.C;
      const size_t ystep = (size_t)p.R * p.C;
      const uint4* s_ = sl;
#pragma unroll 2
      for (int i = 0; i < nit; ++i) {
        float2 f[4];
        unpack_u4_2(*s_, f);
        s_ += RN;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 h = __ffma2_rn(f[k], A[k], Bc[k]);
          f[k] = SILU ? __ffma2_rn(h, tanh2(h), h) : h;
        }
        *reinterpret_cast<uint4*>(y) = pack_u4_2(f);
        y += ystep;
      }
    }
    __syncthreads();      // slab st and ab are free for the next iteration
  }
  if (p.S > 1) cluster.sync();   // no CTA exits while a peer may still read its gpart
}

// smem: bf16 du slab[2][per][CC] | float wsum[8][2][CC] | float pub[2][2][CC] | float tot[2][CC] | float cst[6][CC]
// With u = x*aG + bH (aG = rstd*G, bH = -mean*rstd*G + Hh), du = dy*silu'(u), P = sum du, Qx = sum du*x:
//   Q = sum du*xhat = rstd*Qx - mean*rstd*P ;  dx = K1*du - K2' - x*K3'  with K1 = rstd*G, K3' = rstd^2*s2/n,
//   K2' = rstd*s1/n - mean*rstd^2*s2/n.
// dy is prefetched into the ring (cp.async) and overwritten in place by du; x is streamed through registers in the
// reduction pass and read again in the apply pass (an L2 hit: the same CTA touched it microseconds earlier).
// cst rows: 0 aG/2 (aG when !SILU) | 1 bH/2 | 2 G (gamma*(1+scale)) | 3 K1 | 4 -K2' | 5 -K3'
template <bool SILU>
__global__ void __launch_bounds__(kResThreads, 3) gn_bwd_pipe_kernel(const GnParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int per = p.per;
  const size_t slab_bytes = (size_t)per * p.CC * 2;
  float* wsum = reinterpret_cast<float*>(smraw + 2 * slab_bytes);
  float* pub = wsum + kResWarps * 2 * p.CC;
  float* tot = pub + 4 * p.CC;
  float* cst = tot + 2 * p.CC;
  cg::cluster_group cluster = cg::this_cluster();
  const int cid = blockIdx.x / p.S, rank = blockIdx.x % p.S;
  const int vec = threadIdx.x % p.nvp, row = threadIdx.x / p.nvp;     // nvp = nvec rounded up to a power of two
  const bool active = vec < p.nvec;
  const int cl = active ? vec * 8 : 0;
  const int p0 = rank * per, p1 = min(p.HW, p0 + per);
  const int cpg = p.C / kGroups;
  const int ng = p.CC / cpg;
  const int RN = p.R * p.nvec;
  const int nit = active ? max(0, (p1 - p0 - row + p.R - 1) / p.R) : 0;
  const int sidx = row * p.nvec + vec;                                 // this thread's first slab entry (stride RN)
  const float inv_n = 1.f / ((float)cpg * (float)p.HW);
  const size_t dstep = (size_t)p.R * p.C;
  const bool hasadd = p.dadd != nullptr;
  const bool chan_thread = (int)threadIdx.x < p.CC;
  const int myg = chan_thread ? (int)threadIdx.x / cpg : 0;
  const float half = SILU ? 0.5f : 1.f;

  auto issue = [&](int u, int stage) {
    const int b = u / p.nchunk, c = (u % p.nchunk) * p.CC + cl;
    const __nv_bfloat16* g = p.dy + (size_t)b * p.HW * p.C + c + (size_t)(p0 + row) * p.C;
    uint4* s_ = reinterpret_cast<uint4*>(smraw + stage * slab_bytes) + sidx;
    for (int i = 0; i < nit; ++i) { cp_async16(s_, g); s_ += RN; g += dstep; }
  };

  for (int i = threadIdx.x; i < kResWarps * 2 * p.CC; i += blockDim.x) wsum[i] = 0.f;   // entries a warp never owns stay 0
  __syncthreads();
  if (cid < p.nunits) issue(cid, 0);
  cp_async_commit();
  int it = 0;
  for (int u = cid; u < p.nunits; u += p.ncl, ++it) {
    const int st = it & 1;
    if (u + p.ncl < p.nunits) issue(u + p.ncl, st ^ 1);
    cp_async_commit();
    const int b = u / p.nchunk, chunk = u - b * p.nchunk;
    const int c0 = chunk * p.CC;
    const int c = c0 + cl;
    // this thread's channel
    float G = 0.f, m = 0.f, rs = 0.f, sc1 = 1.f, ga = 0.f, be = 0.f;
    if (chan_thread) {
      const int ch = c0 + threadIdx.x;
      ga = __ldg(p.gamma + ch); be = __ldg(p.beta + ch);
      float sh = 0.f;
      if (p.film) {
        const float* fr = p.film + (size_t)b * p.film_ld + p.film_off;
        sc1 = 1.f + __ldg(fr + ch); sh = __ldg(fr + p.C + ch);
      }
      G = ga * sc1;
      const float Hh = be * sc1 + sh;
      const int g = chunk * ng + myg;
      m = p.mean[b * kGroups + g]; rs = p.rstd[b * kGroups + g];
      cst[threadIdx.x] = half * rs * G;
      cst[p.CC + threadIdx.x] = half * (Hh - m * rs * G);
      cst[2 * p.CC + threadIdx.x] = G;
    }
    const bool in0 = c < p.C0;
    const int xpitch = in0 ? p.C0 : p.C1;
    const size_t xstep = (size_t)p.R * xpitch;
    const __nv_bfloat16* const xg0 = (in0 ? p.x0 + (size_t)b * p.HW * p.C0 + c : p.x1 + (size_t)b * p.HW * p.C1 + (c - p.C0)) +
                                     (size_t)(p0 + row) * xpitch;
    uint4* const sl = reinterpret_cast<uint4*>(smraw + st * slab_bytes) + sidx;

    // the first x loads go out before the wait: they do not depend on the ring
    uint4 vxa = make_uint4(0u, 0u, 0u, 0u), vxb = vxa;
    if (nit > 0) vxa = __ldg(reinterpret_cast<const uint4*>(xg0));
    if (nit > 1) vxb = __ldg(reinterpret_cast<const uint4*>(xg0 + xstep));
    cp_async_wait1();
    __syncthreads();
    float P[8], Qx[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { P[k] = 0.f; Qx[k] = 0.f; }
    if (active) {
      float2 aGh[4], bHh[4];      // halved affine: h = u/2 = x*aGh + bHh
      if (SILU) {
        *reinterpret_cast<float4*>(aGh) = *reinterpret_cast<const float4*>(cst + cl);
        *reinterpret_cast<float4*>(aGh + 2) = *reinterpret_cast<const float4*>(cst + cl + 4);
        *reinterpret_cast<float4*>(bHh) = *reinterpret_cast<const float4*>(cst + p.CC + cl);
        *reinterpret_cast<float4*>(bHh + 2) = *reinterpret_cast<const float4*>(cst + p.CC + cl + 4);
      }
      float2 P2[4], Q2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { P2[k] = make_float2(0.f, 0.f); Q2[k] = make_float2(0.f, 0.f); }
      const float2 kM1 = make_float2(-1.f, -1.f), kHalf = make_float2(0.5f, 0.5f);
      const __nv_bfloat16* xg = xg0 + 2 * xstep;
      uint4* s_ = sl;
      auto one = [&](const uint4& vx, uint4* dst) {
        float2 f[4], d[4];
        unpack_u4_2(vx, f); unpack_u4_2(*dst, d);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (SILU) {
            const float2 h = __ffma2_rn(f[k], aGh[k], bHh[k]);
            const float2 t = tanh2(h);
            const float2 q = __ffma2_rn(t, t, kM1);                        // t^2 - 1
            const float2 w = __ffma2_rn(__fmul2_rn(h, kM1), q, t);         // t + h*(1 - t^2)
            d[k] = __fmul2_rn(d[k], __ffma2_rn(w, kHalf, kHalf));
          }
          P2[k] = __fadd2_rn(P2[k], d[k]); Q2[k] = __ffma2_rn(d[k], f[k], Q2[k]);
        }
        if (SILU) *dst = pack_u4_2(d);
      };
      int i = 0;
      for (; i + 2 <= nit; i += 2) {       // software pipeline: the loads of pixels i+2, i+3 fly while i, i+1 are reduced
        uint4 nxa = vxa, nxb = vxb;
        if (i + 2 < nit) nxa = __ldg(reinterpret_cast<const uint4*>(xg));
        if (i + 3 < nit) nxb = __ldg(reinterpret_cast<const uint4*>(xg + xstep));
        one(vxa, s_); one(vxb, s_ + RN);
        vxa = nxa; vxb = nxb;
        xg += 2 * xstep; s_ += 2 * RN;
      }
      if (i < nit) one(vxa, s_);
#pragma unroll
      for (int k = 0; k < 4; ++k) { P[2 * k] = P2[k].x; P[2 * k + 1] = P2[k].y; Qx[2 * k] = Q2[k].x; Qx[2 * k + 1] = Q2[k].y; }
    }
    warp_partials(wsum, p.CC, p.nvp, active, cl, P, Qx);
    __syncthreads();
    float* pb = pub + st * 2 * p.CC;
    for (int i = threadIdx.x; i < 2 * p.CC; i += blockDim.x) {
      float a = 0.f;
      for (int w = 0; w < kResWarps; ++w) a += wsum[w * 2 * p.CC + i];
      pb[i] = a;
    }
    if (p.S > 1) cluster.sync(); else __syncthreads();
    for (int i = threadIdx.x; i < 2 * p.CC; i += blockDim.x) {
      float a = 0.f;
      if (p.S > 1) { for (int r = 0; r < p.S; ++r) a += cluster.map_shared_rank(pb, r)[i]; }
      else a = pb[i];
      tot[i] = a;
    }
    __syncthreads();
    // tot[0][c] = P_c, tot[1][c] = sum du*x ; Q_c = rstd*(Qx_c - mean*P_c); every channel thread folds its own group
    if (chan_thread) {
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < cpg; ++k) {
        const int lc = myg * cpg + k;
        const float kc = cst[2 * p.CC + lc];
        const float Pc = tot[lc], Qc = rs * (tot[p.CC + lc] - m * Pc);
        s1 = fmaf(kc, Pc, s1); s2 = fmaf(kc, Qc, s2);
      }
      const float K3 = rs * rs * s2 * inv_n;
      cst[3 * p.CC + threadIdx.x] = rs * G;
      cst[4 * p.CC + threadIdx.x] = -(rs * s1 * inv_n - m * K3);     // stored negated: dx = K1*du + (x*(-K3') + (-K2'))
      cst[5 * p.CC + threadIdx.x] = -K3;
      if (rank == 0) {
        const int ch = c0 + threadIdx.x;
        const float Pc = tot[threadIdx.x], Qc = rs * (tot[p.CC + threadIdx.x] - m * Pc);
        if (p.film && p.dfilm) {
          const size_t fo = (size_t)b * p.film_ld + p.film_off;
          p.dfilm[fo + ch] += ga * Qc + be * Pc;        // d scale  (columns owned by this layer & sample)
          p.dfilm[fo + p.C + ch] += Pc;                 // d shift
        }
        if (p.dgamma) atomicAdd(p.dgamma + ch, sc1 * Qc);
        if (p.dbeta) atomicAdd(p.dbeta + ch, sc1 * Pc);
      }
    }
    __syncthreads();

    if (active) {
      float2 K1[4], K2[4], K3[4];
      *reinterpret_cast<float4*>(K1) = *reinterpret_cast<const float4*>(cst + 3 * p.CC + cl);
      *reinterpret_cast<float4*>(K1 + 2) = *reinterpret_cast<const float4*>(cst + 3 * p.CC + cl + 4);
      *reinterpret_cast<float4*>(K2) = *reinterpret_cast<const float4*>(cst + 4 * p.CC + cl);
      *reinterpret_cast<float4*>(K2 + 2) = *reinterpret_cast<const float4*>(cst + 4 * p.CC + cl + 4);
      *reinterpret_cast<float4*>(K3) = *reinterpret_cast<const float4*>(cst + 5 * p.CC + cl);
      *reinterpret_cast<float4*>(K3 + 2) = *reinterpret_cast<const float4*>(cst + 5 * p.CC + cl + 4);
      const bool acc = (p.accumulate_dx >> (in0 ? 0 : 1)) & 1;
      const __nv_bfloat16* xg = xg0;
      __nv_bfloat16* og = (in0 ? p.dx0 + (size_t)b * p.HW * p.C0 + c : p.dx1 + (size_t)b * p.HW * p.C1 + (c - p.C0)) +
                          (size_t)(p0 + row) * xpitch;
      const __nv_bfloat16* ag = hasadd ? p.dadd + (size_t)b * p.HW * p.C + c + (size_t)(p0 + row) * p.C : nullptr;
      const uint4* s_ = sl;
      auto one = [&](const uint4& vx, const uint4& vdu, const uint4& vo, const uint4& va, __nv_bfloat16* dst) {
        float2 f[4], d[4], o[4];
        unpack_u4_2(vx, f); unpack_u4_2(vdu, d);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = __ffma2_rn(K1[k], d[k], __ffma2_rn(f[k], K3[k], K2[k]));   // K2, K3 are negated
        if (acc) {
          float2 t[4];
          unpack_u4_2(vo, t);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = __fadd2_rn(o[k], t[k]);
        }
        if (hasadd) {
          float2 t[4];
          unpack_u4_2(va, t);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = __fadd2_rn(o[k], t[k]);
        }
        *reinterpret_cast<uint4*>(dst) = pack_u4_2(o);
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
    __syncthreads();      // slab st, tot and cst are free for the next iteration
  }
  if (p.S > 1) cluster.sync();   // no CTA exits while a peer may still read its published sums
}


// ------------------------------------------------------------------------------------------------ streaming forward
// When the producing convolution's epilogue has already accumulated per-(image, channel) sum / sum of squares
// (cdae_igemm_desc.stats), the forward GroupNorm needs no reduction pass and nothing resident on chip: every CTA
// rebuilds the group statistics of its sample from <= 1024 channel sums (an L2 hit), folds gamma / beta / FiLM into two
// per-channel constants in shared memory and then streams its pixel range once: 128-bit loads, four rows in flight per
// thread, packed FFMA2, one MUFU per element (tanh form of SiLU), 128-bit stores.  2 B read + 2 B written per element,
// no clusters, no barriers in the streaming loop.
struct GnApplyParams {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  const float* st0; const float* st1;      // fp32 [B][C0][2], [B][C1][2]
  int C0, C1, C, HW;
  const float* gamma; const float* beta; const float* film; int film_ld, film_off;
  __nv_bfloat16* y; float* mean; float* rstd;
  float* ab;                                // optional fp32 [B][C][2]: the per-(image, channel) constants {a, b} with
                                            // u (or u/2 under SiLU) = a x + b, kept for the fused backward (igemm gnb_ab)
  int ppc;                                  // pixels per CTA
  int nvec, R;                              // 16 B vectors per pixel row; pixel rows per CTA pass (threads >= nvec*R idle)
};
constexpr int kApplyThreads = 256;
constexpr int kApplyUnroll = 4;

template <bool SILU>
__global__ void __launch_bounds__(kApplyThreads, 3) gn_apply_fwd_kernel(const GnApplyParams p) {
  extern __shared__ __align__(16) float sm_apply[];
  float* cs = sm_apply;                 // [C] channel sums   -> later A (scale)
  float* cq = sm_apply + p.C;           // [C] channel sumsq  -> later B (shift)
  float* gm = sm_apply + 2 * p.C;       // [32] group mean
  float* gr = gm + kGroups;             // [32] group rstd
  const int b = blockIdx.y;
  const int cpg = p.C / kGroups;
  // streaming geometry first: the first rows are requested BEFORE the statistics prologue, so that their DRAM latency
  // hides the three dependent L2 round trips (channel sums -> group statistics -> per-channel constants)
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const bool active = row < p.R;
  const int cl = vec * 8;
  const bool in0 = cl < p.C0;
  const int xpitch = in0 ? p.C0 : p.C1;
  const __nv_bfloat16* xb = in0 ? p.x0 + (size_t)b * p.HW * p.C0 + cl : p.x1 + (size_t)b * p.HW * p.C1 + (cl - p.C0);
  __nv_bfloat16* yb = p.y + (size_t)b * p.HW * p.C + cl;
  const int p0 = blockIdx.x * p.ppc, p1 = min(p.HW, p0 + p.ppc);
  uint4 v[kApplyUnroll];
  int pix = p0 + row;
  if (active) {
#pragma unroll
    for (int k = 0; k < kApplyUnroll; ++k) {
      const int pk = pix + k * p.R;
      if (pk < p1) v[k] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)pk * xpitch));
    }
  }
  for (int c = threadIdx.x; c < p.C; c += kApplyThreads) {
    const float2 t = c < p.C0 ? __ldg(reinterpret_cast<const float2*>(p.st0) + (size_t)b * p.C0 + c)
                              : __ldg(reinterpret_cast<const float2*>(p.st1) + (size_t)b * p.C1 + (c - p.C0));
    cs[c] = t.x; cq[c] = t.y;
  }
  __syncthreads();
  if (threadIdx.x < kGroups) {
    float a = 0.f, q = 0.f;
    for (int k = 0; k < cpg; ++k) { a += cs[threadIdx.x * cpg + k]; q += cq[threadIdx.x * cpg + k]; }
    const float inv_n = 1.f / ((float)cpg * (float)p.HW);
    const float m = a * inv_n;
    const float var = fmaxf(q * inv_n - m * m, 0.f);
    const float rs = rsqrtf(var + 1e-5f);
    gm[threadIdx.x] = m; gr[threadIdx.x] = rs;
    if (blockIdx.x == 0) { p.mean[b * kGroups + threadIdx.x] = m; p.rstd[b * kGroups + threadIdx.x] = rs; }
  }
  __syncthreads();
  const float half = SILU ? 0.5f : 1.f;
  for (int c = threadIdx.x; c < p.C; c += kApplyThreads) {
    float gk = __ldg(p.gamma + c), hk = __ldg(p.beta + c);
    if (p.film) {
      const float* fr = p.film + (size_t)b * p.film_ld + p.film_off;
      const float sc = __ldg(fr + c), sh = __ldg(fr + p.C + c);
      gk = fmaf(gk, sc, gk); hk = fmaf(hk, sc, hk) + sh;
    }
    const int g = c / cpg;
    const float m = gm[g], rs = gr[g];
    cs[c] = half * rs * gk;
    cq[c] = half * (hk - m * rs * gk);
    if (p.ab && blockIdx.x == 0)
      *reinterpret_cast<float2*>(p.ab + ((size_t)b * p.C + c) * 2) = make_float2(cs[c], cq[c]);
  }
  __syncthreads();
  if (!active) return;
  float2 A[4], Bc[4];
  *reinterpret_cast<float4*>(A) = *reinterpret_cast<const float4*>(cs + cl);
  *reinterpret_cast<float4*>(A + 2) = *reinterpret_cast<const float4*>(cs + cl + 4);
  *reinterpret_cast<float4*>(Bc) = *reinterpret_cast<const float4*>(cq + cl);
  *reinterpret_cast<float4*>(Bc + 2) = *reinterpret_cast<const float4*>(cq + cl + 4);
  const int step = kApplyUnroll * p.R;
  for (; pix < p1; pix += step) {
    // rows of the NEXT batch are requested before this batch is normalised and stored
    uint4 nx[kApplyUnroll];
#pragma unroll
    for (int k = 0; k < kApplyUnroll; ++k) {
      const int pk = pix + step + k * p.R;
      if (pk < p1) nx[k] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)pk * xpitch));
    }
#pragma unroll
    for (int k = 0; k < kApplyUnroll; ++k) {
      const int pk = pix + k * p.R;
      if (pk < p1) {
        float2 f[4];
        unpack_u4_2(v[k], f);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 h = __ffma2_rn(f[e], A[e], Bc[e]);
          f[e] = SILU ? __ffma2_rn(h, tanh2(h), h) : h;
        }
        *reinterpret_cast<uint4*>(yb + (size_t)pk * p.C) = pack_u4_2(f);
      }
    }
#pragma unroll
    for (int k = 0; k < kApplyUnroll; ++k) v[k] = nx[k];
  }
}

// ------------------------------------------------------------------------------------------------ streaming backward
// The data-gradient convolution that produces the gradient w.r.t. the GroupNorm OUTPUT has already (cdae_igemm_desc.gnb_*)
// turned it into du = dy * silu'(u) and accumulated P_c = sum du, Qx_c = sum du*x per (image, channel) in its epilogue - the
// backward twin of the forward statistics epilogue.  What is left is a pure streaming pass with the shape of
// gn_apply_fwd_kernel: every CTA folds the <= 1024 channel sums of its sample into three per-channel constants and streams
// its pixel range once:   dx = K1*du - K2' - x*K3'  (+ dadd) (+ old dx)
//   G = gamma (1 + scale), Qc = rstd (Qx_c - mean P_c), s1 = sum_group G P, s2 = sum_group G Qc, n = HW * C/32
//   K1 = rstd G,  K3' = rstd^2 s2 / n,  K2' = rstd s1 / n - mean K3'
// The first pixel chunk of every sample also emits d gamma / d beta / d FiLM.  2 B (du) + 2 B (x) read, 2 B written per
// element; no clusters, no reduction, no second read.
struct GnBwdParams {
  const __nv_bfloat16* du; const __nv_bfloat16* x0; const __nv_bfloat16* x1; const __nv_bfloat16* dadd;
  __nv_bfloat16* dx0; __nv_bfloat16* dx1;
  int C0, C1, C, HW, B;
  const float* gamma; const float* beta; const float* film; int film_ld, film_off;
  const float* mean; const float* rstd;
  const float* ws;                          // [B][C][2] = {P_c, Qx_c}
  float* dgamma; float* dbeta; float* dfilm;
  int accumulate_dx;
  int ppc, nvec, R;
};
constexpr int kBwdThreads = 256;
constexpr int kBwdUnroll = 2;

// smem: float K1[C] | float K2n[C] | float K3n[C] | float G[C] | float P[C] | float Qc[C]
__global__ void __launch_bounds__(kBwdThreads, 3) gn_bwd_apply_kernel(const GnBwdParams p) {
  extern __shared__ __align__(16) float sm_bwd[];
  float* K1s = sm_bwd;
  float* K2s = K1s + p.C;
  float* K3s = K2s + p.C;
  float* Gs = K3s + p.C;
  float* Ps = Gs + p.C;
  float* Qs = Ps + p.C;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int cpg = p.C / kGroups;
  const float inv_n = 1.f / ((float)cpg * (float)p.HW);
  const int vec = threadIdx.x % p.nvec, row = threadIdx.x / p.nvec;
  const bool active = row < p.R;
  const int cl = vec * 8;
  const bool in0 = cl < p.C0;
  const int xpitch = in0 ? p.C0 : p.C1;
  const __nv_bfloat16* xb = in0 ? p.x0 + (size_t)b * p.HW * p.C0 + cl : p.x1 + (size_t)b * p.HW * p.C1 + (cl - p.C0);
  const __nv_bfloat16* db = p.du + (size_t)b * p.HW * p.C + cl;
  const int p0 = chunk * p.ppc, p1 = min(p.HW, p0 + p.ppc);
  int pix = p0 + row;
  // the first rows are requested before the constant prologue: their DRAM latency hides its dependent L2 round trips
  uint4 vx[kBwdUnroll], vd[kBwdUnroll];
  if (active) {
#pragma unroll
    for (int k = 0; k < kBwdUnroll; ++k) {
      const int pk = pix + k * p.R;
      if (pk < p1) {
        vx[k] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)pk * xpitch));
        vd[k] = __ldg(reinterpret_cast<const uint4*>(db + (size_t)pk * p.C));
      }
    }
  }
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(p.ws) + (size_t)b * p.C + c);
    const int g = c / cpg;
    const float m = __ldg(p.mean + b * kGroups + g), rs = __ldg(p.rstd + b * kGroups + g);
    float G = __ldg(p.gamma + c);
    if (p.film) G *= 1.f + __ldg(p.film + (size_t)b * p.film_ld + p.film_off + c);
    Gs[c] = G; Ps[c] = t.x; Qs[c] = rs * (t.y - m * t.x);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int g = c / cpg;
    const float m = __ldg(p.mean + b * kGroups + g), rs = __ldg(p.rstd + b * kGroups + g);
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      const int lc = g * cpg + k;
      s1 = fmaf(Gs[lc], Ps[lc], s1); s2 = fmaf(Gs[lc], Qs[lc], s2);
    }
    const float K3 = rs * rs * s2 * inv_n;
    K1s[c] = rs * Gs[c];
    K2s[c] = -(rs * s1 * inv_n - m * K3);     // stored negated: dx = K1*du + (x*(-K3') + (-K2'))
    K3s[c] = -K3;
    if (chunk == 0) {                         // one CTA per sample owns the parameter gradients
      const float Pc = Ps[c], Qc = Qs[c];
      const float ga = __ldg(p.gamma + c), be = __ldg(p.beta + c);
      float sc1 = 1.f;
      if (p.film) {
        const size_t fo = (size_t)b * p.film_ld + p.film_off;
        sc1 = 1.f + __ldg(p.film + fo + c);
        if (p.dfilm) {
          p.dfilm[fo + c] += ga * Qc + be * Pc;       // d scale  (columns owned by this layer & sample)
          p.dfilm[fo + p.C + c] += Pc;                // d shift
        }
      }
      if (p.dgamma) atomicAdd(p.dgamma + c, sc1 * Qc);
      if (p.dbeta) atomicAdd(p.dbeta + c, sc1 * Pc);
    }
  }
  __syncthreads();
  if (!active) return;
  float2 K1[4], K2[4], K3[4];
#define LD8(dst, src) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>((src) + cl); \
                      *reinterpret_cast<float4*>((dst) + 2) = *reinterpret_cast<const float4*>((src) + cl + 4)
  LD8(K1, K1s); LD8(K2, K2s); LD8(K3, K3s);
#undef LD8
  const bool acc = (p.accumulate_dx >> (in0 ? 0 : 1)) & 1;
  const bool hasadd = p.dadd != nullptr;
  __nv_bfloat16* ob = in0 ? p.dx0 + (size_t)b * p.HW * p.C0 + cl : p.dx1 + (size_t)b * p.HW * p.C1 + (cl - p.C0);
  const __nv_bfloat16* ab = hasadd ? p.dadd + (size_t)b * p.HW * p.C + cl : nullptr;
  const int step = kBwdUnroll * p.R;
  for (; pix < p1; pix += step) {
    uint4 nx[kBwdUnroll], nd[kBwdUnroll], vo[kBwdUnroll], va[kBwdUnroll];
#pragma unroll
    for (int k = 0; k < kBwdUnroll; ++k) {
      const int pk = pix + step + k * p.R;
      if (pk < p1) {
        nx[k] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)pk * xpitch));
        nd[k] = __ldg(reinterpret_cast<const uint4*>(db + (size_t)pk * p.C));
      }
      const int pc = pix + k * p.R;
      if (pc < p1) {
        if (acc) vo[k] = *reinterpret_cast<const uint4*>(ob + (size_t)pc * xpitch);
        if (hasadd) va[k] = __ldg(reinterpret_cast<const uint4*>(ab + (size_t)pc * p.C));
      }
    }
#pragma unroll
    for (int k = 0; k < kBwdUnroll; ++k) {
      const int pc = pix + k * p.R;
      if (pc < p1) {
        float2 f[4], d[4], o[4];
        unpack_u4_2(vx[k], f); unpack_u4_2(vd[k], d);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = __ffma2_rn(K1[e], d[e], __ffma2_rn(f[e], K3[e], K2[e]));       // K2, K3 are negated
        if (acc) {
          float2 t[4];
          unpack_u4_2(vo[k], t);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = __fadd2_rn(o[e], t[e]);
        }
        if (hasadd) {
          float2 t[4];
          unpack_u4_2(va[k], t);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = __fadd2_rn(o[e], t[e]);
        }
        *reinterpret_cast<uint4*>(ob + (size_t)pc * xpitch) = pack_u4_2(o);
      }
    }
#pragma unroll
    for (int k = 0; k < kBwdUnroll; ++k) { vx[k] = nx[k]; vd[k] = nd[k]; }
  }
}

static inline size_t gn_pipe_smem(const GnParams& p, bool bwd) {
  const size_t slabs = 2 * (size_t)p.per * p.CC * 2;
  const size_t stats = bwd ? sizeof(float) * ((size_t)kResWarps * 2 * p.CC + 4 * p.CC + 2 * p.CC + 6 * p.CC)
                           : sizeof(float) * ((size_t)kResWarps * 2 * p.CC + 4 * kGroups + 2 * p.CC);
  return slabs + stats;
}

// How many clusters of S CTAs (kResThreads threads, `smem` dynamic bytes) can be resident at once.  The persistent grid
// must not exceed it: a cluster that is not co-resident would only start after another one has finished ALL its units.
static int gn_max_clusters(const void* kernel, int S, size_t smem) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, size_t>, int> cache;
  const size_t key_smem = (smem + 1023) / 1024 * 1024;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(kernel, S, key_smem);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(kNumSMs * 4 / S * S));
  cfg.blockDim = dim3((unsigned)kResThreads);
  cfg.dynamicSmemBytes = key_smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = kNumSMs / S / 2;                            // conservative guess: one CTA per SM, half the chip
    if (n < 1) n = 1;
  }
  cache[key] = n;
  return n;
}

static inline int pow2_ceil_i(int v) { int q = 1; while (q < v) q <<= 1; return q; }

// Pick (CC, S) for the pipelined kernels by a small cost model (rounds of the persistent grid x work per round, with
// penalties for cluster barriers, rows that are not whole 32 B sectors and idle lanes).  False: shape does not fit.
static bool gn_pipe_config(GnParams& p, int B, bool bwd, const void* kernel) {
  p.C = p.C0 + p.C1;
  if (p.C % kGroups || p.C0 % 8 || p.C1 % 8) return false;
  const int cpg = p.C / kGroups;
  int base = cpg;                                   // lcm(8, cpg)
  while (base % 8) base += cpg;
  double best = 1e30;
  int bestCC = 0, bestS = 1, bestNcl = 1;
  for (int m = base; m <= p.C; m += base) {
    const int nvec = m / 8, nvp = pow2_ceil_i(nvec);
    if (p.C % m || nvp > kResThreads / 4 || m / cpg > kGroups) continue;       // >= 4 pixel rows per pass
    int S = 1;
    while (S <= 8 && (int64_t)((p.HW + S - 1) / S) * m > kPipeElems) S <<= 1;
    if (S > 8) continue;
    const int per = (p.HW + S - 1) / S;
    GnParams t = p; t.CC = m; t.per = per;
    const size_t smem = gn_pipe_smem(t, bwd);
    if (smem > 200 * 1024) continue;
    const int64_t units = (int64_t)B * (p.C / m);
    int64_t ncl = gn_max_clusters(kernel, S, smem);
    if (ncl > units) ncl = units;
    const int64_t rounds = (units + ncl - 1) / ncl;
    double cost = (double)rounds * ((double)per * m + 3000.0 + (S > 1 ? 3000.0 : 0.0));
    cost *= (double)nvp / nvec;
    if ((m * 2) % 32) cost *= 1.3;
    static const int min_row = getenv("CDAE_GN_MIN_ROW") ? atoi(getenv("CDAE_GN_MIN_ROW")) : 64;
    if (m * 2 < min_row) cost *= 1.25;
    if (cost < best) { best = cost; bestCC = m; bestS = S; bestNcl = (int)ncl; }
  }
  if (!bestCC) return false;
  p.CC = bestCC; p.nchunk = p.C / bestCC; p.S = bestS; p.ncl = bestNcl;
  p.per = (p.HW + bestS - 1) / bestS;
  p.nunits = B * p.nchunk;
  p.nvec = bestCC / 8; p.nvp = pow2_ceil_i(p.nvec); p.R = kResThreads / p.nvp;
  static const bool dbg = getenv("CDAE_GN_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "gn %s B%d C%d HW%d: CC %d S %d per %d units %d clusters %d smem %zu\n", bwd ? "bwd" : "fwd", B, p.C, p.HW,
                   p.CC, p.S, p.per, p.nunits, p.ncl, gn_pipe_smem(p, bwd));
  return true;
}

template <typename K>
static int gn_pipe_launch(K kernel, const GnParams& p, size_t smem, cudaStream_t st, const char* name) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(p.ncl * p.S));
  cfg.blockDim = dim3((unsigned)kResThreads);
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

static inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

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
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(x0 && gamma && beta && y && mean && rstd && (C1 == 0 || x1), "gn_fwd: null pointer");
  GnParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1; p.C0 = C0; p.C1 = C1; p.HW = HW;
  p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld; p.film_off = film_off; p.silu = silu;
  p.y = (__nv_bfloat16*)y; p.mean = mean; p.rstd = rstd;
  {
    auto kern = silu ? gn_fwd_pipe_kernel<true> : gn_fwd_pipe_kernel<false>;
    if (gn_pipe_config(p, B, false, reinterpret_cast<const void*>(kern))) {
      p.v4 = aligned16(gamma) && aligned16(beta) && (!film || (aligned16(film) && film_ld % 4 == 0 && film_off % 4 == 0 && p.C % 4 == 0));
      return gn_pipe_launch(kern, p, gn_pipe_smem(p, false), (cudaStream_t)s, "gn_fwd_pipe_kernel");
    }
  }
  int S;
  const bool wide = (C0 + C1) > 4 * kGnThreads;     // > 1024 channels: 8-channel vectors keep nvec <= 256
  int rc = wide ? gn_config<8>(p, &S) : gn_config<4>(p, &S);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (2 * p.C + 4 * kGroups);
  if (wide) return gn_launch(gn_fwd_kernel<8, 4>, p, B, S, smem, (cudaStream_t)s, "gn_fwd_kernel");
  return gn_launch(gn_fwd_kernel<4, 4>, p, B, S, smem, (cudaStream_t)s, "gn_fwd_kernel");
}

extern "C" int cdae_gn_apply_fwd(const void* x0, int C0, const float* stats0, const void* x1, int C1, const float* stats1,
                                 int B, int HW, const float* gamma, const float* beta, const float* film, int film_ld,
                                 int film_off, int silu, void* y, float* mean, float* rstd, float* ab, cdae_stream s) {
  if (B == 0 || HW == 0) return CDAE_OK;
  // y == nullptr: constants only - the {a, b} table (and mean / rstd) for a consumer that applies the norm while it loads its
  // operand (cdae_igemm_desc.gn_ab); one CTA per image, nothing is streamed
  CDAE_CHECK_ARG(x0 && stats0 && gamma && beta && (y || ab) && mean && rstd && (C1 == 0 || (x1 && stats1)), "gn_apply_fwd: null pointer");
  CDAE_CHECK_ARG((reinterpret_cast<uintptr_t>(ab) & 7) == 0, "gn_apply_fwd: misaligned constant table");
  GnApplyParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1; p.st0 = stats0; p.st1 = stats1;
  p.C0 = C0; p.C1 = C1; p.C = C0 + C1; p.HW = HW;
  p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld; p.film_off = film_off;
  p.y = (__nv_bfloat16*)y; p.mean = mean; p.rstd = rstd; p.ab = ab;
  CDAE_CHECK_SHAPE(p.C % kGroups == 0 && C0 % 8 == 0 && C1 % 8 == 0, "gn_apply_fwd: C0=%d C1=%d (C %% 32, C0/C1 %% 8)", C0, C1);
  CDAE_CHECK_SHAPE(p.C / 8 <= kApplyThreads, "gn_apply_fwd: C=%d too large", p.C);
  CDAE_CHECK_ARG(aligned16(x0) && aligned16(y) && (!x1 || aligned16(x1)) && (reinterpret_cast<uintptr_t>(stats0) & 7) == 0 &&
                 (!stats1 || (reinterpret_cast<uintptr_t>(stats1) & 7) == 0), "gn_apply_fwd: misaligned pointer");
  CDAE_CHECK_SHAPE(B <= 65535, "gn_apply_fwd: batch %d too large", B);
  p.nvec = p.C / 8;
  p.R = kApplyThreads / p.nvec;
  // ~48 KB of input per CTA (a multiple of the rows one pass covers), but enough CTAs to fill the chip twice
  int ppc = (48 * 1024) / (p.C * 2);
  ppc = ppc / (p.R * kApplyUnroll) * (p.R * kApplyUnroll);
  if (ppc < p.R * kApplyUnroll) ppc = p.R * kApplyUnroll;
  while (ppc > p.R * kApplyUnroll && (int64_t)B * ((HW + ppc - 1) / ppc) < 2 * 3 * kNumSMs) ppc -= p.R * kApplyUnroll;
  if (ppc > HW) ppc = HW;
  p.ppc = ppc;
  dim3 grid((unsigned)((HW + ppc - 1) / ppc), (unsigned)B);
  if (!y) { p.ppc = 0; grid.x = 1; }
  const size_t smem = sizeof(float) * (2 * (size_t)p.C + 2 * kGroups);
  if (silu) gn_apply_fwd_kernel<true><<<grid, kApplyThreads, smem, (cudaStream_t)s>>>(p);
  else gn_apply_fwd_kernel<false><<<grid, kApplyThreads, smem, (cudaStream_t)s>>>(p);
  CDAE_CHECK_LAUNCH("gn_apply_fwd_kernel");
  return CDAE_OK;
}

extern "C" int cdae_gn_bwd(const void* dy, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                           const float* gamma, const float* beta, const float* film, int film_ld, int film_off, int silu,
                           const float* mean, const float* rstd, const void* dadd, void* dx0, void* dx1, int accumulate_dx,
                           float* dgamma, float* dbeta, float* dfilm, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(dy && x0 && gamma && beta && mean && rstd && dx0 && (C1 == 0 || (x1 && dx1)), "gn_bwd: null pointer");
  GnParams p;
  memset(&p, 0, sizeof(p));
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1; p.C0 = C0; p.C1 = C1; p.HW = HW;
  p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld; p.film_off = film_off; p.silu = silu;
  p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd);
  p.dy = (const __nv_bfloat16*)dy; p.dadd = (const __nv_bfloat16*)dadd; p.dx0 = (__nv_bfloat16*)dx0; p.dx1 = (__nv_bfloat16*)dx1;
  p.accumulate_dx = accumulate_dx;
  p.dgamma = dgamma; p.dbeta = dbeta; p.dfilm = dfilm;
  {
    auto kern = silu ? gn_bwd_pipe_kernel<true> : gn_bwd_pipe_kernel<false>;
    if (gn_pipe_config(p, B, true, reinterpret_cast<const void*>(kern))) {
      p.v4 = aligned16(gamma) && aligned16(beta) && (!film || (aligned16(film) && film_ld % 4 == 0 && film_off % 4 == 0 && p.C % 4 == 0));
      return gn_pipe_launch(kern, p, gn_pipe_smem(p, true), (cudaStream_t)s, "gn_bwd_pipe_kernel");
    }
  }
  int S;
  const bool wide = (C0 + C1) > 4 * kGnThreads;
  int rc = wide ? gn_config<8>(p, &S) : gn_config<4>(p, &S);
  if (rc) return rc;
  const size_t smem = sizeof(float) * (4 * p.C + 2 * kGroups);
  if (wide) return gn_launch(gn_bwd_kernel<8, 2>, p, B, S, smem, (cudaStream_t)s, "gn_bwd_kernel");
  return gn_launch(gn_bwd_kernel<4, 4>, p, B, S, smem, (cudaStream_t)s, "gn_bwd_kernel");
}

extern "C" int cdae_gn_bwd_apply(const void* du, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                                 const float* gamma, const float* beta, const float* film, int film_ld, int film_off,
                                 const float* mean, const float* rstd, const float* ws, const void* dadd, void* dx0, void* dx1,
                                 int accumulate_dx, float* dgamma, float* dbeta, float* dfilm, cdae_stream s) {
  if (B == 0 || HW == 0) return CDAE_OK;
  CDAE_CHECK_ARG(du && x0 && gamma && beta && mean && rstd && dx0 && ws && (C1 == 0 || (x1 && dx1)), "gn_bwd_apply: null pointer");
  GnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.du = (const __nv_bfloat16*)du; p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1;
  p.dadd = (const __nv_bfloat16*)dadd; p.dx0 = (__nv_bfloat16*)dx0; p.dx1 = (__nv_bfloat16*)dx1;
  p.C0 = C0; p.C1 = C1; p.C = C0 + C1; p.HW = HW; p.B = B;
  p.gamma = gamma; p.beta = beta; p.film = film; p.film_ld = film_ld; p.film_off = film_off;
  p.mean = mean; p.rstd = rstd; p.ws = ws; p.dgamma = dgamma; p.dbeta = dbeta; p.dfilm = dfilm;
  p.accumulate_dx = accumulate_dx;
  CDAE_CHECK_SHAPE(p.C % kGroups == 0 && C0 % 8 == 0 && C1 % 8 == 0, "gn_bwd_apply: C0=%d C1=%d (C %% 32, C0/C1 %% 8)", C0, C1);
  CDAE_CHECK_SHAPE(p.C / 8 <= kBwdThreads && B <= 65535, "gn_bwd_apply: C=%d or B=%d too large", p.C, B);
  CDAE_CHECK_ARG(aligned16(du) && aligned16(x0) && aligned16(dx0) && (!x1 || (aligned16(x1) && aligned16(dx1))) &&
                 (!dadd || aligned16(dadd)) && (reinterpret_cast<uintptr_t>(ws) & 7) == 0, "gn_bwd_apply: misaligned pointer");
  p.nvec = p.C / 8;
  p.R = kBwdThreads / p.nvec;
  const int quantum = kBwdUnroll * p.R;
  int ppc = (32 * 1024) / (p.C * 2);               // ~32 KB of du (and as much of x) per CTA
  ppc = ppc / quantum * quantum;
  if (ppc < quantum) ppc = quantum;
  while (ppc > quantum && (int64_t)B * ((HW + ppc - 1) / ppc) < 2 * 3 * kNumSMs) ppc -= quantum;
  if (ppc > HW) ppc = HW;
  p.ppc = ppc;
  dim3 grid((unsigned)((HW + ppc - 1) / ppc), (unsigned)B);
  const size_t smem = sizeof(float) * (6 * (size_t)p.C);
  gn_bwd_apply_kernel<<<grid, kBwdThreads, smem, (cudaStream_t)s>>>(p);
  CDAE_CHECK_LAUNCH("gn_bwd_apply_kernel");
  return CDAE_OK;
}
