// Fused QKV attention (ref unet.py:239-253): softmax_fp32((q s)(k s)^T) v with s = ch^-1/4, per (batch, head).
// Replaces 2x bmm + _softmax + 2 scale multiplies (and their autograd) by one forward and two backward kernels
// that keep S/P on chip (flash-style, online softmax, recompute in the backward).
//
// Sequence lengths are tiny (T = H*W <= 1024 at the attention resolutions) and attention is < 1 % of the UNet FLOPs,
// so this kernel uses warp-level tensor-core MMAs (mma.sync m16n8k16 bf16, fp32 accumulate) with the whole K/V block
// staged in shared memory; the tcgen05 budget goes to the convolutions (igemm.cu).
// Layout: qkv [B, T, 3*C] bf16 with head h at channels [h*3*ch, (h+1)*3*ch) = [q | k | v]; out [B, T, C] (head h at h*ch).
#include "common.cuh"

namespace cdae {

constexpr int kBQ = 64;   // queries per CTA (4 warps x 16)
constexpr int kBK = 64;   // keys per inner block
constexpr int kPadT = 8;

__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// copy a [rows x CH] tile (row pitch ld in global) into smem row-major with pitch CH+8; rows >= valid are zero
template <int CH>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int ld, int valid, int tid, int nthr) {
  constexpr int V = CH / 8;
  for (int i = tid; i < kBK * V; i += nthr) {
    const int r = i / V, v = i % V;
    bf16x8 x;
    if (r < valid) x = ld8(src + (size_t)r * ld + v * 8);
    else { float z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; x = pack8(z); }
    st8(dst + r * (CH + 8) + v * 8, x);
  }
}
// same tile, stored transposed: dst[c][r] with pitch kBK + 8
template <int CH>
__device__ __forceinline__ void load_tile_t(__nv_bfloat16* dst, const __nv_bfloat16* src, int ld, int valid, int tid, int nthr) {
  constexpr int V = CH / 8;
  for (int i = tid; i < kBK * V; i += nthr) {
    const int r = i % kBK, v = i / kBK;
    __nv_bfloat16 e[8];
    if (r < valid) *reinterpret_cast<bf16x8*>(e) = ld8(src + (size_t)r * ld + v * 8);
    else { for (int k = 0; k < 8; ++k) e[k] = __float2bfloat16(0.f); }
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[(v * 8 + k) * (kBK + kPadT) + r] = e[k];
  }
}

// A fragments (16 rows x CH) straight from global memory (rows beyond `valid` read as zero)
template <int CH>
__device__ __forceinline__ void load_a_frags(uint32_t (*a)[4], const __nv_bfloat16* base, int ld, int row0, int valid, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < CH / 16; ++kk) {
    const int r0 = row0 + g, r1 = row0 + g + 8, c = kk * 16 + t * 2;
    a[kk][0] = r0 < valid ? lds32(base + (size_t)r0 * ld + c) : 0u;
    a[kk][1] = r1 < valid ? lds32(base + (size_t)r1 * ld + c) : 0u;
    a[kk][2] = r0 < valid ? lds32(base + (size_t)r0 * ld + c + 8) : 0u;
    a[kk][3] = r1 < valid ? lds32(base + (size_t)r1 * ld + c + 8) : 0u;
  }
}

// C[16 x 64] = A[16 x CH] * Bt where Bsm is [64 rows(n)][CH (k)] row-major (pitch CH+8)
template <int CH>
__device__ __forceinline__ void gemm_a_bT(float (*c)[4], const uint32_t (*a)[4], const __nv_bfloat16* Bsm, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < kBK / 8; ++nt) {
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    const __nv_bfloat16* brow = Bsm + (nt * 8 + g) * (CH + 8) + t * 2;
#pragma unroll
    for (int kk = 0; kk < CH / 16; ++kk) mma_bf16(c[nt], a[kk], lds32(brow + kk * 16), lds32(brow + kk * 16 + 8));
  }
}
// acc[16 x CH] += P[16 x 64] * B where Btsm is the transposed tile [CH (n)][64 (k)] (pitch 64+8); P given as C-layout regs
template <int CH>
__device__ __forceinline__ void gemm_p_b(float (*acc)[4], const float (*p)[4], const __nv_bfloat16* Btsm, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < kBK / 16; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int nt = 0; nt < CH / 8; ++nt) {
      const __nv_bfloat16* brow = Btsm + (nt * 8 + g) * (kBK + kPadT) + kk * 16 + t * 2;
      mma_bf16(acc[nt], a, lds32(brow), lds32(brow + 8));
    }
  }
}

// ---------------------------------------------------------------- forward
template <int CH>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                       float* __restrict__ lse, int T, int heads, float scale2) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __nv_bfloat16* Ksm = reinterpret_cast<__nv_bfloat16*>(smraw);          // [64][CH+8]
  __nv_bfloat16* Vtsm = Ksm + kBK * (CH + 8);                            // [CH][64+8]
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int ld = 3 * CH * heads;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
  const __nv_bfloat16* kb = qb + CH;
  const __nv_bfloat16* vb = qb + 2 * CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * kBQ + warp * 16;

  uint32_t qa[CH / 16][4];
  load_a_frags<CH>(qa, qb, ld, q0, T, lane);
  float o[CH / 8][4];
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int k0 = 0; k0 < T; k0 += kBK) {
    const int valid = min(kBK, T - k0);
    __syncthreads();
    load_tile<CH>(Ksm, kb + (size_t)k0 * ld, ld, valid, threadIdx.x, blockDim.x);
    load_tile_t<CH>(Vtsm, vb + (size_t)k0 * ld, ld, valid, threadIdx.x, blockDim.x);
    __syncthreads();
    float s[kBK / 8][4];
    gemm_a_bT<CH>(s, qa, Ksm, lane);
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < kBK / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + t * 2 + (e & 1);
        s[nt][e] = col < valid ? s[nt][e] * scale2 : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = __expf(m0 - mx0), c1 = __expf(m1 - mx1);
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < kBK / 8; ++nt) {
      s[nt][0] = __expf(s[nt][0] - mx0); s[nt][1] = __expf(s[nt][1] - mx0);
      s[nt][2] = __expf(s[nt][2] - mx1); s[nt][3] = __expf(s[nt][3] - mx1);
      r0 += s[nt][0] + s[nt][1]; r1 += s[nt][2] + s[nt][3];
    }
    r0 += __shfl_xor_sync(0xffffffffu, r0, 1); r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1); r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
    l0 = l0 * c0 + r0; l1 = l1 * c1 + r1; m0 = mx0; m1 = mx1;
#pragma unroll
    for (int i = 0; i < CH / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    gemm_p_b<CH>(o, s, Vtsm, lane);
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  const int ldo = CH * heads;
  __nv_bfloat16* ob = out + (size_t)b * T * ldo + (size_t)h * CH;
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) {
    const int c = i * 8 + t * 2;
    if (r0 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * ldo + c) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * ldo + c) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
  }
  if (lse && t == 0) {
    if (r0 < T) lse[(size_t)bh * T + r0] = m0 + __logf(l0);
    if (r1 < T) lse[(size_t)bh * T + r1] = m1 + __logf(l1);
  }
}

// ---------------------------------------------------------------- backward, pass 1: dQ  (grid: query blocks)
template <int CH>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                                                          const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                                                          __nv_bfloat16* __restrict__ dqkv, int T, int heads, float scale2) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __nv_bfloat16* Ksm = reinterpret_cast<__nv_bfloat16*>(smraw);          // [64][CH+8]
  __nv_bfloat16* Vsm = Ksm + kBK * (CH + 8);                             // [64][CH+8]
  __nv_bfloat16* Ktsm = Vsm + kBK * (CH + 8);                            // [CH][64+8]
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int ld = 3 * CH * heads, ldo = CH * heads;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
  const __nv_bfloat16* kb = qb + CH;
  const __nv_bfloat16* vb = qb + 2 * CH;
  const __nv_bfloat16* ob = out + (size_t)b * T * ldo + (size_t)h * CH;
  const __nv_bfloat16* dob = dout + (size_t)b * T * ldo + (size_t)h * CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * kBQ + warp * 16;
  const int r0 = q0 + g, r1 = q0 + g + 8;

  uint32_t qa[CH / 16][4], da[CH / 16][4];
  load_a_frags<CH>(qa, qb, ld, q0, T, lane);
  load_a_frags<CH>(da, dob, ldo, q0, T, lane);
  // D = rowsum(dO * O)
  float D0 = 0.f, D1 = 0.f;
#pragma unroll
  for (int kk = 0; kk < CH / 16; ++kk) {
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int c = kk * 16 + t * 2 + hlf * 8;
      if (r0 < T) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ob + (size_t)r0 * ldo + c));
        const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dob + (size_t)r0 * ldo + c));
        D0 += a.x * d.x + a.y * d.y;
      }
      if (r1 < T) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ob + (size_t)r1 * ldo + c));
        const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dob + (size_t)r1 * ldo + c));
        D1 += a.x * d.x + a.y * d.y;
      }
    }
  }
  D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
  D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
  const float L0 = r0 < T ? lse[(size_t)bh * T + r0] : 0.f, L1 = r1 < T ? lse[(size_t)bh * T + r1] : 0.f;

  float dq[CH / 8][4];
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  for (int k0 = 0; k0 < T; k0 += kBK) {
    const int valid = min(kBK, T - k0);
    __syncthreads();
    load_tile<CH>(Ksm, kb + (size_t)k0 * ld, ld, valid, threadIdx.x, blockDim.x);
    load_tile<CH>(Vsm, vb + (size_t)k0 * ld, ld, valid, threadIdx.x, blockDim.x);
    load_tile_t<CH>(Ktsm, kb + (size_t)k0 * ld, ld, valid, threadIdx.x, blockDim.x);
    __syncthreads();
    float s[kBK / 8][4], dp[kBK / 8][4];
    gemm_a_bT<CH>(s, qa, Ksm, lane);
    gemm_a_bT<CH>(dp, da, Vsm, lane);
#pragma unroll
    for (int nt = 0; nt < kBK / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + t * 2 + (e & 1);
        const float L = e < 2 ? L0 : L1, Dd = e < 2 ? D0 : D1;
        const float p = col < valid ? __expf(s[nt][e] * scale2 - L) : 0.f;
        s[nt][e] = p * (dp[nt][e] - Dd) * scale2;      // dS
      }
    }
    gemm_p_b<CH>(dq, s, Ktsm, lane);
  }
  __nv_bfloat16* dqb = dqkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) {
    const int c = i * 8 + t * 2;
    if (r0 < T) *reinterpret_cast<uint32_t*>(dqb + (size_t)r0 * ld + c) = pack_bf16(dq[i][0], dq[i][1]);
    if (r1 < T) *reinterpret_cast<uint32_t*>(dqb + (size_t)r1 * ld + c) = pack_bf16(dq[i][2], dq[i][3]);
  }
}

// ---------------------------------------------------------------- backward, pass 2: dK, dV  (grid: key blocks)
// Works on the transposed problem: S^T = K Q^T so that P^T / dS^T come out in accumulator layout and feed the next MMA.
template <int CH>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                                                           const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                                                           __nv_bfloat16* __restrict__ dqkv, int T, int heads, float scale2) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __nv_bfloat16* Qsm = reinterpret_cast<__nv_bfloat16*>(smraw);          // [64 q][CH+8]
  __nv_bfloat16* dOsm = Qsm + kBK * (CH + 8);                            // [64 q][CH+8]
  __nv_bfloat16* Qtsm = dOsm + kBK * (CH + 8);                           // [CH][64+8]
  __nv_bfloat16* dOtsm = Qtsm + CH * (kBK + kPadT);                      // [CH][64+8]
  float* Lsm = reinterpret_cast<float*>(dOtsm + CH * (kBK + kPadT));     // [64]
  float* Dsm = Lsm + kBK;                                                // [64]
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int ld = 3 * CH * heads, ldo = CH * heads;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
  const __nv_bfloat16* kb = qb + CH;
  const __nv_bfloat16* vb = qb + 2 * CH;
  const __nv_bfloat16* ob = out + (size_t)b * T * ldo + (size_t)h * CH;
  const __nv_bfloat16* dob = dout + (size_t)b * T * ldo + (size_t)h * CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int kr0 = blockIdx.x * kBQ + warp * 16;     // this warp's 16 keys
  const int r0 = kr0 + g, r1 = kr0 + g + 8;

  uint32_t ka[CH / 16][4], va[CH / 16][4];
  load_a_frags<CH>(ka, kb, ld, kr0, T, lane);
  load_a_frags<CH>(va, vb, ld, kr0, T, lane);
  float dk[CH / 8][4], dv[CH / 8][4];
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

  for (int q0 = 0; q0 < T; q0 += kBK) {
    const int valid = min(kBK, T - q0);
    __syncthreads();
    load_tile<CH>(Qsm, qb + (size_t)q0 * ld, ld, valid, threadIdx.x, blockDim.x);
    load_tile<CH>(dOsm, dob + (size_t)q0 * ldo, ldo, valid, threadIdx.x, blockDim.x);
    load_tile_t<CH>(Qtsm, qb + (size_t)q0 * ld, ld, valid, threadIdx.x, blockDim.x);
    load_tile_t<CH>(dOtsm, dob + (size_t)q0 * ldo, ldo, valid, threadIdx.x, blockDim.x);
    if (threadIdx.x < kBK) {
      const int q = q0 + threadIdx.x;
      float Dd = 0.f, L = 0.f;
      if (q < T) {
        L = lse[(size_t)bh * T + q];
        for (int c = 0; c < CH; c += 8) {
          float a[8], d[8];
          unpack8(ld8(ob + (size_t)q * ldo + c), a);
          unpack8(ld8(dob + (size_t)q * ldo + c), d);
#pragma unroll
          for (int k = 0; k < 8; ++k) Dd += a[k] * d[k];
        }
      }
      Lsm[threadIdx.x] = L; Dsm[threadIdx.x] = Dd;
    }
    __syncthreads();
    float st[kBK / 8][4], dpt[kBK / 8][4];
    gemm_a_bT<CH>(st, ka, Qsm, lane);      // S^T  [16 keys x 64 q]
    gemm_a_bT<CH>(dpt, va, dOsm, lane);    // dP^T [16 keys x 64 q]
    float pt[kBK / 8][4];
#pragma unroll
    for (int nt = 0; nt < kBK / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + t * 2 + (e & 1);
        const int krow = e < 2 ? r0 : r1;
        const float p = (col < valid && krow < T) ? __expf(st[nt][e] * scale2 - Lsm[col]) : 0.f;
        pt[nt][e] = p;
        st[nt][e] = p * (dpt[nt][e] - Dsm[col]) * scale2;   // dS^T
      }
    }
    gemm_p_b<CH>(dv, pt, dOtsm, lane);     // dV += P^T dO
    gemm_p_b<CH>(dk, st, Qtsm, lane);      // dK += dS^T Q
  }
  __nv_bfloat16* dkb = dqkv + (size_t)b * T * ld + (size_t)h * 3 * CH + CH;
  __nv_bfloat16* dvb = dkb + CH;
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) {
    const int c = i * 8 + t * 2;
    if (r0 < T) {
      *reinterpret_cast<uint32_t*>(dkb + (size_t)r0 * ld + c) = pack_bf16(dk[i][0], dk[i][1]);
      *reinterpret_cast<uint32_t*>(dvb + (size_t)r0 * ld + c) = pack_bf16(dv[i][0], dv[i][1]);
    }
    if (r1 < T) {
      *reinterpret_cast<uint32_t*>(dkb + (size_t)r1 * ld + c) = pack_bf16(dk[i][2], dk[i][3]);
      *reinterpret_cast<uint32_t*>(dvb + (size_t)r1 * ld + c) = pack_bf16(dv[i][2], dv[i][3]);
    }
  }
}

template <int CH>
static int attn_fwd_launch(const void* qkv, void* out, float* lse, int B, int T, int heads, cudaStream_t st) {
  const size_t smem = sizeof(__nv_bfloat16) * (kBK * (CH + 8) + CH * (kBK + kPadT));
  const float scale2 = 1.0f / sqrtf((float)CH);
  dim3 grid((T + kBQ - 1) / kBQ, B * heads);
  attn_fwd_kernel<CH><<<grid, 128, smem, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, lse, T, heads, scale2);
  CDAE_CHECK_LAUNCH("attn_fwd_kernel");
  return CDAE_OK;
}

template <int CH>
static int attn_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int T,
                           int heads, cudaStream_t st) {
  const float scale2 = 1.0f / sqrtf((float)CH);
  dim3 grid((T + kBQ - 1) / kBQ, B * heads);
  const size_t smem1 = sizeof(__nv_bfloat16) * (2 * kBK * (CH + 8) + CH * (kBK + kPadT));
  const size_t smem2 = sizeof(__nv_bfloat16) * (2 * kBK * (CH + 8) + 2 * CH * (kBK + kPadT)) + sizeof(float) * 2 * kBK;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_bwd_dq_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    attr_done = true;
  }
  attn_bwd_dq_kernel<CH><<<grid, 128, smem1, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)out,
                                                    (const __nv_bfloat16*)dout, lse, (__nv_bfloat16*)dqkv, T, heads, scale2);
  CDAE_CHECK_LAUNCH("attn_bwd_dq_kernel");
  attn_bwd_dkv_kernel<CH><<<grid, 128, smem2, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)out,
                                                     (const __nv_bfloat16*)dout, lse, (__nv_bfloat16*)dqkv, T, heads, scale2);
  CDAE_CHECK_LAUNCH("attn_bwd_dkv_kernel");
  return CDAE_OK;
}

}  // namespace cdae
using namespace cdae;

#define CDAE_ATTN_DISPATCH(CALL)                                                              \
  switch (ch) {                                                                               \
    case 16: return CALL(16); case 32: return CALL(32); case 48: return CALL(48);             \
    case 64: return CALL(64); case 80: return CALL(80); case 96: return CALL(96);             \
    case 112: return CALL(112); case 128: return CALL(128);                                   \
    default: set_error("attention: head dim %d unsupported (multiples of 16 up to 128)", ch); \
      return CDAE_ERR_SHAPE;                                                                  \
  }

extern "C" int cdae_attn_fwd(const void* qkv, void* out, float* lse, int B, int T, int heads, int ch, cdae_stream s) {
  CDAE_CHECK_ARG(qkv && out, "attn_fwd: null pointer");
  if (B == 0 || T == 0) return CDAE_OK;
#define CALL(N) attn_fwd_launch<N>(qkv, out, lse, B, T, heads, (cudaStream_t)s)
  CDAE_ATTN_DISPATCH(CALL)
#undef CALL
}

extern "C" int cdae_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B,
                             int T, int heads, int ch, cdae_stream s) {
  CDAE_CHECK_ARG(qkv && out && dout && lse && dqkv, "attn_bwd: null pointer");
  if (B == 0 || T == 0) return CDAE_OK;
#define CALL(N) attn_bwd_launch<N>(qkv, out, dout, lse, dqkv, B, T, heads, (cudaStream_t)s)
  CDAE_ATTN_DISPATCH(CALL)
#undef CALL
}
