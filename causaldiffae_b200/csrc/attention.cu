// Fused QKV attention (ref unet.py:239-253): softmax_fp32((q s)(k s)^T) v with s = ch^-1/4, per (batch, head).
// Replaces 2x bmm + _softmax + 2 scale multiplies (and their autograd) by one forward and two backward kernels
// that keep S/P on chip (flash-style, online softmax, recompute in the backward).
//
// Sequence lengths are tiny (T = H*W <= 1024 at the attention resolutions) and attention is < 1 % of the UNet FLOPs,
// so these kernels use warp-level tensor-core MMAs (mma.sync m16n8k16 bf16, fp32 accumulate); the tcgen05 budget goes
// to the convolutions (igemm.cu).  What makes them fast is the data path:
//   * every operand tile is copied global -> shared memory ONCE, row-major, by cp.async (16 B, zero fill past T) into a
//     two-stage ring, so the next 64-row block streams in while the current one is multiplied;
//   * fragments come out of shared memory with ldmatrix (x4): the "transposed" operands (V in P.V, K in dS.K, Q / dO in
//     the dK / dV products) use ldmatrix.trans on the same row-major tile - no transposed copies exist;
//   * rows are padded by 16 B, which makes all eight 16 B row segments of an ldmatrix phase hit distinct banks.
// Layout: qkv [B, T, 3*C] bf16 with head h at channels [h*3*ch, (h+1)*3*ch) = [q | k | v]; out [B, T, C] (head h at h*ch).
#include "common.cuh"

namespace cdae {

constexpr int kBK = 64;   // rows of a streamed tile (keys in fwd / dq, queries in dkv)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

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
__device__ __forceinline__ float2 unpack_bf16(uint32_t r) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r));
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, bool ok) {
  const int n = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp4z(uint32_t dst, const void* src, bool ok) {
  const int n = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows [row0, row0+nrows) x CH channels of a [T, ld] matrix -> shared tile with a (CH+8)-element pitch; rows >= T are zero
template <int CH, int NT>
__device__ __forceinline__ void tile_async(uint32_t dst, const __nv_bfloat16* src, int ld, int row0, int nrows, int T, int tid) {
  constexpr int V = CH / 8, P = (CH + 8) * 2;
  for (int i = tid; i < nrows * V; i += NT) {
    const int r = i / V, v = i - r * V;
    const bool ok = row0 + r < T;
    cp16z(dst + r * P + v * 16, src + (size_t)(ok ? row0 + r : 0) * ld + v * 8, ok);
  }
}

// Fragment addresses inside a tile with byte pitch P (lane-dependent parts precomputed by the callers where it pays).
// A operand (16 rows from r0, k chunk kk): regs = {rows 0-7 | k lo, rows 8-15 | k lo, rows 0-7 | k hi, rows 8-15 | k hi}
template <int P>
__device__ __forceinline__ uint32_t a_frag_addr(uint32_t base, int r0, int kk, int lane) {
  return base + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * P + (kk * 16 + (lane >> 4) * 8) * 2;
}
// B operand from a tile stored [n][k] (non-transposed ldmatrix): n tiles (nt, nt+1), k chunk kk -> {b0,b1 | nt, b0,b1 | nt+1}
template <int P>
__device__ __forceinline__ uint32_t b_frag_addr(uint32_t base, int nt, int kk, int lane) {
  return base + ((nt + (lane >> 4)) * 8 + (lane & 7)) * P + (kk * 16 + ((lane >> 3) & 1) * 8) * 2;
}
// B operand from a tile stored [k][n] (ldmatrix.trans): k chunk kk (16 rows), n tiles (nt, nt+1) -> {b0,b1 | nt, b0,b1 | nt+1}
template <int P>
__device__ __forceinline__ uint32_t bt_frag_addr(uint32_t base, int nt, int kk, int lane) {
  return base + (kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * P + ((nt + (lane >> 4)) * 8) * 2;
}

// c[8][4] (16 x 64) = A(16 x CH, rows r0.. of tile `at`) * B^T where tile `bt` is [64][CH]
template <int CH>
__device__ __forceinline__ void gemm_16x64(float (*c)[4], uint32_t at, int r0, uint32_t bt, int lane) {
  constexpr int P = (CH + 8) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < CH / 16; ++kk) {
    uint32_t a[4];
    ldsm4(a, a_frag_addr<P>(at, r0, kk, lane));
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm4(b, b_frag_addr<P>(bt, 2 * np, kk, lane));
      mma_bf16(c[2 * np], a, b[0], b[1]);
      mma_bf16(c[2 * np + 1], a, b[2], b[3]);
    }
  }
}
// same with the A fragments already in registers
template <int CH>
__device__ __forceinline__ void gemm_16x64_reg(float (*c)[4], const uint32_t (*a)[4], uint32_t bt, int lane) {
  constexpr int P = (CH + 8) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < CH / 16; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm4(b, b_frag_addr<P>(bt, 2 * np, kk, lane));
      mma_bf16(c[2 * np], a[kk], b[0], b[1]);
      mma_bf16(c[2 * np + 1], a[kk], b[2], b[3]);
    }
  }
}
// acc[CH/8][4] (16 x CH) += P(16 x 64, accumulator-layout registers) * B where tile `bt` is [64 (k)][CH (n)]
template <int CH>
__device__ __forceinline__ void gemm_p_tile(float (*acc)[4], const float (*p)[4], uint32_t bt, int lane) {
  constexpr int P = (CH + 8) * 2;
#pragma unroll
  for (int kk = 0; kk < kBK / 16; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < CH / 16; ++np) {
      uint32_t b[4];
      ldsm4t(b, bt_frag_addr<P>(bt, 2 * np, kk, lane));
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---------------------------------------------------------------- forward
// CTA = NW warps x 16 queries of one (batch, head); K/V stream through a two-stage ring in blocks of 64 keys.
// smem: stage s = { K[64][CH+8], V[64][CH+8] } x 2; the Q tile is staged through the same bytes before the loop starts.
template <int CH, int NW>
__global__ void __launch_bounds__(NW * 32) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                           float* __restrict__ lse, int T, int heads, float sl2) {
  constexpr int P = (CH + 8) * 2, NT = NW * 32, BQ = NW * 16, kTile = kBK * P;
  static_assert(BQ * P <= 4 * kTile, "Q staging must fit the K/V ring");
  extern __shared__ __align__(16) uint8_t smraw[];
  const uint32_t sm = smem_addr(smraw);
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int ld = 3 * CH * heads;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
  const __nv_bfloat16* kb = qb + CH;
  const __nv_bfloat16* vb = qb + 2 * CH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * BQ;

  tile_async<CH, NT>(sm, qb, ld, q0, BQ, T, tid);
  cp_commit(); cp_wait_all();
  __syncthreads();
  uint32_t qa[CH / 16][4];
#pragma unroll
  for (int kk = 0; kk < CH / 16; ++kk) ldsm4(qa[kk], a_frag_addr<P>(sm, warp * 16, kk, lane));
  __syncthreads();

  float o[CH / 8][4];
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int nkb = (T + kBK - 1) / kBK;
  tile_async<CH, NT>(sm, kb, ld, 0, kBK, T, tid);
  tile_async<CH, NT>(sm + kTile, vb, ld, 0, kBK, T, tid);
  cp_commit();
  for (int ib = 0; ib < nkb; ++ib) {
    cp_wait_all();
    __syncthreads();
    if (ib + 1 < nkb) {
      const uint32_t nx = sm + ((ib + 1) & 1) * 2 * kTile;
      tile_async<CH, NT>(nx, kb, ld, (ib + 1) * kBK, kBK, T, tid);
      tile_async<CH, NT>(nx + kTile, vb, ld, (ib + 1) * kBK, kBK, T, tid);
      cp_commit();
    }
    const uint32_t ks = sm + (ib & 1) * 2 * kTile, vs = ks + kTile;
    const int valid = min(kBK, T - ib * kBK);
    float s[8][4];
    gemm_16x64_reg<CH>(s, qa, ks, lane);
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + t * 2 + (e & 1);
        s[nt][e] = col < valid ? s[nt][e] * sl2 : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = quad_max(mx0); mx1 = quad_max(mx1);
    const float c0 = ex2(m0 - mx0), c1 = ex2(m1 - mx1);
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2(s[nt][0] - mx0); s[nt][1] = ex2(s[nt][1] - mx0);
      s[nt][2] = ex2(s[nt][2] - mx1); s[nt][3] = ex2(s[nt][3] - mx1);
      r0 += s[nt][0] + s[nt][1]; r1 += s[nt][2] + s[nt][3];
    }
    r0 = quad_sum(r0); r1 = quad_sum(r1);
    l0 = l0 * c0 + r0; l1 = l1 * c1 + r1; m0 = mx0; m1 = mx1;
#pragma unroll
    for (int i = 0; i < CH / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    gemm_p_tile<CH>(o, s, vs, lane);
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const int ldo = CH * heads;
  __nv_bfloat16* ob = out + (size_t)b * T * ldo + (size_t)h * CH;
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) {
    const int c = i * 8 + t * 2;
    if (r0 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * ldo + c) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * ldo + c) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
  }
  if (lse && t == 0) {   // natural-log units (the kernels work in base 2 internally)
    if (r0 < T) lse[(size_t)bh * T + r0] = (m0 + __log2f(l0)) * kLn2;
    if (r1 < T) lse[(size_t)bh * T + r1] = (m1 + __log2f(l1)) * kLn2;
  }
}

// ---------------------------------------------------------------- backward, pass 1: dQ and D = rowsum(dO * O)
// CTA = NW warps x 16 queries; the CTA's Q and dO tiles stay in shared memory, K/V stream through the ring.
// smem: Q[BQ][CH+8] | dO[BQ][CH+8] | stage { K[64][CH+8], V[64][CH+8] } x 2 (O is staged through the ring first)
template <int CH, int NW>
__global__ void __launch_bounds__(NW * 32) attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                                                              const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                                                              float* __restrict__ dsum, __nv_bfloat16* __restrict__ dqkv, int T,
                                                              int heads, float scale2) {
  constexpr int P = (CH + 8) * 2, NT = NW * 32, BQ = NW * 16, kTile = kBK * P;
  extern __shared__ __align__(16) uint8_t smraw[];
  const uint32_t qs = smem_addr(smraw), dos = qs + BQ * P, ring = dos + BQ * P;
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int ld = 3 * CH * heads, ldo = CH * heads;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
  const __nv_bfloat16* kb = qb + CH;
  const __nv_bfloat16* vb = qb + 2 * CH;
  const __nv_bfloat16* ob = out + (size_t)b * T * ldo + (size_t)h * CH;
  const __nv_bfloat16* dob = dout + (size_t)b * T * ldo + (size_t)h * CH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * BQ;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float sl2 = scale2 * kLog2e;

  tile_async<CH, NT>(qs, qb, ld, q0, BQ, T, tid);
  tile_async<CH, NT>(dos, dob, ldo, q0, BQ, T, tid);
  tile_async<CH, NT>(ring, ob, ldo, q0, BQ, T, tid);
  cp_commit(); cp_wait_all();
  __syncthreads();
  float D0 = 0.f, D1 = 0.f;
#pragma unroll
  for (int kk = 0; kk < CH / 16; ++kk) {
    uint32_t a[4], d[4];
    ldsm4(a, a_frag_addr<P>(ring, warp * 16, kk, lane));
    ldsm4(d, a_frag_addr<P>(dos, warp * 16, kk, lane));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16(a[j]), y = unpack_bf16(d[j]);
      const float v = x.x * y.x + x.y * y.y;
      if (j & 1) D1 += v; else D0 += v;
    }
  }
  D0 = quad_sum(D0); D1 = quad_sum(D1);
  if (t == 0) {
    if (r0 < T) dsum[(size_t)bh * T + r0] = D0;
    if (r1 < T) dsum[(size_t)bh * T + r1] = D1;
  }
  const float L0 = r0 < T ? lse[(size_t)bh * T + r0] * kLog2e : 0.f, L1 = r1 < T ? lse[(size_t)bh * T + r1] * kLog2e : 0.f;
  __syncthreads();      // every warp has read its O rows: the ring may be overwritten

  float dq[CH / 8][4];
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const int nkb = (T + kBK - 1) / kBK;
  tile_async<CH, NT>(ring, kb, ld, 0, kBK, T, tid);
  tile_async<CH, NT>(ring + kTile, vb, ld, 0, kBK, T, tid);
  cp_commit();
  for (int ib = 0; ib < nkb; ++ib) {
    cp_wait_all();
    __syncthreads();
    if (ib + 1 < nkb) {
      const uint32_t nx = ring + ((ib + 1) & 1) * 2 * kTile;
      tile_async<CH, NT>(nx, kb, ld, (ib + 1) * kBK, kBK, T, tid);
      tile_async<CH, NT>(nx + kTile, vb, ld, (ib + 1) * kBK, kBK, T, tid);
      cp_commit();
    }
    const uint32_t ks = ring + (ib & 1) * 2 * kTile, vs = ks + kTile;
    const int valid = min(kBK, T - ib * kBK);
    float s[8][4], dp[8][4];
    gemm_16x64<CH>(s, qs, warp * 16, ks, lane);
    gemm_16x64<CH>(dp, dos, warp * 16, vs, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + t * 2 + (e & 1);
        const float L = e < 2 ? L0 : L1, Dd = e < 2 ? D0 : D1;
        const float p = col < valid ? ex2(fmaf(s[nt][e], sl2, -L)) : 0.f;
        s[nt][e] = p * (dp[nt][e] - Dd) * scale2;      // dS
      }
    }
    gemm_p_tile<CH>(dq, s, ks, lane);                  // dQ += dS K   (K tile is [k = key][n = ch])
  }
  __nv_bfloat16* dqb = dqkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) {
    const int c = i * 8 + t * 2;
    if (r0 < T) *reinterpret_cast<uint32_t*>(dqb + (size_t)r0 * ld + c) = pack_bf16(dq[i][0], dq[i][1]);
    if (r1 < T) *reinterpret_cast<uint32_t*>(dqb + (size_t)r1 * ld + c) = pack_bf16(dq[i][2], dq[i][3]);
  }
}

// ---------------------------------------------------------------- backward, pass 2: dK, dV  (CTA = NW warps x 16 keys)
// Works on the transposed problem: S^T = K Q^T so that P^T / dS^T come out in accumulator layout and feed the next MMA.
// smem: K[BKC][CH+8] | V[BKC][CH+8] | stage { Q[64][CH+8], dO[64][CH+8] } x 2 | L[2][64] | D[2][64]
template <int CH, int NW>
__global__ void __launch_bounds__(NW * 32) attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                               const float* __restrict__ lse, const float* __restrict__ dsum,
                                                               __nv_bfloat16* __restrict__ dqkv, int T, int heads, float scale2) {
  constexpr int P = (CH + 8) * 2, NT = NW * 32, BKC = NW * 16, kTile = kBK * P;
  extern __shared__ __align__(16) uint8_t smraw[];
  const uint32_t ksm = smem_addr(smraw), vsm = ksm + BKC * P, ring = vsm + BKC * P;
  float* LD = reinterpret_cast<float*>(smraw + 2 * BKC * P + 4 * kTile);     // [stage][L 64 | D 64]
  const uint32_t ldsm = ring + 4 * kTile;
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int ld = 3 * CH * heads, ldo = CH * heads;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * ld + (size_t)h * 3 * CH;
  const __nv_bfloat16* kb = qb + CH;
  const __nv_bfloat16* vb = qb + 2 * CH;
  const __nv_bfloat16* dob = dout + (size_t)b * T * ldo + (size_t)h * CH;
  const float* lrow = lse + (size_t)bh * T;
  const float* drow = dsum + (size_t)bh * T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.x * BKC;
  const int r0 = k0 + warp * 16 + g, r1 = r0 + 8;
  const float sl2 = scale2 * kLog2e;

  auto load_stage = [&](int st, int qrow0) {
    const uint32_t base = ring + st * 2 * kTile;
    tile_async<CH, NT>(base, qb, ld, qrow0, kBK, T, tid);
    tile_async<CH, NT>(base + kTile, dob, ldo, qrow0, kBK, T, tid);
    if (tid < 2 * kBK) {
      const int j = tid & (kBK - 1), q = qrow0 + j;
      const bool ok = q < T;
      const float* src = (tid < kBK ? lrow : drow) + (ok ? q : 0);
      cp4z(ldsm + (st * 2 * kBK + tid) * 4, src, ok);
    }
  };
  tile_async<CH, NT>(ksm, kb, ld, k0, BKC, T, tid);
  tile_async<CH, NT>(vsm, vb, ld, k0, BKC, T, tid);
  load_stage(0, 0);
  cp_commit();

  float dk[CH / 8][4], dv[CH / 8][4];
#pragma unroll
  for (int i = 0; i < CH / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const int nqb = (T + kBK - 1) / kBK;
  for (int ib = 0; ib < nqb; ++ib) {
    cp_wait_all();
    __syncthreads();
    if (ib + 1 < nqb) { load_stage((ib + 1) & 1, (ib + 1) * kBK); cp_commit(); }
    const uint32_t qs = ring + (ib & 1) * 2 * kTile, dos = qs + kTile;
    const float* Ls = LD + (ib & 1) * 2 * kBK;
    const float* Ds = Ls + kBK;
    const int valid = min(kBK, T - ib * kBK);
    float st[8][4], dpt[8][4];
    gemm_16x64<CH>(st, ksm, warp * 16, qs, lane);       // S^T  [16 keys x 64 q]
    gemm_16x64<CH>(dpt, vsm, warp * 16, dos, lane);     // dP^T [16 keys x 64 q]
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int cb = nt * 8 + t * 2;
      const float2 Lc = *reinterpret_cast<const float2*>(Ls + cb), Dc = *reinterpret_cast<const float2*>(Ds + cb);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = cb + (e & 1);
        const int krow = e < 2 ? r0 : r1;
        const float L = ((e & 1) ? Lc.y : Lc.x) * kLog2e, Dd = (e & 1) ? Dc.y : Dc.x;
        const float p = (col < valid && krow < T) ? ex2(fmaf(st[nt][e], sl2, -L)) : 0.f;
        st[nt][e] = p;
        dpt[nt][e] = p * (dpt[nt][e] - Dd) * scale2;    // dS^T
      }
    }
    gemm_p_tile<CH>(dv, st, dos, lane);      // dV += P^T dO    (dO tile is [k = q][n = ch])
    gemm_p_tile<CH>(dk, dpt, qs, lane);      // dK += dS^T Q
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

template <typename K>
static int set_smem(K kernel, size_t smem, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("%s smem attribute (%zu B): %s", name, smem, cudaGetErrorString(e)); return CDAE_ERR_CUDA; }
  return CDAE_OK;
}

template <int CH, int NW>
static int attn_fwd_launch(const void* qkv, void* out, float* lse, int B, int T, int heads, cudaStream_t st) {
  constexpr size_t smem = 4 * kBK * (CH + 8) * 2;
  static int attr_rc = set_smem(attn_fwd_kernel<CH, NW>, smem, "attn_fwd_kernel");   // thread-safe one-time init
  if (attr_rc) return attr_rc;
  const float sl2 = kLog2e / sqrtf((float)CH);
  dim3 grid((T + NW * 16 - 1) / (NW * 16), B * heads);
  attn_fwd_kernel<CH, NW><<<grid, NW * 32, smem, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, lse, T, heads, sl2);
  CDAE_CHECK_LAUNCH("attn_fwd_kernel");
  return CDAE_OK;
}

template <int CH, int NW>
static int attn_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, float* dsum, void* dqkv,
                           int B, int T, int heads, cudaStream_t st) {
  const float scale2 = 1.0f / sqrtf((float)CH);
  constexpr size_t smem1 = (2 * NW * 16 + 4 * kBK) * (CH + 8) * 2;
  constexpr size_t smem2 = smem1 + 4 * kBK * sizeof(float);
  static_assert(smem2 <= 227 * 1024, "attention backward: shared memory budget");
  static int rc1 = set_smem(attn_bwd_dq_kernel<CH, NW>, smem1, "attn_bwd_dq_kernel");
  static int rc2 = set_smem(attn_bwd_dkv_kernel<CH, NW>, smem2, "attn_bwd_dkv_kernel");
  if (rc1) return rc1;
  if (rc2) return rc2;
  dim3 grid((T + NW * 16 - 1) / (NW * 16), B * heads);
  attn_bwd_dq_kernel<CH, NW><<<grid, NW * 32, smem1, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)out,
                                                           (const __nv_bfloat16*)dout, lse, dsum, (__nv_bfloat16*)dqkv, T,
                                                           heads, scale2);
  CDAE_CHECK_LAUNCH("attn_bwd_dq_kernel");
  attn_bwd_dkv_kernel<CH, NW><<<grid, NW * 32, smem2, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, lse, dsum,
                                                            (__nv_bfloat16*)dqkv, T, heads, scale2);
  CDAE_CHECK_LAUNCH("attn_bwd_dkv_kernel");
  return CDAE_OK;
}

}  // namespace cdae
using namespace cdae;

// 8 warps (128 rows) per CTA for long sequences, 4 warps (64 rows) when T <= 64 would leave half a CTA idle
#define CDAE_ATTN_DISPATCH(CALL)                                                              \
  switch (ch) {                                                                               \
    case 16: return CALL(16); case 32: return CALL(32); case 48: return CALL(48);             \
    case 64: return CALL(64); case 80: return CALL(80); case 96: return CALL(96);             \
    case 112: return CALL(112); case 128: return CALL(128);                                   \
    default: set_error("attention: head dim %d unsupported (multiples of 16 up to 128)", ch); \
      return CDAE_ERR_SHAPE;                                                                  \
  }

extern "C" int cdae_attn_fwd(const void* qkv, void* out, float* lse, int B, int T, int heads, int ch, cdae_stream s) {
  if (B == 0 || T == 0) return CDAE_OK;
  CDAE_CHECK_ARG(qkv && out, "attn_fwd: null pointer");
#define CALL(N) (T <= 64 ? attn_fwd_launch<N, 4>(qkv, out, lse, B, T, heads, (cudaStream_t)s) \
                         : attn_fwd_launch<N, 8>(qkv, out, lse, B, T, heads, (cudaStream_t)s))
  CDAE_ATTN_DISPATCH(CALL)
#undef CALL
}

extern "C" int cdae_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* dsum, void* dqkv,
                             int B, int T, int heads, int ch, cdae_stream s) {
  if (B == 0 || T == 0) return CDAE_OK;
  CDAE_CHECK_ARG(qkv && out && dout && lse && dsum && dqkv, "attn_bwd: null pointer");
#define CALL(N) (T <= 64 ? attn_bwd_launch<N, 4>(qkv, out, dout, lse, dsum, dqkv, B, T, heads, (cudaStream_t)s) \
                         : attn_bwd_launch<N, 8>(qkv, out, dout, lse, dsum, dqkv, B, T, heads, (cudaStream_t)s))
  CDAE_ATTN_DISPATCH(CALL)
#undef CALL
}
