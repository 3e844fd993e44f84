// The causal encoder's DAG mask layer as ONE fused kernel per direction (ref nn.py:225-240 MLP, :290-312
// CausalModeling.causal_masking / nonlinearity_add_back_noise, called from unet.py:571-583):
//     z_pre[b,i,:]  = sum_j A[j,i] * u[b,j,:]                                   (A^T u over the n causal variables)
//     z_post[b,i,:] = W2_i * leaky_relu(W1_i * z_pre[b,i,:] + b1_i) + b2_i + u[b,i,:]
// u, z_post: fp32 [B, n, d] (d = latent_dim / n); W1_i [D, d], W2_i [d, D] (D = latent_dim), LeakyReLU slope 0.01.
// The reference runs this as 2 + 5n small ATen launches plus n device<->host copies; the tensors are
// [B, 512]-sized, so the point of fusing is launch count and latency, not FLOPs: fp32 CUDA-core math, weights read
// once per CTA through coalesced 128-bit loads, warp-shuffle reductions.  The backward recomputes the hidden layer.
#include "common.cuh"

namespace cdae {

constexpr int kDagThreads = 256;
constexpr int kDagTB = 8;        // samples per CTA in the forward
constexpr int kDagHT = 32;       // hidden units per CTA in the backward
constexpr int kDagBC = 32;       // samples per batch chunk in the backward
constexpr float kLeaky = 0.01f;

// v[8] holds one partial per sample; returns (to every lane) the sum over the warp of v[lane >> 2]
__device__ __forceinline__ float reduce8(const float* v, int lane) {
  float a[4], b[2];
  const bool h1 = lane & 16, h2 = lane & 8, h3 = lane & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = h1 ? v[j] : v[j + 4], keep = h1 ? v[j + 4] : v[j];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = h2 ? a[j] : a[j + 2], keep = h2 ? a[j + 2] : a[j];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float send = h3 ? b[0] : b[1], keep = h3 ? b[1] : b[0];
  float c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}

// params: device array of 4n pointers {W1_i, b1_i, W2_i, b2_i}
// grid (n, ceil(B / 8)); smem: zp[8][d] | hid[8][D]
__global__ void __launch_bounds__(kDagThreads) dag_fwd_kernel(const float* __restrict__ u, const float* __restrict__ A,
                                                              const float* const* __restrict__ params, float* __restrict__ zpost,
                                                              int B, int n, int d, int D) {
  extern __shared__ __align__(16) float dsm[];
  float* zp = dsm;
  float* hid = dsm + kDagTB * d;
  const int i = blockIdx.x, b0 = blockIdx.y * kDagTB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* W1 = params[4 * i], *b1 = params[4 * i + 1], *W2 = params[4 * i + 2], *b2 = params[4 * i + 3];

  for (int idx = threadIdx.x; idx < kDagTB * d; idx += kDagThreads) {
    const int b = idx / d, k = idx - b * d;
    float z = 0.f;
    if (b0 + b < B)
      for (int j = 0; j < n; ++j) z = fmaf(__ldg(A + j * n + i), u[((size_t)(b0 + b) * n + j) * d + k], z);
    zp[idx] = z;
  }
  __syncthreads();
  // kDagU output rows per warp and pass: their weight rows are requested together (the grid is 4 x B/8 CTAs, so a pass
  // costs one L2 round trip whatever it computes - r2 ncu: 82 us with one row per pass)
  constexpr int kDagU = 4, kWarps = kDagThreads / 32;
  for (int hb = warp; hb < D; hb += kWarps * kDagU) {
    float acc[kDagU][kDagTB];
#pragma unroll
    for (int q = 0; q < kDagU; ++q)
#pragma unroll
      for (int s = 0; s < kDagTB; ++s) acc[q][s] = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      float4 w[kDagU];
#pragma unroll
      for (int q = 0; q < kDagU; ++q) {
        const int h = hb + q * kWarps;
        w[q] = h < D ? __ldg(reinterpret_cast<const float4*>(W1 + (size_t)h * d + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int s = 0; s < kDagTB; ++s) {
        const float4 z = *reinterpret_cast<const float4*>(zp + s * d + k);
#pragma unroll
        for (int q = 0; q < kDagU; ++q)
          acc[q][s] = fmaf(w[q].x, z.x, fmaf(w[q].y, z.y, fmaf(w[q].z, z.z, fmaf(w[q].w, z.w, acc[q][s]))));
      }
    }
#pragma unroll
    for (int q = 0; q < kDagU; ++q) {
      const int h = hb + q * kWarps;
      if (h < D) {
        const float c = reduce8(acc[q], lane) + __ldg(b1 + h);
        if ((lane & 3) == 0) hid[(lane >> 2) * D + h] = c > 0.f ? c : kLeaky * c;
      }
    }
  }
  __syncthreads();
  for (int kb = warp; kb < d; kb += kWarps * kDagU) {
    float acc[kDagU][kDagTB];
#pragma unroll
    for (int q = 0; q < kDagU; ++q)
#pragma unroll
      for (int s = 0; s < kDagTB; ++s) acc[q][s] = 0.f;
    for (int h = lane * 4; h < D; h += 128) {
      float4 w[kDagU];
#pragma unroll
      for (int q = 0; q < kDagU; ++q) {
        const int k = kb + q * kWarps;
        w[q] = k < d ? __ldg(reinterpret_cast<const float4*>(W2 + (size_t)k * D + h)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int s = 0; s < kDagTB; ++s) {
        const float4 z = *reinterpret_cast<const float4*>(hid + s * D + h);
#pragma unroll
        for (int q = 0; q < kDagU; ++q)
          acc[q][s] = fmaf(w[q].x, z.x, fmaf(w[q].y, z.y, fmaf(w[q].z, z.z, fmaf(w[q].w, z.w, acc[q][s]))));
      }
    }
#pragma unroll
    for (int q = 0; q < kDagU; ++q) {
      const int k = kb + q * kWarps;
      if (k < d) {
        const float c = reduce8(acc[q], lane) + __ldg(b2 + k);
        const int s = lane >> 2;
        if ((lane & 3) == 0 && b0 + s < B) {
          const size_t o = ((size_t)(b0 + s) * n + i) * d + k;
          zpost[o] = c + u[o];
        }
      }
    }
  }
}

// Backward.  grid (n, D / 32): CTA = (variable i, 32 hidden units); the batch is walked in chunks of 32 samples.
//   a = zp W1^T + b1, hid = leaky(a), dhid = dout_i W2, da = dhid * leaky'(a)
//   dW2 += dout_i^T hid, db2 += sum_b dout_i, dW1 += da^T zp, db1 += sum_b da, dzp[b,i,:] += da W1  (atomics over hidden tiles)
// grads: device array of 4n pointers {dW1_i, db1_i, dW2_i, db2_i}, ACCUMULATED into (+=; every element has one owner CTA).
// smem: zp[32][d] | dout[32][d] | W1t[32][d+1] | W2t[d][33] | av[32][33] | dav[32][33]
template <int DV>
__global__ void __launch_bounds__(kDagThreads) dag_bwd_kernel(const float* __restrict__ u, const float* __restrict__ A,
                                                              const float* const* __restrict__ params,
                                                              const float* __restrict__ dzpost, float* const* __restrict__ grads,
                                                              float* __restrict__ dzp, int B, int n, int D) {
  constexpr int d = DV;
  constexpr int R1 = DV / 32;            // float4 column chunks per thread for the [32][d] tiles (dW1, dzp)
  constexpr int R2 = DV / 128 > 0 ? DV / 128 : 1;   // row repeats per thread for the [d][32] tile (dW2)
  extern __shared__ __align__(16) float dsm[];
  float* zp = dsm;
  float* dos = zp + kDagBC * d;
  float* W1t = dos + kDagBC * d;               // [32][d + 1]
  float* W2t = W1t + kDagHT * (d + 1);          // [d][33]
  float* hv = W2t + d * (kDagHT + 1);           // hid  [32 samples][33]
  float* dav = hv + kDagBC * (kDagHT + 1);      // da   [32 samples][33]
  const int i = blockIdx.x, h0 = blockIdx.y * kDagHT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* W1 = params[4 * i], *b1 = params[4 * i + 1], *W2 = params[4 * i + 2];

  for (int idx = tid; idx < kDagHT * d; idx += kDagThreads) {
    const int h = idx / d, k = idx - h * d;
    W1t[h * (d + 1) + k] = W1[(size_t)(h0 + h) * d + k];
  }
  for (int idx = tid; idx < d * kDagHT; idx += kDagThreads) {
    const int k = idx / kDagHT, h = idx - k * kDagHT;
    W2t[k * (kDagHT + 1) + h] = W2[(size_t)k * D + h0 + h];
  }
  const float bias1 = __ldg(b1 + h0 + lane);

  float acc1[R1][4];          // dW1[h = tid / 8][k = ((tid % 8) + 8 r) * 4 ..+3]
  float acc2[R2][16];         // dW2[k = tid / 2 + 128 r][h = (tid % 2) * 16 ..+15]
#pragma unroll
  for (int r = 0; r < R1; ++r) acc1[r][0] = acc1[r][1] = acc1[r][2] = acc1[r][3] = 0.f;
#pragma unroll
  for (int r = 0; r < R2; ++r)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc2[r][c] = 0.f;
  float accb1 = 0.f, accb2[R2 > 1 ? R2 : 1];
#pragma unroll
  for (int r = 0; r < R2; ++r) accb2[r] = 0.f;

  for (int bc = 0; bc < B; bc += kDagBC) {
    __syncthreads();
    for (int idx = tid; idx < kDagBC * d; idx += kDagThreads) {
      const int b = idx / d, k = idx - b * d;
      float z = 0.f, g = 0.f;
      if (bc + b < B) {
        for (int j = 0; j < n; ++j) z = fmaf(__ldg(A + j * n + i), u[((size_t)(bc + b) * n + j) * d + k], z);
        g = dzpost[((size_t)(bc + b) * n + i) * d + k];
      }
      zp[idx] = z; dos[idx] = g;
    }
    __syncthreads();
    // a, dhid for (b = warp * 4 + r, h = lane)
    {
      float a[4] = {bias1, bias1, bias1, bias1}, dh[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < d; ++k) {
        const float w1 = W1t[lane * (d + 1) + k], w2 = W2t[k * (kDagHT + 1) + lane];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          a[r] = fmaf(w1, zp[(warp * 4 + r) * d + k], a[r]);
          dh[r] = fmaf(w2, dos[(warp * 4 + r) * d + k], dh[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int b = warp * 4 + r;
        const bool pos = a[r] > 0.f;
        hv[b * (kDagHT + 1) + lane] = pos ? a[r] : kLeaky * a[r];
        dav[b * (kDagHT + 1) + lane] = pos ? dh[r] : kLeaky * dh[r];
      }
    }
    __syncthreads();
    // dW1 tile and dzp chunk share the thread map (row = tid / 8, float4 columns (tid % 8) + 8 r)
    {
      const int row = tid >> 3, cq = tid & 7;
      for (int b = 0; b < kDagBC; ++b) {
        const float da = dav[b * (kDagHT + 1) + row];
#pragma unroll
        for (int r = 0; r < R1; ++r) {
          const float4 z = *reinterpret_cast<const float4*>(zp + b * d + (cq + 8 * r) * 4);
          acc1[r][0] = fmaf(da, z.x, acc1[r][0]); acc1[r][1] = fmaf(da, z.y, acc1[r][1]);
          acc1[r][2] = fmaf(da, z.z, acc1[r][2]); acc1[r][3] = fmaf(da, z.w, acc1[r][3]);
        }
      }
      // dzp[b = row][k] += sum_h da[b][h] W1[h][k]
      if (bc + row < B) {
        float dz[R1][4];
#pragma unroll
        for (int r = 0; r < R1; ++r) dz[r][0] = dz[r][1] = dz[r][2] = dz[r][3] = 0.f;
        for (int h = 0; h < kDagHT; ++h) {
          const float da = dav[row * (kDagHT + 1) + h];
#pragma unroll
          for (int r = 0; r < R1; ++r) {
            const float* w = W1t + h * (d + 1) + (cq + 8 * r) * 4;
            dz[r][0] = fmaf(da, w[0], dz[r][0]); dz[r][1] = fmaf(da, w[1], dz[r][1]);
            dz[r][2] = fmaf(da, w[2], dz[r][2]); dz[r][3] = fmaf(da, w[3], dz[r][3]);
          }
        }
        float* dst = dzp + ((size_t)(bc + row) * n + i) * d;
#pragma unroll
        for (int r = 0; r < R1; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e) atomicAdd(dst + (cq + 8 * r) * 4 + e, dz[r][e]);
      }
    }
    // dW2 tile: rows k = tid / 2 + 128 r, columns (tid % 2) * 16 ..+15 ; db2 on the first hidden tile
    {
      const int hq = (tid & 1) * 16;
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        const int k = (tid >> 1) + 128 * r;
        if (k < d) {
          for (int b = 0; b < kDagBC; ++b) {
            const float g = dos[b * d + k];
            const float* hrow = hv + b * (kDagHT + 1) + hq;
#pragma unroll
            for (int c = 0; c < 16; ++c) acc2[r][c] = fmaf(g, hrow[c], acc2[r][c]);
            if (hq == 0) accb2[r] += g;
          }
        }
      }
    }
    if (tid < kDagHT)
      for (int b = 0; b < kDagBC; ++b) accb1 += dav[b * (kDagHT + 1) + tid];
  }
  // write-out: each gradient element is owned by exactly one thread of one CTA
  float* gW1 = grads[4 * i], *gb1 = grads[4 * i + 1], *gW2 = grads[4 * i + 2], *gb2 = grads[4 * i + 3];
  {
    const int row = tid >> 3, cq = tid & 7;
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      float4* dst = reinterpret_cast<float4*>(gW1 + (size_t)(h0 + row) * d + (cq + 8 * r) * 4);
      float4 v = *dst;
      v.x += acc1[r][0]; v.y += acc1[r][1]; v.z += acc1[r][2]; v.w += acc1[r][3];
      *dst = v;
    }
    const int hq = (tid & 1) * 16;
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      const int k = (tid >> 1) + 128 * r;
      if (k < d) {
        float* dst = gW2 + (size_t)k * D + h0 + hq;
#pragma unroll
        for (int c = 0; c < 16; ++c) dst[c] += acc2[r][c];
        if (blockIdx.y == 0 && hq == 0) gb2[k] += accb2[r];
      }
    }
    if (tid < kDagHT) gb1[h0 + tid] += accb1;
  }
}

// du[b,j,:] = dzpost[b,j,:] + sum_i A[j,i] dzp[b,i,:] ; dzp is cleared for the next call
__global__ void __launch_bounds__(kDagThreads) dag_bwd_finish_kernel(const float* __restrict__ A, const float* __restrict__ dzpost,
                                                                     float* __restrict__ dzp, float* __restrict__ du,
                                                                     const float* __restrict__ du_add, int B, int n, int d) {
  const int64_t total = (int64_t)B * d;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / d;
    const int k = (int)(idx - b * d);
    float z[8];
    for (int i = 0; i < n; ++i) { float* q = dzp + ((size_t)b * n + i) * d + k; z[i] = *q; *q = 0.f; }
    for (int j = 0; j < n; ++j) {
      const size_t o = ((size_t)b * n + j) * d + k;
      float v = du ? dzpost[o] : 0.f;
      if (du_add) v += du_add[o];
      for (int i = 0; i < n; ++i) v = fmaf(__ldg(A + j * n + i), z[i], v);
      if (du) du[o] = v;
    }
  }
}

template <int DV>
static int dag_bwd_launch(const float* u, const float* A, const float* const* params, const float* dzpost,
                          float* const* grads, float* dzp, int B, int n, int D, cudaStream_t st) {
  constexpr size_t smem = sizeof(float) * (2 * kDagBC * DV + kDagHT * (DV + 1) + DV * (kDagHT + 1) + 2 * kDagBC * (kDagHT + 1));
  static cudaError_t attr_err = cudaFuncSetAttribute(dag_bwd_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (attr_err != cudaSuccess) { set_error("dag_bwd smem attribute: %s", cudaGetErrorString(attr_err)); return CDAE_ERR_CUDA; }
  dag_bwd_kernel<DV><<<dim3(n, D / kDagHT), kDagThreads, smem, st>>>(u, A, params, dzpost, grads, dzp, B, n, D);
  CDAE_CHECK_LAUNCH("dag_bwd_kernel");
  return CDAE_OK;
}

}  // namespace cdae
using namespace cdae;

extern "C" int cdae_dag_fwd(const float* u, const float* A, const void* const* params, float* zpost, int B, int n, int d,
                            int D, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(u && A && params && zpost, "dag_fwd: null pointer");
  CDAE_CHECK_SHAPE(n >= 1 && n <= 8 && d % 4 == 0 && D % 4 == 0 && d >= 4 && D >= 4, "dag_fwd: n=%d d=%d D=%d unsupported", n, d, D);
  const size_t smem = sizeof(float) * (size_t)kDagTB * (d + D);
  CDAE_CHECK_SHAPE(smem <= 48 * 1024, "dag_fwd: d + D = %d too large", d + D);
  dag_fwd_kernel<<<dim3(n, (B + kDagTB - 1) / kDagTB), kDagThreads, smem, (cudaStream_t)s>>>(
      u, A, reinterpret_cast<const float* const*>(params), zpost, B, n, d, D);
  CDAE_CHECK_LAUNCH("dag_fwd_kernel");
  return CDAE_OK;
}

extern "C" int cdae_dag_bwd(const float* u, const float* A, const void* const* params, const float* dzpost,
                            void* const* grads, float* dzp_ws, float* du, const float* du_add, int B, int n, int d, int D,
                            cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(u && A && params && dzpost && grads && dzp_ws, "dag_bwd: null pointer");
  CDAE_CHECK_SHAPE(n >= 1 && n <= 8 && D % kDagHT == 0, "dag_bwd: n=%d D=%d unsupported", n, D);
  const float* const* pp = reinterpret_cast<const float* const*>(params);
  float* const* gp = reinterpret_cast<float* const*>(grads);
  cudaStream_t st = (cudaStream_t)s;
  int rc;
  switch (d) {
    case 64: rc = dag_bwd_launch<64>(u, A, pp, dzpost, gp, dzp_ws, B, n, D, st); break;
    case 128: rc = dag_bwd_launch<128>(u, A, pp, dzpost, gp, dzp_ws, B, n, D, st); break;
    case 256: rc = dag_bwd_launch<256>(u, A, pp, dzpost, gp, dzp_ws, B, n, D, st); break;
    default: set_error("dag_bwd: per-variable width d=%d unsupported (64, 128, 256)", d); return CDAE_ERR_SHAPE;
  }
  if (rc) return rc;
  const int64_t total = (int64_t)B * d;
  int blocks = (int)((total + kDagThreads - 1) / kDagThreads);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  dag_bwd_finish_kernel<<<blocks, kDagThreads, 0, st>>>(A, dzpost, dzp_ws, du, du_add, B, n, d);
  CDAE_CHECK_LAUNCH("dag_bwd_finish_kernel");
  return CDAE_OK;
}
