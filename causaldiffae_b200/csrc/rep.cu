// The [B, 512]-sized representation path of CausalDiffAE in hand-written fp32 kernels (no cuBLAS / cuDNN / ATen):
//   timestep embedding + time_embed / up_emb / emb_layers GEMMs         (ref nn.py:551-569, unet.py:545-554,148-154,616)
//   GaussianConvEncoder: [conv3x3 s2 -> BatchNorm2d(batch stats) -> LeakyReLU] x L -> fc_mu / softplus(fc_var)
//                                                                        (ref nn.py:15-110), forward and backward
//   reparameterisation, classifier-free keep mask, the closed-form KL of representation_loss and its gradient
//                                                                        (ref nn.py:440-467, unet.py:590-613,
//                                                                         gaussian_diffusion.py:718-766)
// The path is 0.03 % of the step's FLOPs: what matters is launch count and accuracy (fp32 CUDA-core math, <= 1e-4 against
// the oracle), not tensor cores.  One register-tiled SGEMM kernel with pluggable operand loaders covers every GEMM-shaped
// piece: dense (any strides), SiLU-on-load, and the implicit im2col view of a stride-2 3x3 convolution with the previous
// layer's BatchNorm + LeakyReLU applied on load; its epilogue adds the bias, applies softplus, accumulates BatchNorm column
// statistics in double, or scatters (col2im) with atomics for the data gradient.  Split-K over grid.z with fp32 atomics.
#include "common.cuh"

namespace cdae {

constexpr float kLeakyRep = 0.01f;

struct ConvGeo {
  int64_t sb, sc, sh, sw;     // strides (elements) of the gathered / scattered activation tensor
  int Cin, H, W, OH, OW;      // stride 2, padding 1, 3x3
  const float* ab;            // per input channel {a, b}: value = leaky(a*x + b) (BatchNorm + LeakyReLU of the producer) or null
};

struct SgemmP {
  const float* A; int64_t a_sm, a_sk; int a_mode;   // 0 dense, 1 dense + SiLU, 2 im2col rows=pixels cols=k, 3 im2col transposed
  const float* B; int64_t b_sk, b_sn; int b_mode;   // 0 dense, 1 dense + SiLU
  float* C; int64_t c_sm, c_sn; int c_mode;         // 0 store, 1 atomic add (+=), 2 col2im scatter-add (rows = pixels, cols = k)
  const float* bias;                                // + bias[n] (added by K split 0 only)
  int act_out;                                      // 1: softplus(v) + 1e-8
  double* colstats;                                 // [N][2] += {sum, sum of squares} over the rows of C as stored
  int M, N, K, k_per_split;
  ConvGeo g;
};

__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// conv k = (ci, kh, kw) -> ci, kh, kw
__device__ __forceinline__ void decode_k(int k, int& ci, int& kh, int& kw) {
  ci = k / 9; const int t = k - ci * 9; kh = t / 3; kw = t - kh * 3;
}
__device__ __forceinline__ void decode_m(const ConvGeo& g, int m, int& b, int& oh, int& ow) {
  const int p = g.OH * g.OW;
  b = m / p; const int r = m - b * p; oh = r / g.OW; ow = r - oh * g.OW;
}
__device__ __forceinline__ float gather_px(const ConvGeo& g, const float* src, int b, int oh, int ow, int ci, int kh, int kw) {
  const int ih = 2 * oh + kh - 1, iw = 2 * ow + kw - 1;
  if ((unsigned)ih >= (unsigned)g.H || (unsigned)iw >= (unsigned)g.W) return 0.f;
  float v = src[b * g.sb + ci * g.sc + ih * g.sh + iw * g.sw];
  if (g.ab) {
    v = fmaf(__ldg(g.ab + 2 * ci), v, __ldg(g.ab + 2 * ci + 1));
    v = v > 0.f ? v : kLeakyRep * v;
  }
  return v;
}

template <int BM, int BN>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmP p) {
  constexpr int BK = 16, TM = BM / 16, TN = BN / 16;
  constexpr int EA = BM * BK / 256, EB = BN * BK / 256;
  __shared__ __align__(16) float smem_ab[BK * (BM + 4) + BK * (BN + 4)];
  float (*As)[BM + 4] = reinterpret_cast<float (*)[BM + 4]>(smem_ab);
  float (*Bs)[BN + 4] = reinterpret_cast<float (*)[BN + 4]>(smem_ab + BK * (BM + 4));
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb = blockIdx.z * p.k_per_split;
  const int ke = min(p.K, kb + p.k_per_split);
  // loader maps: "k-fast" = consecutive threads walk k (row-major A / the im2col view), else consecutive threads walk m
  const bool a_kfast = (p.a_mode == 2) || (p.a_mode < 2 && p.a_sk == 1);
  const bool b_nfast = (p.b_sn == 1);
  int a_i[EA], a_k[EA], b_k[EB], b_n[EB];
#pragma unroll
  for (int j = 0; j < EA; ++j) {
    const int e = tid + 256 * j;
    if (a_kfast) { a_k[j] = e % BK; a_i[j] = e / BK; } else { a_i[j] = e % BM; a_k[j] = e / BM; }
  }
#pragma unroll
  for (int j = 0; j < EB; ++j) {
    const int e = tid + 256 * j;
    if (b_nfast) { b_n[j] = e % BN; b_k[j] = e / BN; } else { b_k[j] = e % BK; b_n[j] = e / BK; }
  }
  // im2col: the GEMM index that does not move with the K loop is decoded once
  int r_b[EA], r_h[EA], r_w[EA];
  if (p.a_mode == 2) {
#pragma unroll
    for (int j = 0; j < EA; ++j) decode_m(p.g, min(m0 + a_i[j], p.M - 1), r_b[j], r_h[j], r_w[j]);
  } else if (p.a_mode == 3) {
#pragma unroll
    for (int j = 0; j < EA; ++j) decode_k(min(m0 + a_i[j], p.M - 1), r_b[j], r_h[j], r_w[j]);   // (ci, kh, kw)
  }
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto load_a = [&](int k0, float* ra) {
#pragma unroll
    for (int j = 0; j < EA; ++j) {
      const int m = m0 + a_i[j], k = k0 + a_k[j];
      float v = 0.f;
      if (m < p.M && k < ke) {
        if (p.a_mode < 2) {
          v = p.A[m * p.a_sm + k * p.a_sk];
          if (p.a_mode == 1) v = silu_exact(v);
        } else if (p.a_mode == 2) {
          int ci, kh, kw;
          decode_k(k, ci, kh, kw);
          v = gather_px(p.g, p.A, r_b[j], r_h[j], r_w[j], ci, kh, kw);
        } else {
          int b, oh, ow;
          decode_m(p.g, k, b, oh, ow);
          v = gather_px(p.g, p.A, b, oh, ow, r_b[j], r_h[j], r_w[j]);
        }
      }
      ra[j] = v;
    }
  };
  auto load_b = [&](int k0, float* rb) {
#pragma unroll
    for (int j = 0; j < EB; ++j) {
      const int k = k0 + b_k[j], n = n0 + b_n[j];
      float v = 0.f;
      if (k < ke && n < p.N) {
        v = p.B[k * p.b_sk + n * p.b_sn];
        if (p.b_mode == 1) v = silu_exact(v);
      }
      rb[j] = v;
    }
  };

  // The tiles are small and the grids of this path often do not fill the chip, so one K step costs a global-load round
  // trip, not its FMAs (ncu r2: ~1.2 us per step with ONE tile in flight).  kPD tiles are kept in flight in registers.
  constexpr int kPD = 3;
  float ra[kPD][EA], rb[kPD][EB];
#pragma unroll
  for (int s = 0; s < kPD; ++s)
    if (kb + s * BK < ke) { load_a(kb + s * BK, ra[s]); load_b(kb + s * BK, rb[s]); }
  for (int kbase = kb; kbase < ke; kbase += kPD * BK) {
#pragma unroll
    for (int s = 0; s < kPD; ++s) {
      const int k0 = kbase + s * BK;
      if (k0 >= ke) break;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < EA; ++j) As[a_k[j]][a_i[j]] = ra[s][j];
#pragma unroll
      for (int j = 0; j < EB; ++j) Bs[b_k[j]][b_n[j]] = rb[s][j];
      __syncthreads();
      if (k0 + kPD * BK < ke) { load_a(k0 + kPD * BK, ra[s]); load_b(k0 + kPD * BK, rb[s]); }
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // ---------------------------------------------------------------- epilogue
  float cs[TN], cq[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) cs[j] = cq[j] = 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    int sb_ = 0, soh = 0, sow = 0;
    if (p.c_mode == 2) decode_m(p.g, m, sb_, soh, sow);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias && blockIdx.z == 0) v += __ldg(p.bias + n);
      if (p.act_out == 1) v = (v > 20.f ? v : log1pf(expf(v))) + 1e-8f;
      if (p.c_mode == 0) {
        p.C[m * p.c_sm + n * p.c_sn] = v;
      } else if (p.c_mode == 1) {
        atomicAdd(p.C + m * p.c_sm + n * p.c_sn, v);
      } else {
        int ci, kh, kw;
        decode_k(n, ci, kh, kw);
        const int ih = 2 * soh + kh - 1, iw = 2 * sow + kw - 1;
        if ((unsigned)ih < (unsigned)p.g.H && (unsigned)iw < (unsigned)p.g.W)
          atomicAdd(p.C + sb_ * p.g.sb + ci * p.g.sc + ih * p.g.sh + iw * p.g.sw, v);
      }
      cs[j] += v; cq[j] = fmaf(v, v, cq[j]);
    }
  }
  if (p.colstats) {
    // column sums over the tile rows: 16 row-threads per column through shared memory, one double atomic per column
    __syncthreads();
    float* red = smem_ab;                   // 2 * 16 * BN floats of the operand tiles, which are dead by now
    static_assert(2 * 16 * BN <= BK * (BM + 4) + BK * (BN + 4), "statistics scratch does not fit");
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      red[ty * BN + tx * TN + j] = cs[j];
      red[16 * BN + ty * BN + tx * TN + j] = cq[j];
    }
    __syncthreads();
    if (tid < BN && n0 + tid < p.N) {
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) { s += (double)red[r * BN + tid]; q += (double)red[16 * BN + r * BN + tid]; }
      atomicAdd(p.colstats + 2 * (n0 + tid), s);
      atomicAdd(p.colstats + 2 * (n0 + tid) + 1, q);
    }
  }
}

// ---------------------------------------------------------------- small elementwise pieces
// timestep_embedding (ref nn.py:551-569): out[b, j] = cos(t_b f_j), out[b, half + j] = sin(t_b f_j); freqs come from the
// host (computed exactly like the reference and uploaded once), t is int64 or fp32 (rescale_timesteps)
__global__ void temb_kernel(const void* t, int t_is_float, int t_stride, const int64_t* __restrict__ map, float scale,
                            const float* __restrict__ freqs, float* __restrict__ out, int B, int half, int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim, j = idx - b * dim;
  float tv;
  if (t_is_float) {
    tv = reinterpret_cast<const float*>(t)[b * t_stride];
  } else {
    int64_t ti = reinterpret_cast<const int64_t*>(t)[b * t_stride];
    if (map) ti = map[ti];                       // _WrappedModel: respaced step -> original step (ref respace.py:119-124)
    tv = (float)ti;
    if (scale != 0.f) tv *= scale;               // rescale_timesteps: * 1000 / T in fp32
  }
  float v = 0.f;
  if (j < half) v = cosf(tv * freqs[j]);
  else if (j < 2 * half) v = sinf(tv * freqs[j - half]);
  out[idx] = v;
}

// y (bf16) = silu(x) (or x): the A operand of the tensor-core FiLM projection
__global__ void silu_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n, int silu) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = __float2bfloat16_rn(silu ? silu_exact(v) : v);
  }
}

// g *= silu'(x)
__global__ void silu_bwd_kernel(float* __restrict__ g, const float* __restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i], s = 1.0f / (1.0f + expf(-xv));
    g[i] *= s * (1.0f + xv * (1.0f - s));
  }
}

// out[b, :] += table[idx[b], :] (label_emb, ref unet.py:549-551); backward: dtable[idx[b], :] += g[b, :]
__global__ void embed_rows_kernel(float* __restrict__ out, const float* __restrict__ table, const int64_t* __restrict__ idx,
                                  int B, int D, int backward, float* __restrict__ dtable) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, k = i - b * D;
  if (!backward) out[i] += table[idx[b] * D + k];
  else atomicAdd(dtable + idx[b] * D + k, out[i]);
}

// BatchNorm2d bookkeeping of one encoder layer (ref nn.py:52-61 nn.BatchNorm2d defaults: eps 1e-5, momentum 0.1).
// train: batch statistics from the double column sums the conv epilogue accumulated (biased variance normalises, unbiased
// variance updates running_var), running buffers advanced; eval: running statistics.  Writes ab[c] = {a, b} with
// bn(x) = a x + b, and ms[c] = {mean, rstd} for the backward.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                   int64_t* __restrict__ nbt, int train, float* __restrict__ ab, float* __restrict__ ms, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float mean, var;
    if (train) {
      const double m = stats[2 * c] / count;
      double v = stats[2 * c + 1] / count - m * m;
      if (v < 0.0) v = 0.0;
      mean = (float)m; var = (float)v;
      const double unb = count > 1.0 ? v * count / (count - 1.0) : v;
      rmean[c] = 0.9f * rmean[c] + 0.1f * mean;
      rvar[c] = 0.9f * rvar[c] + 0.1f * (float)unb;
    } else {
      mean = rmean[c]; var = rvar[c];
    }
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    const float a = gamma[c] * rstd;
    ab[2 * c] = a; ab[2 * c + 1] = beta[c] - mean * a;
    ms[2 * c] = mean; ms[2 * c + 1] = rstd;
  }
  if (train && nbt && c == 0) *nbt += 1;
}

// last encoder layer: hfeat[b, c*P + p] = leaky(a_c raw[b, p, c] + b_c)   (NHWC raw -> the reference's NCHW flatten order)
// backward: dact[b, p, c] = dhfeat[b, c*P + p]
__global__ void enc_head_kernel(const float* __restrict__ raw, const float* __restrict__ ab, float* __restrict__ hfeat, int B,
                                int P, int C, int backward) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P * C) return;
  const int c = i % C, p = (i / C) % P, b = i / (C * P);
  if (!backward) {
    const float v = fmaf(ab[2 * c], raw[i], ab[2 * c + 1]);
    hfeat[(b * C + c) * P + p] = v > 0.f ? v : kLeakyRep * v;
  } else {
    hfeat[i] = raw[(b * C + c) * P + p];      // here: raw = dhfeat (flatten order), hfeat = dact (NHWC)
  }
}

// BatchNorm + LeakyReLU backward, pass 1: per channel sums of g = dact * leaky'(a x + b) and g * xhat  (double atomics)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dact, const float* __restrict__ raw,
                                                            const float* __restrict__ ab, const float* __restrict__ ms,
                                                            double* __restrict__ sums, int64_t M, int C) {
  // thread = channel (consecutive threads = consecutive channels of one NHWC row); grid.y walks row chunks
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = ab[2 * c], bb = ab[2 * c + 1], mean = ms[2 * c], rstd = ms[2 * c + 1];
  const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f, q = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float x = raw[r * C + c];
    float g = dact[r * C + c];
    if (fmaf(a, x, bb) <= 0.f) g *= kLeakyRep;
    s += g; q = fmaf(g, (x - mean) * rstd, q);
  }
  atomicAdd(sums + 2 * c, (double)s);
  atomicAdd(sums + 2 * c + 1, (double)q);
}

// pass 2: draw = gamma rstd (g - sum_g / M - xhat sum_gx / M), written over dact; block (c-tile, row-chunk 0) also emits
// dgamma += sum_gx, dbeta += sum_g
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ dact, const float* __restrict__ raw,
                                                           const float* __restrict__ ab, const float* __restrict__ ms,
                                                           const float* __restrict__ gamma, const double* __restrict__ sums,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M,
                                                           int C) {
  const int64_t total = M * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float a = ab[2 * c], bb = ab[2 * c + 1], mean = ms[2 * c], rstd = ms[2 * c + 1];
    const float x = raw[i];
    float g = dact[i];
    if (fmaf(a, x, bb) <= 0.f) g *= kLeakyRep;
    const float sg = (float)(sums[2 * c] / (double)M), sgx = (float)(sums[2 * c + 1] / (double)M);
    dact[i] = gamma[c] * rstd * (g - sg - (x - mean) * rstd * sgx);
    if (i < C) {
      atomicAdd(dgamma + c, (float)sums[2 * c + 1]);
      atomicAdd(dbeta + c, (float)sums[2 * c]);
    }
  }
}

// reparameterisation + classifier-free keep mask + closed-form KL (ref nn.py:460-467, unet.py:590-613,
// gaussian_diffusion.py:718-766).  One CTA per sample.
//   z      = (zp + sqrt(0.001 var) xi) keep            zp = z_post of the DAG layer (or mu without causal modelling)
//   zp_out = zp keep
//   kld[b] = 0.5 sum(-log var + var + mu^2 - 1) + [causal] 0.5 sum_i sum_k (zp_out[b,i,k] - c[b,i])^2
// backward (dz from up_emb, dkld[b] from the loss, optional extra dmu/dvar/dzp from autograd users):
//   dzp  = keep (dz + dkld (zp_out - c))        dvar = dz keep xi 0.5 sqrt(0.001 / var) + dkld 0.5 (1 - 1/var)
//   dmu  = dkld mu  (the DAG layer's du is added by its own backward)
__global__ void __launch_bounds__(128) latent_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ var,
                                                         const float* __restrict__ zp, const float* __restrict__ xi,
                                                         const float* __restrict__ keep, const float* __restrict__ c,
                                                         float* __restrict__ z, float* __restrict__ zp_out,
                                                         float* __restrict__ kld, int D, int n, int causal, float var_scale) {
  const int b = blockIdx.x;
  const float kp = keep ? keep[b] : 1.f;
  const int d = D / n;
  float acc = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const int64_t o = (int64_t)b * D + k;
    const float m = mu[o], v = var[o], p = zp[o];
    const float zo = p * kp;
    z[o] = (p + sqrtf(v * var_scale) * xi[o]) * kp;
    if (zp_out) zp_out[o] = zo;
    float t = 0.5f * (-logf(v) + v + m * m - 1.f);
    if (causal && c) { const float dlt = zo - c[b * n + k / d]; t += 0.5f * dlt * dlt; }
    acc += t;
  }
  if (kld) {
    __shared__ float red[4];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) kld[b] = red[0] + red[1] + red[2] + red[3];
  }
}

__global__ void __launch_bounds__(128) latent_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ var,
                                                         const float* __restrict__ zp, const float* __restrict__ xi,
                                                         const float* __restrict__ keep, const float* __restrict__ c,
                                                         const float* __restrict__ dz, const float* __restrict__ dkld,
                                                         const float* __restrict__ dzp_ext, const float* __restrict__ dmu_ext,
                                                         const float* __restrict__ dvar_ext, float* __restrict__ dzp,
                                                         float* __restrict__ dmu, float* __restrict__ dvar, int D, int n,
                                                         int causal, float var_scale) {
  const int b = blockIdx.x;
  const float kp = keep ? keep[b] : 1.f;
  const float gk = dkld ? dkld[b] : 0.f;
  const int d = D / n;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const int64_t o = (int64_t)b * D + k;
    const float m = mu[o], v = var[o], p = zp[o], g = dz ? dz[o] : 0.f;
    float gp = g * kp;
    if (causal && c) gp += gk * (p * kp - c[b * n + k / d]) * kp;
    if (dzp_ext) gp += dzp_ext[o] * kp;           // extra gradient w.r.t. the MASKED z_post (autograd users)
    dzp[o] = gp;
    float gv = g * kp * xi[o] * 0.5f * sqrtf(var_scale / v) + gk * 0.5f * (1.f - 1.f / v);
    if (dvar_ext) gv += dvar_ext[o];
    dvar[o] = gv;
    float gm = gk * m;
    if (dmu_ext) gm += dmu_ext[o];
    dmu[o] = gm;
  }
}

// softplus backward for fc_var: g *= sigmoid(pre) where var = softplus(pre) + 1e-8  ->  sigmoid(pre) = 1 - exp(-(var - 1e-8))
__global__ void softplus_bwd_kernel(float* __restrict__ g, const float* __restrict__ var, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    g[i] *= 1.0f - expf(-(var[i] - 1e-8f));
}

// loss assembly of one training step (ref gaussian_diffusion.py:849-855, train_util.py:266): per-sample
//   loss[b] = mse[b] + klw * kld_rep   with kld_rep = kld[b], or sum(kld keep) / sum(keep) under classifier-free masking,
//   total   = mean_b(loss[b] * w[b])
// and the gradient scales the backward kernels need: gscale[b] = w[b] / B (d total / d mse[b]) and dkld[b].
// One CTA.  stats (optional, += for the logger): {sum loss w, sum mse w, sum kld_rep w, B, quartile sums ...} see host.
__global__ void __launch_bounds__(256) step_loss_kernel(const float* __restrict__ mse, const float* __restrict__ kld,
                                                        const float* __restrict__ keep, const float* __restrict__ w,
                                                        const float* __restrict__ klw_p, const int64_t* __restrict__ t,
                                                        int num_timesteps, int B, float* __restrict__ loss,
                                                        float* __restrict__ gscale, float* __restrict__ dkld,
                                                        float* __restrict__ total, float* __restrict__ logsums) {
  __shared__ float s_a[256], s_b[256], s_c[256];
  const int tid = threadIdx.x;
  const float klw = klw_p ? *klw_p : 0.f;
  float num = 0.f, den = 0.f, wsum = 0.f;
  for (int b = tid; b < B; b += 256) {
    if (kld && keep) { num += kld[b] * keep[b]; den += keep[b]; }
    wsum += w[b];
  }
  s_a[tid] = num; s_b[tid] = den; s_c[tid] = wsum;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s_a[tid] += s_a[tid + o]; s_b[tid] += s_b[tid + o]; s_c[tid] += s_c[tid + o]; }
    __syncthreads();
  }
  num = s_a[0]; den = s_b[0]; wsum = s_c[0];
  __syncthreads();
  const float invB = 1.0f / (float)B;
  const float kmask = (kld && keep) ? num / den : 0.f;
  float tl = 0.f;
  for (int b = tid; b < B; b += 256) {
    const float kr = kld ? (keep ? kmask : kld[b]) : 0.f;
    const float l = mse[b] + klw * kr;
    if (loss) loss[b] = l;
    gscale[b] = w[b] * invB;
    if (dkld) dkld[b] = keep ? klw * wsum * invB * keep[b] / den : klw * w[b] * invB;
    tl += l * w[b];
    if (logsums) {
      // logger sums (train_util.py:401-407 log_loss_dict): [0..2] sum of loss*w, mse*w, kld*w; [3] count;
      // [4 + 4*key + q] per-quartile sums, [16 + q] per-quartile counts
      const int q = min(3, max(0, (int)(4 * t[b] / num_timesteps)));
      atomicAdd(logsums + 0, l * w[b]); atomicAdd(logsums + 1, mse[b] * w[b]); atomicAdd(logsums + 2, kr * w[b]);
      atomicAdd(logsums + 3, 1.f);
      atomicAdd(logsums + 4 + q, l * w[b]); atomicAdd(logsums + 8 + q, mse[b] * w[b]); atomicAdd(logsums + 12 + q, kr * w[b]);
      atomicAdd(logsums + 16 + q, 1.f);
    }
  }
  s_a[tid] = tl;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_a[tid] += s_a[tid + o];
    __syncthreads();
  }
  if (tid == 0 && total) *total = s_a[0] * invB;
}

// Standard-normal (or Bernoulli keep-mask) draws from a counter-based generator: Philox4x32-10 keyed by *seed, counter =
// (*offset + element group); Box-Muller on the four 32-bit outputs.  seed / offset live in device memory and the offset is
// advanced by the launch itself, so the draw is replayable inside a CUDA graph (each replay continues the stream).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ out, int64_t n, const uint64_t* __restrict__ state,
                                                    int bernoulli, float keep_prob) {
  const uint64_t seed = state[0], off = state[1];
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g * 4 < n; g += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = off + (uint64_t)g;
    uint32_t r[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    float v[4];
    if (bernoulli) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = ((r[i] >> 8) * (1.0f / 16777216.0f)) < keep_prob ? 1.f : 0.f;
    } else {
      const float u0 = ((r[0] >> 8) + 1) * (1.0f / 16777216.0f), u1 = (r[1] >> 8) * (1.0f / 16777216.0f);
      const float u2 = ((r[2] >> 8) + 1) * (1.0f / 16777216.0f), u3 = (r[3] >> 8) * (1.0f / 16777216.0f);
      const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
      float s0, c0, s1, c1;
      sincospif(2.0f * u1, &s0, &c0);
      sincospif(2.0f * u3, &s1, &c1);
      v[0] = ra * c0; v[1] = ra * s0; v[2] = rb * c1; v[3] = rb * s1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (g * 4 + i < n) out[g * 4 + i] = v[i];
  }
}
__global__ void randn_tick_kernel(uint64_t* state, uint64_t groups) { state[1] += groups; }

// inverted dropout in place, 8 bf16 per thread: one Philox call = eight 16-bit uniforms
__global__ void __launch_bounds__(256) dropout_kernel(uint4* __restrict__ x, int64_t ngroups, const uint64_t* __restrict__ state,
                                                      uint64_t layer_off, const float* __restrict__ pp) {
  const float p = *pp;
  if (p <= 0.f) return;                                        // eval mode: identity
  const uint64_t seed = state[0], base = state[1] + layer_off;
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  const float scale = 1.0f / (1.0f - p);
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < ngroups; g += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = base + (uint64_t)g;
    uint32_t r[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5eedu, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    uint4 v = x[g];
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool k0 = (r[i] & 0xffffu) >= thr, k1 = (r[i] >> 16) >= thr;
      const float lo = k0 ? __uint_as_float(w[i] << 16) * scale : 0.f;
      const float hi = k1 ? __uint_as_float(w[i] & 0xffff0000u) * scale : 0.f;
      const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    x[g] = v;
  }
}
// the sampler's step counters (int64 for the embedding kernel, int32 for cdae_ddim_step) move together on the device
__global__ void step_tick_kernel(int64_t* a, int32_t* b, int delta) {
  if (a) *a += delta;
  if (b) *b += delta;
}

}  // namespace cdae
using namespace cdae;

// out[n] ~ N(0,1) (bernoulli = 0) or 1[u < keep_prob] (bernoulli = 1); state: device uint64 {seed, offset}, offset += ceil(n/4)
extern "C" int cdae_randn(float* out, int64_t n, void* state, int bernoulli, float keep_prob, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(out && state, "randn: null pointer");
  const int64_t groups = (n + 3) / 4;
  int64_t blocks = (groups + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  randn_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(out, n, reinterpret_cast<const uint64_t*>(state), bernoulli, keep_prob);
  CDAE_CHECK_LAUNCH("randn_kernel");
  randn_tick_kernel<<<1, 1, 0, (cudaStream_t)s>>>(reinterpret_cast<uint64_t*>(state), (uint64_t)groups);
  CDAE_CHECK_LAUNCH("randn_tick_kernel");
  return CDAE_OK;
}

extern "C" int cdae_dropout(void* x, int64_t n, const void* state, int64_t layer_offset, const float* p, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(x && state && p, "dropout: null pointer");
  CDAE_CHECK_SHAPE(n % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "dropout: n %% 8 and 16-byte alignment required");
  const int64_t groups = n / 8;
  int64_t blocks = (groups + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  dropout_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(reinterpret_cast<uint4*>(x), groups,
                                                          reinterpret_cast<const uint64_t*>(state), (uint64_t)layer_offset, p);
  CDAE_CHECK_LAUNCH("dropout_kernel");
  return CDAE_OK;
}

extern "C" int cdae_step_tick(int64_t* step64, int32_t* step32, int delta, cdae_stream s) {
  step_tick_kernel<<<1, 1, 0, (cudaStream_t)s>>>(step64, step32, delta);
  CDAE_CHECK_LAUNCH("step_tick_kernel");
  return CDAE_OK;
}

extern "C" int cdae_sgemm(const cdae_sgemm_desc* d, cdae_stream s) {
  CDAE_CHECK_ARG(d && d->A && d->B && d->C, "sgemm: null pointer");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return CDAE_OK;
  CDAE_CHECK_ARG(d->a_mode >= 0 && d->a_mode <= 3 && d->b_mode >= 0 && d->b_mode <= 1 && d->c_mode >= 0 && d->c_mode <= 2,
                 "sgemm: bad operand mode");
  CDAE_CHECK_ARG(!(d->colstats && (d->c_mode != 0 || d->splits > 1)), "sgemm: column statistics need a plain store and one K split");
  CDAE_CHECK_ARG(!(d->act_out && (d->c_mode != 0 || d->splits > 1)), "sgemm: the output activation needs a plain store and one K split");
  SgemmP p;
  p.A = d->A; p.a_sm = d->a_sm; p.a_sk = d->a_sk; p.a_mode = d->a_mode;
  p.B = d->B; p.b_sk = d->b_sk; p.b_sn = d->b_sn; p.b_mode = d->b_mode;
  p.C = d->C; p.c_sm = d->c_sm; p.c_sn = d->c_sn; p.c_mode = d->c_mode;
  p.bias = d->bias; p.act_out = d->act_out; p.colstats = d->colstats;
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.g.sb = d->g_sb; p.g.sc = d->g_sc; p.g.sh = d->g_sh; p.g.sw = d->g_sw;
  p.g.Cin = d->g_cin; p.g.H = d->g_h; p.g.W = d->g_w; p.g.OH = d->g_oh; p.g.OW = d->g_ow; p.g.ab = d->g_ab;
  const bool small = ((int64_t)((d->M + 63) / 64) * ((d->N + 63) / 64)) < 64;
  const int bm = small ? 32 : 64;
  const int tiles = ((d->M + bm - 1) / bm) * ((d->N + bm - 1) / bm);
  int splits = d->splits;
  if (splits <= 0) {       // auto: fill the machine when the output grid alone does not (only legal with atomics)
    splits = 1;
    if (d->c_mode != 0 && tiles < kNumSMs) {
      splits = (2 * kNumSMs + tiles - 1) / tiles;
      const int maxs = (d->K + 63) / 64;
      if (splits > maxs) splits = maxs;
      if (splits < 1) splits = 1;
    }
  }
  CDAE_CHECK_ARG(splits == 1 || d->c_mode != 0, "sgemm: split-K needs an accumulating output mode");
  int kps = (d->K + splits - 1) / splits;
  kps = (kps + 15) / 16 * 16;
  splits = (d->K + kps - 1) / kps;
  p.k_per_split = kps;
  dim3 grid((d->N + bm - 1) / bm, (d->M + bm - 1) / bm, splits);
  if (small) sgemm_kernel<32, 32><<<grid, 256, 0, (cudaStream_t)s>>>(p);
  else sgemm_kernel<64, 64><<<grid, 256, 0, (cudaStream_t)s>>>(p);
  CDAE_CHECK_LAUNCH("sgemm_kernel");
  return CDAE_OK;
}

extern "C" int cdae_timestep_embedding(const void* t, int t_is_float, int t_stride, const int64_t* map, float scale,
                                       const float* freqs, float* out, int B, int dim, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(t && freqs && out && dim >= 2 && (t_stride == 0 || t_stride == 1), "timestep_embedding: bad arguments");
  const int n = B * dim;
  temb_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)s>>>(t, t_is_float, t_stride, map, scale, freqs, out, B, dim / 2, dim);
  CDAE_CHECK_LAUNCH("temb_kernel");
  return CDAE_OK;
}

extern "C" int cdae_silu_cast(const float* x, void* y_bf16, int64_t n, int silu, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(x && y_bf16, "silu_cast: null pointer");
  int64_t blocks = (n + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  silu_cast_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(x, reinterpret_cast<__nv_bfloat16*>(y_bf16), n, silu);
  CDAE_CHECK_LAUNCH("silu_cast_kernel");
  return CDAE_OK;
}

extern "C" int cdae_silu_bwd(float* g, const float* x, int64_t n, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(g && x, "silu_bwd: null pointer");
  int64_t blocks = (n + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  silu_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(g, x, n);
  CDAE_CHECK_LAUNCH("silu_bwd_kernel");
  return CDAE_OK;
}

extern "C" int cdae_softplus_bwd(float* g, const float* var, int64_t n, cdae_stream s) {
  if (n == 0) return CDAE_OK;
  CDAE_CHECK_ARG(g && var, "softplus_bwd: null pointer");
  int64_t blocks = (n + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  softplus_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(g, var, n);
  CDAE_CHECK_LAUNCH("softplus_bwd_kernel");
  return CDAE_OK;
}

extern "C" int cdae_embed_rows(float* x, const float* table, const int64_t* idx, int B, int D, int backward, float* dtable,
                               cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(x && idx && (backward ? dtable != nullptr : table != nullptr), "embed_rows: null pointer");
  embed_rows_kernel<<<(B * D + 255) / 256, 256, 0, (cudaStream_t)s>>>(x, table, idx, B, D, backward, dtable);
  CDAE_CHECK_LAUNCH("embed_rows_kernel");
  return CDAE_OK;
}

extern "C" int cdae_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, int64_t* num_batches_tracked, int train, float* ab, float* mean_rstd, int C,
                                cdae_stream s) {
  CDAE_CHECK_ARG(gamma && beta && running_mean && running_var && ab && mean_rstd && (!train || stats), "bn_finalize: null pointer");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)s>>>(stats, count, gamma, beta, running_mean, running_var,
                                                                   num_batches_tracked, train, ab, mean_rstd, C);
  CDAE_CHECK_LAUNCH("bn_finalize_kernel");
  return CDAE_OK;
}

extern "C" int cdae_enc_head(const float* in, const float* ab, float* out, int B, int P, int C, int backward, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(in && out && (backward || ab), "enc_head: null pointer");
  const int n = B * P * C;
  enc_head_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)s>>>(in, ab, out, B, P, C, backward);
  CDAE_CHECK_LAUNCH("enc_head_kernel");
  return CDAE_OK;
}

extern "C" int cdae_bn_lrelu_bwd(float* dact, const float* raw, const float* ab, const float* mean_rstd, const float* gamma,
                                 double* sums, float* dgamma, float* dbeta, int64_t M, int C, cdae_stream s) {
  if (M == 0) return CDAE_OK;
  CDAE_CHECK_ARG(dact && raw && ab && mean_rstd && gamma && sums && dgamma && dbeta, "bn_lrelu_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)s;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
  if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return CDAE_ERR_CUDA; }
  const int tx = C < 256 ? (C + 31) / 32 * 32 : 256;
  int chunks = (int)((M + 63) / 64); if (chunks > 512) chunks = 512; if (chunks < 1) chunks = 1;
  bn_bwd_reduce_kernel<<<dim3((C + tx - 1) / tx, chunks), tx, 0, st>>>(dact, raw, ab, mean_rstd, sums, M, C);
  CDAE_CHECK_LAUNCH("bn_bwd_reduce_kernel");
  int64_t blocks = (M * C + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  bn_bwd_apply_kernel<<<(int)blocks, 256, 0, st>>>(dact, raw, ab, mean_rstd, gamma, sums, dgamma, dbeta, M, C);
  CDAE_CHECK_LAUNCH("bn_bwd_apply_kernel");
  return CDAE_OK;
}

extern "C" int cdae_latent_fwd(const float* mu, const float* var, const float* zp, const float* xi, const float* keep,
                               const float* c, float* z, float* zp_out, float* kld, int B, int D, int n, int causal,
                               float var_scale, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(mu && var && zp && xi && z && n >= 1 && D % n == 0, "latent_fwd: bad arguments");
  latent_fwd_kernel<<<B, 128, 0, (cudaStream_t)s>>>(mu, var, zp, xi, keep, c, z, zp_out, kld, D, n, causal, var_scale);
  CDAE_CHECK_LAUNCH("latent_fwd_kernel");
  return CDAE_OK;
}

extern "C" int cdae_latent_bwd(const float* mu, const float* var, const float* zp, const float* xi, const float* keep,
                               const float* c, const float* dz, const float* dkld, const float* dzp_ext, const float* dmu_ext,
                               const float* dvar_ext, float* dzp, float* dmu, float* dvar, int B, int D, int n, int causal,
                               float var_scale, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(mu && var && zp && xi && dzp && dmu && dvar && n >= 1 && D % n == 0, "latent_bwd: bad arguments");
  latent_bwd_kernel<<<B, 128, 0, (cudaStream_t)s>>>(mu, var, zp, xi, keep, c, dz, dkld, dzp_ext, dmu_ext, dvar_ext, dzp, dmu,
                                                    dvar, D, n, causal, var_scale);
  CDAE_CHECK_LAUNCH("latent_bwd_kernel");
  return CDAE_OK;
}

extern "C" int cdae_step_loss(const float* mse, const float* kld, const float* keep, const float* w, const float* kl_weight,
                              const int64_t* t, int num_timesteps, int B, float* loss, float* gscale, float* dkld, float* total,
                              float* logsums, cdae_stream s) {
  if (B == 0) return CDAE_OK;
  CDAE_CHECK_ARG(mse && w && gscale && (!logsums || t), "step_loss: null pointer");
  step_loss_kernel<<<1, 256, 0, (cudaStream_t)s>>>(mse, kld, keep, w, kl_weight, t, num_timesteps, B, loss, gscale, dkld, total,
                                                  logsums);
  CDAE_CHECK_LAUNCH("step_loss_kernel");
  return CDAE_OK;
}
