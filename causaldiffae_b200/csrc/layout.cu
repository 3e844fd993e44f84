// Layout / packing / small reduction kernels (all HBM bound, 128-bit accesses where the layout allows).
#include "common.cuh"

namespace cdae {

// NCHW fp32 -> NHWC bf16 zero-padded to Cpad channels. One thread per pixel (reads coalesced per channel plane).
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C, int HW,
                                        int Cpad) {
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, pix = i % HW;
    __nv_bfloat16* o = out + i * Cpad;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
      float f[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = (c0 + k < C) ? x[(n * C + c0 + k) * HW + pix] : 0.f;
      st8(o + c0, pack8(f));
    }
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int N, int C, int HW, int ld) {
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, pix = i % HW;
    for (int c = 0; c < C; ++c) out[(n * C + c) * HW + pix] = __bfloat162float(x[i * ld + c]);
  }
}

// fp32 master [cout][taps][cin] -> bf16 [cout_pad][taps][cin_pad] (row pitch fwd_ld) and/or transposed bf16
// [cin_pad][taps][cout_pad] (row pitch tr_ld).  One 32x32 (co x ci) tile of one tap per block: the fp32 read and both
// bf16 writes are coalesced (the transposed one goes through a padded shared-memory tile).  entry._pad holds the first
// global tile index of the entry (prefix sum written by the host), found here by binary search.
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ arena, __nv_bfloat16* __restrict__ dst,
                                                           const cdae_pack_entry* __restrict__ entries, int n_entries) {
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (entries[mid]._pad <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const cdae_pack_entry e = entries[lo];
  const int local = blockIdx.x - e._pad;
  const int tci = (e.cin_pad + 31) / 32, tco = (e.cout_pad + 31) / 32;
  const int ci0 = (local % tci) * 32, co0 = ((local / tci) % tco) * 32, t = local / (tci * tco);
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int co = co0 + ty + r * 8, ci = ci0 + tx;
    float v = 0.f;
    if (co < e.cout && ci < e.cin) v = arena[e.src_off + ((int64_t)co * e.taps + t) * e.cin + ci];
    tile[ty + r * 8][tx] = v;
    if (e.dst_fwd_off >= 0 && co < e.cout_pad && ci < e.cin_pad) {
      const int64_t ld = e.fwd_ld > 0 ? e.fwd_ld : (int64_t)e.taps * e.cin_pad;
      dst[e.dst_fwd_off + (int64_t)co * ld + (int64_t)t * e.cin_pad + ci] = __float2bfloat16_rn(v);
    }
  }
  if (e.dst_tr_off >= 0) {
    __syncthreads();
    const int64_t ld = e.tr_ld > 0 ? e.tr_ld : (int64_t)e.taps * e.cout_pad;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ci = ci0 + ty + r * 8, co = co0 + tx;
      if (ci < e.cin_pad && co < e.cout_pad)
        dst[e.dst_tr_off + (int64_t)ci * ld + (int64_t)t * e.cout_pad + co] = __float2bfloat16_rn(tile[tx][ty + r * 8]);
    }
  }
}

__global__ void upsample2x_kernel(const bf16x8* __restrict__ x, bf16x8* __restrict__ out, int N, int H, int W, int CV) {
  const int64_t total = (int64_t)N * 2 * H * 2 * W * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ox = (int)(r % (2 * W)); r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int64_t n = r / (2 * H);
    out[i] = x[((n * H + (oy >> 1)) * W + (ox >> 1)) * CV + cv];
  }
}

__global__ void sumpool2x_kernel(const bf16x8* __restrict__ dy, bf16x8* __restrict__ dx, int N, int H, int W, int CV, int acc) {
  const int64_t total = (int64_t)N * H * W * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int64_t n = r / H;
    float a[8];
    if (acc) unpack8(dx[i], a); else { for (int k = 0; k < 8; ++k) a[k] = 0.f; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
      unpack8(dy[((n * 2 * H + 2 * y + (q >> 1)) * 2 * W + 2 * x + (q & 1)) * CV + cv], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += f[k];
    }
    dx[i] = pack8(a);
  }
}

__global__ void zero_insert2x_kernel(const bf16x8* __restrict__ x, bf16x8* __restrict__ out, int N, int H, int W, int CV) {
  const int64_t total = (int64_t)N * 2 * H * 2 * W * CV;
  bf16x8 z; { float f[8] = {0, 0, 0, 0, 0, 0, 0, 0}; z = pack8(f); }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    int64_t r = i / CV;
    const int ox = (int)(r % (2 * W)); r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int64_t n = r / (2 * H);
    out[i] = ((ox | oy) & 1) ? z : x[((n * H + (oy >> 1)) * W + (ox >> 1)) * CV + cv];
  }
}

// column sums: block = 8 vector lanes (64 channels) x 32 rows
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                     int64_t rows, int C, int ld, int out_ld) {
  const int vl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = blockIdx.x * 64 + vl * 8;
  x += (size_t)blockIdx.z * rows * ld;          // group g: its own [rows, C] matrix and its own output row
  out += (size_t)blockIdx.z * out_ld;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < C) {
    for (int64_t r = blockIdx.y * 32 + rl; r < rows; r += (int64_t)gridDim.y * 32) {
      float f[8];
      unpack8(ld8(x + r * ld + c), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += f[k];
    }
  }
  __shared__ float sm[32][65];
#pragma unroll
  for (int k = 0; k < 8; ++k) sm[rl][vl * 8 + k] = a[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int r = 0; r < 32; ++r) t += sm[r][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < C) atomicAdd(out + cc, t);
  }
}

static inline int grid_for(int64_t n) {
  int64_t b = ceil_div(n, 256), cap = (int64_t)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace cdae
using namespace cdae;

extern "C" int cdae_nchw_to_nhwc_pad(const float* x, void* out, int N, int C, int H, int W, int Cpad, cdae_stream s) {
  CDAE_CHECK_ARG(x && out, "nchw_to_nhwc_pad: null pointer");
  CDAE_CHECK_SHAPE(Cpad % 8 == 0 && Cpad >= C, "nchw_to_nhwc_pad: Cpad %d", Cpad);
  if (N == 0) return CDAE_OK;
  nchw_to_nhwc_pad_kernel<<<grid_for((int64_t)N * H * W), 256, 0, (cudaStream_t)s>>>(x, (__nv_bfloat16*)out, N, C, H * W, Cpad);
  CDAE_CHECK_LAUNCH("nchw_to_nhwc_pad_kernel");
  return CDAE_OK;
}

extern "C" int cdae_nhwc_to_nchw(const void* x, float* out, int N, int C, int H, int W, int ld, cdae_stream s) {
  CDAE_CHECK_ARG(x && out, "nhwc_to_nchw: null pointer");
  if (N == 0) return CDAE_OK;
  nhwc_to_nchw_kernel<<<grid_for((int64_t)N * H * W), 256, 0, (cudaStream_t)s>>>((const __nv_bfloat16*)x, out, N, C, H * W, ld);
  CDAE_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return CDAE_OK;
}

extern "C" int cdae_pack_weights(const float* arena, void* bf16_arena, const cdae_pack_entry* entries_dev, int n_entries,
                                 int64_t total_tiles, cdae_stream s) {
  CDAE_CHECK_ARG(arena && bf16_arena && entries_dev, "pack_weights: null pointer");
  if (n_entries == 0 || total_tiles == 0) return CDAE_OK;
  pack_weights_kernel<<<(unsigned)total_tiles, 256, 0, (cudaStream_t)s>>>(arena, (__nv_bfloat16*)bf16_arena, entries_dev, n_entries);
  CDAE_CHECK_LAUNCH("pack_weights_kernel");
  return CDAE_OK;
}

extern "C" int cdae_upsample2x(const void* x, void* out, int N, int H, int W, int C, cdae_stream s) {
  CDAE_CHECK_ARG(x && out, "upsample2x: null pointer");
  CDAE_CHECK_SHAPE(C % 8 == 0, "upsample2x: C %% 8");
  if (N == 0) return CDAE_OK;
  upsample2x_kernel<<<grid_for((int64_t)N * 4 * H * W * (C / 8)), 256, 0, (cudaStream_t)s>>>((const bf16x8*)x, (bf16x8*)out, N, H, W, C / 8);
  CDAE_CHECK_LAUNCH("upsample2x_kernel");
  return CDAE_OK;
}

extern "C" int cdae_sumpool2x(const void* dy, void* dx, int N, int H, int W, int C, int accumulate, cdae_stream s) {
  CDAE_CHECK_ARG(dy && dx, "sumpool2x: null pointer");
  CDAE_CHECK_SHAPE(C % 8 == 0, "sumpool2x: C %% 8");
  if (N == 0) return CDAE_OK;
  sumpool2x_kernel<<<grid_for((int64_t)N * H * W * (C / 8)), 256, 0, (cudaStream_t)s>>>((const bf16x8*)dy, (bf16x8*)dx, N, H, W, C / 8, accumulate);
  CDAE_CHECK_LAUNCH("sumpool2x_kernel");
  return CDAE_OK;
}

extern "C" int cdae_zero_insert2x(const void* x, void* out, int N, int H, int W, int C, cdae_stream s) {
  CDAE_CHECK_ARG(x && out, "zero_insert2x: null pointer");
  CDAE_CHECK_SHAPE(C % 8 == 0, "zero_insert2x: C %% 8");
  if (N == 0) return CDAE_OK;
  zero_insert2x_kernel<<<grid_for((int64_t)N * 4 * H * W * (C / 8)), 256, 0, (cudaStream_t)s>>>((const bf16x8*)x, (bf16x8*)out, N, H, W, C / 8);
  CDAE_CHECK_LAUNCH("zero_insert2x_kernel");
  return CDAE_OK;
}

extern "C" int cdae_colsum(const void* x, float* out, int64_t rows, int C, int ld, int groups, int out_ld, cdae_stream s) {
  CDAE_CHECK_ARG(x && out, "colsum: null pointer");
  CDAE_CHECK_SHAPE(C % 8 == 0 && ld % 8 == 0 && groups >= 1 && groups <= 65535, "colsum: C, ld %% 8; 1 <= groups <= 65535");
  if (rows == 0) return CDAE_OK;
  int gx = (C + 63) / 64;
  int64_t gy = ceil_div(rows, 32 * 16);
  int64_t cap = (int64_t)kNumSMs * 4 / ((int64_t)gx * groups); if (cap < 1) cap = 1;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  colsum_kernel<<<dim3(gx, (unsigned)gy, (unsigned)groups), 256, 0, (cudaStream_t)s>>>((const __nv_bfloat16*)x, out, rows, C, ld,
                                                                                        out_ld);
  CDAE_CHECK_LAUNCH("colsum_kernel");
  return CDAE_OK;
}
