// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA operand loads).
//
//   forward / data-gradient / plain GEMM :  igemm_kernel   (A = activations, K-major via 4-D TMA boxes with zero fill
//                                                            doing the conv padding; B = packed weights, K-major)
//   weight gradient                      :  wgrad_kernel   (A = dY^T, B = X^T: both MN-major straight from NHWC)
//
// Replaces aten::convolution / convolution_backward (cuDNN) and addmm at the call sites listed in include/cdae.h.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue
// (TMEM -> registers -> swizzled smem slab -> TMA store), one TMEM lane quarter each; igemm adds a 7th warp that
// prefetches residual slabs.
#include <cstdlib>
#include <mutex>

#include "sm100.cuh"

namespace cdae {
using namespace sm100;

// ------------------------------------------------------------------------------------------------ host: tensor maps
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides, bool swizzle128) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDAE_ERR_CUDA; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r, rank,
              (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
              (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0],
              rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return CDAE_ERR_CUDA;
  }
  return CDAE_OK;
}

static inline int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// pixel-tile geometry: BW*BH*BNI == pixels, every factor a power of two
static void tile_geometry(int pixels, int OHt, int OWt, int* BW, int* BH, int* BNI) {
  int bw = pow2_ceil(OWt); if (bw > pixels) bw = pixels;
  int bh = pow2_ceil(OHt); if (bh > pixels / bw) bh = pixels / bw;
  *BW = bw; *BH = bh; *BNI = pixels / (bw * bh);
}

// ------------------------------------------------------------------------------------------------ forward kernel
struct alignas(64) IgemmKParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmO;   // output tile store  (out_mode 0): box {slab channels, BW, BH, BNI}
  CUtensorMap tmR;   // residual tile load (same box)
  cdae_seg seg[CDAE_MAX_SEG];
  int nseg, nkb;
  int BW, BH, BNI, tilesW, tilesH;
  int in_stride;
  int Nimg, OHt, OWt;
  int OH, OW, cout, out_mode, has_resid, ldo;
  void* out;
  const float* bias;
  const float* bias2;
  const float* bias_img; int bias_img_ld;   // optional per-(image, channel) additive term: + bias_img[n * ld + co]
  float* stats;      // optional fp32 [Nimg][cout][2]: per-(image, channel) sum / sum of squares of the stored bf16 output
  // GroupNorm-backward fusion of a data-gradient launch (cdae_igemm_desc.gnb_*): the output dy is the gradient w.r.t. the
  // OUTPUT of a GroupNorm(+FiLM)(+SiLU) whose input x = concat(gnb_x0, gnb_x1) has the same pixel grid.  The epilogue
  // fetches the matching x slab over the residual TMA path, turns dy into du = dy * silu'(u) (u/2 = a x + b from the forward
  // pass's per-(image, channel) constant table), stores du, and accumulates sum du / sum du*x per (image, channel).
  CUtensorMap tmR2;  // x slab of the second source
  float* gnb_ws;     // fp32 [Nimg][cout][2] += {sum du, sum du*x}
  const float* gnb_ab;
  const __nv_bfloat16* gnb_x0; const __nv_bfloat16* gnb_x1;
  int gnb_c0, gnb_ld0, gnb_ld1, gnb_silu, lgBW, lgBH;
  int nboxes, ntn;   // number of 128-pixel boxes and of N tiles
  // halo kernel (3x3 stride 1): K is walked source-chunk-major; each 64-channel chunk brings ONE halo tile per box
  struct { int src, c0, nchunk, ntap, wk0; } hs[8];
  int nhs, tap_stride, flip;
  // GroupNorm+SiLU on load (cdae_igemm_desc.gn_*): per halo segment the first table column of its channels, or -1
  const float* gn_ab; int gn_c; int gn_col[8];
};

constexpr int kATileBytes = 128 * 128;  // 128 pixels x 64 bf16

// Barrier addresses and buffers the epilogue roles need (identical in the tap-streaming and the halo kernel).
struct EpiCtx {
  uint32_t tmem_base, stg_base;
  uint32_t tfull0, tempty0, sready0, sfree0;   // shared addresses of the first barrier of each kind (8 B apart)
  int total_tiles, boxes_per_img;
  uint32_t abs_base;                           // [NS][ABI] x 512 B: {a, b} of the 64 slab channels per image of the box
  int abi;
};

__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// warps 2..5: tcgen05.ld -> +bias (+bias2) (+residual) -> bf16 -> 128B-swizzled smem slab -> TMA store.  Nothing here
// waits on global memory: stores are asynchronous bulk copies, the residual slab is prefetched by the staging warp
// into the very buffer the result is written to (added in place).
template <int BN, int MT, int NS>
__device__ __forceinline__ void igemm_epilogue(const IgemmKParams& p, const EpiCtx& cx, int warp, int lane) {
  constexpr int SLABW = BN < 64 ? BN : 64;                    // channels per output slab
  constexpr int kSlabStride = 128 * 128;                      // staging buffers are 16 KB apart (1024 B aligned)
  constexpr uint32_t kAccCols = MT * BN;
  static_assert(NS >= 3, "the statistics pass reads slab i while slab i+1 is produced: buffer i must not be recycled before the next named barrier");
  const uint32_t tmem_base = cx.tmem_base, stg_base = cx.stg_base;
  const int total_tiles = cx.total_tiles, boxes_per_img = cx.boxes_per_img;
  auto tfull_bar = [&](int a) { return cx.tfull0 + 8u * a; };
  auto tempty_bar = [&](int a) { return cx.tempty0 + 8u * a; };
  auto sready_bar = [&](int b) { return cx.sready0 + 8u * b; };
  auto sfree_bar = [&](int b) { return cx.sfree0 + 8u * b; };
  // ---------------------------------------------------------------- epilogue (warps 2..5; TMEM lane quarter = warp % 4)
  const int q = warp & 3;
  const int r = q * 32 + lane;                     // pixel row inside a 128-pixel box
  const bool elected = (threadIdx.x == 64);
  const float* __restrict__ bias = p.bias;
  const float* __restrict__ bias2 = p.bias2;
  int it = 0, sidx = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int as = it & 1;
    const uint32_t aph = (it >> 1) & 1;
    const int tm = tile / p.ntn, n0 = (tile % p.ntn) * BN;
    mbar_wait(tfull_bar(as), aph);
    tc_fence_after();
    if (p.out_mode == 0) {
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int box = tm * MT + m;
        if (box >= p.nboxes) break;
        const int w0 = (box % p.tilesW) * p.BW, h0 = ((box / p.tilesW) % p.tilesH) * p.BH;
        const int nn0 = (box / boxes_per_img) * p.BNI;
        // GroupNorm statistics of the consumer (gn_apply_fwd_kernel) ride along: which of this warp's 32 pixel rows
        // are real output pixels, and the image they belong to (BW*BH >= 32, checked on the host)
        uint32_t vmask = 0;
        int nimg = 0;
        const bool gnb = SLABW == 64 && p.gnb_ws != nullptr;
        if (SLABW == 64 && (p.stats != nullptr || gnb)) {
          const int pb = p.BW * p.BH;
          const bool row_ok = (nn0 + r / pb < p.Nimg) && (h0 + (r / p.BW) % p.BH < p.OHt) && (w0 + r % p.BW < p.OWt);
          vmask = __ballot_sync(0xffffffffu, row_ok);
          nimg = nn0 + (q * 32) / pb;
        }
#pragma unroll 1
        for (int c = 0; c < BN; c += SLABW) {
          const int co0 = n0 + c;
          if (co0 >= p.cout) break;
          const int buf = sidx % NS;
          const uint32_t sph = (sidx / NS) & 1;
          uint32_t acc[SLABW];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccCols + m * BN + c);
          __syncwarp();
          if (SLABW >= 32) {
#pragma unroll
            for (int h = 0; h < SLABW / 32; ++h) tmem_ld32(taddr + 32 * h, acc + 32 * h);
          } else {
            tmem_ld16(taddr, acc);
          }
          tmem_ld_wait();
          mbar_wait(sready_bar(buf), sph);           // staging buffer drained (and residual slab landed)
          const uint32_t row = stg_base + buf * kSlabStride + r * (SLABW * 2);
          // GroupNorm-backward fusion: forward constants of this warp's image, staged by the manager warp
          const uint32_t abrow = cx.abs_base + (uint32_t)(buf * cx.abi + (q * 32) / (p.BW * p.BH)) * 512u;
          const bool row_valid = (vmask >> lane) & 1u;
          float gsum[SLABW / 8 > 0 ? SLABW / 8 : 1];
#pragma unroll
          for (int j = 0; j < SLABW / 8; ++j) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc[j * 8 + e]);
            if (co0 + j * 8 < p.cout) {
              if (bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + co0 + j * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + co0 + j * 8 + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (bias2) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias2 + co0 + j * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias2 + co0 + j * 8 + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (p.bias_img) {        // additive timestep conditioning (use_scale_shift_norm=False): + emb_out[image, channel]
                const int nrow = min(nn0 + r / (p.BW * p.BH), p.Nimg - 1);
                const float* bi = p.bias_img + (size_t)nrow * p.bias_img_ld + co0 + j * 8;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bi));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bi + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
            }
            // 16 B chunk j of row r sits at chunk (j ^ (r & 7)) under the 128B swizzle (SLABW == 64); narrower slabs
            // are stored unswizzled (their tensor maps use SWIZZLE_NONE)
            const uint32_t a = SLABW == 64 ? row + (uint32_t)((j ^ (r & 7)) << 4) : row + (uint32_t)(j << 4);
            if (gnb) {
              float rv[8];
              unpack8(lds8(a), rv);                      // x of this pixel row (the slab the manager warp fetched)
              if (p.gnb_silu) {                          // du = dy * silu'(u), u/2 = a x + b
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float4 c4 = lds128f(abrow + (uint32_t)(j * 8 + 2 * e) * 8u);      // {a0, b0, a1, b1}
                  const float2 h = make_float2(fmaf(rv[2 * e], c4.x, c4.y), fmaf(rv[2 * e + 1], c4.z, c4.w));
                  const float2 d = dsilu_half(h);
                  v[2 * e] *= d.x; v[2 * e + 1] *= d.y;
                }
              }
              // column sums over this warp's 32 pixel rows of {du, du*x} (du as stored: rounded to bf16), entirely in
              // registers: a halving butterfly - every xor step trades half of the values for the partner's other half -
              // leaves lane l with ONE fully reduced value: P (bit 4 clear) or Q (bit 4 set) of channel j*8 + ((l >> 1) & 7)
              float val[16];
              {
                const bf16x8 pk = pack8(v);
                unpack8(pk, val);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  if (!row_valid) val[e] = 0.f;
                  val[8 + e] = val[e] * rv[e];
                }
              }
              const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
              float k8[8], k4[4], k2[2];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                k8[i] = (h4 ? val[i + 8] : val[i]) + __shfl_xor_sync(0xffffffffu, h4 ? val[i] : val[i + 8], 16);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                k4[i] = (h3 ? k8[i + 4] : k8[i]) + __shfl_xor_sync(0xffffffffu, h3 ? k8[i] : k8[i + 4], 8);
#pragma unroll
              for (int i = 0; i < 2; ++i)
                k2[i] = (h2 ? k4[i + 2] : k4[i]) + __shfl_xor_sync(0xffffffffu, h2 ? k4[i] : k4[i + 2], 4);
              float k1 = (h1 ? k2[1] : k2[0]) + __shfl_xor_sync(0xffffffffu, h1 ? k2[0] : k2[1], 2);
              k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
              gsum[j] = k1;
            } else if (p.has_resid) {
              float rv[8];
              unpack8(lds8(a), rv);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += rv[e];
            }
            sts8(a, pack8(v));
          }
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (elected) {
            tma_store_4d(&p.tmO, stg_base + buf * kSlabStride, co0, w0, h0, nn0);
            bulk_commit();
            bulk_wait_read<NS - 2>();                 // every store but the newest NS-2 has finished reading smem
            if (sidx >= NS - 2) mbar_arrive(sfree_bar((sidx - (NS - 2)) % NS));
          }
          if (gnb) {
            // lanes with bit 0 clear own the pair's result: 8 channels (one per 8-channel chunk) of P or Q
            if (vmask != 0 && !(lane & 1)) {
              float* dst = p.gnb_ws + ((size_t)nimg * p.cout + co0 + ((lane >> 1) & 7)) * 2 + (lane >> 4);
#pragma unroll
              for (int j = 0; j < SLABW / 8; ++j) atomicAdd(dst + j * 16, gsum[j]);
            }
          } else if (SLABW == 64 && vmask != 0) {
            // column sums of the slab as stored (bf16): lane = channel pair, this warp's 32 rows; one 128-bit red
            // per lane into stats[nimg][co0 + 2*lane .. +1][sum, sumsq].  The buffer is recycled no earlier than the
            // named barrier of a later slab, i.e. after every thread has left this loop.
            const uint32_t sb = stg_base + buf * kSlabStride + (uint32_t)(q * 32) * 128u + (uint32_t)((lane & 3) << 2);
            const uint32_t cj = (uint32_t)(lane >> 2);
            float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
              uint32_t u = lds32(sb + (uint32_t)i * 128u + ((cj ^ (uint32_t)(i & 7)) << 4));
              if (!((vmask >> i) & 1u)) u = 0u;
              const float2 f = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
              s2 = __fadd2_rn(s2, f);
              q2 = __ffma2_rn(f, f, q2);
            }
            float* dst = p.stats + ((size_t)nimg * p.cout + co0 + 2 * lane) * 2;
            atomicAdd(reinterpret_cast<float4*>(dst), make_float4(s2.x, q2.x, s2.y, q2.y));
          }
          ++sidx;
        }
      }
    } else {
      // NCHW fp32 (final eps conv, a handful of channels): consecutive lanes are consecutive pixels -> coalesced
      const int bw = r % p.BW, bh = (r / p.BW) % p.BH, bn = r / (p.BW * p.BH);
      constexpr int CH = BN < 32 ? 16 : 32;
      float* __restrict__ o = reinterpret_cast<float*>(p.out);
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int box = tm * MT + m;
        const int n = (box / boxes_per_img) * p.BNI + bn;
        const int oy = ((box / p.tilesW) % p.tilesH) * p.BH + bh, ox = (box % p.tilesW) * p.BW + bw;
        const bool row_ok = (n < p.Nimg) && (oy < p.OHt) && (ox < p.OWt);
#pragma unroll 1
        for (int c = 0; c < BN; c += CH) {
          uint32_t acc[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccCols + m * BN + c);
          __syncwarp();
          if (CH == 32) tmem_ld32(taddr, acc); else tmem_ld16(taddr, acc);
          tmem_ld_wait();
          const int co0 = n0 + c;
          if (row_ok && co0 < p.cout) {
            if (p.out_mode == 2) {
              // fp32 row-major [pixel][ldo]: GEMM-shaped outputs that stay fp32 (the FiLM vectors, their gradient)
              float* orow = o + (((size_t)n * p.OH + oy) * p.OW + ox) * p.ldo + co0;
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                if (co0 + j + 3 < p.cout) {
                  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + co0 + j));
                  *reinterpret_cast<float4*>(orow + j) = make_float4(__uint_as_float(acc[j]) + b4.x, __uint_as_float(acc[j + 1]) + b4.y,
                                                                     __uint_as_float(acc[j + 2]) + b4.z, __uint_as_float(acc[j + 3]) + b4.w);
                } else {
                  for (int e = 0; e < 4 && co0 + j + e < p.cout; ++e)
                    orow[j + e] = __uint_as_float(acc[j + e]) + (bias ? __ldg(bias + co0 + j + e) : 0.f);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                const int co = co0 + j;
                if (co < p.cout)
                  o[(((size_t)n * p.cout + co) * p.OH + oy) * p.OW + ox] = __uint_as_float(acc[j]) + (bias ? __ldg(bias + co) : 0.f);
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty_bar(as));
  }
  if (elected) bulk_wait_all();                      // smem must stay valid until the last store has read it
}

// staging-buffer manager (one thread): waits until a slab buffer has been drained by its TMA store, then either
// TMA-loads the residual slab into it or just marks it ready.
template <int BN, int MT, int NS>
__device__ __forceinline__ void igemm_stage_manager(const IgemmKParams& p, const EpiCtx& cx, int lane) {
  constexpr int SLABW = BN < 64 ? BN : 64;
  constexpr int kSlabBytes = 128 * SLABW * 2;
  constexpr int kSlabStride = 128 * 128;
  const uint32_t stg_base = cx.stg_base;
  const int total_tiles = cx.total_tiles, boxes_per_img = cx.boxes_per_img;
  auto sready_bar = [&](int b) { return cx.sready0 + 8u * b; };
  auto sfree_bar = [&](int b) { return cx.sfree0 + 8u * b; };
  if (p.out_mode != 0) return;
  const bool gnb = SLABW == 64 && p.gnb_ws != nullptr;
  if (!gnb && lane != 0) return;           // only the GroupNorm-backward fusion has work for the other 31 lanes
  int sidx = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tm = tile / p.ntn, n0 = (tile % p.ntn) * BN;
    for (int m = 0; m < MT; ++m) {
      const int box = tm * MT + m;
      if (box >= p.nboxes) break;
      const int w0 = (box % p.tilesW) * p.BW, h0 = ((box / p.tilesW) % p.tilesH) * p.BH;
      const int nn0 = (box / boxes_per_img) * p.BNI;
      for (int c = 0; c < BN; c += SLABW) {
        const int co0 = n0 + c;
        if (co0 >= p.cout) break;
        const int buf = sidx % NS;
        const uint32_t sph = (sidx / NS) & 1;
        mbar_wait(sfree_bar(buf), sph ^ 1);
        if (gnb) {
          // forward constants {a, b} of the slab's 64 channels, for every image of the box: 512 B per image, one 128-bit
          // load per lane; visible to the epilogue through the release of lane 0's arrive below
          if (p.gnb_silu) {
            for (int bi = 0; bi < p.BNI && bi < cx.abi; ++bi) {
              if (nn0 + bi < p.Nimg) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.gnb_ab + ((size_t)(nn0 + bi) * p.cout + co0) * 2) + lane);
                sts128f(cx.abs_base + (uint32_t)(buf * cx.abi + bi) * 512u + (uint32_t)lane * 16u, v);
              }
            }
          }
          __syncwarp();
          if (lane == 0) {
            const bool in1 = co0 >= p.gnb_c0;
            mbar_expect_tx(sready_bar(buf), kSlabBytes);
            tma_load_4d(stg_base + buf * kSlabStride, in1 ? &p.tmR2 : &p.tmR, sready_bar(buf), in1 ? co0 - p.gnb_c0 : co0, w0, h0,
                        nn0);
          }
        } else if (p.has_resid) {
          mbar_expect_tx(sready_bar(buf), kSlabBytes);
          tma_load_4d(stg_base + buf * kSlabStride, &p.tmR, sready_bar(buf), co0, w0, h0, nn0);
        } else {
          mbar_arrive(sready_bar(buf));
        }
        ++sidx;
      }
    }
  }
}

// One CTA per SM loops over output tiles of MT x 128 pixels by BN channels (persistent).  Warp roles (224 threads):
//   warp 0    TMA producer: A (4-D NHWC boxes, zero fill = conv padding) and B (weights) into a STAGES-deep smem ring
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer; TWO accumulator sets in TMEM, so the epilogue of
//             tile i overlaps the main loop of tile i+1 (mbarrier pairs tfull / tempty)
//   warps 2-5 epilogue: tcgen05.ld -> +bias (+bias2) (+residual) -> bf16 -> 128B-swizzled smem slab -> TMA store.
//             Nothing in the epilogue waits on global memory: stores are asynchronous bulk copies, the residual
//             slab is prefetched by warp 6 into the very staging buffer the result is written to (added in place).
//   warp 6    staging-buffer manager: waits until a slab buffer has been drained by its TMA store, then either
//             TMA-loads the residual slab into it or just marks it ready.
template <int BN, int MT, int STAGES, int NS>
__global__ void __launch_bounds__(224, 1) igemm2_kernel(const __grid_constant__ IgemmKParams p) {
  constexpr int kBTileBytes = BN * 128;
  constexpr int kStageBytes = MT * kATileBytes + kBTileBytes;
  constexpr int SLABW = BN < 64 ? BN : 64;                    // channels per output slab
  constexpr int kSlabBytes = 128 * SLABW * 2;
  constexpr int kSlabStride = 128 * 128;                      // staging buffers are 16 KB apart (1024 B aligned)
  constexpr uint32_t kAccCols = MT * BN;                      // one accumulator set
  constexpr uint32_t kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64 : 2 * kAccCols <= 128 ? 128
                                 : 2 * kAccCols <= 256 ? 256 : 512;
  static_assert(2 * kAccCols <= 512, "accumulators exceed TMEM");
  static_assert(NS >= 2, "need at least two staging buffers");
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + STAGES * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + NS * kSlabStride);
  // bars: [0,S) full | [S,2S) empty | [2S,2S+2) tmem_full | [2S+2,2S+4) tmem_empty | NS sready | NS sfree
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + 2 * NS);
  constexpr int kABI = 4;                                     // images per 128-pixel box (>= 32 pixels each)
  const uint32_t abs_base = (smem_u32(tmem_slot) + 16 + 15) & ~15u;      // [NS][kABI] x 512 B, GroupNorm-backward fusion
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t stg_base = smem_u32(stg);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto sready_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 4 + b); };
  auto sfree_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 4 + NS + b); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    for (int b = 0; b < NS; ++b) { mbar_init(sready_bar(b), 1); mbar_init(sfree_bar(b), 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m = (p.nboxes + MT - 1) / MT;
  const int total_tiles = tiles_m * p.ntn;
  const int boxes_per_img = p.tilesW * p.tilesH;
  const EpiCtx cx{tmem_base, stg_base, tfull_bar(0), tempty_bar(0), sready_bar(0), sfree_bar(0), total_tiles, boxes_per_img,
                  abs_base, kABI};

  if (warp == 0) {
    // producer: the whole warp walks the loop, one elected lane issues the TMA loads (see elect_one)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tm = tile / p.ntn, n0 = (tile % p.ntn) * BN;
      int cw[MT], chh[MT], cn[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int box = tm * MT + m;
        cw[m] = (box % p.tilesW) * p.BW * p.in_stride;
        chh[m] = ((box / p.tilesW) % p.tilesH) * p.BH * p.in_stride;
        cn[m] = (box / boxes_per_img) * p.BNI;          // boxes past the end land beyond N: TMA zero-fills them
      }
      for (int sg = 0; sg < p.nseg; ++sg) {
        const cdae_seg g = p.seg[sg];
        const CUtensorMap* tma = &p.tmA[g.src];
        for (int ch = 0; ch < g.nchunk; ++ch) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t a_dst = smem_base + s * kStageBytes;
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), kStageBytes);
#pragma unroll
            for (int m = 0; m < MT; ++m)
              tma_load_4d(a_dst + m * kATileBytes, tma, full_bar(s), g.c0 + ch * 64, cw[m] + g.dw, chh[m] + g.dh, cn[m]);
            tma_load_2d(a_dst + MT * kATileBytes, &p.tmB, full_bar(s), g.wk + ch * 64, n0);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loop on warp-uniform values; one elected lane issues (see elect_one)
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(tempty_bar(as), aph ^ 1);          // epilogue has drained this accumulator set
      tc_fence_after();
      for (int kbl = 0; kbl < p.nkb; ++kbl) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * kStageBytes;
        const uint64_t bdesc = smem_desc_kmajor_sw128(a_addr + MT * kATileBytes);
        uint64_t adesc[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) adesc[m] = smem_desc_kmajor_sw128(a_addr + m * kATileBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              umma_f16(tmem_base + (uint32_t)(as * kAccCols + m * BN), adesc[m] + 2 * k, bdesc + 2 * k, kIdesc, (kbl | k) != 0);
          }
          umma_commit(empty_bar(s));
          if (kbl == p.nkb - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < 6) {
    igemm_epilogue<BN, MT, NS>(p, cx, warp, lane);
  } else {
    igemm_stage_manager<BN, MT, NS>(p, cx, lane);
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

template <int BN, int MT, int STAGES, int NS>
static int launch_igemm2(const IgemmKParams& kp, cudaStream_t st) {
  constexpr int smem = STAGES * (MT * kATileBytes + BN * 128) + NS * 128 * 128 + 1024 + 256 + NS * 4 * 512 + 32;
  static_assert(smem <= 227 * 1024, "igemm2: shared memory budget");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(igemm2_kernel<BN, MT, STAGES, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (attr_err != cudaSuccess) { set_error("igemm2 smem attribute: %s", cudaGetErrorString(attr_err)); return CDAE_ERR_CUDA; }
  const int total = ((kp.nboxes + MT - 1) / MT) * kp.ntn;
  const int grid = total < kNumSMs ? total : kNumSMs;
  igemm2_kernel<BN, MT, STAGES, NS><<<grid, 224, smem, st>>>(kp);
  CDAE_CHECK_LAUNCH("igemm2_kernel");
  return CDAE_OK;
}

// ------------------------------------------------------------------------------------------------ halo kernel (3x3, stride 1)
// Same tile loop, TMEM double buffering and epilogue as igemm2_kernel, but the nine filter taps of a 64-channel chunk
// read ONE halo tile per pixel box: box = 8 wide x 16 high output pixels, halo tile = 10 x 18 input pixels x 64 channels
// (180 rows of 128 B, 128B-swizzled by TMA; out-of-image rows are zero filled = the conv padding).  Tap (dh, dw) is the
// K-major operand whose descriptor starts (1+dh)*10 + (1+dw) rows into the tile with 8-row groups 1280 B apart - the
// swizzle is a pure function of the shared-memory address, so row-shifted windows need no re-layout (tools/exp_desc.cu).
// Shared-memory fill traffic for A falls 6.4x (180 instead of 9 x 128 rows per chunk), which is what bounds the
// tap-streaming kernel (operand reads + TMA fills > 128 B/clk/SM).  A and B tiles run in separate rings with their own
// producer warps: an A stage lives for nine B stages.
//   warp 0 A producer | warp 1 MMA | warps 2-5 epilogue | warp 6 staging manager | warp 7 B producer
constexpr int kGnWarps = 8;       // transform warps of the GroupNorm-on-load variants (cdae_igemm_desc.gn_*)
constexpr int kHaloRows = 180;
constexpr int kHaloBytes = kHaloRows * 128;            // 23040
constexpr int kHaloStride = 23 * 1024;                 // tiles 1024 B aligned

__device__ __forceinline__ uint64_t smem_desc_kmajor_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// GN: kGnWarps more warps normalise + activate every halo tile in place between the TMA load and the MMAs (see igemm3t_kernel)
template <int BN, int MT, int AST, int BST, int NS, bool GN = false>
__global__ void __launch_bounds__(GN ? 256 + 32 * kGnWarps : 256, 1) igemm3_kernel(const __grid_constant__ IgemmKParams p) {
  constexpr int kBTileBytes = BN * 128;
  constexpr int kAStage = MT * kHaloStride;
  constexpr int kSlabStride = 128 * 128;
  constexpr uint32_t kAccCols = MT * BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64 : 2 * kAccCols <= 128 ? 128
                                 : 2 * kAccCols <= 256 ? 256 : 512;
  static_assert(2 * kAccCols <= 512, "accumulators exceed TMEM");
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bsm = smem + AST * kAStage;
  uint8_t* stg = bsm + BST * kBTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + NS * kSlabStride);
  // bars: afull[AST] aempty[AST] bfull[BST] bempty[BST] tfull[2] tempty[2] sready[NS] sfree[NS] aready[AST]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * AST + 2 * BST + 4 + 2 * NS);
  constexpr int kABI = 1;                                     // the halo box is 8 x 16 pixels of ONE image
  const uint32_t abs_base = (smem_u32(tmem_slot) + 16 + 15) & ~15u;      // [NS][1] x 512 B, GroupNorm-backward fusion
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(bsm), stg_base = smem_u32(stg);
  const uint32_t bar_base = smem_u32(bars);
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (AST + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * AST + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * AST + BST + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * AST + 2 * BST + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * AST + 2 * BST + 2 + a); };
  auto sready_bar = [&](int b) { return bar_base + 8u * (2 * AST + 2 * BST + 4 + b); };
  auto sfree_bar = [&](int b) { return bar_base + 8u * (2 * AST + 2 * BST + 4 + NS + b); };
  auto aready = [&](int s) { return bar_base + 8u * (2 * AST + 2 * BST + 4 + 2 * NS + s); };   // GN: tile transformed

  if (threadIdx.x == 0) {
    for (int s = 0; s < AST; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); mbar_init(aready(s), kGnWarps); }
    for (int s = 0; s < BST; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    for (int b = 0; b < NS; ++b) { mbar_init(sready_bar(b), 1); mbar_init(sfree_bar(b), 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m = (p.nboxes + MT - 1) / MT;
  const int total_tiles = tiles_m * p.ntn;
  const int boxes_per_img = p.tilesW * p.tilesH;
  const EpiCtx cx{tmem_base, stg_base, tfull_bar(0), tempty_bar(0), sready_bar(0), sfree_bar(0), total_tiles, boxes_per_img,
                  abs_base, kABI};

  if (warp == 0) {
    // halo producer: the whole warp walks the loop, one elected lane issues the TMA loads (see elect_one)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tm = tile / p.ntn;
      int cw[MT], chh[MT], cn[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int box = tm * MT + m;
        cw[m] = (box % p.tilesW) * 8 - 1;
        chh[m] = ((box / p.tilesW) % p.tilesH) * 16 - 1;
        cn[m] = box / boxes_per_img;                    // boxes past the end land beyond N: TMA zero-fills them
      }
      for (int h = 0; h < p.nhs; ++h) {
        const CUtensorMap* tma = &p.tmA[p.hs[h].src];
        const int c0 = p.hs[h].c0, nch = p.hs[h].nchunk;
        for (int j = 0; j < nch; ++j) {
          mbar_wait(aempty(s), ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(afull(s), MT * kHaloBytes);
#pragma unroll
            for (int m = 0; m < MT; ++m)
              tma_load_4d(a_base + s * kAStage + m * kHaloStride, tma, afull(s), c0 + j * 64, cw[m], chh[m], cn[m]);
          }
          __syncwarp();
          if (++s == AST) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 7) {
    // weight producer
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n0 = (tile % p.ntn) * BN;
      for (int h = 0; h < p.nhs; ++h) {
        const int nch = p.hs[h].nchunk, ntap = p.hs[h].ntap, wk0 = p.hs[h].wk0;
        for (int j = 0; j < nch; ++j) {
          for (int t = 0; t < ntap; ++t) {
            mbar_wait(bempty(s), ph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(bfull(s), kBTileBytes);
              tma_load_2d(b_base + s * kBTileBytes, &p.tmB, bfull(s), wk0 + t * p.tap_stride + j * 64, n0);
            }
            __syncwarp();
            if (++s == BST) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loop on warp-uniform values; one elected lane issues (see elect_one)
    int sa = 0, sb = 0, it = 0;
    uint32_t pha = 0, phb = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(tempty_bar(as), aph ^ 1);          // epilogue has drained this accumulator set
      tc_fence_after();
      uint32_t first = 1;
      for (int h = 0; h < p.nhs; ++h) {
        const int nch = p.hs[h].nchunk, ntap = p.hs[h].ntap;
        for (int j = 0; j < nch; ++j) {
          mbar_wait(GN ? aready(sa) : afull(sa), pha);
          for (int t = 0; t < ntap; ++t) {
            // tap index -> window origin inside the halo tile (flip: data-gradient taps are negated)
            const int ti = ntap == 9 ? (p.flip ? 8 - t : t) : 4;
            const uint32_t row0 = (uint32_t)((ti / 3) * 10 + (ti % 3));
            mbar_wait(bfull(sb), phb);
            tc_fence_after();
            const uint64_t bdesc = smem_desc_kmajor_sw128(b_base + sb * kBTileBytes);
            uint64_t adesc[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m)
              adesc[m] = smem_desc_kmajor_sw128_sbo(a_base + sa * kAStage + m * kHaloStride + row0 * 128, 1280);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int m = 0; m < MT; ++m)
                  umma_f16(tmem_base + (uint32_t)(as * kAccCols + m * BN), adesc[m] + 2 * k, bdesc + 2 * k, kIdesc,
                           (first && k == 0) ? 0u : 1u);
              }
              umma_commit(bempty(sb));
              if (t == ntap - 1) umma_commit(aempty(sa));
            }
            first = 0;
            __syncwarp();
            if (++sb == BST) { sb = 0; phb ^= 1; }
          }
          if (++sa == AST) { sa = 0; pha ^= 1; }
        }
      }
      if (elect_one()) umma_commit(tfull_bar(as));
      __syncwarp();
    }
  } else if (warp < 6) {
    igemm_epilogue<BN, MT, NS>(p, cx, warp, lane);
  } else if (warp == 6) {
    igemm_stage_manager<BN, MT, NS>(p, cx, lane);
  } else if (GN && warp >= 8) {
    // ---------------------------------------------------------------- GroupNorm + SiLU on the halo tiles, in place
    const int tt = (int)threadIdx.x - 256;
    const uint32_t ck = (uint32_t)(tt & 7);                  // logical 16-byte chunk: channels ck*8 .. +7 of the 64
    const int r0 = tt >> 3;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tm = tile / p.ntn;
      int cw[MT], chh[MT], cn[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int box = tm * MT + m;
        cw[m] = (box % p.tilesW) * 8 - 1;
        chh[m] = ((box / p.tilesW) % p.tilesH) * 16 - 1;
        cn[m] = box / boxes_per_img;
      }
      for (int h = 0; h < p.nhs; ++h) {
        const int nch = p.hs[h].nchunk, col0 = p.gn_col[h];
        for (int j = 0; j < nch; ++j) {
          float2 ab[MT][8];
          if (col0 >= 0) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              if (cn[m] >= p.Nimg) continue;
              const float4* q = reinterpret_cast<const float4*>(p.gn_ab + ((size_t)cn[m] * p.gn_c + col0 + j * 64 + ck * 8) * 2);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 v = __ldg(q + e);
                ab[m][2 * e] = make_float2(v.x, v.y); ab[m][2 * e + 1] = make_float2(v.z, v.w);
              }
            }
          }
          mbar_wait(afull(s), ph);
          if (col0 >= 0) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              if (cn[m] >= p.Nimg) continue;                 // a box past the end: all zeros, stays so
              const uint32_t tb = a_base + s * kAStage + m * kHaloStride;
              for (int r = r0; r < kHaloRows; r += 4 * kGnWarps) {
                const int hh = r / 10, ww = r - hh * 10;
                if ((unsigned)(chh[m] + hh) >= (unsigned)p.OHt || (unsigned)(cw[m] + ww) >= (unsigned)p.OWt) continue;
                const uint32_t a = tb + (uint32_t)r * 128u + ((ck ^ (uint32_t)(r & 7)) << 4);
                float f[8];
                unpack8(lds8(a), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float hv = fmaf(f[e], ab[m][e].x, ab[m][e].y);
                  f[e] = fmaf(hv, tanh_fast(hv), hv);
                }
                sts8(a, pack8(f));
              }
            }
            fence_proxy_async();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(aready(s));
          if (++s == AST) { s = 0; ph ^= 1; }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

template <int BN, int MT, int AST, int BST, int NS, bool GN = false>
static int launch_igemm3(const IgemmKParams& kp, cudaStream_t st) {
  constexpr int smem = AST * MT * kHaloStride + BST * BN * 128 + NS * 128 * 128 + 1024 + 512 + NS * 512 + 32 + 64;
  static_assert(smem <= 227 * 1024, "igemm3: shared memory budget");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(igemm3_kernel<BN, MT, AST, BST, NS, GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (attr_err != cudaSuccess) { set_error("igemm3 smem attribute: %s", cudaGetErrorString(attr_err)); return CDAE_ERR_CUDA; }
  const int total = ((kp.nboxes + MT - 1) / MT) * kp.ntn;
  const int grid = total < kNumSMs ? total : kNumSMs;
  igemm3_kernel<BN, MT, AST, BST, NS, GN><<<grid, GN ? 256 + 32 * kGnWarps : 256, smem, st>>>(kp);
  CDAE_CHECK_LAUNCH("igemm3_kernel");
  return CDAE_OK;
}

// ------------------------------------------------------------------------------------------------ transposed halo kernel
// igemm3t_kernel: the 3x3 stride-1 conv for layers whose output-channel tile is 128 (the 64x64 level of the UNet: 28 % of the
// forward FLOPs).  igemm3_kernel runs them as M = 128 pixels x N = 128 channels: every 64-cycle MMA reads 4 KB of A and 4 KB
// of B - the whole 128 B/clk shared-memory port, which TMA fills and the epilogue staging share (tensor pipe 53 %, r1 ncu).
// Here the roles are swapped:   D[co (M = 128), pixel (N = 256)] = W[co, k] . X[pixel, k]
//   A = weight tile (128 rows x 64 k, K-major, plain SBO 1024)             16 KB per (tap, chunk)
//   B = the halo tile, now 8 wide x 32 high output pixels = 10 x 34 input pixels x 64 channels (340 rows, 42.5 KB per
//       chunk), tap (dh, dw) = the same row-shifted descriptor with SBO 1280 as in igemm3_kernel, 32 row groups
// so an MMA is 128 x 256 x 16: 4 KB + 8 KB of operands per 128 cycles = 96 B/clk, like the N = 256 tiles that reach 92 %.
// The accumulator has CHANNELS on the TMEM lanes: an epilogue thread owns one channel and walks 128 pixels of a slab, so the
// bias is a register, the GroupNorm statistics of the consumer (stats) and of the backward (gnb: sum du, sum du*x with the
// {a, b} constants in two registers) are plain per-thread sums - no transposed shared-memory pass, no shuffles.  The price is
// the transposing store: 16-bit writes into the [pixel][64 ch] slab (conflict-free: a warp writes 64 contiguous bytes).
// Roles: warp 0 halo producer | warp 1 MMA | warps 2-5 epilogue (lane quarter = warp % 4; quarters {0,1} / {2,3} form two
// PAIRS, each owning one 64-channel slab at a time) | warp 6 staging manager | warp 7 weight producer.
constexpr int kHaloTRows = 340;
constexpr int kHaloTBytes = kHaloTRows * 128;          // 43520
constexpr int kHaloTStride = 43 * 1024;                // 1024 B aligned

__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t u;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(addr));
  return (uint32_t)u;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v));
}
__device__ __forceinline__ uint32_t bf16_bits(float f) {
  const __nv_bfloat16 h = __float2bfloat16_rn(f);
  return (uint32_t)(*reinterpret_cast<const uint16_t*>(&h));
}

// GN: kGnWarps more warps (8..) normalise + activate every halo tile in place between the TMA load and the MMAs
// (cdae_igemm_desc.gn_*): thread = one 16-byte chunk column (8 channels, constants in registers) x every (4 kGnWarps)-th row.
// XF = 2 (cdae_igemm_desc.up2x): the sources are at HALF the output resolution; the halo producer loads the 6 x 18 low-resolution
// pixels under a 10 x 34 halo tile and the transform warps expand them (nearest neighbour: row (r + 1) >> 1, column
// (c + 1) >> 1) - F.interpolate(scale_factor=2) of unet.py:69-79 without the 4x tensor in HBM.
constexpr int kLoTRows = 108;                          // 6 x 18 low-resolution pixels
constexpr int kLoTBytes = kLoTRows * 128;              // 13824
constexpr int kLoTStride = 14 * 1024;
template <int AST, int BST, int NS, int XF>
__global__ void __launch_bounds__(XF ? 256 + 32 * kGnWarps : 256, 1) igemm3t_kernel(const __grid_constant__ IgemmKParams p) {
  constexpr bool GN = XF == 1, UP = XF == 2;
  constexpr int kWTileBytes = 128 * 128;                 // 128 output channels x 64 k
  constexpr int kSlabStride = 128 * 128;
  constexpr uint32_t kAccCols = 256;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, 256, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem + AST * kHaloTStride;
  uint8_t* stg = wsm + BST * kWTileBytes;
  uint8_t* losm = stg + NS * kSlabStride;                 // UP: low-resolution tiles [AST]
  uint64_t* bars = reinterpret_cast<uint64_t*>(losm + (UP ? AST * kLoTStride : 0));
  // bars: hfull[AST] hempty[AST] wfull[BST] wempty[BST] tfull[2] tempty[2] sready[NS] sfree[NS] hready[AST] lempty[AST]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * AST + 2 * BST + 4 + 2 * NS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t h_base = smem_u32(smem), w_base = smem_u32(wsm), stg_base = smem_u32(stg);
  const uint32_t bar_base = smem_u32(bars);
  auto hfull = [&](int s) { return bar_base + 8u * s; };
  auto hempty = [&](int s) { return bar_base + 8u * (AST + s); };
  auto wfull = [&](int s) { return bar_base + 8u * (2 * AST + s); };
  auto wempty = [&](int s) { return bar_base + 8u * (2 * AST + BST + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * AST + 2 * BST + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * AST + 2 * BST + 2 + a); };
  auto sready_bar = [&](int b) { return bar_base + 8u * (2 * AST + 2 * BST + 4 + b); };
  auto sfree_bar = [&](int b) { return bar_base + 8u * (2 * AST + 2 * BST + 4 + NS + b); };
  auto hready = [&](int s) { return bar_base + 8u * (2 * AST + 2 * BST + 4 + 2 * NS + s); };   // XF: tile transformed
  auto lempty = [&](int s) { return bar_base + 8u * (3 * AST + 2 * BST + 4 + 2 * NS + s); };   // UP: low-res tile consumed
  const uint32_t lo_base = smem_u32(losm);

  if (threadIdx.x == 0) {
    for (int s = 0; s < AST; ++s) {
      mbar_init(hfull(s), 1); mbar_init(hempty(s), 1); mbar_init(hready(s), kGnWarps); mbar_init(lempty(s), kGnWarps);
    }
    for (int s = 0; s < BST; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    for (int b = 0; b < NS; ++b) { mbar_init(sready_bar(b), 1); mbar_init(sfree_bar(b), 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile = (8 x 32 pixel box) x (128-channel block); boxes walk w fastest, then h, then the image
  const int tilesH = p.OHt / 32;
  const int boxes_per_img = p.tilesW * tilesH;
  const int nbox = boxes_per_img * p.Nimg;
  const int total_tiles = nbox * p.ntn;

  if (warp == 0) {
    // halo producer: the whole warp walks the loop, one elected lane issues the TMA loads (see elect_one)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int box = tile / p.ntn;
      const int cw = (box % p.tilesW) * 8 - 1, chh = ((box / p.tilesW) % tilesH) * 32 - 1, cn = box / boxes_per_img;
      for (int h = 0; h < p.nhs; ++h) {
        const CUtensorMap* tma = &p.tmA[p.hs[h].src];
        const int c0 = p.hs[h].c0, nch = p.hs[h].nchunk;
        for (int j = 0; j < nch; ++j) {
          mbar_wait(UP ? lempty(s) : hempty(s), ph ^ 1);
          if (elect_one()) {
            if (UP) {      // (cw, chh) = high-resolution halo origin (odd): the low-resolution origin is its arithmetic half
              mbar_expect_tx(hfull(s), kLoTBytes);
              tma_load_4d(lo_base + s * kLoTStride, tma, hfull(s), c0 + j * 64, (cw - 1) / 2, (chh - 1) / 2, cn);
            } else {
              mbar_expect_tx(hfull(s), kHaloTBytes);
              tma_load_4d(h_base + s * kHaloTStride, tma, hfull(s), c0 + j * 64, cw, chh, cn);
            }
          }
          __syncwarp();
          if (++s == AST) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 7) {
    // weight producer
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n0 = (tile % p.ntn) * 128;
      for (int h = 0; h < p.nhs; ++h) {
        const int nch = p.hs[h].nchunk, ntap = p.hs[h].ntap, wk0 = p.hs[h].wk0;
        for (int j = 0; j < nch; ++j) {
          for (int t = 0; t < ntap; ++t) {
            mbar_wait(wempty(s), ph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(wfull(s), kWTileBytes);
              tma_load_2d(w_base + s * kWTileBytes, &p.tmB, wfull(s), wk0 + t * p.tap_stride + j * 64, n0);
            }
            __syncwarp();
            if (++s == BST) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loop on warp-uniform values; one elected lane issues (see elect_one)
    int sa = 0, sb = 0, it = 0;
    uint32_t pha = 0, phb = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      uint32_t first = 1;
      for (int h = 0; h < p.nhs; ++h) {
        const int nch = p.hs[h].nchunk, ntap = p.hs[h].ntap;
        for (int j = 0; j < nch; ++j) {
          mbar_wait(XF ? hready(sa) : hfull(sa), pha);
          for (int t = 0; t < ntap; ++t) {
            const int ti = ntap == 9 ? (p.flip ? 8 - t : t) : 4;
            const uint32_t row0 = (uint32_t)((ti / 3) * 10 + (ti % 3));
            mbar_wait(wfull(sb), phb);
            tc_fence_after();
            const uint64_t adesc = smem_desc_kmajor_sw128(w_base + sb * kWTileBytes);
            const uint64_t bdesc = smem_desc_kmajor_sw128_sbo(h_base + sa * kHaloTStride + row0 * 128, 1280);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base + (uint32_t)(as * kAccCols), adesc + 2 * k, bdesc + 2 * k, kIdesc, (first && k == 0) ? 0u : 1u);
              umma_commit(wempty(sb));
              if (t == ntap - 1) umma_commit(hempty(sa));
            }
            first = 0;
            __syncwarp();
            if (++sb == BST) { sb = 0; phb ^= 1; }
          }
          if (++sa == AST) { sa = 0; pha ^= 1; }
        }
      }
      if (elect_one()) umma_commit(tfull_bar(as));
      __syncwarp();
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- epilogue: thread = output channel
    const int q = warp & 3, hp = q >> 1;                  // TMEM lane quarter; pair (= 64-channel slab half) of this warp
    const int cl = (q & 1) * 32 + lane;                   // channel inside the slab
    const bool elected = ((q & 1) == 0) && lane == 0;     // one thread per pair issues the stores
    const bool gnb = p.gnb_ws != nullptr;
    const bool need_x = gnb || p.has_resid;
    const uint32_t chunk = (uint32_t)(cl >> 3), sub = (uint32_t)(cl & 7) * 2u;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const int box = tile / p.ntn, n0 = (tile % p.ntn) * 128;
      const int w0 = (box % p.tilesW) * 8, h0 = ((box / p.tilesW) % tilesH) * 32, nimg = box / boxes_per_img;
      const int c = n0 + 64 * hp + cl;
      const bool cvalid = c < p.cout;
      float bsum = 0.f;
      if (cvalid) {
        if (p.bias) bsum += __ldg(p.bias + c);
        if (p.bias2) bsum += __ldg(p.bias2 + c);
        if (p.bias_img) bsum += __ldg(p.bias_img + (size_t)nimg * p.bias_img_ld + c);
      }
      float ga = 0.f, gb = 0.f;
      const __nv_bfloat16* dummy = nullptr; (void)dummy;
      if (gnb && p.gnb_silu && cvalid) {
        const float2 t2 = __ldg(reinterpret_cast<const float2*>(p.gnb_ab) + (size_t)nimg * p.cout + c);
        ga = t2.x; gb = t2.y;
      }
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;                           // {sum, sumsq} (stats) or {sum du, sum du*x} (gnb) of this channel
#pragma unroll 1
      for (int sb = 0; sb < 2; ++sb) {                    // 16-row sub-box = one [128 pixels][64 channels] slab per pair
        const int sidx = it * 4 + sb * 2 + hp;
        const int buf = sidx % NS;
        mbar_wait(sready_bar(buf), (sidx / NS) & 1);
        const uint32_t slab = stg_base + buf * kSlabStride;
#pragma unroll 1
        for (int cb4 = 0; cb4 < 4; ++cb4) {
          uint32_t acc[32];
          __syncwarp();
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccCols + sb * 128 + cb4 * 32), acc);
          tmem_ld_wait();
          // three passes over the 32 pixels so that the shared-memory reads are all in flight before the first use and the
          // 16-bit stores go out back to back (one pass per pixel serialises on the LDS latency: 36k cycles per tile, ncu r2)
          const uint32_t rowbase = slab + (uint32_t)(cb4 * 32) * 128u + sub;
          uint32_t xr[32];
          if (need_x) {
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[i] = lds16(rowbase + (uint32_t)i * 128u + ((chunk ^ (uint32_t)(i & 7)) << 4));
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v = __uint_as_float(acc[i]) + bsum;
            if (gnb) {
              const float x = __uint_as_float(xr[i] << 16);
              if (p.gnb_silu) {
                const float hh = fmaf(x, ga, gb), t = tanh_fast(hh);
                v *= 0.5f * (1.f + t + hh * (1.f - t * t));
              }
              const uint32_t vb = bf16_bits(v);
              const float vr = __uint_as_float(vb << 16);
              s1 += vr; s2 = fmaf(vr, x, s2);
              acc[i] = vb;
            } else {
              if (p.has_resid) v += __uint_as_float(xr[i] << 16);
              const uint32_t vb = bf16_bits(v);
              if (p.stats) { const float vr = __uint_as_float(vb << 16); s1 += vr; s2 = fmaf(vr, vr, s2); }
              acc[i] = vb;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) sts16(rowbase + (uint32_t)i * 128u + ((chunk ^ (uint32_t)(i & 7)) << 4), acc[i]);
        }
        fence_proxy_async();
        named_bar_sync(1 + hp, 64);
        if (elected) {
          tma_store_4d(&p.tmO, slab, n0 + 64 * hp, w0, h0 + 16 * sb, nimg);
          bulk_commit();
          bulk_wait_read<0>();
          mbar_arrive(sfree_bar(buf));
        }
      }
      if (cvalid && (gnb || p.stats)) {
        float* dst = (gnb ? p.gnb_ws : p.stats) + ((size_t)nimg * p.cout + c) * 2;
        atomicAdd(reinterpret_cast<float2*>(dst), make_float2(s1, s2));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
    if (elected) bulk_wait_all();
  } else if (warp == 6) {
    // ---------------------------------------------------------------- staging manager: residual / GroupNorm-input slabs
    if (lane == 0) {
      const bool gnb = p.gnb_ws != nullptr;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int box = tile / p.ntn, n0 = (tile % p.ntn) * 128;
        const int w0 = (box % p.tilesW) * 8, h0 = ((box / p.tilesW) % tilesH) * 32, nimg = box / boxes_per_img;
        for (int k = 0; k < 4; ++k) {
          const int sidx = it * 4 + k, sb = k >> 1, hp = k & 1;
          const int buf = sidx % NS;
          const int co0 = n0 + 64 * hp;
          mbar_wait(sfree_bar(buf), ((sidx / NS) & 1) ^ 1);
          if (gnb) {
            const bool in1 = co0 >= p.gnb_c0;
            mbar_expect_tx(sready_bar(buf), kSlabStride);
            tma_load_4d(stg_base + buf * kSlabStride, in1 ? &p.tmR2 : &p.tmR, sready_bar(buf), in1 ? co0 - p.gnb_c0 : co0, w0,
                        h0 + 16 * sb, nimg);
          } else if (p.has_resid) {
            mbar_expect_tx(sready_bar(buf), kSlabStride);
            tma_load_4d(stg_base + buf * kSlabStride, &p.tmR, sready_bar(buf), co0, w0, h0 + 16 * sb, nimg);
          } else {
            mbar_arrive(sready_bar(buf));
          }
        }
      }
    }
  } else if (UP && warp >= 8) {
    // ---------------------------------------------------------------- nearest x2 expansion of the low-resolution tile
    const int tt = (int)threadIdx.x - 256;
    const uint32_t ck = (uint32_t)(tt & 7);
    const int r0 = tt >> 3;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int h = 0; h < p.nhs; ++h) {
        const int nch = p.hs[h].nchunk;
        for (int j = 0; j < nch; ++j) {
          mbar_wait(hfull(s), ph);                           // low-resolution tile landed
          mbar_wait(hempty(s), ph ^ 1);                      // the MMAs are done with the previous contents of the halo tile
          const uint32_t src = lo_base + s * kLoTStride, dst = h_base + s * kHaloTStride;
          for (int r = r0; r < kHaloTRows; r += 4 * kGnWarps) {
            const int hh = r / 10, ww = r - hh * 10;
            const int lr = ((hh + 1) >> 1) * 6 + ((ww + 1) >> 1);
            sts8(dst + (uint32_t)r * 128u + ((ck ^ (uint32_t)(r & 7)) << 4),
                 lds8(src + (uint32_t)lr * 128u + ((ck ^ (uint32_t)(lr & 7)) << 4)));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { mbar_arrive(hready(s)); mbar_arrive(lempty(s)); }
          if (++s == AST) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (GN && warp >= 8) {
    // ---------------------------------------------------------------- GroupNorm + SiLU on the halo tile, in place
    // walks the halo stages in the producer's order; segments without a table (the 1x1-skip sources) pass through
    const int tt = (int)threadIdx.x - 256;                   // 0 .. 32 kGnWarps - 1
    const uint32_t ck = (uint32_t)(tt & 7);                  // logical 16-byte chunk: channels ck*8 .. +7 of the 64
    const int r0 = tt >> 3;                                  // rows r0, r0 + 4 kGnWarps, ...
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int box = tile / p.ntn;
      const int cw = (box % p.tilesW) * 8 - 1, chh = ((box / p.tilesW) % tilesH) * 32 - 1, cn = box / boxes_per_img;
      for (int h = 0; h < p.nhs; ++h) {
        const int nch = p.hs[h].nchunk, col0 = p.gn_col[h];
        for (int j = 0; j < nch; ++j) {
          float2 ab[8];
          if (col0 >= 0) {                                   // constants first: they do not depend on the tile
            const float4* q = reinterpret_cast<const float4*>(p.gn_ab + ((size_t)cn * p.gn_c + col0 + j * 64 + ck * 8) * 2);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 v = __ldg(q + e);
              ab[2 * e] = make_float2(v.x, v.y); ab[2 * e + 1] = make_float2(v.z, v.w);
            }
          }
          mbar_wait(hfull(s), ph);
          if (col0 >= 0) {
            const uint32_t tb = h_base + s * kHaloTStride;
            for (int r = r0; r < kHaloTRows; r += 4 * kGnWarps) {
              const int hh = r / 10, ww = r - hh * 10;
              if ((unsigned)(chh + hh) >= (unsigned)p.OHt || (unsigned)(cw + ww) >= (unsigned)p.OWt) continue;   // padding stays 0
              const uint32_t a = tb + (uint32_t)r * 128u + ((ck ^ (uint32_t)(r & 7)) << 4);
              float f[8];
              unpack8(lds8(a), f);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float hv = fmaf(f[e], ab[e].x, ab[e].y);
                f[e] = fmaf(hv, tanh_fast(hv), hv);
              }
              sts8(a, pack8(f));
            }
            fence_proxy_async();                             // generic-proxy stores -> visible to the MMA's async-proxy reads
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(hready(s));
          if (++s == AST) { s = 0; ph ^= 1; }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int AST, int BST, int NS, int XF = 0>
static int launch_igemm3t(const IgemmKParams& kp, cudaStream_t st) {
  constexpr int smem = AST * kHaloTStride + BST * 128 * 128 + NS * 128 * 128 + (XF == 2 ? AST * kLoTStride : 0) + 1024 + 512;
  static_assert(smem <= 227 * 1024, "igemm3t: shared memory budget");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(igemm3t_kernel<AST, BST, NS, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (attr_err != cudaSuccess) { set_error("igemm3t smem attribute: %s", cudaGetErrorString(attr_err)); return CDAE_ERR_CUDA; }
  const int total = kp.tilesW * (kp.OHt / 32) * kp.Nimg * kp.ntn;
  const int grid = total < kNumSMs ? total : kNumSMs;
  igemm3t_kernel<AST, BST, NS, XF><<<grid, XF ? 256 + 32 * kGnWarps : 256, smem, st>>>(kp);
  CDAE_CHECK_LAUNCH("igemm3t_kernel");
  return CDAE_OK;
}

// Does the segment list describe "3x3 stride-1 conv over some sources (+ optional 1x1 taps over others)" in the packing
// the engine uses (weight column = wk0 + tap*stride + channel)?  Fills kp.hs / tap_stride / flip when it does.
static bool halo_plan(const cdae_igemm_desc* d, IgemmKParams& kp) {
  if (d->in_stride != 1 || d->H % 16 != 0 || d->W % 8 != 0) return false;
  int stride = -1, flip = -1, nhs = 0;
  bool any9 = false;
  for (int s = 0; s < d->nsrc; ++s) {
    const cdae_seg* first = nullptr; const cdae_seg* last = nullptr; const cdae_seg* center = nullptr;
    int cnt = 0;
    for (int i = 0; i < d->nseg; ++i) {
      const cdae_seg& g = d->seg[i];
      if (g.src != s) continue;
      ++cnt;
      if (g.dh == -1 && g.dw == -1) first = &g;
      if (g.dh == 1 && g.dw == 1) last = &g;
      if (g.dh == 0 && g.dw == 0) center = &g;
    }
    if (cnt == 0) continue;
    if (nhs >= 8) return false;
    if (cnt == 1) {
      if (!center) return false;
      kp.hs[nhs].src = s; kp.hs[nhs].c0 = center->c0; kp.hs[nhs].nchunk = center->nchunk; kp.hs[nhs].ntap = 1;
      kp.hs[nhs].wk0 = center->wk;
      ++nhs;
      continue;
    }
    if (cnt != 9 || !first || !last || !center) return false;
    const int fl = first->wk < last->wk ? 0 : 1;
    const int wk0 = fl ? last->wk : first->wk;
    const int diff = fl ? first->wk - last->wk : last->wk - first->wk;
    if (diff % 8) return false;
    const int st = diff / 8;
    if ((stride >= 0 && st != stride) || (flip >= 0 && fl != flip)) return false;
    stride = st; flip = fl;
    for (int i = 0; i < d->nseg; ++i) {
      const cdae_seg& g = d->seg[i];
      if (g.src != s) continue;
      if (g.dh < -1 || g.dh > 1 || g.dw < -1 || g.dw > 1) return false;
      const int ti = fl ? (1 - g.dh) * 3 + (1 - g.dw) : (g.dh + 1) * 3 + (g.dw + 1);
      if (g.wk != wk0 + ti * st || g.c0 != center->c0 || g.nchunk != center->nchunk) return false;
    }
    kp.hs[nhs].src = s; kp.hs[nhs].c0 = center->c0; kp.hs[nhs].nchunk = center->nchunk; kp.hs[nhs].ntap = 9;
    kp.hs[nhs].wk0 = wk0;
    ++nhs;
    any9 = true;
  }
  if (!any9) return false;
  kp.nhs = nhs; kp.tap_stride = stride; kp.flip = flip;
  return true;
}

}  // namespace cdae

using namespace cdae;

extern "C" int cdae_igemm(const cdae_igemm_desc* d, cdae_stream s) {
  if (d && (d->N == 0 || d->H == 0 || d->W == 0)) return CDAE_OK;        // empty batch: nothing to compute
  CDAE_CHECK_ARG(d && d->out && d->wgt && d->nsrc >= 1 && d->nsrc <= 4, "igemm: bad descriptor");
  CDAE_CHECK_ARG(d->nseg >= 1 && d->nseg <= CDAE_MAX_SEG, "igemm: nseg %d out of range", d->nseg);
  CDAE_CHECK_SHAPE(d->in_stride == 1 || d->in_stride == 2, "igemm: in_stride %d", d->in_stride);
  CDAE_CHECK_SHAPE(d->wk % 8 == 0, "igemm: weight K %d must be a multiple of 8", d->wk);
  // strided output placement (sps = 2): tile pixel (y, x) is stored at (y*sps + ooh, x*sps + oow) of the [OH, OW] output - one
  // parity class of a stride-2 data gradient (conv_transpose).  Expressed in the TMA maps of the output / residual alone.
  const int sps = d->sps > 0 ? d->sps : 1;
  CDAE_CHECK_SHAPE((sps == 1 && d->ooh == 0 && d->oow == 0) ||
                       (sps == 2 && d->ooh >= 0 && d->ooh < 2 && d->oow >= 0 && d->oow < 2 && d->out_mode == 0 && !d->stats &&
                        !d->gnb_ws && !d->bias_img),
                   "igemm: output placement sps=%d ooh=%d oow=%d unsupported (stride 2 needs a plain NHWC launch)", d->sps, d->ooh,
                   d->oow);
  IgemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  const int es = d->in_stride;
  const int up = d->up2x ? 2 : 1;                  // sources at half the output resolution (nearest x2 on load)
  CDAE_CHECK_SHAPE(up == 1 || es == 1, "igemm: up2x needs a stride-1 conv");
  const int OHt = up == 2 ? 2 * d->H : (d->H + es - 1) / es, OWt = up == 2 ? 2 * d->W : (d->W + es - 1) / es;
  static const bool no_halo = getenv("CDAE_NO_HALO") != nullptr;
  const bool halo = !no_halo && halo_plan(d, kp) && OHt % 16 == 0 && OWt % 8 == 0;
  if (halo) { kp.BW = 8; kp.BH = 16; kp.BNI = 1; }
  else tile_geometry(128, OHt, OWt, &kp.BW, &kp.BH, &kp.BNI);
  kp.tilesW = (OWt + kp.BW - 1) / kp.BW;
  kp.tilesH = (OHt + kp.BH - 1) / kp.BH;
  const int tilesN = (d->N + kp.BNI - 1) / kp.BNI;
  kp.in_stride = es; kp.Nimg = d->N; kp.OHt = OHt; kp.OWt = OWt;
  kp.OH = d->OH; kp.OW = d->OW; kp.cout = d->cout; kp.out_mode = d->out_mode; kp.ldo = d->ldo;
  CDAE_CHECK_SHAPE(d->out_mode >= 0 && d->out_mode <= 2, "igemm: out_mode %d", d->out_mode);
  CDAE_CHECK_SHAPE(d->out_mode != 2 || (d->ldo % 4 == 0 && d->ldo >= d->cout && (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 &&
                                        (!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0)),
                   "igemm: fp32 row-major output needs ldo %% 4 == 0, ldo >= cout and 16-byte aligned out / bias");
  kp.out = d->out; kp.bias = d->bias; kp.bias2 = d->bias2; kp.has_resid = d->resid != nullptr;
  kp.bias_img = d->bias_img; kp.bias_img_ld = d->bias_img_ld;
  CDAE_CHECK_SHAPE(!d->bias_img || (d->out_mode == 0 && d->bias_img_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(d->bias_img) & 15) == 0),
                   "igemm: the per-image bias needs NHWC output, a pitch %% 4 == 0 and a 16-byte aligned pointer");
  kp.stats = d->stats;
  if (d->gnb_ws) {
    CDAE_CHECK_ARG(d->gnb_x0 && (!d->gnb_silu || d->gnb_ab), "igemm: GroupNorm-backward fusion needs x0 (and the constant table with SiLU)");
    CDAE_CHECK_SHAPE(kp.out_mode == 0 && !d->resid && !d->stats && !d->bias && d->cout % 64 == 0 && kp.BW * kp.BH >= 32 &&
                         d->gnb_c0 % 64 == 0 && d->gnb_c0 > 0 && d->gnb_c0 <= d->cout && (d->gnb_c0 == d->cout || d->gnb_x1) &&
                         d->gnb_ld0 % 8 == 0 && (!d->gnb_x1 || d->gnb_ld1 % 8 == 0) &&
                         (reinterpret_cast<uintptr_t>(d->gnb_ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->gnb_ab) & 15) == 0,
                     "igemm: GroupNorm-backward fusion needs a plain NHWC data-gradient launch (no bias / residual / statistics), "
                     "cout %% 64 == 0, 64-channel aligned sources and >= 32 output pixels per image");
    kp.gnb_ws = d->gnb_ws; kp.gnb_ab = d->gnb_ab; kp.gnb_silu = d->gnb_silu;
    kp.gnb_x0 = reinterpret_cast<const __nv_bfloat16*>(d->gnb_x0); kp.gnb_x1 = reinterpret_cast<const __nv_bfloat16*>(d->gnb_x1);
    kp.gnb_c0 = d->gnb_c0; kp.gnb_ld0 = d->gnb_ld0; kp.gnb_ld1 = d->gnb_ld1;
    for (kp.lgBW = 0; (1 << kp.lgBW) < kp.BW; ++kp.lgBW) {}
    for (kp.lgBH = 0; (1 << kp.lgBH) < kp.BH; ++kp.lgBH) {}
  }
  CDAE_CHECK_SHAPE(!d->stats || (kp.out_mode == 0 && d->cout % 64 == 0 && kp.BW * kp.BH >= 32 &&
                                 (reinterpret_cast<uintptr_t>(d->stats) & 15) == 0),
                   "igemm: channel statistics need NHWC output, cout %% 64 == 0 and >= 32 output pixels per image");
  CDAE_CHECK_SHAPE(kp.out_mode != 0 || (d->cout % 8 == 0 && d->ldo % 8 == 0), "igemm: NHWC output needs cout, ldo %% 8 == 0");
  CDAE_CHECK_SHAPE(kp.out_mode == 1 || (d->OH == OHt * sps && d->OW == OWt * sps),
                   "igemm: output dims %dx%d do not match the tile grid %dx%d", d->OH, d->OW, OHt, OWt);
  CDAE_CHECK_SHAPE(!d->resid || (d->ldr % 8 == 0 && kp.out_mode == 0), "igemm: residual needs NHWC output and pitch %% 8");
  const int nboxes = kp.tilesW * kp.tilesH * tilesN;
  int bn = d->bn, mt = 1;
  if (bn == 0) {
    const int c = d->cout;
    if (c >= 256 && c % 256 == 0 && (int64_t)nboxes * (c / 256) >= kNumSMs) bn = 256;
    else if (c >= 192 && c % 192 == 0 && (int64_t)nboxes * (c / 192) >= kNumSMs) bn = 192;
    else bn = c >= 128 ? 128 : c >= 64 ? 64 : c > 16 ? 32 : 16;
  }
  // 128-channel tiles of the halo kernel are shared-memory bound (M = N = 128): run them transposed (channels on the TMEM
  // lanes, 256 pixels per MMA) when the image tiles into 8 x 32 boxes
  static const bool no_t = getenv("CDAE_NO_IGEMM3T") != nullptr;
  // (measured, profiles/r2_igemm_bench_t*.log: never slower than the pixel-major N = 256 / 192 tiles either, and the
  // statistics epilogue is free in this orientation, so every eligible layer takes it)
  const bool use_t = halo && !no_t && d->cout % 128 == 0 && kp.out_mode == 0 && OHt % 32 == 0 && OWt % 8 == 0 && d->bn == 0;
  if (use_t) bn = 128;
  for (int i = 0; i < d->nsrc; ++i) {
    CDAE_CHECK_ARG(d->src[i], "igemm: null source %d", i);
    CDAE_CHECK_SHAPE(d->src_c[i] % 8 == 0, "igemm: source %d channels %d must be a multiple of 8", i, d->src_c[i]);
    const uint64_t C = d->src_c[i];
    uint64_t dims[4] = {C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[3] = {C * 2, C * 2 * d->W, C * 2 * (uint64_t)d->W * d->H};
    uint32_t box[4] = {64, (uint32_t)(kp.BW * es), (uint32_t)(kp.BH * es), (uint32_t)kp.BNI};
    if (halo) { box[1] = 10; box[2] = use_t ? 34 : 18; box[3] = 1; }     // output box + one pixel of halo on every side
    if (up == 2) { box[1] = 6; box[2] = 18; }                           // the low-resolution pixels under a 10 x 34 halo tile
    uint32_t est[4] = {1, (uint32_t)es, (uint32_t)es, 1};
    int rc = make_tmap_bf16(&kp.tmA[i], d->src[i], 4, dims, str, box, est);
    if (rc) return rc;
  }
  CDAE_CHECK_SHAPE(!(d->stats || d->gnb_ws) || bn >= 64, "igemm: channel statistics need an N tile of at least 64 (bn %d)", bn);
  const int ntn = (d->cout + bn - 1) / bn;
  if (bn <= 128 && (int64_t)((nboxes + 1) / 2) * ntn >= kNumSMs) mt = 2;
  {
    uint64_t dims[2] = {(uint64_t)d->wk, (uint64_t)d->wrows};
    uint64_t str[1] = {(uint64_t)d->wk * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    int rc = make_tmap_bf16(&kp.tmB, d->wgt, 2, dims, str, box, nullptr);
    if (rc) return rc;
  }
  if (kp.out_mode == 0) {
    const uint32_t slabw = bn < 64 ? bn : 64;
    uint32_t box[4] = {slabw, (uint32_t)kp.BW, (uint32_t)kp.BH, (uint32_t)kp.BNI};
    {
      const uint64_t L = d->ldo;
      uint64_t dims[4] = {(uint64_t)d->cout, (uint64_t)OWt, (uint64_t)OHt, (uint64_t)d->N};
      uint64_t str[3] = {L * 2 * sps, L * 2 * d->OW * sps, L * 2 * (uint64_t)d->OW * d->OH};
      char* base = reinterpret_cast<char*>(d->out) + ((size_t)d->ooh * d->OW + d->oow) * L * 2;
      int rc = make_tmap_bf16(&kp.tmO, base, 4, dims, str, box, nullptr, slabw == 64);
      if (rc) return rc;
    }
    if (d->resid) {
      const uint64_t L = d->ldr;
      uint64_t dims[4] = {(uint64_t)d->cout, (uint64_t)OWt, (uint64_t)OHt, (uint64_t)d->N};
      uint64_t str[3] = {L * 2 * sps, L * 2 * d->OW * sps, L * 2 * (uint64_t)d->OW * d->OH};
      const char* base = reinterpret_cast<const char*>(d->resid) + ((size_t)d->ooh * d->OW + d->oow) * L * 2;
      int rc = make_tmap_bf16(&kp.tmR, base, 4, dims, str, box, nullptr, slabw == 64);
      if (rc) return rc;
    }
    if (d->gnb_ws) {      // x slabs of the GroupNorm input arrive over the residual path: one map per concatenated source
      for (int si = 0; si < (d->gnb_x1 ? 2 : 1); ++si) {
        const uint64_t L = si ? d->gnb_ld1 : d->gnb_ld0;
        const uint64_t Cs = si ? (uint64_t)(d->cout - d->gnb_c0) : (uint64_t)d->gnb_c0;
        uint64_t dims[4] = {Cs, (uint64_t)d->OW, (uint64_t)d->OH, (uint64_t)d->N};
        uint64_t str[3] = {L * 2, L * 2 * d->OW, L * 2 * (uint64_t)d->OW * d->OH};
        int rc = make_tmap_bf16(si ? &kp.tmR2 : &kp.tmR, si ? d->gnb_x1 : d->gnb_x0, 4, dims, str, box, nullptr, slabw == 64);
        if (rc) return rc;
      }
    }
  }
  int nkb = 0;
  for (int i = 0; i < d->nseg; ++i) {
    const cdae_seg& g = d->seg[i];
    CDAE_CHECK_ARG(g.src >= 0 && g.src < d->nsrc && g.nchunk >= 1, "igemm: bad segment %d", i);
    CDAE_CHECK_SHAPE(g.wk + g.nchunk * 64 <= d->wk + 56, "igemm: segment %d overruns weight K", i);
    kp.seg[i] = g;
    nkb += g.nchunk;
  }
  kp.nseg = d->nseg; kp.nkb = nkb; kp.nboxes = nboxes; kp.ntn = ntn;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  if (up == 2) {
    CDAE_CHECK_SHAPE(use_t && !d->gn_ab && !d->gnb_ws && !d->resid,
                     "igemm: up2x needs a plain 3x3 stride-1 layer with cout %% 128 == 0 whose output tiles into 8x32 boxes");
    return launch_igemm3t<2, 3, 3, 2>(kp, st);
  }
  if (d->gn_ab) {
    CDAE_CHECK_SHAPE(halo && !d->gnb_ws && d->gn_c > 0 && (reinterpret_cast<uintptr_t>(d->gn_ab) & 15) == 0 && d->gn_c % 4 == 0,
                     "igemm: GroupNorm on load needs a 3x3 stride-1 layer on an image that tiles into 8x16 boxes (the halo "
                     "kernels), a 16-byte aligned table and gn_c %% 4 == 0");
    kp.gn_ab = d->gn_ab; kp.gn_c = d->gn_c;
    for (int h = 0; h < kp.nhs; ++h) {
      const int off = d->gn_off[kp.hs[h].src];
      CDAE_CHECK_SHAPE(off < 0 || (kp.hs[h].ntap == 9 && (off + kp.hs[h].c0) % 8 == 0 &&
                                   off + kp.hs[h].c0 + kp.hs[h].nchunk * 64 <= d->gn_c),
                       "igemm: GroupNorm on load: source %d (columns %d..) does not fit the table of %d channels", kp.hs[h].src, off,
                       d->gn_c);
      kp.gn_col[h] = off < 0 ? -1 : off + kp.hs[h].c0;
    }
    // three halo stages: load -> transform -> MMA are all in flight (2/5/3 and 2/4/4 measured 1-3 % slower, r2_gnload_bench)
    if (use_t) return launch_igemm3t<3, 3, 3, 1>(kp, st);
    switch (bn) {          // the pixel-major halo kernel (16x16 levels, narrow heads): one more halo stage where it fits
      case 16: return launch_igemm3<16, 2, 3, 8, 3, true>(kp, st);
      case 32: return launch_igemm3<32, 2, 3, 8, 3, true>(kp, st);
      case 64: return mt == 2 ? launch_igemm3<64, 2, 3, 4, 3, true>(kp, st) : launch_igemm3<64, 1, 3, 8, 3, true>(kp, st);
      case 128: return mt == 2 ? launch_igemm3<128, 2, 2, 5, 3, true>(kp, st) : launch_igemm3<128, 1, 3, 6, 3, true>(kp, st);
      case 192: return launch_igemm3<192, 1, 3, 4, 3, true>(kp, st);
      case 256: return launch_igemm3<256, 1, 3, 3, 3, true>(kp, st);
      default: set_error("igemm: unsupported bn %d", bn); return CDAE_ERR_SHAPE;
    }
  }
  if (use_t) {
    static const char* tcfg = getenv("CDAE_T_CFG");           // staging experiments: halo stages / weight stages / slabs
    if (tcfg && tcfg[0] == '3') return launch_igemm3t<3, 3, 3>(kp, st);
    if (tcfg && tcfg[0] == '4') return launch_igemm3t<2, 4, 4>(kp, st);
    return launch_igemm3t<2, 5, 3>(kp, st);        // measured best on every cfg2 shape (profiles/r2_igemm_bench_t253_stats.log)
  }
  if (halo) {
    switch (bn) {
      case 16: return launch_igemm3<16, 2, 2, 8, 3>(kp, st);
      case 32: return launch_igemm3<32, 2, 2, 8, 3>(kp, st);
      case 64: return mt == 2 ? launch_igemm3<64, 2, 2, 8, 3>(kp, st) : launch_igemm3<64, 1, 3, 8, 3>(kp, st);
      case 128: return mt == 2 ? launch_igemm3<128, 2, 2, 5, 3>(kp, st) : launch_igemm3<128, 1, 3, 6, 3>(kp, st);
      case 192: return launch_igemm3<192, 1, 2, 5, 3>(kp, st);
      case 256: return launch_igemm3<256, 1, 2, 4, 3>(kp, st);
      default: set_error("igemm: unsupported bn %d", bn); return CDAE_ERR_SHAPE;
    }
  }
  switch (bn) {
    case 16: return launch_igemm2<16, 1, 8, 3>(kp, st);
    case 32: return launch_igemm2<32, 1, 8, 3>(kp, st);
    case 64: return mt == 2 ? launch_igemm2<64, 2, 4, 3>(kp, st) : launch_igemm2<64, 1, 7, 3>(kp, st);
    case 128: return mt == 2 ? launch_igemm2<128, 2, 3, 3>(kp, st) : launch_igemm2<128, 1, 5, 3>(kp, st);
    case 192: return launch_igemm2<192, 1, 4, 3>(kp, st);
    case 256: return launch_igemm2<256, 1, 3, 3>(kp, st);
    default: set_error("igemm: unsupported bn %d", bn); return CDAE_ERR_SHAPE;
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
namespace cdae {

struct alignas(64) WgradKParams {
  CUtensorMap tmDy;   // dY  [N, OH, OW, ldy]   box {64, bw, bh, bni}
  CUtensorMap tmX;    // src [N, H, W, src_c]   box {64, bw*es, bh*es, bni}
  int bw, bh, bni, tilesW, tilesH, ntiles;   // 64-pixel K tiles over the dY pixel grid
  int es, ksize, c0;
  int tiles_per_split;
  float* dw; int dw_ld, taps, ci_off, cin_real, cout;
  float* dbias;        // fused bias gradient (column sums of dY) or nullptr
};

constexpr int kWgBoxBytes = 64 * 128;  // 64 pixels x 64 channels bf16

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) wgrad_kernel(const __grid_constant__ WgradKParams p) {
  constexpr int kABytes = 2 * kWgBoxBytes;            // 128 output channels
  constexpr int kBBytes = (BN / 64) * kWgBoxBytes;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BN + 8 <= 128 ? 128 : BN + 8 <= 256 ? 256 : 512;   // + 8 columns for the bias gradient
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN, 1, 1);
  constexpr uint32_t kIdescOnes = make_idesc_bf16(128, 8, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + STAGES * kStageBytes;                      // 16 rows x 128 B of bf16 1.0
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 2048);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);

  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int co0 = blockIdx.x * 128;
  const int ci_tiles = gridDim.y / p.taps;
  const int tap = blockIdx.y / ci_tiles, ci0 = (blockIdx.y % ci_tiles) * BN;
  // The bias gradient (column sums of dY) rides along as an N = 8 MMA against a tile of ones.  Every (tap, ci) column of
  // CTAs sees the same dY tiles, so pixel tile t is summed by column t % gridDim.y: the extra MMAs are spread evenly
  // instead of making one column of CTAs the straggler of the launch.
  const int bcols = gridDim.y, bcol = blockIdx.y;
  const int dh = p.ksize == 3 ? tap / 3 - 1 : 0, dw = p.ksize == 3 ? tap % 3 - 1 : 0;
  const int t_begin = blockIdx.z * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split; if (t_end > p.ntiles) t_end = p.ntiles;
  const int nkb = t_end - t_begin;

  if (nkb > 0) {
    if (warp == 0) {
      // producer: the whole warp walks the loop, one elected lane issues the TMA loads (see elect_one)
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int t = t_begin + kb;
        const int tw = t % p.tilesW, th = (t / p.tilesW) % p.tilesH, tn = t / (p.tilesW * p.tilesH);
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t a_dst = smem_base + s * kStageBytes;
        const int ow = tw * p.bw, oh = th * p.bh, nn = tn * p.bni;
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), kStageBytes);
          tma_load_4d(a_dst, &p.tmDy, full_bar(s), co0, ow, oh, nn);
          tma_load_4d(a_dst + kWgBoxBytes, &p.tmDy, full_bar(s), co0 + 64, ow, oh, nn);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(a_dst + kABytes + j * kWgBoxBytes, &p.tmX, full_bar(s), p.c0 + ci0 + j * 64, ow * p.es + dw,
                        oh * p.es + dh, nn);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      // the whole warp walks the loop on warp-uniform values; one elected lane issues (see elect_one)
      int bias_started = 0, s = 0;
      uint32_t ph = 0;
      const uint64_t odesc = smem_desc_mnmajor_sw128(smem_u32(ones), 0, 1024);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * kStageBytes;
        // MN-major: 64-channel blocks kWgBoxBytes apart (LBO), 8-pixel K groups 1024 B apart (SBO)
        const uint64_t adesc = smem_desc_mnmajor_sw128(a_addr, kWgBoxBytes, 1024);
        const uint64_t bdesc = smem_desc_mnmajor_sw128(a_addr + kABytes, kWgBoxBytes, 1024);
        const bool with_bias = p.dbias && (t_begin + kb) % bcols == bcol;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // K = 16 pixels = two 1024 B atoms = +128 in the >>4 address field
            umma_f16(tmem_base, adesc + 128 * k, bdesc + 128 * k, kIdesc, (kb | k) != 0);
          if (with_bias) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base + BN, adesc + 128 * k, odesc, kIdescOnes, (bias_started | k) != 0);
          }
          umma_commit(empty_bar(s));
          if (kb == nkb - 1) umma_commit(tmem_full_bar);
        }
        if (with_bias) bias_started = 1;
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    } else {
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
      const bool do_bias = p.dbias != nullptr && t_begin + ((bcol - t_begin % bcols) + bcols) % bcols < t_end;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t acc[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, acc);
        tmem_ld_wait();
        float* row = p.dw + ((size_t)co * p.taps + tap) * p.dw_ld + p.ci_off + ci0 + c;
        const int lim = p.cin_real - (ci0 + c);
        if (co >= p.cout) {
          // padded output-channel row: nothing to accumulate
        } else if (lim >= 32 && (p.dw_ld % 4 == 0) && ((p.ci_off & 3) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            atomicAdd(reinterpret_cast<float4*>(row + j),
                      make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                  __uint_as_float(acc[j + 3])));
        } else {
          for (int j = 0; j < 32 && j < lim; ++j) atomicAdd(row + j, __uint_as_float(acc[j]));
        }
      }
      if (do_bias) {
        uint32_t acc[16];
        __syncwarp();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)BN, acc);   // columns 0..7 hold the same sum
        tmem_ld_wait();
        if (co < p.cout) atomicAdd(p.dbias + co, __uint_as_float(acc[0]));
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

template <int BN, int STAGES>
static int launch_wgrad(const WgradKParams& kp, dim3 grid, cudaStream_t st) {
  constexpr int smem = STAGES * (2 * kWgBoxBytes + (BN / 64) * kWgBoxBytes) + 2048 + 1024 + 256;
  static_assert(smem <= 227 * 1024, "wgrad: shared memory budget");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (attr_err != cudaSuccess) { set_error("wgrad smem attribute: %s", cudaGetErrorString(attr_err)); return CDAE_ERR_CUDA; }
  wgrad_kernel<BN, STAGES><<<grid, 192, smem, st>>>(kp);
  CDAE_CHECK_LAUNCH("wgrad_kernel");
  return CDAE_OK;
}

// ------------------------------------------------------------------------------------------------ 3x3 stride-1 weight gradient
// One CTA = (128 output channels) x (BN input channels) x (ONE KERNEL ROW: the three taps dw = -1,0,+1) over a split of
// the pixel tiles.  Per 8x8-pixel K tile the CTA loads dY once (64 px x 128 co) and ONE halo tile of the source
// (8 rows x 10 columns x BN channels); the three taps read that halo tile through row-shifted MN-major descriptors
// (start address + (dw+1)*128 B, 8-pixel groups 1280 B apart - the 128B swizzle is a pure function of the shared-memory
// address, see tools/exp_desc.cu), each into its own TMEM accumulator.  L2 -> SM traffic per MMA drops 2.7x against the
// one-tap kernel.  The bias gradient (column sums of dY, aten::convolution_backward's third output) rides along as an
// N = 8 MMA against a tile of ones.
struct alignas(64) Wgrad3KParams {
  CUtensorMap tmDy;   // dY  [N, OH, OW, ldy]  box {64, 8, 8, 1}
  CUtensorMap tmX;    // src [N, H, W, C]      box {64, 10, 8, 1}
  int tilesW, tilesH, ntiles, tiles_per_split;
  int c0;
  float* dw; int dw_ld, ci_off, cin_real, cout;
  float* dbias;
};

constexpr int kW3DyBytes = 2 * kWgBoxBytes;        // 64 px x 128 co
constexpr int kW3XBlk = 80 * 128;                  // 8 rows x 10 columns of pixels x 64 channels

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) wgrad3_kernel(const __grid_constant__ Wgrad3KParams p) {
  constexpr int kXBytes = (BN / 64) * kW3XBlk;
  constexpr int kStageBytes = kW3DyBytes + ((kXBytes + 1023) & ~1023);
  constexpr uint32_t kTmemCols = 3 * BN + 8 <= 256 ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN, 1, 1);
  constexpr uint32_t kIdescOnes = make_idesc_bf16(128, 8, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + STAGES * kStageBytes;                      // 16 rows x 128 B of bf16 1.0
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 2048);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);

  const int co0 = blockIdx.x * 128;
  const int ci_tiles = gridDim.y / 3;
  const int krow = blockIdx.y / ci_tiles, ci0 = (blockIdx.y % ci_tiles) * BN;     // kernel row 0..2 -> dh = krow - 1
  const int bcols = gridDim.y, bcol = blockIdx.y;       // bias gradient: pixel tile t is summed by column t % gridDim.y
  const int t_begin = blockIdx.z * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split; if (t_end > p.ntiles) t_end = p.ntiles;
  const int nkb = t_end - t_begin;

  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nkb > 0) {
    if (warp == 0) {
      // producer: the whole warp walks the loop, one elected lane issues the TMA loads (see elect_one)
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int t = t_begin + kb;
        const int tw = t % p.tilesW, th = (t / p.tilesW) % p.tilesH, tn = t / (p.tilesW * p.tilesH);
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t a_dst = smem_base + s * kStageBytes;
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), kW3DyBytes + kXBytes);
          tma_load_4d(a_dst, &p.tmDy, full_bar(s), co0, tw * 8, th * 8, tn);
          tma_load_4d(a_dst + kWgBoxBytes, &p.tmDy, full_bar(s), co0 + 64, tw * 8, th * 8, tn);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(a_dst + kW3DyBytes + j * kW3XBlk, &p.tmX, full_bar(s), p.c0 + ci0 + j * 64, tw * 8 - 1,
                        th * 8 + krow - 1, tn);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    } else if (warp == 1) {
      // the whole warp walks the loop on warp-uniform values; one elected lane issues (see elect_one)
      int bias_started = 0, s = 0;
      uint32_t ph = 0;
      const uint64_t odesc = smem_desc_mnmajor_sw128(smem_u32(ones), 0, 1024);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * kStageBytes;
        // A = dY^T: 64-channel blocks kWgBoxBytes apart (LBO), 8-pixel K groups 1024 B apart (SBO)
        const uint64_t adesc = smem_desc_mnmajor_sw128(a_addr, kWgBoxBytes, 1024);
        // B = shifted source window: pixel (h, w) of the tile sits at halo row h*10 + w + tap
        uint64_t bdesc[3];
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) bdesc[tap] = smem_desc_mnmajor_sw128(a_addr + kW3DyBytes + tap * 128, kW3XBlk, 1280);
        const bool with_bias = p.dbias && (t_begin + kb) % bcols == bcol;
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
            for (int k = 0; k < 4; ++k)   // K = 16 pixels = two image rows of the tile: +2 SBO steps per MMA
              umma_f16(tmem_base + tap * BN, adesc + 128 * k, bdesc[tap] + 160 * k, kIdesc, (kb | k) != 0);
          }
          if (with_bias) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 3 * BN, adesc + 128 * k, odesc, kIdescOnes, (bias_started | k) != 0);
          }
          umma_commit(empty_bar(s));
          if (kb == nkb - 1) umma_commit(tmem_full_bar);
        }
        if (with_bias) bias_started = 1;
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    } else {
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
      const bool do_bias = p.dbias != nullptr && t_begin + ((bcol - t_begin % bcols) + bcols) % bcols < t_end;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int tap = 0; tap < 3; ++tap) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t acc[32];
          __syncwarp();
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * BN + c), acc);
          tmem_ld_wait();
          float* row = p.dw + ((size_t)co * 9 + krow * 3 + tap) * p.dw_ld + p.ci_off + ci0 + c;
          const int lim = p.cin_real - (ci0 + c);
          if (co >= p.cout) {
            // padded output-channel row: nothing to accumulate
          } else if (lim >= 32 && (p.dw_ld % 4 == 0) && ((p.ci_off & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(row + j),
                        make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                    __uint_as_float(acc[j + 3])));
          } else {
            for (int j = 0; j < 32 && j < lim; ++j) atomicAdd(row + j, __uint_as_float(acc[j]));
          }
        }
      }
      if (do_bias) {
        uint32_t acc[16];
        __syncwarp();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(3 * BN), acc);   // columns 0..7 hold the same sum
        tmem_ld_wait();
        if (co < p.cout) atomicAdd(p.dbias + co, __uint_as_float(acc[0]));
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

template <int BN, int STAGES>
static int launch_wgrad3(const Wgrad3KParams& kp, dim3 grid, cudaStream_t st) {
  constexpr int xb = ((BN / 64) * kW3XBlk + 1023) & ~1023;
  constexpr int smem = STAGES * (kW3DyBytes + xb) + 2048 + 1024 + 256;
  static_assert(smem <= 227 * 1024, "wgrad3: shared memory budget");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(wgrad3_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (attr_err != cudaSuccess) { set_error("wgrad3 smem attribute: %s", cudaGetErrorString(attr_err)); return CDAE_ERR_CUDA; }
  wgrad3_kernel<BN, STAGES><<<grid, 192, smem, st>>>(kp);
  CDAE_CHECK_LAUNCH("wgrad3_kernel");
  return CDAE_OK;
}

static int wgrad3(const cdae_wgrad_desc* d, cudaStream_t st) {
  Wgrad3KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.tilesW = (d->OW + 7) / 8; kp.tilesH = (d->OH + 7) / 8;
  kp.ntiles = kp.tilesW * kp.tilesH * d->N;
  kp.c0 = d->c0;
  kp.dw = d->dw; kp.dw_ld = d->dw_ld; kp.ci_off = d->ci_off;
  kp.cin_real = d->cin_real > 0 ? d->cin_real : d->cin; kp.cout = d->cout;
  kp.dbias = d->dbias;
  {
    const uint64_t C = d->ldy;
    uint64_t dims[4] = {C, (uint64_t)d->OW, (uint64_t)d->OH, (uint64_t)d->N};
    uint64_t str[3] = {C * 2, C * 2 * d->OW, C * 2 * (uint64_t)d->OW * d->OH};
    uint32_t box[4] = {64, 8, 8, 1};
    int rc = make_tmap_bf16(&kp.tmDy, d->dy, 4, dims, str, box, nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t C = d->src_c;
    uint64_t dims[4] = {C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[3] = {C * 2, C * 2 * d->W, C * 2 * (uint64_t)d->W * d->H};
    uint32_t box[4] = {64, 10, 8, 1};
    int rc = make_tmap_bf16(&kp.tmX, d->src, 4, dims, str, box, nullptr);
    if (rc) return rc;
  }
  const int bn = d->cin > 64 ? 128 : 64;
  const int co_tiles = (d->cout + 127) / 128, ci_tiles = (d->cin + bn - 1) / bn;
  int splits = d->splits;
  if (splits <= 0) {
    const int base = co_tiles * ci_tiles * 3;
    splits = base >= kNumSMs ? 1 : kNumSMs / base;
    if (splits > kp.ntiles) splits = kp.ntiles;
    if (splits < 1) splits = 1;
  }
  kp.tiles_per_split = (kp.ntiles + splits - 1) / splits;
  splits = (kp.ntiles + kp.tiles_per_split - 1) / kp.tiles_per_split;
  dim3 grid(co_tiles, ci_tiles * 3, splits);
  if (bn == 128) return launch_wgrad3<128, 5>(kp, grid, st);
  return launch_wgrad3<64, 7>(kp, grid, st);
}

}  // namespace cdae

extern "C" int cdae_wgrad(const cdae_wgrad_desc* d, cdae_stream s) {
  if (d && (d->N == 0 || d->OH == 0 || d->OW == 0)) return CDAE_OK;      // empty batch: dW += 0
  CDAE_CHECK_ARG(d && d->dy && d->src && d->dw, "wgrad: bad descriptor");
  CDAE_CHECK_SHAPE(d->ksize == 1 || d->ksize == 3, "wgrad: ksize %d", d->ksize);
  CDAE_CHECK_SHAPE(d->in_stride == 1 || d->in_stride == 2, "wgrad: in_stride %d", d->in_stride);
  CDAE_CHECK_SHAPE(d->ldy % 8 == 0 && d->src_c % 8 == 0, "wgrad: pitches must be multiples of 8");
  static const bool no_w3 = getenv("CDAE_WGRAD_V1") != nullptr;
  // one-kernel-row kernel (N = 128 per tap): wins wherever the one-tap kernel cannot run its N = 256 tile
  if (d->ksize == 3 && d->in_stride == 1 && d->H == d->OH && d->W == d->OW && !no_w3 && d->cin % 256 != 0)
    return wgrad3(d, reinterpret_cast<cudaStream_t>(s));
  WgradKParams kp;
  memset(&kp, 0, sizeof(kp));
  tile_geometry(64, d->OH, d->OW, &kp.bw, &kp.bh, &kp.bni);
  kp.tilesW = (d->OW + kp.bw - 1) / kp.bw;
  kp.tilesH = (d->OH + kp.bh - 1) / kp.bh;
  const int tilesN = (d->N + kp.bni - 1) / kp.bni;
  kp.ntiles = kp.tilesW * kp.tilesH * tilesN;
  kp.es = d->in_stride; kp.ksize = d->ksize; kp.c0 = d->c0;
  kp.dw = d->dw; kp.dw_ld = d->dw_ld; kp.taps = d->ksize * d->ksize; kp.ci_off = d->ci_off;
  kp.cin_real = d->cin_real > 0 ? d->cin_real : d->cin; kp.cout = d->cout;
  kp.dbias = d->dbias;
  {
    const uint64_t C = d->ldy;
    uint64_t dims[4] = {C, (uint64_t)d->OW, (uint64_t)d->OH, (uint64_t)d->N};
    uint64_t str[3] = {C * 2, C * 2 * d->OW, C * 2 * (uint64_t)d->OW * d->OH};
    uint32_t box[4] = {64, (uint32_t)kp.bw, (uint32_t)kp.bh, (uint32_t)kp.bni};
    int rc = make_tmap_bf16(&kp.tmDy, d->dy, 4, dims, str, box, nullptr);
    if (rc) return rc;
  }
  {
    const uint64_t C = d->src_c;
    const int es = d->in_stride;
    uint64_t dims[4] = {C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[3] = {C * 2, C * 2 * d->W, C * 2 * (uint64_t)d->W * d->H};
    uint32_t box[4] = {64, (uint32_t)(kp.bw * es), (uint32_t)(kp.bh * es), (uint32_t)kp.bni};
    uint32_t est[4] = {1, (uint32_t)es, (uint32_t)es, 1};
    int rc = make_tmap_bf16(&kp.tmX, d->src, 4, dims, str, box, est);
    if (rc) return rc;
  }
  // N = 256 takes the MMA off the shared-memory bandwidth limit (A 4 KB + B 8 KB per 128 cycles); use it when Cin allows
  const int bn = (d->cin >= 256 && d->cin % 256 == 0) ? 256 : d->cin > 64 ? 128 : 64;
  const int co_tiles = (d->cout + 127) / 128, ci_tiles = (d->cin + bn - 1) / bn;
  int splits = d->splits;
  if (splits <= 0) {
    // one CTA per SM (the 192 KB ring excludes co-residency): fill exactly one wave when the tile count allows
    const int base = co_tiles * ci_tiles * kp.taps;
    splits = base >= kNumSMs ? 1 : kNumSMs / base;
    if (splits > kp.ntiles) splits = kp.ntiles;
    if (splits < 1) splits = 1;
  }
  kp.tiles_per_split = (kp.ntiles + splits - 1) / splits;
  splits = (kp.ntiles + kp.tiles_per_split - 1) / kp.tiles_per_split;
  dim3 grid(co_tiles, ci_tiles * kp.taps, splits);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(s);
  if (bn == 256) return launch_wgrad<256, 4>(kp, grid, st);
  if (bn == 128) return launch_wgrad<128, 6>(kp, grid, st);
  return launch_wgrad<64, 6>(kp, grid, st);
}
